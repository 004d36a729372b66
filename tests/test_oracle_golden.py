"""CPU: the restatement (oracle/auncel_oracle.c) against the fixtures generated from
the unmodified reference (tests/golden/make_golden.py).  Bit-exact throughout."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import PARAMS, golden_case, golden_traces


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module", params=["l2_d16", "ip_d24"])
def case(request):
    c, g, xb, q = golden_case(request.param)
    orc = O.OracleIndex(c["d"], c["nlist"], c["metric"])
    orc.set_centroids(g["centroids"])
    orc.add(xb)
    return c, g, xb, q, orc


def test_interdis_and_lists(case):
    c, g, xb, q, orc = case
    assert sha(orc.interdis) == str(g["interdis_sha"])
    assert sha(orc.assign(xb)) == str(g["assign_sha"])
    _, off, _ = orc.csr()
    assert np.array_equal(np.diff(off), g["list_sizes"])
    assert np.array_equal(orc.arcos, g["arcos"])


def test_coarse_full_ranking(case):
    c, g, xb, q, orc = case
    dis, keys = orc.coarse(q[:8], c["nlist"])
    assert np.array_equal(dis, g["coarse_dis"]) and np.array_equal(keys, g["coarse_keys"])


@pytest.mark.parametrize("nprobe", [1, 4, 16])
def test_fixed_nprobe(case, nprobe):
    c, g, xb, q, orc = case
    D, I = orc.search_fixed(q, c["k"], nprobe)
    assert np.array_equal(D, g[f"fixed_D_{nprobe}"]) and np.array_equal(I, g[f"fixed_I_{nprobe}"])


def test_max_codes(case):
    c, g, xb, q, orc = case
    D, I = orc.search_fixed(q, c["k"], 16, max_codes=300)
    assert np.array_equal(D, g["fixed_D_16_mc300"]) and np.array_equal(I, g["fixed_I_16_mc300"])


def test_calibration_traces(case):
    c, g, xb, q, orc = case
    ts = int(g["ts"])
    orc.calibrate(q[:ts], g["gt_D"][:ts])
    ref = golden_traces(g)
    assert len(orc.traces) == len(ref)
    for a, b in zip(orc.traces, ref):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_bounded_search(case, pi):
    c, g, xb, q, orc = case
    ts, ses = int(g["ts"]), int(g["ses"])
    orc.traces = golden_traces(g)
    orc.multipler, orc.std_m = PARAMS[pi][0], PARAMS[pi][1]
    D, I, mynp, trec = orc.search_bounded(q[ts:], c["k"], c["qk"], g[f"b{pi}_acc"], gt_D=g["gt_D"],
                                          offset=ts, profile=True)
    assert orc.last_err == 0
    assert np.array_equal(mynp[ts:], g[f"b{pi}_my_nprobe"])
    assert np.array_equal(D, g[f"b{pi}_D"]) and np.array_equal(I, g[f"b{pi}_I"])
    assert np.array_equal(trec[ts:], g[f"b{pi}_t_recalls"])


def test_merge_tables_semantics():
    # IndexShards.cpp:44-105: heads-of-rows heap merge, -1 labels end a row
    rng = np.random.RandomState(0)
    for metric in (O.L2, O.IP):
        nshard, n, k = 3, 17, 6
        Dall = np.sort(rng.rand(nshard, n, k).astype(np.float32), axis=2)
        if metric == O.IP:
            Dall = Dall[:, :, ::-1].copy()
        Iall = rng.randint(0, 1000, size=(nshard, n, k)).astype(np.int64)
        Iall[1, :, 4:] = -1
        Iall[2, 3, :] = -1
        D, I = O.merge_tables(metric, Dall, Iall)
        for i in range(n):
            cand = [(Dall[s, i, j], Iall[s, i, j]) for s in range(nshard) for j in range(k)
                    if Iall[s, i, :j + 1].min() >= 0]
            cand.sort(key=lambda t: t[0] if metric == O.L2 else -t[0])
            exp = [c[0] for c in cand[:k]]
            assert np.allclose(D[i, :len(exp)], exp)
