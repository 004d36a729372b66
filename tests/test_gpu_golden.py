"""GPU: the CUDA path, through the C ABI, against the fixtures generated from the unmodified
reference (tests/golden/make_golden.py).  Distances and my_nprobe bit-exact; ids identical
except at exact distance ties (north_star tolerance)."""
import hashlib

import numpy as np
import pytest

import auncel_b200 as ab
from tests.util import PARAMS, assert_results_match, golden_case, golden_traces

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module", params=["l2_d16", "ip_d24"])
def case(request):
    c, g, xb, q = golden_case(request.param)
    ix = ab.IndexIVFFlat(c["d"], c["nlist"], c["metric"])
    ix.set_centroids(g["centroids"], compute_interdis=True)
    ix.add(xb)
    return c, g, xb, q, ix


def test_interdis_assign_lists(case):
    c, g, xb, q, ix = case
    assert sha(ix.interdis_cem()) == str(g["interdis_sha"])
    assert sha(ix.assign(xb)) == str(g["assign_sha"])
    assert np.array_equal(ix.list_sizes(), g["list_sizes"])
    assert ix.ntotal == len(xb)


def test_coarse_full_ranking(case):
    c, g, xb, q, ix = case
    dis, keys = ix.coarse_search(q[:8], c["nlist"])
    assert np.array_equal(dis, g["coarse_dis"])
    assert np.array_equal(keys, g["coarse_keys"])


@pytest.mark.parametrize("nprobe", [1, 4, 16])
def test_fixed_nprobe(case, nprobe):
    c, g, xb, q, ix = case
    ix.nprobe, ix.max_codes = nprobe, 0
    D, I = ix.search(q, c["k"])
    assert np.array_equal(D, g[f"fixed_D_{nprobe}"])
    assert_results_match(D, I, g[f"fixed_D_{nprobe}"], g[f"fixed_I_{nprobe}"], what=f"nprobe={nprobe}")
    assert (I == g[f"fixed_I_{nprobe}"]).mean() > 0.999


def test_max_codes(case):
    c, g, xb, q, ix = case
    ix.nprobe, ix.max_codes = 16, 300
    D, I = ix.search(q, c["k"])
    ix.max_codes = 0
    assert np.array_equal(D, g["fixed_D_16_mc300"])
    assert_results_match(D, I, g["fixed_D_16_mc300"], g["fixed_I_16_mc300"], what="max_codes")


def test_exhaustive_is_ground_truth(case):
    c, g, xb, q, ix = case
    ix.nprobe = c["nlist"]
    D, I = ix.search(q[:64], c["k"])
    assert np.array_equal(D, g["gt_D"][:64])


def test_calibration_traces(case):
    c, g, xb, q, ix = case
    ts = int(g["ts"])
    es = ab.Error_sys(ix, len(q), c["k"])
    es.set_gt(g["gt_D"], g["gt_I"])
    es.sys_train(ts, q)
    got, ref = ix.traces(), golden_traces(g)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_bounded_search(case, pi):
    c, g, xb, q, ix = case
    ts, ses = int(g["ts"]), int(g["ses"])
    mult, stdm, eb = PARAMS[pi]
    ix.set_error_model(golden_traces(g), mult, stdm)
    es = ab.Error_sys(ix, len(q), c["k"])
    es.set_gt(g["gt_D"], g["gt_I"])
    es.is_trained = True
    es.set_topk(c["qk"])
    es.set_queries(ses, q, g[f"b{pi}_acc"], ts + ses)
    es.profile = True
    D, I = es.search(ts)
    assert ix.stats()["err_bits"] == 0
    assert np.array_equal(es.my_nprobe[ts:], g[f"b{pi}_my_nprobe"])
    assert np.array_equal(D, g[f"b{pi}_D"])
    assert_results_match(D, I, g[f"b{pi}_D"], g[f"b{pi}_I"], what=f"bounded {pi}")
    assert np.array_equal(es.t_recalls[ts:], g[f"b{pi}_t_recalls"])
    # latency mode (eval/bound.cpp:390-396): one query per call gives the same answer
    es.set_queries(ses, q, g[f"b{pi}_acc"], ts + ses)
    for i in range(ts, ts + 16):
        D1, I1 = es.search(i, 1)
        assert np.array_equal(D1[0], D[i - ts])
    # stale my_nprobe is replayed (IndexIVF.cpp:615-632): searching again changes nothing
    D2, I2 = es.search(ts, 16)
    assert np.array_equal(D2, D[:16])


def test_small_pool_budget_same_answer(case):
    c, g, xb, q, ix = case
    ts, ses = int(g["ts"]), int(g["ses"])
    ix.set_error_model(golden_traces(g), *PARAMS[2][:2])
    acc = g["b2_acc"][ts:]
    ix.set_pool_budget(1 << 20)
    D, I, mynp = ix.search_bounded(q[ts:], c["k"], c["qk"], acc)
    ix.set_pool_budget(1 << 30)
    assert np.array_equal(mynp, g["b2_my_nprobe"])
    assert np.array_equal(D, g["b2_D"])


@pytest.mark.parametrize("pi", [0, 2])
def test_bounded_search_tensor_core_filter(case, pi):
    """Rounds served by the tcgen05 TF32 filter + exact rerank give the same bits."""
    c, g, xb, q, ix = case
    ts, ses = int(g["ts"]), int(g["ses"])
    mult, stdm, eb = PARAMS[pi]
    ix.set_error_model(golden_traces(g), mult, stdm)
    ix.set_option("tensor_core_filter", 2)
    try:
        D, I, mynp = ix.search_bounded(q[ts:], c["k"], c["qk"], g[f"b{pi}_acc"][ts:])
        st = ix.stats()
        assert st["tc_rounds"] > 0, st  # (an overflowing round falls back to the exact scan: tc_fallbacks)
        assert np.array_equal(mynp, g[f"b{pi}_my_nprobe"])
        assert np.array_equal(D, g[f"b{pi}_D"])
        assert_results_match(D, I, g[f"b{pi}_D"], g[f"b{pi}_I"], what="tc bounded")
        ix.nprobe = 16
        D, I = ix.search(q, c["k"])
        assert ix.stats()["tc_rounds"] > 0
        assert np.array_equal(D, g["fixed_D_16"])
        assert_results_match(D, I, g["fixed_D_16"], g["fixed_I_16"], what="tc fixed")
    finally:
        ix.set_option("tensor_core_filter", 1)
