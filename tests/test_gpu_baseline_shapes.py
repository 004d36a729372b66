"""GPU vs the UNMODIFIED reference (oracle/_ref) at the BASELINE.json shapes, with the tcgen05
TF32 candidate filter forced on for every round it can serve.

d = 96 IP (DEEP, k = 100), d = 128 L2 (SIFT), d = 200 normalised IP (TEXT), d = 960 L2 (GIST):
these are the shapes where the filter accumulates over several 32-float k-chunks
(tcfilter.cu: (c | k) != 0), where the last chunk is partial and filled by TMA out-of-bounds
zeros (d = 200), and where the query tile shrinks (d = 960).  Every case asserts
`tc_rounds > 0`, distances / my_nprobe / traces bit-equal and labels equal up to exact ties
(Auncel/IndexIVFFlat.cpp:117-137, tests/test_lowlevel_ivf.cpp:82-220 pattern).

The adversarial cases attack the filter's error bound (index.cu: c1/c2/c3): a false negative
would silently drop a true neighbour, so results are compared with the exact reference on data
chosen to maximise the TF32 truncation error relative to the top-k threshold.
"""
import os

import numpy as np
import pytest

import auncel_b200 as ab
from oracle import oracle as O
from tests.util import assert_results_match, mixture

pytestmark = pytest.mark.gpu

NB, NLIST, TS, SES = 200_000, 1024, 300, 300
THREADS = max(1, min(32, os.cpu_count() or 1))

SHAPES = {
    "deep96_ip": dict(d=96, metric=O.IP, normalize=True, K=100, qk=10),
    "sift128_l2": dict(d=128, metric=O.L2, normalize=False, K=100, qk=10),
    "text200_ip": dict(d=200, metric=O.IP, normalize=True, K=100, qk=10),
    "gist960_l2": dict(d=960, metric=O.L2, normalize=False, K=100, qk=10),
}


def _need_ref():
    if not O.have_ref():
        pytest.skip("oracle/_ref/libauncel_ref.so not present")
    O.RefIndex.set_blas_threshold(1 << 30)  # exact-difference coarse path on both sides (utils.cpp:622)


@pytest.fixture(scope="module", params=list(SHAPES))
def pair(request):
    _need_ref()
    c = SHAPES[request.param]
    d, metric = c["d"], c["metric"]
    nb = NB if d < 900 else NB // 2  # keeps the host copy of GIST below 400 MB
    xb = mixture(11, nb, d, c["normalize"])
    xq = mixture(22, TS + SES + 200, d, c["normalize"])
    ix = ab.IndexIVFFlat(d, NLIST, metric)
    ix.set_tune_mode()
    ix.train(xb[:: max(1, nb // (64 * NLIST))], niter=4)
    ix.set_tune_off()
    ix.add(xb)
    R = O.RefIndex(d, NLIST, metric)
    R.set_centroids(ix.centroids())
    R.add(xb, ids=np.arange(nb, dtype=np.int64), list_no=ix.assign(xb))
    assert np.array_equal(R.list_sizes(), ix.list_sizes())
    assert np.array_equal(R.interdis(), ix.interdis_cem())
    # queries the reference can serve in IP mode: first list holds >= K vectors, similarity <= 1
    ok = np.ones(len(xq), bool)
    if metric == O.IP:
        dis, keys = R.coarse(xq, 1)
        ok = (R.list_sizes()[keys[:, 0]] >= c["K"]) & (dis[:, 0] <= 1.0)
    q = xq[ok][: TS + SES]
    assert len(q) == TS + SES
    ix.set_option("tensor_core_filter", 2)
    yield c, ix, R, xb, q
    R.close()


@pytest.mark.parametrize("nprobe", [16, 64])
def test_fixed_nprobe_tc(pair, nprobe):
    c, ix, R, xb, q = pair
    ix.nprobe = nprobe
    D, I = ix.search(q[TS:], c["K"])
    st = ix.stats()
    assert st["tc_rounds"] > 0, st
    Dr, Ir = R.search_fixed(q[TS:], c["K"] + 1, nprobe, threads=THREADS)
    Dn, Dr, Ir = Dr[:, -1], np.ascontiguousarray(Dr[:, :-1]), np.ascontiguousarray(Ir[:, :-1])
    assert np.array_equal(D, Dr)
    assert_results_match(D, I, Dr, Ir, what=f"fixed nprobe={nprobe}", D_next=Dn)
    assert (I == Ir).mean() > 0.999


def test_calibration_and_bounded_tc(pair):
    c, ix, R, xb, q = pair
    K, qk = c["K"], c["qk"]
    # ground truth: exhaustive search; a sample of rows is checked against the reference
    ix.nprobe = NLIST
    gD, gI = ix.search(q, K)
    Dr, Ir = R.search_fixed(q[:24], K, NLIST, threads=THREADS)
    assert np.array_equal(gD[:24], Dr)
    es = ab.Error_sys(ix, TS + SES, K)
    es.set_gt(gD, gI)
    es.sys_train(TS, q)
    R.es_create(gD, gI)
    R.sys_train(TS, q)
    got, ref = ix.traces(), R.traces()
    assert len(got) == len(ref) == 8
    for a, b in zip(got, ref):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for mult, stdm, eb in [(7.9, 6.0, 0.1), (2.5, 2.0, 0.2)]:
        acc = np.full(TS + SES, 1.0 - eb, np.float32)
        acc[::5] = 1.0 - eb / 2
        R.set_queries(qk, SES, q, acc, mult, stdm, profile=True)
        Dr, Ir = R.es_search(TS, SES, threads=THREADS)
        ref_np, ref_tr = R.my_nprobe(TS, SES), R.t_recalls(TS, SES)
        R.clear_my_nprobe()
        es.set_topk(qk)
        es.setparam(mult, stdm)
        es.set_queries(SES, q, acc, TS + SES)
        es.profile = True
        D, I = es.search(TS)
        st = ix.stats()
        assert st["tc_rounds"] > 0 and st["err_bits"] == 0, st
        assert np.array_equal(es.my_nprobe[TS:], ref_np), (mult, stdm, eb)
        assert np.array_equal(D, Dr)
        assert_results_match(D, I, Dr, Ir, what=f"bounded {mult},{stdm},{eb}")
        assert np.array_equal(es.t_recalls[TS:], ref_tr)


def test_tc_rounds_audited(pair):
    """Direct statement of the filter's contract: with option "tc_audit" every tensor-core round is
    redone by the exact scan and the two candidate pools are compared slot by slot -- no pair the
    reference's strict test (IndexIVFFlat.cpp:129) accepts may be missing from the filter's output."""
    c, ix, R, xb, q = pair
    ix.set_option("tc_audit", 1)
    try:
        for nprobe in (32, 256):
            ix.nprobe = nprobe
            ix.search(q[TS:], c["K"])
            st = ix.stats()
            assert st["tc_rounds"] > 0 and st["tc_audit_slots"] > 0, st
            assert st["tc_audit_bad"] == 0, st
    finally:
        ix.set_option("tc_audit", 0)


def test_streamed_query_tiles(pair):
    """d > 256: the filter streams the query tile through its stage ring (256 queries per list pass) instead of
    keeping 32 queries resident.  Forced on for every filter round (option tc_stream_min = 0), audited against
    the exact scan, results equal to the resident-tile run bit for bit."""
    c, ix, R, xb, q = pair
    if c["d"] <= 256:
        pytest.skip("query tiles stay resident at d <= 256")
    ix.nprobe = 64
    D1, I1 = ix.search(q[TS:], c["K"])
    ix.set_option("tc_stream_min", 0)
    ix.set_option("tc_audit", 1)
    try:
        D2, I2 = ix.search(q[TS:], c["K"])
        st = ix.stats()
        assert st["tc_rounds"] > 0 and st["tc_audit_slots"] > 0 and st["tc_audit_bad"] == 0, st
        assert np.array_equal(D1, D2)
        assert_results_match(D2, I2, D1, I1, what="streamed vs resident query tiles")
    finally:
        ix.set_option("tc_audit", 0)
        ix.set_option("tc_stream_min", 96)


@pytest.mark.parametrize("kernel", [2, 3])
def test_alternative_filter_kernels(pair, kernel):
    """Option "tc_kernel" = 2 (tcfilter2.cu: queries resident in TMEM, TS-mode tcgen05.mma, the lists
    streaming through a 208 KB ring whose length exceeds many tiles -- the lap-ahead case) and = 3
    (tcfilter3.cu: CTA pairs, tcgen05 cta_group::2, cross-CTA barriers): audited against the exact
    scan, and the search results equal those of the default kernel bit for bit."""
    c, ix, R, xb, q = pair
    if kernel == 2 and c["d"] > 256:
        pytest.skip("tcfilter2.cu serves d <= 256")
    ix.nprobe = 64
    D1, I1 = ix.search(q[TS:], c["K"])
    ix.set_option("tc_kernel", kernel)
    ix.set_option("tc_audit", 1)
    try:
        D2, I2 = ix.search(q[TS:], c["K"])
        st = ix.stats()
        assert st["tc_rounds"] > 0 and st["tc_audit_slots"] > 0 and st["tc_audit_bad"] == 0, st
        assert np.array_equal(D1, D2)
        assert_results_match(D2, I2, D1, I1, what=f"tc_kernel {kernel} vs 1")
    finally:
        ix.set_option("tc_audit", 0)
        ix.set_option("tc_kernel", 0)


# ----------------------------------------------------------------------------- adversarial
def _low_bits(x):
    """Set the 13 mantissa bits the tensor core ignores: the largest truncation error a TF32
    operand can carry, with one sign for every coordinate so the errors add up coherently."""
    y = np.abs(x).astype(np.float32)
    return (y.view(np.uint32) | np.uint32(0x1FFF)).view(np.float32)


def _adversarial_sets(d):
    rng = np.random.default_rng(5)
    n, nq = 40_000, 256
    base = rng.standard_normal((n, d), dtype=np.float32)
    qs = rng.standard_normal((nq, d), dtype=np.float32)
    out = {}
    # (1) common offset: ||q|| ||v|| grows against the top-k threshold, the bound must widen with it
    for off in (3.0, 30.0, 300.0):
        out[f"offset{off:g}"] = (base + np.float32(off), qs + np.float32(off))
    # (2) worst-case truncation: all operands positive with the ignored bits set
    out["lowbits"] = (_low_bits(base + 2.0), _low_bits(qs + 2.0))
    out["lowbits_far"] = (_low_bits(base * 0.05 + 20.0), _low_bits(qs * 0.05 + 20.0))
    # (3) near-duplicates of the queries: true neighbours a few ulps away, and exact copies (ties)
    dup = base.copy()
    for i in range(nq):
        for j in range(6):
            v = qs[i].copy()
            v[(i + j) % d] = np.nextafter(v[(i + j) % d], np.float32(np.inf if j % 2 else -np.inf))
            dup[(i * 97 + j * 1013) % n] = v
        dup[(i * 131 + 7) % n] = qs[i]
    out["near_dup"] = (dup + np.float32(10.0), qs + np.float32(10.0))
    out["near_dup_big"] = ((dup + np.float32(10.0)) * np.float32(1000.0), (qs + np.float32(10.0)) * np.float32(1000.0))
    return out


@pytest.mark.parametrize("metric", [O.L2, O.IP])
@pytest.mark.parametrize("d", [128, 200])
def test_tc_filter_adversarial(metric, d):
    _need_ref()
    nlist, K = 64, 100
    used_tc = 0
    for name, (xb, xq) in _adversarial_sets(d).items():
        cent = xb[:: len(xb) // nlist][:nlist].copy()
        ix = ab.IndexIVFFlat(d, nlist, metric)
        ix.set_centroids(cent, compute_interdis=False)
        ix.add(xb)
        R = O.RefIndex(d, nlist, metric)
        R.set_centroids(cent)
        R.add(xb, ids=np.arange(len(xb), dtype=np.int64), list_no=ix.assign(xb))
        for nprobe in (12, 48):
            Dr, Ir = R.search_fixed(xq, K + 1, nprobe, threads=THREADS)  # one extra column: ties at the k-th edge
            Dn, Dr, Ir = Dr[:, K], np.ascontiguousarray(Dr[:, :K]), np.ascontiguousarray(Ir[:, :K])
            for mode in (2, 0):
                ix.set_option("tensor_core_filter", mode)
                ix.set_pool_budget((2 << 20) if mode == 2 else (1 << 30))  # small budget: many rounds, many thresholds
                ix.nprobe = nprobe
                ix.set_option("tc_audit", 1 if mode == 2 else 0)
                D, I = ix.search(xq, K)
                st = ix.stats()
                assert st["tc_audit_bad"] == 0, (name, nprobe, st)
                assert np.array_equal(D, Dr), (name, nprobe, mode, st)
                assert_results_match(D, I, Dr, Ir, what=f"{name} nprobe={nprobe} tc={mode}", D_next=Dn)
                if mode == 2:
                    assert st["rounds"] > 1 and st["tc_rounds"] + st["tc_fallbacks"] > 0, (name, st)
                    used_tc += st["tc_rounds"] > 0
        R.close()
        del ix
    assert used_tc > 0
