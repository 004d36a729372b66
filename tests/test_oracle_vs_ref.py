"""CPU, only where /root/reference was compiled (oracle/_ref): the restatement against
the live reference on inputs the fixtures do not cover (odd d, tiny lists, k > list)."""
import numpy as np
import pytest

from auncel_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built here")


def test_low_level_kernels():
    lib = O.ref()
    x = synth.clustered(1, 64, 40)
    y = synth.clustered(2, 64, 40)
    for d in (1, 2, 3, 4, 5, 7, 8, 13, 32, 39, 40):
        for i in range(64):
            a, b = np.ascontiguousarray(x[i, :d]), np.ascontiguousarray(y[i, :d])
            assert O.fvec_L2sqr(a, b) == lib.ref_fvec_L2sqr(O._p(a, O._f), O._p(b, O._f), d)
            assert O.fvec_inner_product(a, b) == lib.ref_fvec_inner_product(O._p(a, O._f), O._p(b, O._f), d)
    err = O.C.c_int(0)
    for a, b, c in [(1.0, 2.0, 3.0), (0.5, 0.5, 1e-3), (3.25, 7.5, 0.125), (1e4, 2e4, 5e3)]:
        assert O.orc().orc_cosine_theorem(a, b, c, O.C.byref(err)) == lib.ref_cosine_theorem(a, b, c)


@pytest.mark.parametrize("metric,d", [(O.L2, 10), (O.L2, 32), (O.IP, 12)])
def test_fixed_search_small(metric, d):
    O.RefIndex.set_blas_threshold(1 << 30)
    nlist, nb, nq, k = 32, 1000, 50, 10
    xb = synth.clustered(5, nb, d, 20, 0.3, normalize=metric == O.IP)
    xq = synth.clustered(6, nq, d, 20, 0.3, normalize=metric == O.IP)
    R = O.RefIndex(d, nlist, metric)
    R.train(xb, niter=3)
    R.add(xb)
    orc = O.OracleIndex(d, nlist, metric)
    orc.set_centroids(R.centroids())
    orc.add(xb)
    for nprobe in (1, 4, 32):
        D1, I1 = R.search_fixed(xq, k, nprobe)
        D2, I2 = orc.search_fixed(xq, k, nprobe)
        assert np.array_equal(D1, D2) and np.array_equal(I1, I2)
    # k larger than what nprobe=1 can return: -1 / FLT_MAX padding (Heap.h:295-322)
    D1, I1 = R.search_fixed(xq, 100, 1)
    D2, I2 = orc.search_fixed(xq, 100, 1)
    assert np.array_equal(D1, D2) and np.array_equal(I1, I2)
    assert (I1 == -1).any()
    R.close()


def test_shards_match_merge_tables():
    O.RefIndex.set_blas_threshold(1 << 30)
    d, nlist, nb, nq, k = 16, 32, 3000, 40, 10
    xb = synth.clustered(7, nb, d, 20, 0.3)
    xq = synth.clustered(8, nq, d, 20, 0.3)
    full = O.RefIndex(d, nlist, O.L2)
    full.train(xb, niter=3)
    full.add(xb)
    cent = full.centroids()
    subs = []
    for s in range(3):
        r = O.RefIndex(d, nlist, O.L2)
        r.set_centroids(cent)
        O._ck(O.ref().ref_copy_subset_to(full.h, r.h, 1, 3, s))  # id % 3 == s
        subs.append(r)
    Dm, Im = O.ref_shards_search(subs, xq, k, 4)
    Df, If = full.search_fixed(xq, k, 4)
    assert np.array_equal(Dm, Df)  # tests/test_merge.cpp invariant
    allD = np.stack([r.search_fixed(xq, k, 4)[0] for r in subs])
    allI = np.stack([r.search_fixed(xq, k, 4)[1] for r in subs])
    D2, I2 = O.merge_tables(O.L2, allD, allI)
    assert np.array_equal(D2, Dm) and np.array_equal(I2, Im)
    for r in subs + [full]:
        r.close()


@pytest.mark.parametrize("metric,d,nb,k,qk,seed", [(O.L2, 12, 40000, 16, 4, 7), (O.L2, 40, 30000, 30, 10, 8),
                                                    (O.IP, 20, 90000, 10, 3, 9)])
def test_bounded_search_live(metric, d, nb, k, qk, seed):
    """Calibration (Error_sys::sys_train -> traces) and error-bounded search of the restatement
    against the unmodified reference run right here, on configurations the stored fixtures do not
    hold: other dimensions, heap widths, query_topk and seeds.  Same flow as tests/golden/make_golden.py."""
    O.RefIndex.set_blas_threshold(1 << 30)
    nlist, ts, ses = 1024, 200, 100
    norm = metric == O.IP
    xb = synth.clustered(seed, nb, d, 400 if not norm else 2500, 0.32 if not norm else 0.45, normalize=norm)
    xq = synth.clustered(seed + 100, ts + ses, d, 400 if not norm else 2500, 0.32 if not norm else 0.45, normalize=norm)
    R = O.RefIndex(d, nlist, metric)
    R.train(xb, niter=4)
    cent = R.centroids()
    R.add(xb)
    ok = np.ones(len(xq), bool)
    if norm:  # the reference throws when the first list holds < k vectors or a similarity exceeds 1
        sizes = R.list_sizes()
        dis, keys = R.coarse(xq, 1)
        ok = (sizes[keys[:, 0]] >= k) & (dis[:, 0] <= 1.0)
    cal = np.flatnonzero(ok[:ts])
    tst = np.flatnonzero(ok[ts:]) + ts
    cal, tst = cal[: len(cal) // 10 * 10], tst[: len(tst) // 10 * 10]
    ts, ses = len(cal), len(tst)
    assert ts >= 50 and ses >= 30
    q = xq[np.concatenate([cal, tst])]
    gD, gI = R.search_fixed(q, k, nlist)
    R.es_create(gD, gI)
    R.sys_train(ts, q)
    ref_traces = R.traces()

    orc = O.OracleIndex(d, nlist, metric)
    orc.set_centroids(cent)
    orc.add(xb)
    assert np.array_equal(orc.assign(xb[:5000]), R.assign(xb[:5000]))
    orc.calibrate(q[:ts], gD[:ts])
    assert len(orc.traces) == len(ref_traces)
    for a, b in zip(orc.traces, ref_traces):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for mult, stdm, eb in [(1.0, 1.0, 0.1), (4.0, 3.0, 0.05), (7.9, 6.0, 0.2)]:
        acc = np.full(ts + ses, 1 - eb, np.float32)
        acc[::4] = 1 - eb / 3
        R.set_queries(qk, ses, q, acc, mult, stdm, profile=True)
        D, I = R.es_search(ts, ses, -1)
        orc.multipler, orc.std_m = mult, stdm
        D2, I2, mynp, trec = orc.search_bounded(q[ts:], k, qk, acc, gt_D=gD, offset=ts, profile=True)
        assert orc.last_err == 0
        assert np.array_equal(mynp[ts:], R.my_nprobe(ts, ses)), (mult, stdm, eb)
        assert np.array_equal(D2, D) and np.array_equal(I2, I)
        assert np.array_equal(trec[ts:], R.t_recalls(ts, ses))
        R.clear_my_nprobe()
    R.close()


@pytest.mark.parametrize("metric", [O.L2, O.IP])
def test_range_search_live(metric):
    """IndexIVF::range_search (IndexIVF.cpp:741-860): lims, distances and labels, in scan order."""
    O.RefIndex.set_blas_threshold(1 << 30)
    d, nlist, nb = 20, 64, 8000
    norm = metric == O.IP
    xb = synth.clustered(3, nb, d, 40, normalize=norm)
    xq = synth.clustered(9, 45, d, 40, normalize=norm)
    R = O.RefIndex(d, nlist, metric)
    R.train(xb, niter=2)
    R.add(xb)
    orc = O.OracleIndex(d, nlist, metric)
    orc.set_centroids(R.centroids())
    orc.add(xb)
    gD, _ = R.search_fixed(xq, 20, nlist)
    for col in (0, 3, 19):
        rad = float(np.median(gD[:, col]))
        for nprobe in (1, 9, nlist):
            l1, D1, I1 = R.range_search(xq, rad, nprobe)
            l2, D2, I2 = orc.range_search(xq, rad, nprobe)
            assert np.array_equal(l1, l2) and np.array_equal(D1, D2) and np.array_equal(I1, I2)
    assert l1[-1] > 0
    R.close()


def test_time_search_live():
    """Error_sys::time_search (profile.cpp:229-244) under the test clock (every IndexIVF::time()
    call = +1 s, ref_driver.cpp) against the restatement's clock model (1e6 us per list): the cut of
    IndexIVF.cpp:545-549 -- and the reference leaving error_pro::time_tune set afterwards (:242), so
    that the NEXT tuned search is cut as well."""
    O.RefIndex.set_blas_threshold(1 << 30)
    d, nlist, nb, k, qk = 16, 1024, 30000, 10, 4
    xb = synth.clustered(3, nb, d, 300)
    xq = synth.clustered(9, 60, d, 300)
    R = O.RefIndex(d, nlist, O.L2)
    R.train(xb, niter=2)
    R.add(xb)
    orc = O.OracleIndex(d, nlist, O.L2)
    orc.set_centroids(R.centroids())
    orc.add(xb)
    gD, gI = R.search_fixed(xq, k, nlist)
    R.es_create(gD, gI)
    R.sys_train(30, xq)
    orc.calibrate(xq[:30], gD[:30])
    bud = np.array([(i % 11 + 2) * 1300.0 for i in range(60)], np.float32)  # "ms" of the 1 s/call clock
    R.set_queries(qk, 30, xq, bud, 1.0, 1.0)
    D, I = R.es_time_search(30, 30, keep_flag=True)
    D2, I2 = orc.search_timed(xq[30:], k, bud, 1_000_000, 0, offset=30)
    assert np.array_equal(D, D2) and np.array_equal(I, I2)
    assert not np.array_equal(D, gD[30:])  # the budget did cut the scan
    # sticky flag: a tuned search now reads require_acc both as recall target and as budget
    acc = np.full(60, 0.9, np.float32)
    acc[::2] = 4000.0  # every other query: a target no recall estimate reaches -> only the time cut stops it
    R.set_queries(qk, 30, xq, acc, 2.0, 1.0)
    D, I = R.es_search(30, 30, virtual_clock=True)
    orc.multipler, orc.std_m = 2.0, 1.0
    D2, I2, mynp, _ = orc.search_bounded(xq[30:], k, qk, acc, gt_D=gD, offset=30, time_model=(1_000_000, 0))
    assert np.array_equal(D, D2) and np.array_equal(I, I2)
    assert np.array_equal(mynp[30:], R.my_nprobe(30, 30))
    O.ref().ref_es_time_search  # (flag reset for later users of this handle)
    R.close()
