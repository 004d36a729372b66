"""GPU: the two remaining stop rules / result containers of the path against the
CPU restatement (which tests/test_oracle_vs_ref.py pins to the unmodified reference):

* IndexIVF::range_search (IndexIVF.cpp:741-860, scan_codes_range IndexIVFFlat.cpp:139-155): lims,
  distances and labels equal INCLUDING the order inside a query (probe rank, then in-list order);
* Error_sys::time_search (profile.cpp:229-244; cut IndexIVF.cpp:545-549) on the modelled clock,
  and the flag error_pro::time_tune staying set for the next tuned search (:242)."""
import numpy as np
import pytest

import auncel_b200 as ab
from auncel_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _pair(metric, d, nlist, nb, n_centers=40):
    norm = metric == O.IP
    xb = synth.clustered(3, nb, d, n_centers, 0.3, normalize=norm)
    cent = synth.clustered(53, nlist, d, n_centers, 0.3, normalize=norm)
    orc = O.OracleIndex(d, nlist, metric)
    orc.set_centroids(cent)
    orc.add(xb)
    ix = ab.IndexIVFFlat(d, nlist, metric)
    ix.set_centroids(cent)
    ix.add(xb)
    return xb, orc, ix


@pytest.mark.parametrize("metric,d", [(O.L2, 20), (O.L2, 7), (O.IP, 100), (O.L2, 128)])
def test_range_search_equals_oracle(metric, d):
    nlist, nb = 48, 6000
    xb, orc, ix = _pair(metric, d, nlist, nb)
    xq = synth.clustered(9, 61, d, 40, 0.3, normalize=metric == O.IP)
    gD, _ = orc.search_fixed(xq, 30, nlist)
    for col in (0, 5, 29):
        rad = float(np.median(gD[:, col]))
        for nprobe in (1, 6, nlist):
            ix.nprobe = nprobe
            lims, D, I = ix.range_search(xq, rad)
            l2, D2, I2 = orc.range_search(xq, rad, nprobe)
            assert np.array_equal(lims, l2), (col, nprobe)
            assert np.array_equal(D, D2) and np.array_equal(I, I2), (col, nprobe)
            st = ix.stats()
            assert st["ndis"] == orc.last_stats["ndis"] and st["nlist"] == orc.last_stats["nlist"]
    assert lims[-1] > 0
    # a radius nothing passes, and an empty batch
    lims, D, I = ix.range_search(xq, -1.0 if metric == O.L2 else 1e30)
    assert lims[-1] == 0 and len(D) == 0 and np.all(lims == 0)
    lims, D, I = ix.range_search(xq[:0], 1.0)
    assert lims.tolist() == [0]


def test_range_search_equals_reference():
    if not O.have_ref():
        pytest.skip("oracle/_ref not present")
    O.RefIndex.set_blas_threshold(1 << 30)
    d, nlist, nb = 32, 64, 20000
    xb, orc, ix = _pair(O.L2, d, nlist, nb)
    R = O.RefIndex(d, nlist, O.L2)
    R.set_centroids(ix.centroids())
    R.add(xb)
    xq = synth.clustered(9, 200, d, 40, 0.3)
    gD, _ = R.search_fixed(xq, 10, nlist)
    ix.nprobe = 16
    lims, D, I = ix.range_search(xq, float(np.median(gD[:, 9])))
    l2, D2, I2 = R.range_search(xq, float(np.median(gD[:, 9])), 16)
    assert np.array_equal(lims, l2) and np.array_equal(D, D2) and np.array_equal(I, I2)
    R.close()


@pytest.mark.parametrize("model", [(1_000_000, 0), (3, 117), (0, 2500)])
def test_time_search_equals_oracle(model):
    d, nlist, nb, k, qk = 16, 1024, 30000, 10, 4
    xb, orc, ix = _pair(O.L2, d, nlist, nb, n_centers=300)
    xq = synth.clustered(9, 60, d, 300, 0.3)
    gD, gI = orc.search_fixed(xq, k, nlist)
    unit = {1_000_000: 1300.0, 3: 0.004, 0: 0.05}[model[0]]
    bud = np.array([(i % 11 + 2) * unit for i in range(60)], np.float32)
    ix.set_time_model(*model)
    es = ab.Error_sys(ix, 60, k)
    es.set_gt(gD, gI)
    es.sys_train(30, xq)
    orc.calibrate(xq[:30], gD[:30])
    es.set_topk(qk)
    es.set_queries(30, xq, bud, 60)
    D, I = es.time_search(30)
    D2, I2 = orc.search_timed(xq[30:], k, bud, model[0], model[1], offset=30)
    assert np.array_equal(D, D2) and np.array_equal(I, I2)
    assert not np.array_equal(D, gD[30:]), "the budget must cut the scan short"
    assert ix.stats()["ndis"] == orc.last_stats["ndis"]
    # latency mode: one query per call gives the same rows
    for i in range(30, 36):
        D1, I1 = es.time_search(i, 1)
        assert np.array_equal(D1[0], D[i - 30])
    # the flag stays set (profile.cpp:242): the next tuned search is cut by the budget as well
    acc = np.full(60, 0.9, np.float32)
    acc[::2] = 4000.0 * unit / 1300.0
    es.setparam(2.0, 1.0)
    es.set_queries(30, xq, acc, 60)
    D, I = es.search(30)
    orc.multipler, orc.std_m = 2.0, 1.0
    D2, I2, mynp, _ = orc.search_bounded(xq[30:], k, qk, acc, gt_D=gD, offset=30, time_model=model)
    assert np.array_equal(D, D2) and np.array_equal(I, I2)
    assert np.array_equal(es.my_nprobe[30:], mynp[30:])
    ix.time_tune = False
    es.set_queries(30, xq, np.full(60, 0.9, np.float32), 60)
    D, I = es.search(30)
    D2, I2, mynp, _ = orc.search_bounded(xq[30:], k, qk, np.full(60, 0.9, np.float32), gt_D=gD, offset=30)
    assert np.array_equal(D, D2) and np.array_equal(es.my_nprobe[30:], mynp[30:])
