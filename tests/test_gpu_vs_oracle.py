"""GPU vs the CPU restatement on seeded inputs the fixtures do not cover: dimensions that
are not a multiple of 4 or 32, tiny and empty lists, k larger than a list, ragged batches,
precomputed assignment, add in several calls, shards merge."""
import numpy as np
import pytest

import auncel_b200 as ab
from auncel_b200 import synth
from oracle import oracle as O
from tests.util import assert_results_match

pytestmark = pytest.mark.gpu


def build_pair(metric, d, nlist, nb, seed=3, n_centers=20):
    norm = metric == O.IP
    xb = synth.clustered(seed, nb, d, n_centers, 0.3, normalize=norm)
    cent = synth.clustered(seed + 50, nlist, d, n_centers, 0.3, normalize=norm)
    orc = O.OracleIndex(d, nlist, metric)
    orc.set_centroids(cent)
    orc.add(xb)
    ix = ab.IndexIVFFlat(d, nlist, metric)
    ix.set_centroids(cent)
    ix.add(xb)
    return xb, cent, orc, ix


@pytest.mark.parametrize("metric,d", [(O.L2, 7), (O.L2, 32), (O.L2, 100), (O.IP, 12), (O.IP, 200), (O.L2, 960)])
def test_fixed_various_dims(metric, d):
    nlist, nb = 48, 3000 if d < 500 else 1500
    xb, cent, orc, ix = build_pair(metric, d, nlist, nb)
    xq = synth.clustered(9, 37, d, 20, 0.3, normalize=metric == O.IP)
    assert np.array_equal(ix.assign(xb), orc.assign(xb))
    for nprobe, k in [(1, 1), (3, 10), (48, 100), (7, 128)]:
        ix.nprobe = nprobe
        D, I = ix.search(xq, k)
        D2, I2 = orc.search_fixed(xq, k, nprobe)
        assert np.array_equal(D, D2), (nprobe, k)
        assert_results_match(D, I, D2, I2, what=f"d={d} nprobe={nprobe} k={k}")


def test_empty_and_tiny_lists_and_padding():
    d, nlist = 16, 64
    xb = synth.clustered(4, 40, d, 5, 0.2)  # fewer vectors than lists: most lists empty
    cent = synth.clustered(5, nlist, d, 5, 0.5)
    orc = O.OracleIndex(d, nlist, O.L2)
    orc.set_centroids(cent)
    orc.add(xb)
    ix = ab.IndexIVFFlat(d, nlist, O.L2)
    ix.set_centroids(cent)
    ix.add(xb)
    xq = synth.clustered(6, 5, d, 5, 0.2)
    for nprobe in (1, 8, 64):
        ix.nprobe = nprobe
        D, I = ix.search(xq, 50)
        D2, I2 = orc.search_fixed(xq, 50, nprobe)
        assert np.array_equal(D, D2) and np.array_equal(I, I2)
    assert (I == -1).any() and D.max() == np.finfo(np.float32).max
    D, I = ix.search(xq[:0], 5)
    assert D.shape == (0, 5)


def test_add_in_pieces_with_ids_and_precomputed():
    d, nlist = 20, 32
    xb = synth.clustered(7, 2000, d)
    cent = synth.clustered(8, nlist, d)
    orc = O.OracleIndex(d, nlist, O.L2)
    orc.set_centroids(cent)
    ix = ab.IndexIVFFlat(d, nlist, O.L2)
    ix.set_centroids(cent)
    ids = np.arange(2000, dtype=np.int64)[::-1] * 3 + 7
    pre = orc.assign(xb)
    pre[5::97] = -1  # skipped vectors (IndexIVFFlat.cpp:65-66)
    for a, b in [(0, 700), (700, 701), (701, 2000)]:
        ix.add_core(xb[a:b], ids[a:b], pre[a:b])
        orc.add(xb[a:b], ids[a:b], pre[a:b])
    assert ix.ntotal == 2000
    xq = synth.clustered(10, 33, d)
    ix.nprobe = 5
    D, I = ix.search(xq, 10)
    D2, I2 = orc.search_fixed(xq, 10, 5)
    assert np.array_equal(D, D2) and np.array_equal(I, I2)


def test_error_messages():
    ix = ab.IndexIVFFlat(8, 16)
    with pytest.raises(ab.FaissException):
        ix.add(np.zeros((3, 8), np.float32))  # not trained (IndexIVFFlat.cpp:45)
    ix.set_centroids(synth.clustered(1, 16, 8))
    ix.add(synth.clustered(2, 100, 8))
    with pytest.raises(ab.FaissException):
        ix.search(np.zeros((1, 8), np.float32), 1000)
    with pytest.raises(ab.FaissException):
        ix.search_bounded(np.zeros((1, 8), np.float32), 10, 5, np.ones(1, np.float32))  # no error model
    with pytest.raises(ab.FaissException):
        ab.Error_sys(ix, 15, 10)


def test_bounded_vs_oracle_small_nlist():
    """nlist = 64 (8 stages max, 4 traces): the whole tune path incl. calibration."""
    for metric, d in [(O.L2, 12), (O.IP, 16)]:
        nlist, nb, k, qk = 64, 20000, 16, 4
        xb, cent, orc, ix = build_pair(metric, d, nlist, nb, n_centers=40)
        xq = synth.clustered(21, 300, d, 40, 0.3, normalize=metric == O.IP)
        if metric == O.IP:
            sizes = ix.list_sizes()
            dis, keys = orc.coarse(xq, 1)
            xq = xq[(sizes[keys[:, 0]] >= k) & (dis[:, 0] <= 1.0)]
        xq = xq[:len(xq) // 10 * 10]
        n = len(xq)
        ts = n // 2 // 10 * 10
        gD, gI = orc.search_fixed(xq, k, nlist)
        orc.calibrate(xq[:ts], gD[:ts])
        es = ab.Error_sys(ix, n, k)
        es.set_gt(gD, gI)
        es.sys_train(ts, xq)
        for a, b in zip(ix.traces(), orc.traces):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
        for mult, stdm, eb in [(1.0, 1.0, 0.1), (3.0, 0.5, 0.3)]:
            acc = np.full(n, 1 - eb, np.float32)
            orc.multipler, orc.std_m = mult, stdm
            D2, I2, np2, tr2 = orc.search_bounded(xq[ts:], k, qk, acc, gt_D=gD, offset=ts)
            assert orc.last_err == 0
            ix.set_params(mult, stdm)
            D, I, np1 = ix.search_bounded(xq[ts:], k, qk, acc[ts:])
            assert np.array_equal(np1, np2[ts:])
            assert np.array_equal(D, D2)
            assert_results_match(D, I, D2, I2, what="bounded small")
            assert ix.stats()["ndis"] == orc.last_stats["ndis"]
            assert ix.stats()["nlist"] == orc.last_stats["nlist"]


def test_merge_tables_and_shards():
    d, nlist, nb, k = 16, 32, 4000, 10
    xb, cent, orc, ix = build_pair(O.L2, d, nlist, nb)
    xq = synth.clustered(31, 50, d)
    subs = []
    for s in range(3):
        sub = ab.IndexIVFFlat(d, nlist, O.L2)
        sub.set_centroids(cent, compute_interdis=False)
        ix.copy_subset_to(sub, 1, 3, s)  # id % 3 == s
        sub.nprobe = 4
        subs.append(sub)
    assert sum(s.ntotal for s in subs) == nb
    allD = np.stack([s.search(xq, k)[0] for s in subs])
    allI = np.stack([s.search(xq, k)[1] for s in subs])
    D, I = ab.merge_tables(O.L2, allD, allI)
    D2, I2 = O.merge_tables(O.L2, allD, allI)
    assert np.array_equal(D, D2) and np.array_equal(I, I2)
    ix.nprobe = 4
    Df, If = ix.search(xq, k)
    assert np.array_equal(D, Df)  # tests/test_merge.cpp invariant: shards == single index
    # type 2 (proportional slices) also partitions the index
    subs2 = []
    for s in range(2):
        sub = ab.IndexIVFFlat(d, nlist, O.L2)
        sub.set_centroids(cent, compute_interdis=False)
        ix.copy_subset_to(sub, 2, s * nb // 2, (s + 1) * nb // 2)
        sub.nprobe = 4
        subs2.append(sub)
    assert sum(s.ntotal for s in subs2) == nb
    allD = np.stack([s.search(xq, k)[0] for s in subs2])
    allI = np.stack([s.search(xq, k)[1] for s in subs2])
    D, I = ab.merge_tables(O.L2, allD, allI)
    assert np.array_equal(D, Df)
    # rows with missing results: -1 labels end a row (IndexShards.cpp:62-66)
    allI[1, :, 3:] = -1
    D, I = ab.merge_tables(O.L2, allD, allI)
    D2, I2 = O.merge_tables(O.L2, allD, allI)
    assert np.array_equal(D, D2) and np.array_equal(I, I2)


@pytest.mark.parametrize("metric", [O.L2, O.IP])
@pytest.mark.parametrize("nshard,k", [(2, 10), (5, 33), (8, 100), (17, 7), (64, 128)])
def test_merge_tables_tie_order(metric, nshard, k):
    """Equal distances arriving from different shards leave the merge in the order of the
    reference's size-nshard heap (IndexShards.cpp:86-101 with Heap.h push/pop), not merely sorted:
    labels must be identical, ties included."""
    rng = np.random.default_rng(nshard * 1000 + k)
    n = 257
    vals = rng.integers(0, 12, size=(nshard, n, k)).astype(np.float32)  # few distinct values: ties everywhere
    vals.sort(axis=2)
    if metric == O.IP:
        vals = vals[:, :, ::-1].copy()
    labels = rng.integers(0, 1 << 40, size=(nshard, n, k)).astype(np.int64)
    cut = rng.integers(0, k + 1, size=(nshard, n))
    for s in range(nshard):
        for q in range(0, n, 3):
            labels[s, q, cut[s, q]:] = -1  # rows that end early; some shards contribute nothing
    tr = rng.integers(0, 1000, size=nshard).astype(np.int64)
    D, I = ab.merge_tables(metric, vals, labels, tr)
    D2, I2 = O.merge_tables(metric, vals, labels, tr)
    assert np.array_equal(D, D2)
    assert np.array_equal(I, I2)


@pytest.mark.parametrize("metric,nlist", [(O.L2, 4096), (O.IP, 4096), (O.L2, 16384)])
def test_partial_centroid_ranking_large_nlist(metric, nlist):
    """nlist = 4096 with the partial ranking forced on (option partial_rank = 2): only the best 1024
    centroids are ranked up front, rows are completed when a round reads past them, ties are replayed as
    before.  Distances, my_nprobe and labels against the restatement, which ranks everything like the
    reference (IndexFlat::search with k = nlist).  nlist = 16384 (BASELINE config 5): set_online reads
    ranks up to nlist / 8 + 20 = 2068, past what a partial ranking holds -- the engine must rank those rows
    completely (regression: it read unranked keys and faulted)."""
    d, nb, k, qk = 16, 150000 if nlist == 4096 else 500000, 20, 5
    norm = metric == O.IP
    xb = synth.clustered(3, nb, d, 600, 0.3, normalize=norm)
    cent = synth.clustered(53, nlist, d, 600, 0.3, normalize=norm)
    orc = O.OracleIndex(d, nlist, metric)
    orc.set_centroids(cent)
    orc.add(xb)
    ix = ab.IndexIVFFlat(d, nlist, metric)
    ix.set_centroids(cent)
    ix.add(xb)
    xq = synth.clustered(21, 400, d, 600, 0.3, normalize=norm)
    if norm:
        sizes = ix.list_sizes()
        dis, keys = orc.coarse(xq, 1)
        xq = xq[(sizes[keys[:, 0]] >= k) & (dis[:, 0] <= 1.0)]
    xq = xq[:len(xq) // 20 * 20]
    n = len(xq)
    ts = n // 2
    gD, gI = orc.search_fixed(xq, k, nlist)
    orc.calibrate(xq[:ts], gD[:ts])
    ix.set_option("partial_rank", 2)
    es = ab.Error_sys(ix, n, k)
    es.set_gt(gD, gI)
    es.sys_train(ts, xq)
    for a, b in zip(ix.traces(), orc.traces):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for mult, stdm, eb in [(7.9, 6.0, 0.1), (40.0, 1.0, 0.05)]:  # the second one drives my_nprobe past 1024
        acc = np.full(n, 1 - eb, np.float32)
        es.set_topk(qk)
        es.setparam(mult, stdm)
        es.set_queries(n - ts, xq, acc, n)
        D, I = es.search(ts)
        orc.multipler, orc.std_m = mult, stdm
        D2, I2, mynp, _ = orc.search_bounded(xq[ts:], k, qk, acc, gt_D=gD, offset=ts)
        assert np.array_equal(es.my_nprobe[ts:], mynp[ts:]), (mult, stdm)
        assert np.array_equal(D, D2)
        assert_results_match(D, I, D2, I2, what=f"partial rank {mult}")
        # the same through the full ranking
        ix.set_option("partial_rank", 0)
        es.set_queries(n - ts, xq, acc, n)
        D3, I3 = es.search(ts)
        ix.set_option("partial_rank", 2)
        assert np.array_equal(D3, D) and np.array_equal(I3, I)
    assert mynp[ts:].max() > 1024 or nlist > 4096


def test_tie_exactly_at_the_stop_stage():
    """Equal centroid distances that matter only because a query's stop stage falls between them.
    767 distinct centroids around the queries, then 128 far-away PAIRS of identical centroids (ranks
    767/768, 769/770, ...): with multipler = 900 a query that is satisfied after its first list stops at
    stage 900, i.e. between the twins at ranks 899 and 900 -- which of the two (the one that holds the
    vectors or its empty duplicate) is scanned is decided by the reference's heap order alone.  No query has
    a tie below rank nlist/8 + 21, so nothing is replayed up front: this is the lazy stop-stage replay of
    coarse.cu (collect_ties_kernel / heap_order_kernel, straddle mode)."""
    d, nlist, k, qk = 8, 1024, 16, 4
    rng = np.random.default_rng(17)
    near = rng.standard_normal((767, d)).astype(np.float32)
    far = (rng.standard_normal((128, d)) * 3.0 + 60.0).astype(np.float32)
    cent = np.concatenate([near, np.repeat(far, 2, axis=0), (far[:1] + 500.0)]).astype(np.float32)
    assert cent.shape == (nlist, d)
    xb = np.concatenate([rng.standard_normal((6, d)).astype(np.float32),
                         (np.repeat(far, 40, axis=0) + 0.05 * rng.standard_normal((5120, d))).astype(np.float32)])
    xq = rng.standard_normal((1400, d)).astype(np.float32) * 0.5
    orc = O.OracleIndex(d, nlist, O.L2)
    orc.set_centroids(cent)
    orc.add(xb)
    ix = ab.IndexIVFFlat(d, nlist, O.L2)
    ix.set_centroids(cent)
    ix.add(xb)
    cd, ck = orc.coarse(xq, nlist)
    assert (cd[:, 1:767] == cd[:, :766]).any(1).mean() < 0.1, "(nearly) no ties among the near centroids"
    assert (cd[:, 768::2][:, :64] == cd[:, 767::2][:, :64]).all(), "twins are equidistant"
    gD, gI = orc.search_fixed(xq, k, nlist)
    orc.calibrate(xq[:200], gD[:200])
    es = ab.Error_sys(ix, 1400, k)
    es.set_gt(gD, gI)
    es.sys_train(200, xq)
    for mult in (900.0, 901.0, 902.0):  # stop stage between twins / between two pairs / between twins
        acc = np.full(1400, 0.5, np.float32)
        es.set_topk(qk)
        es.setparam(mult, 1.0)
        es.set_queries(1200, xq, acc, 1400)
        D, I = es.search(200)
        st = ix.stats()
        orc.multipler, orc.std_m = mult, 1.0
        D2, I2, mynp, _ = orc.search_bounded(xq[200:], k, qk, acc, gt_D=gD, offset=200)
        b = mynp[200:].astype(np.int64)
        inside = (b > 767) & (b < nlist)
        straddle = inside & (cd[200:][np.arange(1200), np.minimum(b, nlist - 1)] == cd[200:][np.arange(1200), np.minimum(b, nlist - 1) - 1])
        assert (straddle.sum() > 1000) == (mult != 901.0), "the construction puts stop stages between twins"
        assert np.array_equal(es.my_nprobe[200:], mynp[200:])
        assert st["err_bits"] == 0
        assert st["ndis"] == orc.last_stats["ndis"] and st["nlist"] == orc.last_stats["nlist"], "which twin was scanned"
        assert np.array_equal(D, D2)
        assert_results_match(D, I, D2, I2, what=f"stop-stage ties, multipler {mult}")


def test_shard_group_world_of_one():
    """auncel_shard_group_* with a single shard: no NCCL, the packed table goes straight to the merge."""
    from auncel_b200 import distributed as AD
    d, nlist, nb, k = 16, 32, 4000, 10
    xb, cent, orc, ix = build_pair(O.L2, d, nlist, nb)
    xq = synth.clustered(31, 50, d)
    ix.nprobe = 4
    Df, If = ix.search(xq, k)
    g = AD.NcclShardGroup(ix)
    D, I = g.search(xq, k)
    assert np.array_equal(D, Df) and np.array_equal(I, If)
    st = g.stats()
    assert st["world"] == 1 and st["allgather_bytes"] == 50 * k * 12


def test_train_kmeans_runs_and_searches():
    d, nlist = 16, 64
    xb = synth.clustered(41, 20000, d, 30)
    ix = ab.IndexIVFFlat(d, nlist)
    ix.set_tune_mode()
    ix.train(xb, niter=5)
    ix.set_tune_off()
    assert ix.is_trained
    ix.add(xb)
    cent = ix.centroids()
    orc = O.OracleIndex(d, nlist, O.L2)
    orc.set_centroids(cent)
    assert np.array_equal(orc.interdis, ix.interdis_cem())
    sizes = ix.list_sizes()
    assert sizes.sum() == 20000 and (sizes > 0).mean() > 0.9


def _lattice(seed, n, d, levels=3):
    """small-integer coordinates: squared distances are small integers -> ties everywhere"""
    return np.floor(synth.uniform(seed, (n, d)) * levels).astype(np.float32)


@pytest.mark.parametrize("metric", [O.L2, O.IP])
def test_coarse_ties_follow_reference_heap_order(metric):
    """Equal coarse distances: the reference's probe order is its heap's pop order
    (utils.cpp:417-490 + Heap.h:295-322); the GPU replays it (coarse.cu heap_order_kernel)."""
    d, nlist = 8, 256
    cent = _lattice(1, nlist, d, 4)
    xq = _lattice(2, 64, d, 4)
    orc = O.OracleIndex(d, nlist, metric)
    orc.centroids = cent
    ix = ab.IndexIVFFlat(d, nlist, metric)
    ix.set_centroids(cent, compute_interdis=False)
    for nprobe in (1, 7, 64, 256):
        dis, keys = ix.coarse_search(xq, nprobe)
        odis, okeys = orc.coarse(xq, nprobe)
        assert np.array_equal(dis, odis)
        assert np.array_equal(keys, okeys), nprobe
        assert (odis[:, 1:] == odis[:, :-1]).any() or nprobe == 1


def test_bounded_search_with_coarse_ties():
    """my_nprobe parity when many centroids are equidistant from the query."""
    d, nlist, nb, k, qk = 8, 64, 6000, 16, 4
    cent = _lattice(3, nlist, d, 5) + 0.25 * synth.clustered(4, nlist, d, 10, 0.05)
    cent = np.round(cent * 4) / 4  # quarter-integer grid: still many exact ties
    xb = (_lattice(5, nb, d, 5) + 0.01 * synth.clustered(6, nb, d, 10, 0.3)).astype(np.float32)
    xq = _lattice(7, 200, d, 5)
    orc = O.OracleIndex(d, nlist, O.L2)
    orc.set_centroids(cent)
    orc.add(xb)
    ix = ab.IndexIVFFlat(d, nlist, O.L2)
    ix.set_centroids(cent)
    ix.add(xb, )
    assert np.array_equal(ix.assign(xb), orc.assign(xb))
    cd, ck = orc.coarse(xq, nlist)
    assert (cd[:, 1:] == cd[:, :-1]).any(1).mean() > 0.5
    gD, gI = orc.search_fixed(xq, k, nlist)
    orc.calibrate(xq[:100], gD[:100])
    ix.set_error_model(orc.traces, 1.5, 1.0)
    orc.multipler, orc.std_m = 1.5, 1.0
    acc = np.full(200, 0.9, np.float32)
    D2, I2, np2, _ = orc.search_bounded(xq[100:], k, qk, acc, offset=100)
    D1, I1, np1 = ix.search_bounded(xq[100:], k, qk, acc[100:])
    assert np.array_equal(np1, np2[100:])
    assert np.array_equal(D1, D2)
    ix.nprobe = 5
    D, I = ix.search(xq, k)
    D2, I2 = orc.search_fixed(xq, k, 5)
    assert np.array_equal(D, D2)


def test_cpp_api_mirror_demo():
    """examples/bound_demo.cpp: the reference-API mirror (include/auncel/faiss_api.h) end to end."""
    import os
    import subprocess
    from auncel_b200 import build as b
    exe = b.build_examples()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, cwd=os.path.dirname(exe))
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "DEMO OK" in r.stdout
    assert "Error bound is guaranteed" in r.stdout


@pytest.mark.parametrize("metric,nlist", [(O.L2, 256), (O.L2, 300), (O.IP, 1000), (O.L2, 1024), (O.L2, 2048),
                                          (O.IP, 4096), (O.L2, 5000), (O.L2, 9000)])
def test_coarse_full_ranking_large_nlist(metric, nlist):
    """Full centroid ranking through the register/shuffle/shared-memory sort (coarse.cu
    rank_rows_reg_kernel, P >= 256) and the pipelined tie replay, incl. heap sizes that are not a
    power of two; quantised coordinates force plenty of equal distances."""
    d = 6
    rng = np.random.RandomState(nlist)
    cent = (rng.randint(-6, 7, size=(nlist, d)) / 4.0).astype(np.float32)
    cent[: nlist // 2] += (rng.rand(nlist // 2, d) * 1e-3).astype(np.float32)  # half of them distinct
    xq = (rng.randint(-6, 7, size=(9, d)) / 4.0).astype(np.float32)
    orc = O.OracleIndex(d, nlist, metric)
    orc.centroids = cent
    ix = ab.IndexIVFFlat(d, nlist, metric)
    ix.set_centroids(cent, compute_interdis=False)
    for nprobe in (nlist, max(1, nlist // 3)):
        dis, keys = ix.coarse_search(xq, nprobe)
        odis, okeys = orc.coarse(xq, nprobe)
        assert np.array_equal(dis, odis), (nlist, nprobe)
        assert np.array_equal(keys, okeys), (nlist, nprobe)


@pytest.mark.parametrize("metric,d,nb,k,qk,seed", [(O.L2, 12, 40000, 16, 4, 7), (O.L2, 40, 30000, 30, 10, 8),
                                                    (O.IP, 20, 90000, 10, 3, 9)])
def test_calibration_and_bounded_search_other_configs(metric, d, nb, k, qk, seed):
    """Calibration traces and error-bounded search (D, my_nprobe) on configurations the fixtures do
    not hold -- the ones tests/test_oracle_vs_ref.py::test_bounded_search_live pins the restatement on
    against the live reference (other d, heap widths, query_topk; L2 and IP)."""
    nlist, ts, ses = 1024, 200, 100
    norm = metric == O.IP
    nc, sg = (2500, 0.45) if norm else (400, 0.32)
    xb = synth.clustered(seed, nb, d, nc, sg, normalize=norm)
    xq = synth.clustered(seed + 100, ts + ses, d, nc, sg, normalize=norm)
    cent = synth.clustered(seed + 200, nlist, d, nc, sg, normalize=norm)
    orc = O.OracleIndex(d, nlist, metric)
    orc.set_centroids(cent)
    orc.add(xb)
    ix = ab.IndexIVFFlat(d, nlist, metric)
    ix.set_centroids(cent, compute_interdis=True)
    ix.add(xb)
    ok = np.ones(len(xq), bool)
    if norm:  # queries the reference can run (arccos domain, first list >= k)
        sizes = ix.list_sizes()
        dis, keys = orc.coarse(xq, 1)
        ok = (sizes[keys[:, 0]] >= k) & (dis[:, 0] <= 1.0)
    cal = np.flatnonzero(ok[:ts])
    tst = np.flatnonzero(ok[ts:]) + ts
    cal, tst = cal[: len(cal) // 10 * 10], tst[: len(tst) // 10 * 10]
    ts, ses = len(cal), len(tst)
    assert ts >= 50 and ses >= 30
    q = xq[np.concatenate([cal, tst])]
    gD, gI = orc.search_fixed(q, k, nlist)
    orc.calibrate(q[:ts], gD[:ts])
    es = ab.Error_sys(ix, len(q), k)
    es.set_gt(gD, gI)
    es.sys_train(ts, q)
    got = ix.traces()
    assert len(got) == len(orc.traces)
    for a, b in zip(got, orc.traces):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for mult, stdm, eb in [(1.0, 1.0, 0.1), (4.0, 3.0, 0.05), (7.9, 6.0, 0.2)]:
        acc = np.full(ts + ses, 1 - eb, np.float32)
        acc[::4] = 1 - eb / 3
        orc.multipler, orc.std_m = mult, stdm
        D2, I2, mynp, _ = orc.search_bounded(q[ts:], k, qk, acc, gt_D=gD, offset=ts, profile=False)
        ix.set_params(mult, stdm)
        D, I, np_gpu = ix.search_bounded(q[ts:], k, qk, acc[ts:])
        assert ix.stats()["err_bits"] == 0 and orc.last_err == 0
        assert np.array_equal(np_gpu, mynp[ts:]), (mult, stdm, eb)
        assert np.array_equal(D, D2)
        assert_results_match(D, I, D2, I2, what=f"bounded d={d} {mult},{stdm},{eb}")
