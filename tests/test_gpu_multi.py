"""GPU, >= 2 devices: shards (NCCL all_gather + device merge_tables) equal the single index
(tests/test_merge.cpp invariant); replicas answer disjoint query slices identically."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def _worker(rank, world, port, out):
    import torch.distributed as dist

    import auncel_b200 as ab
    from auncel_b200 import distributed as AD
    from auncel_b200 import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        d, nlist, nb, k = 24, 64, 20000, 10
        xb = synth.clustered(3, nb, d, 30)
        cent = synth.clustered(53, nlist, d, 30)
        xq = synth.clustered(21, 300, d, 30)
        ids = np.arange(nb, dtype=np.int64)
        full = ab.IndexIVFFlat(d, nlist, ab.METRIC_L2, device=rank)
        full.set_centroids(cent)
        full.add(xb)
        full.nprobe = 6
        Dref, Iref = full.search(xq, k)
        # shard: the vectors with id % world == rank, same centroids (copy_subset_to type 1)
        mine = AD.shard_mask(ids, world, rank)
        shard = ab.IndexIVFFlat(d, nlist, ab.METRIC_L2, device=rank)
        shard.set_centroids(cent, compute_interdis=False)
        shard.add_with_ids(xb[mine], ids[mine])
        shard.nprobe = 6
        sg = AD.ShardGroup(shard, ab.METRIC_L2)
        xq_t = torch.from_numpy(xq).cuda(rank)
        D, I = sg.search_device(xq_t, k)
        torch.cuda.synchronize()
        assert np.array_equal(D.cpu().numpy(), Dref)
        assert (I.cpu().numpy() == Iref).mean() > 0.999
        # the same inside the library: ncclAllGather of the packed tables + merge on the index stream
        nsg = AD.NcclShardGroup(shard)
        for kk, npb in [(k, 6), (100, 17), (1, 1)]:
            full.nprobe = shard.nprobe = npb
            Dr_k, Ir_k = full.search(xq, kk)
            D2, I2 = nsg.search_device(xq_t, kk)
            assert np.array_equal(D2.cpu().numpy(), Dr_k)
            assert (I2.cpu().numpy() == Ir_k).mean() > 0.999
            D3, I3 = nsg.search(xq, kk)  # host-pointer entry point
            assert np.array_equal(D3, Dr_k) and np.array_equal(I3, I2.cpu().numpy())
            st = nsg.stats()
            assert st["world"] == world and st["allgather_bytes"] == world * (((300 * kk * 4 + 7) // 8) * 8 + 300 * kk * 8)
            assert st["nccl_version"] > 0
        full.nprobe = shard.nprobe = 6
        # replicas
        rg = AD.ReplicaGroup(full)
        base, Dr, Ir = rg.search(xq, k)
        assert np.array_equal(Dr, Dref[base:base + len(Dr)]) and np.array_equal(Ir, Iref[base:base + len(Ir)])
        Dg, Ig = rg.search_gathered(xq, k)
        assert np.array_equal(Dg, Dref)
        # error-bounded search over shards, dist/ semantics: every worker calibrates and terminates
        # on its own slice; tables merged at the end.  Both ranks rebuild both shards to check.
        from oracle import oracle as O
        k2, qk = 16, 4
        xq2 = synth.clustered(22, 200, d, 30)
        per_shard = []
        for r in range(world):
            m_r = AD.shard_mask(ids, world, r)
            sh = ab.IndexIVFFlat(d, nlist, ab.METRIC_L2, device=rank)
            sh.set_centroids(cent)
            sh.add_with_ids(xb[m_r], ids[m_r])
            sh.nprobe = nlist
            gD, gI = sh.search(xq2, k2)
            es = ab.Error_sys(sh, 200, k2)
            es.set_gt(gD, gI)
            es.sys_train(100, xq2)
            es.set_topk(qk)
            es.setparam(2.0, 1.0)
            es.set_queries(100, xq2, np.full(200, 0.9, np.float32), 200)
            per_shard.append(es)
        mine_es = per_shard[rank]
        bg = AD.BoundedShardGroup(mine_es, ab.METRIC_L2)
        Dm, Im = bg.search(100, 100)
        tabs = []
        for r in range(world):
            per_shard[r].set_queries(100, xq2, np.full(200, 0.9, np.float32), 200)
            tabs.append(per_shard[r].search(100, 100))
        De, Ie = O.merge_tables(O.L2, np.stack([t[0] for t in tabs]), np.stack([t[1] for t in tabs]))
        assert np.array_equal(Dm, De) and np.array_equal(Im, Ie)
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_shards_and_replicas_two_gpus():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}


def _bounded_worker(rank, world, port, out):
    """Error-bounded search over shards with single-index semantics (csrc/shard_rounds.cu): calibration
    traces, distances, my_nprobe and t_recalls of the sharded run equal the one-index run bit for bit."""
    import torch.distributed as dist

    import auncel_b200 as ab
    from auncel_b200 import distributed as AD
    from tests.util import assert_results_match, mixture
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        for metric, d, nlist, nb, K, qk, nq in [(ab.METRIC_L2, 32, 256, 60_000, 100, 10, 2600),
                                                (ab.METRIC_INNER_PRODUCT, 96, 1024, 150_000, 100, 10, 3000),
                                                (ab.METRIC_L2, 24, 128, 30_000, 16, 4, 400)]:
            norm = metric == ab.METRIC_INNER_PRODUCT
            xb = mixture(3, nb, d, norm)
            xq = mixture(4, nq + 300, d, norm)
            ids = np.arange(nb, dtype=np.int64) * 3 + 1
            full = ab.IndexIVFFlat(d, nlist, metric, device=rank)
            full.set_tune_mode()
            full.train(xb[:: max(1, nb // (40 * nlist))], niter=3)
            full.set_tune_off()
            full.add_with_ids(xb, ids)
            cent = full.centroids()
            if norm:  # queries the reference serves in IP mode: first list holds >= K vectors
                full.nprobe = 1
                D1, I1 = full.search(xq, K)
                xq = xq[(I1[:, K - 1] >= 0) & (D1[:, 0] <= 1.0)]
            ts, ses = 200, min(nq, len(xq) - 200) // 10 * 10  # (Error_sys wants multiples of ten)
            q = np.ascontiguousarray(xq[: ts + ses])
            full.nprobe = nlist
            gD, gI = full.search(q, K)
            mine = AD.shard_mask(ids, world, rank)
            shard = ab.IndexIVFFlat(d, nlist, metric, device=rank)
            shard.set_centroids(cent)
            shard.add_with_ids(xb[mine], ids[mine])
            nsg = AD.NcclShardGroup(shard)
            nsg.set_bounded(True)
            res = []
            for tc in (1, 2, 0):
                row = []
                for ix in (full, shard):
                    ix.set_option("tensor_core_filter", tc)
                    es = ab.Error_sys(ix, ts + ses, K)
                    es.set_gt(gD, gI)
                    es.sys_train(ts, q)
                    es.set_topk(qk)
                    es.setparam(3.0, 2.0)
                    acc = np.full(ts + ses, 0.9, np.float32)
                    acc[::3] = 0.97
                    es.set_queries(ses, q, acc, ts + ses)
                    es.profile = True
                    D, I = es.search(ts)
                    st = ix.stats()
                    assert st["err_bits"] == 0, st
                    row.append((ix.traces(), D, I, es.my_nprobe[ts:].copy(), es.t_recalls[ts:].copy(), st))
                (tr_f, D_f, I_f, np_f, rc_f, st_f), (tr_s, D_s, I_s, np_s, rc_s, st_s) = row
                assert len(tr_f) == len(tr_s) > 0
                for a, b in zip(tr_f, tr_s):
                    for x, y in zip(a, b):
                        assert np.array_equal(x, y)
                assert np.array_equal(np_f, np_s), (tc, int((np_f != np_s).sum()))
                assert np.array_equal(D_f, D_s)
                assert np.array_equal(rc_f, rc_s)
                assert_results_match(D_s, I_s, D_f, I_f, what=f"sharded bounded tc={tc}")
                assert (I_s == I_f).mean() > 0.999
                xs = nsg.exchange_stats()
                assert xs["exchanges"] == st_s["rounds"] > 0 and xs["entries_all"] >= xs["entries_sent"] > 0, xs
                if tc == 2 and nq >= 2000:
                    assert st_s["tc_rounds"] > 0
                res.append(np_s)
            assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
            assert len(np.unique(res[0])) > 2  # the bound does vary the stop stage
            # switching the exchange off restores the per-shard behaviour
            nsg.set_bounded(False)
            es = ab.Error_sys(shard, ts + ses, K)
            es.set_gt(gD, gI)
            es.sys_train(ts, q)
            assert nsg.exchange_stats()["exchanges"] == xs["exchanges"]  # untouched: no exchange happened
            del nsg
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_bounded_shards_single_index_semantics(world):
    """world 4: a pair's union can exceed 512 entries only from 6 shards on at K = 100; 4 shards still cover
    the overflow pool with more than two contributors per pair."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bounded_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {r: "ok" for r in range(world)}
