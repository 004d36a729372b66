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
