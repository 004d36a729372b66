"""2-GPU diagnosis of the sharded rounds: where does the sharded run leave the single-index run?"""
import os
import socket
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, world, port):
    import torch.distributed as dist

    import auncel_b200 as ab
    from auncel_b200 import distributed as AD
    from tests.util import mixture
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    metric, d, nlist, nb, K, qk, nq = ab.METRIC_L2, 32, 256, 60_000, 100, 10, int(os.environ.get("NQ", 600))
    xb = mixture(3, nb, d)
    xq = mixture(4, nq, d)
    ids = np.arange(nb, dtype=np.int64)
    full = ab.IndexIVFFlat(d, nlist, metric, device=rank)
    full.set_tune_mode()
    full.train(xb[:: max(1, nb // (40 * nlist))], niter=3)
    full.set_tune_off()
    full.add_with_ids(xb, ids)
    cent = full.centroids()
    full.nprobe = nlist
    gD, gI = full.search(xq, K)
    mine = AD.shard_mask(ids, world, rank)
    shard = ab.IndexIVFFlat(d, nlist, metric, device=rank)
    shard.set_centroids(cent)
    shard.add_with_ids(xb[mine], ids[mine])
    if rank == 0:
        print("interdis equal", np.array_equal(full.interdis_cem(), shard.interdis_cem()), flush=True)
    nsg = AD.NcclShardGroup(shard)
    nsg.set_bounded(True)
    for tc in (0, 1, 2):
        full.set_option("tensor_core_filter", tc)
        shard.set_option("tensor_core_filter", tc)
        Df, If = full.calibrate(xq, K, gD)
        rs_f = full.round_stats()
        Ds, Is = shard.calibrate(xq, K, gD)
        rs_s = shard.round_stats()
        if rank == 0:
            bad = np.flatnonzero((Df != Ds).any(1))
            print(f"tc={tc} calibrate: D rows differing {len(bad)}/{nq}  labels equal {(If == Is).mean():.4f}  "
                  f"rounds {len(rs_f)} vs {len(rs_s)}  xchg {nsg.exchange_stats()}", flush=True)
            print("  rounds full ", [(int(r['r0']), int(r['w']), int(r['active']), int(r['tc'])) for r in rs_f], flush=True)
            print("  rounds shard", [(int(r['r0']), int(r['w']), int(r['active']), int(r['tc'])) for r in rs_s], flush=True)
            for q in bad[:3]:
                j = np.flatnonzero(Df[q] != Ds[q])
                print("   q", q, "first diff col", j[:5], Df[q, j[:3]], Ds[q, j[:3]], "gt", gD[q, j[:3]], flush=True)
            trf, trs = full.traces(), shard.traces()
        else:
            trf, trs = full.traces(), shard.traces()
        eq = [all(np.array_equal(x, y) for x, y in zip(a, b)) for a, b in zip(trf, trs)]
        if rank == 0:
            print("  traces equal per stage", eq, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)
