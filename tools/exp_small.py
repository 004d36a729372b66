"""GPU tool: mean latency of small error-bounded batches on the bench workload (device ms and wall ms)."""
import argparse, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,8,64")
ap.add_argument("--calls", type=int, default=200)
a0 = ap.parse_args()
a = argparse.Namespace(shape="sift", nb=10_000_000, ncal=5000, nq=10000, nlist=4096, eb=0.1)
S = B.build_everything(a, 0, 0)
ix, dev = S["ix"], S["dev"]
ix.set_params(*B.HYPER[0.1])
for b in [int(x) for x in a0.batches.split(",")]:
    acc = torch.full((b,), 0.9, device=dev)
    npb = torch.zeros(b, dtype=torch.int64, device=dev)
    D = torch.empty(b, 100, device=dev)
    I = torch.empty(b, 100, dtype=torch.int64, device=dev)
    ms, wall, rounds, nps = [], [], [], []
    for c in range(a0.calls + 3):
        q = S["qtest"][(c * b) % (a.nq - b + 1):][:b].contiguous()
        npb.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ix.search_bounded_device(q, 100, 10, acc, npb, D, I)
        w = time.perf_counter() - t0
        if c >= 3:
            st = ix.stats()
            ms.append(st["search_ms"]); wall.append(w * 1e3); rounds.append(st["rounds"]); nps.append(float(npb.float().mean()))
    ms = np.array(ms)
    print(f"batch {b}: device ms mean {ms.mean():.3f} p50 {np.median(ms):.3f} p90 {np.percentile(ms, 90):.3f}  wall {np.mean(wall):.3f}  "
          f"rounds {np.mean(rounds):.2f}  my_nprobe {np.mean(nps):.0f}  W0={os.environ.get('AUNCEL_W0', '-')}", flush=True)
