"""GPU tool: batch-size sweep of the error-bounded search on the bench workload (BASELINE config
"batch 1 vs batch 10k latency/throughput sweep"): latency per call, QPS, bytes/s of the scan."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="sift")
ap.add_argument("--nb", type=int, default=10_000_000)
ap.add_argument("--nlist", type=int, default=4096)
ap.add_argument("--batches", default="1,8,64,512,4096,10000")
ap.add_argument("--out", default="")
a0 = ap.parse_args()
a = argparse.Namespace(shape=a0.shape, nb=a0.nb, ncal=5000, nq=10000, nlist=a0.nlist, eb=0.1)
S = B.build_everything(a, 0, 0)
ix, W = S["ix"], S["W"]
d = W.SHAPES[a.shape]["d"]
dev = S["dev"]
ix.set_params(*B.HYPER[0.1])
rows = []
for bs in [int(v) for v in a0.batches.split(",")]:
    q = S["qtest"][:bs].contiguous()
    acc = torch.full((bs,), 0.9, device=dev)
    npb = torch.zeros(bs, dtype=torch.int64, device=dev)
    D = torch.empty(bs, 100, device=dev)
    I = torch.empty(bs, 100, dtype=torch.int64, device=dev)
    reps = 30 if bs <= 64 else 5
    lat, agg = [], {}
    for r in range(reps + 3):
        npb.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ix.search_bounded_device(q, 100, 10, acc, npb, D, I)
        dt = time.perf_counter() - t0
        st = ix.stats()
        if r >= 3:
            lat.append(dt)
            for k, v in st.items():
                agg[k] = agg.get(k, 0.0) + v / reps
    ms = 1e3 * float(np.median(lat))
    uniq = (agg["tc_uniq"] + agg["simt_uniq"]) * 4 * d
    row = dict(batch=bs, wall_ms=ms, device_ms=agg["search_ms"], qps=bs / (ms / 1e3), rounds=agg["rounds"],
               scan_ms=agg["scan_ms"], ndis=agg["ndis"], alg_gbs=agg["ndis"] * 4 * d / (agg["scan_ms"] / 1e3) / 1e9,
               compulsory_gbs=uniq / (agg["scan_ms"] / 1e3) / 1e9, tc_rounds=agg["tc_rounds"], coarse_ms=agg["coarse_ms"],
               mean_my_nprobe=float(npb.float().mean()))
    rows.append(row)
    print(json.dumps(row), flush=True)
if a0.out:
    json.dump(rows, open(a0.out, "w"), indent=1)
