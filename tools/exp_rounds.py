"""GPU experiment: round schedule / tensor-core take-over knobs on the bench workload (stderr: per-round log)."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
a = argparse.Namespace(shape="sift", nb=10_000_000, ncal=5000, nq=10000, nlist=4096, eb=0.1)
S = B.build_everything(a, 0, 0)
ix, dev = S["ix"], S["dev"]
ix.set_params(*B.HYPER[0.1])
n = a.nq
acc = torch.full((n,), 0.9, device=dev)
npb = torch.zeros(n, dtype=torch.int64, device=dev)
D = torch.empty(n, 100, device=dev)
I = torch.empty(n, 100, dtype=torch.int64, device=dev)
for rep in range(3):
    npb.zero_()
    ix.search_bounded_device(S["qtest"], 100, 10, acc, npb, D, I)
st = ix.stats()
print(json.dumps({k: st[k] for k in ("search_ms", "rounds", "tc_rounds", "tc_candidates", "tc_fallbacks", "tc_ms", "simt_ms", "scan_ms", "coarse_ms")}))
