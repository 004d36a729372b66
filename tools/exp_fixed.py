"""GPU experiment: fixed-nprobe search on the DEEP-shaped config 3 workload (one GPU), per-round log on stderr."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auncel_b200 as ab
from auncel_b200 import workload as W
dev = torch.device("cuda:0")
nb, nq, nlist, K = 10_000_000, 10000, 4096, 100
base = W.make_vectors("deep", nb, 321, dev)
q = W.make_vectors("deep", nq, 654, dev)
ix = W.build_index(ab, "deep", base, nlist, 0, niter=6, tune=False)
D = torch.empty(nq, K, device=dev); I = torch.empty(nq, K, device=dev, dtype=torch.int64)
for nprobe in (16, 64, 256):
    ix.nprobe = nprobe
    for _ in range(3):
        ix.search_device(q, K, D, I)
    st = ix.stats()
    print(json.dumps({"nprobe": nprobe, **{k: st[k] for k in ("search_ms", "rounds", "tc_rounds", "tc_ms", "simt_ms", "coarse_ms")}}))
