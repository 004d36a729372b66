"""GPU tool for ncu: one representative bulk scan launch (10k queries x nprobe lists, K=100)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auncel_b200 as ab
from auncel_b200 import workload as W
ap = argparse.ArgumentParser()
ap.add_argument("--nb", type=int, default=10_000_000)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--nprobe", type=int, default=64)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tc", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
base = W.make_vectors("sift", a.nb, 123, dev)
q = W.make_vectors("sift", a.nq, 789, dev)
ix = W.build_index(ab, "sift", base, 4096, 0, niter=2)
D = torch.empty(a.nq, a.k, device=dev); I = torch.empty(a.nq, a.k, device=dev, dtype=torch.int64)
ix.nprobe = a.nprobe
ix.set_option("tensor_core_filter", a.tc)
for _ in range(a.reps):
    ix.search_device(q, a.k, D, I)
    st = ix.stats()
    print("search_ms %.2f scan_ms %.2f rounds %d ndis %.3g -> %.1f T lane-ops/s | tc: rounds %d ms %.3f ndis %.3g cand %d fallbacks %d" % (
        st["search_ms"], st["scan_ms"], st["rounds"], st["ndis"], st["ndis"] * 128 * 3 / st["scan_ms"] / 1e9,
        st["tc_rounds"], st["tc_ms"], st["tc_ndis"], st["tc_candidates"], st["tc_fallbacks"]))
