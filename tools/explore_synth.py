"""GPU tool: recall curve of fixed-nprobe search on the synthetic workload, and phase timings.
usage: python tools/explore_synth.py [--nb N] [--nlist L] [--sig LR:ISO ...]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auncel_b200 as ab  # noqa: E402
from auncel_b200 import workload as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="sift")
ap.add_argument("--nb", type=int, default=10_000_000)
ap.add_argument("--nq", type=int, default=1000)
ap.add_argument("--nlist", type=int, default=4096)
ap.add_argument("--niter", type=int, default=10)
ap.add_argument("--sig", nargs="*", default=["0.55:0.25"])
ap.add_argument("--k", type=int, default=100)
a = ap.parse_args()
dev = torch.device("cuda:0")


def sync():
    torch.cuda.synchronize()
    return time.time()


for sg in a.sig:
    lr, iso = (float(v) for v in sg.split(":"))
    gen = dict(sigma_lr=lr, sigma_iso=iso)
    t0 = sync()
    base = W.make_vectors(a.shape, a.nb, 123, dev, gen)
    q = W.make_vectors(a.shape, a.nq, 456, dev, gen)
    t1 = sync()
    ix = W.build_index(ab, a.shape, base, a.nlist, 0, niter=a.niter)
    t2 = sync()
    gD, gI = W.ground_truth(ix, q, a.k)
    t3 = sync()
    sizes = ix.list_sizes()
    print(f"sig={sg} gen {t1-t0:.1f}s build {t2-t1:.1f}s gt {t3-t2:.1f}s ({ix.stats()}) "
          f"list sizes min/mean/max {sizes.min()}/{sizes.mean():.0f}/{sizes.max()}", flush=True)
    gDn = gD.cpu().numpy()
    D = torch.empty_like(gD)
    I = torch.empty_like(gI)
    for nprobe in (1, 4, 16, 64, 256):
        ix.nprobe = nprobe
        t4 = sync()
        ix.search_device(q, a.k, D, I)
        t5 = sync()
        st = ix.stats()
        r10 = W.recall_at(gDn, D.cpu().numpy(), 10, W.SHAPES[a.shape]["metric"]).mean()
        print(f"   nprobe {nprobe:4d}: recall@10 {r10:.3f}  {1e3*(t5-t4):8.2f} ms  scan {st['scan_ms']:.2f} ms "
              f"rounds {st['rounds']:.0f} ndis {st['ndis']:.3g} -> {st['ndis']*4*W.SHAPES[a.shape]['d']/1e9/max(st['scan_ms'],1e-9)*1e3:.0f} GB/s alg", flush=True)
    del ix, base
    torch.cuda.empty_cache()
