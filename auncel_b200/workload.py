"""Synthetic SIFT/DEEP/TEXT/GIST-shaped workloads (BASELINE.json configs) for bench.py and the
full-size property tests.  torch is used here only to generate and hold data on the device;
every search goes through the C ABI.

There is no real dataset and no network.  Vectors are drawn from a mixture: a centre, a
low-intrinsic-dimension component (rank 16, like tests/common.py:82-96 of the reference) and
an isotropic component, tuned (tools/explore_synth.py) so that fixed-nprobe recall@10 behaves
like SIFT (benchs/README.md:232-241 of the reference: ~0.4-0.65 at nprobe=1, ~0.95 at 64).
"""
import numpy as np

SHAPES = {
    # name: d, metric (1 = L2, 0 = IP), normalize
    "sift": dict(d=128, metric=1, normalize=False),
    "deep": dict(d=96, metric=0, normalize=True),
    "text": dict(d=200, metric=0, normalize=True),
    "gist": dict(d=960, metric=1, normalize=False),
}

GEN = dict(n_centers=1024, center_scale=1.0, lowrank=16, sigma_lr=1.2, sigma_iso=0.5)


def make_vectors(shape, n, seed, device, gen=None, chunk=1 << 20):
    """n x d float32 torch tensor on `device` (mixture parameters are seed-independent)."""
    import torch
    g = dict(GEN)
    if gen:
        g.update(gen)
    d = SHAPES[shape]["d"]
    gp = torch.Generator(device=device)
    gp.manual_seed(977)  # mixture parameters
    centers = torch.randn(g["n_centers"], d, generator=gp, device=device) * g["center_scale"]
    basis = torch.randn(g["lowrank"], d, generator=gp, device=device) / np.sqrt(g["lowrank"])
    gd = torch.Generator(device=device)
    gd.manual_seed(seed)
    out = torch.empty(n, d, device=device, dtype=torch.float32)
    for i0 in range(0, n, chunk):
        m = min(chunk, n - i0)
        which = torch.randint(0, g["n_centers"], (m,), generator=gd, device=device)
        x = centers[which]
        x += g["sigma_lr"] * (torch.randn(m, g["lowrank"], generator=gd, device=device) @ basis)
        x += g["sigma_iso"] * torch.randn(m, d, generator=gd, device=device)
        if SHAPES[shape]["normalize"]:
            x = x / x.norm(dim=1, keepdim=True)
        out[i0:i0 + m] = x
    return out


def build_index(ab, shape, base_t, nlist, device_index, niter=10, train_seed=5, tune=True):
    """train (k-means on <= 256*nlist sampled points, Clustering.cpp:24-35) + add, all on device."""
    import torch
    d, metric = SHAPES[shape]["d"], SHAPES[shape]["metric"]
    ix = ab.IndexIVFFlat(d, nlist, metric, device=device_index)
    n = base_t.shape[0]
    ntrain = min(n, 256 * nlist)
    gp = torch.Generator(device=base_t.device)
    gp.manual_seed(train_seed)
    sel = torch.randperm(n, generator=gp, device=base_t.device)[:ntrain]
    xt = base_t[sel].cpu().numpy()
    if tune:
        ix.set_tune_mode()
    ix.train(xt, niter=niter)
    ix.set_tune_off()
    ix.add_device(base_t)
    return ix


def ground_truth(ix, q_t, k):
    """exhaustive search (nprobe = nlist) with the exact kernel == brute force with
    fvec_L2sqr / fvec_inner_product, which is what kscaling's 1e-5 match needs."""
    import torch
    n = q_t.shape[0]
    D = torch.empty(n, k, device=q_t.device, dtype=torch.float32)
    I = torch.empty(n, k, device=q_t.device, dtype=torch.int64)
    saved = ix.nprobe
    ix.nprobe = ix.nlist
    bs = 2048
    for i0 in range(0, n, bs):
        m = min(bs, n - i0)
        ix.search_device(q_t[i0:i0 + m], k, D[i0:i0 + m], I[i0:i0 + m])
    ix.nprobe = saved
    return D, I


def recall_at(gt_D, D, qk, metric):
    """eval/bound.cpp:117-128 (inter_sec)/topk per query; numpy arrays."""
    t = gt_D[:, qk - 1][:, None]
    hit = (D[:, :qk] <= t + 1e-6) if metric == 1 else (D[:, :qk] >= t - 1e-6)
    return hit.sum(1) / float(qk)
