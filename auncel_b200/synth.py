"""Deterministic synthetic vector sets (no RNG-library dependence).

The reference's tests regenerate their data from seeds (tests/common.py:17-96,
tests/test_lowlevel_ivf.cpp:82-110); there is no real dataset in this environment.
Values come from a splitmix64 counter hash evaluated with numpy integer arithmetic, so
the same (seed, shape) gives bit-identical float32 arrays on every machine and numpy
version -- golden fixtures only need to store outputs.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return z ^ (z >> np.uint64(31))


def uniform(seed, shape):
    """float64 in [0,1), exactly reproducible."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) + (np.uint64(seed) << np.uint64(40))
        h = _splitmix64(ctr)
    return ((h >> np.uint64(11)).astype(np.float64) / float(1 << 53)).reshape(shape)


def normalish(seed, shape):
    """sum of 4 uniforms, centred and scaled to unit variance (Irwin-Hall)."""
    u = uniform(seed, (4,) + tuple(shape))
    return (u.sum(0) - 2.0) * np.sqrt(3.0)


def clustered(seed, n, d, n_centers=64, sigma=0.35, rank=0, normalize=False):
    """n x d float32: mixture centre + noise.  rank>0 adds low-intrinsic-dimension noise
    (tests/common.py:82-96 style); normalize=True gives unit vectors (TEXT-shaped IP)."""
    centers = uniform(seed * 7 + 1, (n_centers, d)) * 2.0 - 1.0
    which = (uniform(seed * 7 + 2, (n,)) * n_centers).astype(np.int64)
    x = centers[which] + sigma * normalish(seed * 7 + 3, (n, d))
    if rank > 0:
        basis = normalish(seed * 7 + 4, (rank, d)) / np.sqrt(rank)
        x = x + normalish(seed * 7 + 5, (n, rank)) @ basis
    if normalize:
        x = x / np.sqrt((x * x).sum(1, keepdims=True))
    return np.ascontiguousarray(x, dtype=np.float32)


def brute_force_gt(metric_l2, xb, xq, k, dist_fn):
    """exact GT with a caller-supplied per-pair distance (used with the oracle's
    SSE-order kernels so kscaling's 1e-5 match holds, IVF_pro.cpp:72-82)."""
    raise NotImplementedError
