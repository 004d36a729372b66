"""Shared helpers: golden-case inputs (regenerated from auncel_b200.synth) and the
comparison rules of BASELINE.json's north_star (ids identical except at distance ties
within 1e-5 relative, distances within 1e-4 relative)."""
import os

import numpy as np

from auncel_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IP, L2 = 0, 1

CASES = {
    "l2_d16": dict(metric=L2, d=16, nlist=1024, nb=60000, ts=400, ses=200, k=20, qk=5,
                   n_centers=300, sigma=0.30),
    "ip_d24": dict(metric=IP, d=24, nlist=1024, nb=120000, ts=400, ses=200, k=12, qk=4,
                   n_centers=3000, sigma=0.45),
}
PARAMS = [(1.0, 1.0, 0.1), (2.5, 2.0, 0.2), (7.9, 6.0, 0.1)]

_cache = {}


def golden_case(name):
    """-> (cfg, golden npz dict, xb, q) with q = the golden's selected queries."""
    if name not in _cache:
        c = CASES[name]
        g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        norm = c["metric"] == IP
        xb = synth.clustered(101, c["nb"], c["d"], c["n_centers"], c["sigma"], normalize=norm)
        xq = synth.clustered(202, c["ts"] + c["ses"], c["d"], c["n_centers"], c["sigma"],
                             normalize=norm)
        _cache[name] = (c, g, xb, xq[g["sel"]])
    return _cache[name]


def golden_traces(g):
    return [(g[f"trace_phi_{t}"], g[f"trace_U_{t}"], g[f"trace_sigma_{t}"])
            for t in range(int(g["n_traces"]))]


def assert_results_match(D, I, Dref, Iref, rtol_d=1e-4, tie_rtol=1e-5, what="", D_next=None):
    """north_star tolerance: distances within 1e-4 relative; ids identical except where
    the reference row has another entry within 1e-5 relative (a tie).  D_next (optional, one
    value per row): the reference's NEXT distance after the row's last one -- a tie between the
    k-th and the (k+1)-th neighbour is a tie too, it just is not visible inside the row."""
    D, Dref = np.asarray(D, np.float64), np.asarray(Dref, np.float64)
    assert D.shape == Dref.shape, what
    big = np.abs(Dref) > 1e37
    assert np.array_equal(big, np.abs(D) > 1e37), what + ": padding differs"
    ok = np.abs(D - Dref) <= rtol_d * np.maximum(np.abs(Dref), 1e-30)
    assert np.all(ok | big), f"{what}: distance mismatch, max rel {np.max(np.abs(D-Dref)[~big]/np.abs(Dref)[~big])}"
    diff = (np.asarray(I) != np.asarray(Iref))
    if diff.any():
        rows, cols = np.nonzero(diff)
        for r, c in zip(rows, cols):
            row = Dref[r]
            near = np.abs(row - row[c]) <= tie_rtol * max(abs(row[c]), 1e-30)
            edge = D_next is not None and abs(float(D_next[r]) - row[c]) <= tie_rtol * max(abs(row[c]), 1e-30)
            assert near.sum() > 1 or big[r, c] or edge, (f"{what}: id mismatch at ({r},{c}) not a tie: ours {I[r][c]} ref {Iref[r][c]} "
                                                        f"D {row[max(0, c - 2):c + 2]} next {None if D_next is None else D_next[r]}")


def recall_at(gt_D, D, qk, metric):
    """eval/bound.cpp:117-128 (inter_sec): distance-threshold recall of the first qk."""
    t = gt_D[:, qk - 1][:, None]
    if metric == L2:
        hit = D[:, :qk] <= t + 1e-6
    else:
        hit = D[:, :qk] >= t - 1e-6
    return hit.sum(1) / float(qk)


def mixture(seed, n, d, normalize=False, n_centers=1024, lowrank=16, sigma_lr=1.2, sigma_iso=0.5):
    """Host copy of auncel_b200.workload.make_vectors' mixture (centre + rank-16 + isotropic
    noise) on numpy's PCG64, for CPU-vs-GPU parity tests at the BASELINE.json shapes."""
    gp = np.random.default_rng(977)
    centers = gp.standard_normal((n_centers, d), dtype=np.float32)
    basis = gp.standard_normal((lowrank, d), dtype=np.float32) / np.float32(np.sqrt(lowrank))
    g = np.random.default_rng(seed)
    out = np.empty((n, d), np.float32)
    for i0 in range(0, n, 1 << 16):
        m = min(1 << 16, n - i0)
        x = centers[g.integers(0, n_centers, m)]
        x = x + np.float32(sigma_lr) * (g.standard_normal((m, lowrank), dtype=np.float32) @ basis)
        x = x + np.float32(sigma_iso) * g.standard_normal((m, d), dtype=np.float32)
        if normalize:
            x = x / np.sqrt((x * x).sum(1, keepdims=True))
        out[i0:i0 + m] = x
    return out
