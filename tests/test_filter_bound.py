"""CPU restatement of the tensor-core filter's test (auncel_b200/csrc/tcfilter.cu, scheduler + epilogue) in numpy:
for pairs whose EXACT reference distance beats the threshold, the filter -- fed with a TF32-truncated dot product
accumulated in float32 -- must let the pair through.  The GPU tests audit the real kernel slot by slot; this one
pins the algebra of the reformulated test (dot + kq > 0.5 |v|^2 (1 - c2), margin on the list's largest norm) and
its rounding slack on adversarial inputs without a GPU."""
import numpy as np
import pytest

F = np.float32


def tf32(x):
    """what the tensor core reads of an fp32 operand: the low 13 mantissa bits are ignored"""
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def exact_l2(q, v):
    """exact.cuh / utils_simd.cpp:391-443: four lane accumulators, separately rounded sub, mul, add"""
    s = np.zeros((q.shape[0], 4), np.float32)
    for k in range(0, q.shape[1], 4):
        t = (q[:, k:k + 4] - v[:, k:k + 4]).astype(np.float32)
        s = (s + (t * t).astype(np.float32)).astype(np.float32)
    return ((s[:, 0] + s[:, 1]).astype(np.float32) + (s[:, 2] + s[:, 3]).astype(np.float32)).astype(np.float32)


def exact_ip(q, v):
    s = np.zeros((q.shape[0], 4), np.float32)
    for k in range(0, q.shape[1], 4):
        s = (s + (q[:, k:k + 4] * v[:, k:k + 4]).astype(np.float32)).astype(np.float32)
    return ((s[:, 0] + s[:, 1]).astype(np.float32) + (s[:, 2] + s[:, 3]).astype(np.float32)).astype(np.float32)


def tf32_dot(q, v):
    a, b = tf32(q), tf32(v)
    acc = np.zeros(q.shape[0], np.float32)
    for k in range(q.shape[1]):  # float32 accumulation, one product at a time (no better than the hardware's)
        acc = (acc + (a[:, k].astype(np.float64) * b[:, k].astype(np.float64)).astype(np.float32)).astype(np.float32)
    return acc


def norms(x):
    return (x.astype(np.float64) ** 2).sum(1).astype(np.float32)  # row_norms_kernel: double accumulation


def cases(rng, n, d):
    base = rng.standard_normal((n, d)).astype(np.float32)
    yield "gaussian", base, (base + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
    off = (base + 100.0).astype(np.float32)  # large common offset: |q||v| >> distance
    yield "offset", off, (off + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    yield "near-duplicates", off, (off * (1 + 2.0 ** -12)).astype(np.float32)
    ones = (np.abs(base).view(np.uint32) | np.uint32(0x1FFF)).view(np.float32)  # all ignored mantissa bits set, one sign
    yield "max truncation", ones, (ones * F(1.01)).astype(np.float32)


@pytest.mark.parametrize("d", [96, 128, 200, 960])
def test_filter_never_drops_a_pair_the_exact_test_accepts(d):
    rng = np.random.default_rng(d)
    n = 4000
    c1 = F(2.0) * (F(1.02) / F(512.0) + F(d) / F(2097152.0))  # index.cu: ta.c1 .. ta.c3
    c2, c3 = F(1.0 / 1048576.0), F(1.0 / 16384.0)
    slack = F(1.0 / 1048576.0)
    for name, q, v in cases(rng, n, d):
        nq, nv = norms(q), norms(v)
        nmax = nv  # tightest case: the row itself is the largest of its list (a larger maximum only widens the margin)
        snmax = np.sqrt(nmax, dtype=np.float32)
        dot = tf32_dot(q, v)
        # L2: threshold one ulp above the exact distance -> the reference's strict test accepts the pair
        dist = exact_l2(q, v)
        tau = np.nextafter(dist, F(np.inf), dtype=np.float32)
        rhs = (tau + c3 * np.abs(tau) - nq * (F(1) - c2)).astype(np.float32)
        kq = (F(0.5) * (rhs + c1 * np.sqrt(nq, dtype=np.float32) * snmax) + (nq + nmax) * slack).astype(np.float32)
        nvh = (F(0.5) * (nv * (F(1) - c2))).astype(np.float32)
        ok = (dot + kq).astype(np.float32) > nvh
        assert ok.all(), (name, "L2", int((~ok).sum()))
        # inner product: threshold one ulp below the exact similarity
        sim = exact_ip(q, v)
        tau = np.nextafter(sim, F(-np.inf), dtype=np.float32)
        sq = (np.sqrt(nq, dtype=np.float32) * snmax).astype(np.float32)
        kq = (tau - c3 * np.abs(tau) - F(0.5) * c1 * sq - sq * slack).astype(np.float32)
        ok = dot > kq
        assert ok.all(), (name, "IP", int((~ok).sum()))


def test_filter_rejects_far_pairs():
    """the bound is a filter, not a pass-through: unrelated pairs fail it against a realistic threshold"""
    rng = np.random.default_rng(7)
    d, n = 128, 4000
    q = rng.standard_normal((n, d)).astype(np.float32)
    v = rng.standard_normal((n, d)).astype(np.float32)
    c1 = F(2.0) * (F(1.02) / F(512.0) + F(d) / F(2097152.0))
    c2, c3, slack = F(1.0 / 1048576.0), F(1.0 / 16384.0), F(1.0 / 1048576.0)
    nq, nv = norms(q), norms(v)
    nmax = F(nv.max())
    tau = F(0.5) * exact_l2(q, v)  # the pair is twice as far as the threshold
    rhs = (tau + c3 * np.abs(tau) - nq * (F(1) - c2)).astype(np.float32)
    kq = (F(0.5) * (rhs + c1 * np.sqrt(nq, dtype=np.float32) * np.sqrt(nmax)) + (nq + nmax) * slack).astype(np.float32)
    nvh = (F(0.5) * (nv * (F(1) - c2))).astype(np.float32)
    assert not ((tf32_dot(q, v) + kq).astype(np.float32) > nvh).any()
