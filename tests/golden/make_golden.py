"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The reference publishes no golden vectors for this path; these fixtures
are outputs of the reference's own code -- train (k-means), add, IndexIVF::search,
Error_sys::sys_train and Error_sys::search -- on inputs that tests regenerate from
auncel_b200.synth (bit-reproducible), so only outputs are stored.

faiss::distance_compute_blas_threshold is raised so the coarse quantizer takes the
exact-difference path (utils.cpp:417-490) in batch mode too -- the path the paper's
latency mode (one query per call, eval/bound.cpp:390-396) always takes.  BLAS-path
coarse distances depend on the BLAS build and cannot be golden.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from auncel_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: metric, d, nlist, nb, ts, ses, k(max_topk), query_topk, n_centers, sigma
    "l2_d16": dict(metric=O.L2, d=16, nlist=1024, nb=60000, ts=400, ses=200, k=20, qk=5,
                   n_centers=300, sigma=0.30, niter=6),
    "ip_d24": dict(metric=O.IP, d=24, nlist=1024, nb=120000, ts=400, ses=200, k=12, qk=4,
                   n_centers=3000, sigma=0.45, niter=6),
}
PARAMS = [(1.0, 1.0, 0.1), (2.5, 2.0, 0.2), (7.9, 6.0, 0.1)]  # (multipler, std_m, error bound)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs(c):
    norm = c["metric"] == O.IP
    xb = synth.clustered(101, c["nb"], c["d"], c["n_centers"], c["sigma"], normalize=norm)
    xq = synth.clustered(202, c["ts"] + c["ses"], c["d"], c["n_centers"], c["sigma"], normalize=norm)
    return xb, xq


def usable_queries(c, R, xq):
    """IP: the reference throws (arcos domain, IVF_pro.cpp:180) when the first probed
    list holds < k vectors or a similarity exceeds 1; keep queries it can run."""
    if c["metric"] != O.IP:
        return np.ones(len(xq), bool)
    sizes = R.list_sizes()
    dis, keys = R.coarse(xq, 1)
    return (sizes[keys[:, 0]] >= c["k"]) & (dis[:, 0] <= 1.0)


def make(name, c):
    O.RefIndex.set_blas_threshold(1 << 30)
    xb, xq = inputs(c)
    R = O.RefIndex(c["d"], c["nlist"], c["metric"])
    R.train(xb, niter=c["niter"])
    cent = R.centroids()
    R.add(xb)
    ok = usable_queries(c, R, xq)
    # keep calibration / test split sizes multiples of 10 (profile.cpp:31-32)
    cal = np.flatnonzero(ok[: c["ts"]])
    tst = np.flatnonzero(ok[c["ts"]:]) + c["ts"]
    cal = cal[: len(cal) // 10 * 10]
    tst = tst[: len(tst) // 10 * 10]
    sel = np.concatenate([cal, tst])
    ts, ses = len(cal), len(tst)
    q = xq[sel]
    k, qk = c["k"], c["qk"]

    out = dict(sel=sel, ts=ts, ses=ses, centroids=cent, interdis_sha=sha(R.interdis()),
               interdis_head=R.interdis()[:4096], list_sizes=R.list_sizes(),
               assign_sha=sha(R.assign(xb)))
    cd, ck = R.coarse(q[:8], c["nlist"])
    out.update(coarse_dis=cd, coarse_keys=ck)
    for nprobe in (1, 4, 16):
        D, I = R.search_fixed(q, k, nprobe)
        out[f"fixed_D_{nprobe}"], out[f"fixed_I_{nprobe}"] = D, I
    D, I = R.search_fixed(q, k, 16, max_codes=300)
    out["fixed_D_16_mc300"], out["fixed_I_16_mc300"] = D, I
    gD, gI = R.search_fixed(q, k, c["nlist"])  # exhaustive = ground truth
    out.update(gt_D=gD, gt_I=gI)

    R.es_create(gD, gI)
    R.sys_train(ts, q)
    tr = R.traces()
    out["n_traces"] = len(tr)
    for t, (phi, U, sg) in enumerate(tr):
        out[f"trace_phi_{t}"], out[f"trace_U_{t}"], out[f"trace_sigma_{t}"] = phi, U, sg
    out["arcos"] = R.arcos()
    for pi, (mult, stdm, eb) in enumerate(PARAMS):
        acc = np.full(ts + ses, 1 - eb, np.float32)
        acc[1::3] = 1 - eb / 2  # per-query targets differ (effect_error.cpp:277-281 style)
        R.set_queries(qk, ses, q, acc, mult, stdm, profile=True)
        D, I = R.es_search(ts, ses, -1)
        out[f"b{pi}_acc"] = acc
        out[f"b{pi}_D"], out[f"b{pi}_I"] = D, I
        out[f"b{pi}_my_nprobe"] = R.my_nprobe(ts, ses)
        out[f"b{pi}_t_recalls"] = R.t_recalls(ts, ses)
        # latency mode: one query per call (eval/bound.cpp:390-396) must agree
        R.clear_my_nprobe()
        D1, I1 = R.es_search(ts, ses, 1)
        assert np.array_equal(D, D1) and np.array_equal(I, I1)
    R.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ts", ts, "ses", ses, "mean my_nprobe",
          [float(out[f"b{i}_my_nprobe"].mean()) for i in range(len(PARAMS))],
          "trace sizes", [len(t[0]) for t in tr])


if __name__ == "__main__":
    for name, c in CASES.items():
        make(name, c)
