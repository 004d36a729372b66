"""GPU-box tool: how far is the B200 path (exact-difference coarse distances, = the reference's latency
mode and the parity harness) from the reference in its DEFAULT batch mode, where IndexFlat::search
switches to sgemm_ + norms for >= 20 queries (Auncel/utils.cpp:622-655)?  BASELINE config 1: IVF-Flat
nlist=1024, 1M x 128 SIFT-shaped, 10k queries, query_topk=10, error bound 0.1.  Prints one JSON line."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, default=10000)
a0 = ap.parse_args()
a = argparse.Namespace(shape="sift", nb=1_000_000, ncal=5000, nq=a0.nq, nlist=1024, eb=0.1)
S = B.build_everything(a, 0, 0)
ix, W, dev = S["ix"], S["W"], S["dev"]
ix.set_params(*B.HYPER[0.1])
n = a.nq
acc = torch.full((n,), 0.9, device=dev)
npb = torch.zeros(n, dtype=torch.int64, device=dev)
D = torch.empty(n, 100, device=dev)
I = torch.empty(n, 100, dtype=torch.int64, device=dev)
ix.search_bounded_device(S["qtest"], 100, 10, acc, npb, D, I)
Dg, Ig, npg = D.cpu().numpy(), I.cpu().numpy(), npb.cpu().numpy()
gt = S["gD"][a.ncal:]

R, _ = B.build_reference(a, S)
cores = os.cpu_count() or 1
out = {"config": "IVF-Flat nlist=1024, 1M x 128 sift-shaped, %d queries, max_topk=100, query_topk=10, eb 0.1, "
                 "(multipler,std_m)=(7.9,6.0)" % n, "has_blas": bool(O.RefIndex.has_blas()), "cores": cores}
for name, thr in (("exact_coarse", 1 << 30), ("blas_coarse_default", 20)):
    O.RefIndex.set_blas_threshold(thr)
    R.clear_my_nprobe() if name != "exact_coarse" else None
    t0 = time.time()
    # 64 queries per IndexIVF::search call: above distance_compute_blas_threshold (20), so with the default
    # threshold the coarse quantizer really takes the sgemm_ path (utils.cpp:644-655)
    dt, Dr, npr = B.cpu_sample_search(a, S, R, O, n, cores, chunk=64)
    npr = npr.astype(np.int64)
    rec_r = W.recall_at(gt, Dr, 10, 1)
    rec_g = W.recall_at(gt, Dg, 10, 1)
    same_np = npr == npg
    rel = np.abs(Dr - Dg) / np.maximum(np.abs(Dr), 1e-30)
    out[name] = {
        "reference_qps": n / dt,
        "my_nprobe_mismatch_rate": float(1.0 - same_np.mean()), "my_nprobe_mismatches": int((~same_np).sum()),
        "mean_my_nprobe_ref": float(npr.mean()), "mean_my_nprobe_gpu": float(npg.mean()),
        "distances_bit_equal_rows": float((Dr == Dg).all(1).mean()),
        "max_rel_distance_diff_where_my_nprobe_equal": float(rel[same_np].max()) if same_np.any() else None,
        "recall@10_ref": float(rec_r.mean()), "recall@10_gpu": float(rec_g.mean()),
        "satisfied_frac_ref": float((rec_r >= 0.9 - 1e-6).mean()), "satisfied_frac_gpu": float((rec_g >= 0.9 - 1e-6).mean())}
O.RefIndex.set_blas_threshold(1 << 30)
R.close()
print(json.dumps(out))
