"""debug: adversarial near_dup case, GPU vs reference, row-level diff"""
import sys
import numpy as np
sys.path.insert(0, ".")
import auncel_b200 as ab
from oracle import oracle as O
from tests.test_gpu_baseline_shapes import _adversarial_sets

O.RefIndex.set_blas_threshold(1 << 30)
d, nlist, K = 128, 64, 100
name = sys.argv[1] if len(sys.argv) > 1 else "near_dup"
nprobe = int(sys.argv[2]) if len(sys.argv) > 2 else 12
xb, xq = _adversarial_sets(d)[name]
cent = xb[:: len(xb) // nlist][:nlist].copy()
ix = ab.IndexIVFFlat(d, nlist, 1)
ix.set_centroids(cent, compute_interdis=False)
ix.add(xb)
R = O.RefIndex(d, nlist, 1)
R.set_centroids(cent)
R.add(xb, ids=np.arange(len(xb), dtype=np.int64), list_no=ix.assign(xb))
import os
TH = max(1, min(32, os.cpu_count() or 1))
print("threads", TH)
seq = [(12, 2, 1, 2 << 20), (12, 0, 0, 1 << 30), (48, 2, 1, 2 << 20), (48, 0, 0, 1 << 30), (48, 2, 1, 2 << 20)]
for nprobe, mode, audit, budget in seq:
    Dr, Ir = R.search_fixed(xq, K + 1, nprobe, threads=TH)
    Dr1, Ir1 = R.search_fixed(xq, K + 1, nprobe, threads=1)
    print("ref threads vs serial equal:", np.array_equal(Dr, Dr1), np.array_equal(Ir, Ir1))
    ix.set_option("tensor_core_filter", mode)
    ix.set_option("tc_audit", audit)
    ix.set_pool_budget(budget)
    ix.nprobe = nprobe
    D, I = ix.search(xq, K)
    st = ix.stats()
    bad = []
    for r in range(len(xq)):
        for c in np.nonzero(I[r] != Ir[r, :K])[0]:
            row = Dr[r]
            if (np.abs(row - row[c]) <= 1e-5 * abs(row[c])).sum() <= 1:
                bad.append((r, int(c)))
    print("nprobe", nprobe, "mode", mode, "audit", audit, "budget", budget, "rounds", st["rounds"], "tc", st["tc_rounds"], "fb", st["tc_fallbacks"],
          "Dequal", np.array_equal(D, Dr[:, :K]), "nontie id mismatches", bad[:6])
    for r, c in bad[:2]:
        print("   row", r, "col", c, "ours", I[r, c], D[r, c], "ref", Ir[r, c], Dr[r, c], "next", Dr[r, c + 1],
              "dist(ours id) recomputed", float(((xq[r].astype(np.float64) - xb[I[r, c]].astype(np.float64)) ** 2).sum()),
              "ours id in ref row:", int(I[r, c]) in set(Ir[r].tolist()),
              "same vector:", np.array_equal(xb[I[r, c]], xb[Ir[r, c]]))
