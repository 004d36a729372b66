import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
a = argparse.Namespace(shape="sift", nb=2_000_000, ncal=10, nq=2000, nlist=4096, eb=0.1)
import auncel_b200 as ab
from auncel_b200 import workload as W
dev = torch.device("cuda:0")
base = W.make_vectors("sift", a.nb, 123, dev)
ix = W.build_index(ab, "sift", base, 4096, 0, niter=10)
q = W.make_vectors("sift", a.nq, 789, dev).cpu().numpy()
cd, ck = ix.coarse_search(q, 4096)
ties = (cd[:, 1:] == cd[:, :-1])
print("queries with any tie:", ties.any(1).sum(), "of", len(q), "; ties within first 600 ranks:", ties[:, :600].any(1).sum())
cent = ix.centroids()
u = np.unique(cent, axis=0)
print("distinct centroids", len(u), "of", len(cent))
sizes = ix.list_sizes()
print("empty lists", (sizes == 0).sum())
qi, r = np.argwhere(ties)[0]
i, j = ck[qi, r], ck[qi, r + 1]
print("example tie: query", qi, "rank", r, "centroids", i, j, "dist", cd[qi, r], cd[qi, r + 1], "identical vectors", np.array_equal(cent[i], cent[j]), "sizes", sizes[i], sizes[j])
print(cent[i][:6], cent[j][:6])
