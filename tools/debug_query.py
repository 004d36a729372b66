"""GPU tool: per-stage comparison GPU vs restatement for one test query of the bench workload."""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--q", type=int, default=39)
ap.add_argument("--nb", type=int, default=10_000_000)
ap.add_argument("--ncal", type=int, default=5000)
ap.add_argument("--nq", type=int, default=10000)
a0 = ap.parse_args()
a = argparse.Namespace(shape="sift", nb=a0.nb, ncal=a0.ncal, nq=a0.nq, nlist=4096, eb=0.1)
S = B.build_everything(a, 0, 0)
ix = S["ix"]
K, qk = 100, 10
qi = a0.q
xq = S["qtest"][qi:qi + 1].cpu().numpy()
# oracle over the same lists: only the lists this query can reach matter -> build a CSR with all lists (host copy)
xb = S["base"].cpu().numpy()
asg = ix.assign(xb)
orc = O.OracleIndex(128, 4096, O.L2)
orc.set_centroids(ix.centroids())
orc.add(xb, list_no=asg)
orc.traces = ix.traces()
orc.multipler, orc.std_m = 7.9, 6.0
acc = np.full(1, 0.9, np.float32)
D2, I2, np2, _ = orc.search_bounded(xq, K, qk, acc, dump_q=0)
dump = orc.last_dump
ix.set_params(7.9, 6.0)
D1, I1, np1 = ix.search_bounded(xq, K, qk, acc)
print("oracle my_nprobe", np2, "gpu", np1, "gpu batch-of-1")
# same query inside the original batch position
acc_all = np.full(a.nq, 0.9, np.float32)
Db, Ib, npb = ix.search_bounded(S["qtest"].cpu().numpy(), K, qk, acc_all)
print("gpu in full batch:", npb[qi])
Db2, Ib2, npb2 = ix.search_bounded(S["qtest"][:200].cpu().numpy(), K, qk, acc_all[:200])
print("gpu in batch of 200:", npb2[qi] if qi < 200 else None)
cd, ck = ix.coarse_search(xq, 64)
od, ok = orc.coarse(xq, 64)
print("coarse equal", np.array_equal(cd, od), np.array_equal(ck, ok))
lib = O.orc()
ntr, toff, phi, U, sg = orc._model_arrays()
dtb = np.zeros(4096 // 8 + 20, np.float32); c2c = np.zeros_like(dtb); err = O.C.c_int(0)
fd, fk = orc.coarse(xq, 4096)
lib.orc_set_online(O.L2, 4096, O._p(fd, O._f), O._p(fk, O._l), O._p(orc.interdis, O._f), O._p(orc.arcos, O._f), 500,
                   O._p(c2c, O._f), O._p(dtb, O._f), O.C.byref(err))
for s in range(1, 40):
    ix.nprobe = s
    Dg, Ig = ix.search(xq, K)
    Do, Io = orc.search_fixed(xq, K, s)
    ind = 0
    while s > (1 << ind):
        ind += 1
    Dg_s = np.ascontiguousarray(Dg[0])
    pre = lib.orc_cur_num(O._p(orc.arcos, O._f), 500, ntr, O._p(toff, O._l), O._p(phi, O._f), O._p(U, O._f), O._p(sg, O._f),
                          orc.std_m, O._p(Dg_s, O._f), O._p(dtb, O._f), ind, qk, O.C.byref(err))
    print(s, "fixed D equal", np.array_equal(Dg, Do), "ext gpu", Dg[0, K - 1], "oracle dump (pre,recall,ext,mynp)", dump[s - 1], "pre(from gpu D)", pre)
