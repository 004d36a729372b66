"""GPU tool: bounded search GPU vs restatement vs reference at nlist=4096, K=100."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auncel_b200 as ab
from auncel_b200 import workload as W
from oracle import oracle as O

nb, nlist, d, K, qk = 1_000_000, 4096, 128, 100, 10
ncal, nq = 200, 100
dev = torch.device("cuda:0")
base = W.make_vectors("sift", nb, 123, dev)
qcal = W.make_vectors("sift", ncal, 456, dev)
qt = W.make_vectors("sift", nq, 789, dev)
ix = W.build_index(ab, "sift", base, nlist, 0, niter=4)
q_all = torch.cat([qcal, qt])
gD, gI = W.ground_truth(ix, q_all, K)
gD = gD.cpu().numpy(); gI = gI.cpu().numpy()
es = ab.Error_sys(ix, ncal + nq, K)
es.set_gt(gD, gI)
es.sys_train(ncal, qcal.cpu().numpy())
xb = base.cpu().numpy()
asg = ix.assign(xb)
orc = O.OracleIndex(d, nlist, O.L2)
orc.set_centroids(ix.centroids())
orc.add(xb, list_no=asg)
orc.traces = ix.traces()
print("interdis equal", np.array_equal(orc.interdis, ix.interdis_cem()))
for mult, stdm in [(7.9, 6.0), (1.0, 1.0)]:
    acc = np.full(ncal + nq, 0.9, np.float32)
    orc.multipler, orc.std_m = mult, stdm
    D2, I2, np2, _ = orc.search_bounded(qt.cpu().numpy(), K, qk, acc, gt_D=gD, offset=ncal)
    ix.set_params(mult, stdm)
    D1, I1, np1 = ix.search_bounded(qt.cpu().numpy(), K, qk, acc[ncal:])
    bad = np.flatnonzero(np1 != np2[ncal:])
    print(mult, stdm, "my_nprobe equal", len(bad) == 0, "D equal", np.array_equal(D1, D2), "err", orc.last_err, ix.stats()["err_bits"])
    print("  gpu", np1[:12], "\n  orc", np2[ncal:][:12])
    if len(bad):
        b = int(bad[0])
        print("  first bad query", b, np1[b], np2[ncal + b])
        _ = orc.search_bounded(qt.cpu().numpy()[b:b+1], K, qk, acc, gt_D=gD, offset=ncal + b, dump_q=0)
        dump = orc.last_dump
        rows = np.flatnonzero(dump[:, 0] > -2)
        print("  oracle stages (pre, recall, ext, mynp):")
        for r in rows[:40]:
            print("   ", r + 1, dump[r])
# reference
if O.have_ref():
    O.RefIndex.set_blas_threshold(1 << 30)
    R = O.RefIndex(d, nlist, O.L2)
    R.set_centroids(ix.centroids())
    R.add(xb, list_no=asg)
    q = np.concatenate([qcal[:10].cpu().numpy(), qt.cpu().numpy()])
    g2 = np.concatenate([gD[:10], gD[ncal:]])
    R.es_create(g2, np.zeros_like(g2, dtype=np.int64))
    R.sys_train(10, q)
    R.set_traces(ix.traces())
    acc = np.full(10 + nq, 0.9, np.float32)
    R.set_queries(qk, nq, q, acc, 7.9, 6.0)
    D3, I3 = R.es_search(10, nq, -1)
    np3 = R.my_nprobe(10, nq)
    orc.multipler, orc.std_m = 7.9, 6.0
    D2, I2, np2, _ = orc.search_bounded(qt.cpu().numpy(), K, qk, np.full(ncal + nq, 0.9, np.float32), gt_D=gD, offset=ncal)
    print("ref vs orc: my_nprobe equal", np.array_equal(np3, np2[ncal:]), "D equal", np.array_equal(D3, D2))
    print("  ref", np3[:12])
    R.clear_my_nprobe()
    D4, I4 = R.es_search(10, nq, threads=8)
    print("ref threads vs ref: ", np.array_equal(R.my_nprobe(10, nq), np3), np.array_equal(D4, D3))
