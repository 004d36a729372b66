"""GPU tool: the bench step with the shared-memory-query filter kernel (1) and the TMEM-resident one (2):
step time and the duration of every tensor-core launch."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="sift")
a0 = ap.parse_args()
a = argparse.Namespace(shape=a0.shape, nb=10_000_000, ncal=5000, nq=10000, nlist=4096, eb=0.1)
S = B.build_everything(a, 0, int(os.environ.get("QRANK", "0")))
ix, dev = S["ix"], S["dev"]
n = a.nq
q = S["qtest"]
acc = torch.full((n,), 0.9, device=dev)
npb = torch.zeros(n, dtype=torch.int64, device=dev)
D = torch.empty(n, 100, device=dev)
I = torch.empty(n, 100, dtype=torch.int64, device=dev)
ix.set_params(*B.HYPER[0.1])
ref = None
for kern in [int(k) for k in os.environ.get("KERNELS", "1,2,1,2").split(",")]:
    ix.set_option("tc_kernel", kern)
    ms = []
    for rep in range(6):
        npb.zero_()
        ix.search_bounded_device(q, 100, 10, acc, npb, D, I)
        ms.append(ix.stats()["search_ms"])
    st = ix.stats()
    torch.cuda.synchronize()
    res = (D.clone(), npb.clone())
    same = None if ref is None else (bool(torch.equal(res[0], ref[0])), bool(torch.equal(res[1], ref[1])))
    ref = ref or res
    print("kernel", kern, "search_ms", [round(x, 3) for x in ms], "tc_ms", round(st["tc_ms"], 3), "same as first", same,
          "tc launches", [(int(r["r0"]), int(r["w"]), round(r["tc_ms"], 3)) for r in ix.round_stats() if r.get("tc")], flush=True)
