"""GPU tool: one error-bounded search of the bench workload between cudaProfilerStart/Stop
(use with `ncu --profile-from-start off ...`)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--batch", type=int, default=0)
a0 = ap.parse_args()
a = argparse.Namespace(shape="sift", nb=10_000_000, ncal=5000, nq=a0.nq, nlist=4096, eb=0.1)
S = B.build_everything(a, 0, int(os.environ.get("QRANK", "0")))
ix, dev = S["ix"], S["dev"]
ix.set_params(*B.HYPER[0.1])
n = a0.batch or a.nq
q = S["qtest"][:n].contiguous()
acc = torch.full((n,), 0.9, device=dev)
npb = torch.zeros(n, dtype=torch.int64, device=dev)
D = torch.empty(n, 100, device=dev)
I = torch.empty(n, 100, dtype=torch.int64, device=dev)
for rep in range(2):
    npb.zero_()
    ix.search_bounded_device(q, 100, 10, acc, npb, D, I)
torch.cuda.synchronize()
torch.cuda.profiler.start()
npb.zero_()
ix.search_bounded_device(q, 100, 10, acc, npb, D, I)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(ix.stats())
