"""GPU tool: one bench step with a -DTC2_TIMING build (AUNCEL_LIB=_variants/libauncel_tc2time.so): wait-cycle attribution of tc_filter2."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
a = argparse.Namespace(shape="sift", nb=10_000_000, ncal=5000, nq=10000, nlist=4096, eb=0.1)
S = B.build_everything(a, 0, 0)
ix, dev = S["ix"], S["dev"]
n = a.nq
acc = torch.full((n,), 0.9, device=dev)
npb = torch.zeros(n, dtype=torch.int64, device=dev)
D = torch.empty(n, 100, device=dev)
I = torch.empty(n, 100, dtype=torch.int64, device=dev)
ix.set_params(*B.HYPER[0.1])
ix.set_option("tc_kernel", int(os.environ.get("KERNEL", "2")))
print("=== timed step", flush=True)
ix.search_bounded_device(S["qtest"], 100, 10, acc, npb, D, I)
torch.cuda.synchronize()
print([(int(r["r0"]), int(r["w"]), round(r["tc_ms"], 3)) for r in ix.round_stats() if r.get("tc")])
