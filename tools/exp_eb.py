"""GPU tool: the bench step at another error bound, several calls, per-round timing."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
ap = argparse.ArgumentParser()
ap.add_argument("--eb", type=float, default=0.05)
a0 = ap.parse_args()
a = argparse.Namespace(shape="sift", nb=10_000_000, ncal=5000, nq=10000, nlist=4096, eb=0.1)
S = B.build_everything(a, 0, int(os.environ.get("QRANK", "0")))
ix, dev = S["ix"], S["dev"]
n = a.nq
q = S["qtest"]
acc = torch.full((n,), 1.0 - a0.eb, device=dev)
npb = torch.zeros(n, dtype=torch.int64, device=dev)
D = torch.empty(n, 100, device=dev)
I = torch.empty(n, 100, dtype=torch.int64, device=dev)
for eb in (0.1, a0.eb, a0.eb, a0.eb):
    ix.set_params(*B.HYPER[eb])
    acc.fill_(1.0 - eb)
    npb.zero_()
    ix.search_bounded_device(q, 100, 10, acc, npb, D, I)
    st = ix.stats()
    print(eb, "search_ms", round(st["search_ms"], 3), "rounds", st["rounds"], "tc", st["tc_rounds"], "fallbacks", st["tc_fallbacks"],
          "coarse", round(st["coarse_ms"], 3), "scan", round(st["scan_ms"], 3), flush=True)
    for r in ix.round_stats():
        print("   ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
