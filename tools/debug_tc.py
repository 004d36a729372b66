import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auncel_b200 as ab
from tests.util import PARAMS, golden_case, golden_traces
for name in ["l2_d16", "ip_d24"]:
    c, g, xb, q = golden_case(name)
    ix = ab.IndexIVFFlat(c["d"], c["nlist"], c["metric"])
    ix.set_centroids(g["centroids"]); ix.add(xb)
    ts = int(g["ts"])
    for pi in [0, 2]:
        mult, stdm, eb = PARAMS[pi]
        ix.set_error_model(golden_traces(g), mult, stdm)
        for mode in [1, 2]:
            ix.set_option("tensor_core_filter", mode)
            D, I, mynp = ix.search_bounded(q[ts:], c["k"], c["qk"], g[f"b{pi}_acc"][ts:])
            st = ix.stats()
            print(name, pi, "mode", mode, "np eq", np.array_equal(mynp, g[f"b{pi}_my_nprobe"]), "D eq", np.array_equal(D, g[f"b{pi}_D"]),
                  {k: st[k] for k in ["rounds", "scan_pairs", "tc_rounds", "tc_candidates", "tc_fallbacks", "ndis"]})
    ix.set_option("tensor_core_filter", 2)
    ix.nprobe = 16
    D, I = ix.search(q, c["k"])
    st = ix.stats()
    print(name, "fixed16 D eq", np.array_equal(D, g["fixed_D_16"]), {k: st[k] for k in ["rounds", "scan_pairs", "tc_rounds", "tc_candidates", "tc_fallbacks", "ndis"]})
