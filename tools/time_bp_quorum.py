#!/usr/bin/env python3
"""bp-weighted permuted growth with general thresholds (the c3 shape, bp instead of counting): weight-sorted copy with and
without the coverage as the tie order among equal weights (PGX_GM_COVSORT=0).  Kernel time via the library's timing."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth
N, G, P = 5_000_000, 512, 100
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 3)
orders = synth.random_orders(P, G, seed=synth.SEED_BASE + 3)
pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
cov = [c for c, _ in pairs]
thr = np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])
res = {}
for mode in ("weight+coverage", "weight only"):
    if mode == "weight only":
        os.environ["PGX_GM_COVSORT"] = "0"
    else:
        os.environ.pop("PGX_GM_COVSORT", None)
    a = pb.DeviceAbacus(N, G)
    a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
    out = a.permuted_growth(orders, cov, thr, weighted=True)   # builds the copies
    a.set_timing(True); a.kernel_time_ms()
    for _ in range(3):
        out = a.permuted_growth(orders, cov, thr, weighted=True)
    ms, n = a.kernel_time_ms()
    res[mode] = (ms / 3, int(out.astype(np.uint64).sum() % (1 << 61)), a.last_launch_info())
    a.close()
print(json.dumps({"shape": [N, G, P], "pairs": pairs, "kernel_ms": {k: round(v[0], 3) for k, v in res.items()},
                  "same_result": res["weight+coverage"][1] == res["weight only"][1], "launch": res["weight+coverage"][2]}))
