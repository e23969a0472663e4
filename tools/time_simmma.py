#!/usr/bin/env python3
"""k_sim_mma timing experiments at 10M x 1024: normal, no expansion (PGX_SIM_DEBUG=2), no MMAs (=3)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth
dev = torch.device("cuda", 0)
N, G = 10_000_000, 1024
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 4, device=dev)
a = pb.DeviceAbacus(N, G, device=0)
a.adopt_device(bitmap.data_ptr(), None, keepalive=bitmap)
a.set_timing(True)
os.environ["PGX_SIM"] = "mma"
for dbg in (None, "2", "3", "4", "5"):
    if dbg: os.environ["PGX_SIM_DEBUG"] = dbg
    else: os.environ.pop("PGX_SIM_DEBUG", None)
    a.similarity(); a.kernel_time_ms()
    ts = []
    for _ in range(4):
        a.similarity(); ts.append(a.kernel_time_ms()[0])
    print(json.dumps({"debug": dbg or "normal", "ms": round(float(np.median(ts)), 3)}), flush=True)
