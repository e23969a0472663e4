#!/usr/bin/env python3
"""A few launches of the fused pass for one shape (for ncu captures):  python tools/one_scan.py N G [count|bp] [T]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth
N, G = int(sys.argv[1]), int(sys.argv[2])
bp = len(sys.argv) > 3 and sys.argv[3] == "bp"
T = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda", 0)
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 7, device=dev)
a = pb.DeviceAbacus(N, G, device=0)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
out = torch.zeros(a.fused_out_words(max(T, 1)), dtype=torch.int64, device=dev)
cov = [1, 2, 4][:T]
for _ in range(3):
    a.fused_pass_async(out.data_ptr(), cov, None, weighted=bp, hist_count=not bp, hist_weight=bp)
torch.cuda.synchronize()
print(a.last_launch_info())
