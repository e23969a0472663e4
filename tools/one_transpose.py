#!/usr/bin/env python3
"""Two derivations of the group-major copy for one shape (for ncu captures):  python tools/one_transpose.py N G"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth
N, G = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 4, device=dev)
for _ in range(2):
    a = pb.DeviceAbacus(N, G, device=0)
    a.adopt_device(bitmap.data_ptr(), None, keepalive=bitmap)
    a.similarity(row_begin=0, row_end=0)
    a.close()
