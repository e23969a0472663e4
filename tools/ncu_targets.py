#!/usr/bin/env python3
"""One launch of each secondary hot-path kernel at the BASELINE config 3 / 4 shapes, for `ncu --set full`:
  ncu --set full --clock-control none --import-source on -k regex:'k_gm_quorum|k_gm_similarity' -c 8 \
      -o gpurun_out/r1_secondary python tools/ncu_targets.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth

N, G, P = 5_000_000, 512, 20
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 3)
torch.cuda.synchronize()
a = pb.DeviceAbacus(N, G)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
orders = synth.random_orders(P, G, seed=synth.SEED_BASE + 3)
pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
cov = [c for c, _ in pairs]
thr = np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])
a.permuted_growth(orders, [1], None)                     # k_gm_quorum<1,0,1,false>: q = 0 only
a.permuted_growth(orders, cov, thr)                      # k_gm_quorum<10,2,1,false>
a.permuted_growth(orders, [1], None, weighted=True)      # k_gm_quorum<1,0,1,true>  (weight-sorted copy)
a.permuted_growth(orders, cov, thr, weighted=True)       # k_gm_quorum<10,2,1,true>
a.close(); del bitmap, weight
torch.cuda.empty_cache()
N, G = 10_000_000, 1024
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 4)
torch.cuda.synchronize()
a = pb.DeviceAbacus(N, G)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
a.similarity(weighted=False, row_begin=0, row_end=256)   # k_gm_similarity<false,true> (carry-save)
a.close()
print("done")
