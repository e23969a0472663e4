#!/usr/bin/env python3
"""Small invocations of every round-2 kernel for compute-sanitizer (memcheck / racecheck / synccheck):
k_scan (direct epilogue, dynamic tiles), k_scan_priv (+ hist-only), k_scan_vert, k_scan<quorum>, k_transpose_reg (all tile
widths, permuted), k_gm_quorum on the coverage-sorted copy, k_gm_union, k_sim_mma, k_gm_similarity (bp), k_build, CSR."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth

for N, G in ((9000, 44), (6000, 130), (5000, 300), (4200, 1024), (3000, 2100)):
    bits, bitmap, weight = synth.numpy_table(N, G, seed=N + G)
    pairs = [(1, 0.0), (2, 0.5), (3, 0.9)]
    cov = [c for c, _ in pairs]
    thr = np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weight)
        a.hist(count=True, weight=True, countable=True)
        a.hist(count=True, weight=False)
        a.hist_ordered_growth([1], None, weighted=False, hist_count=True, hist_weight=False)
        a.hist_ordered_growth([1, 2], None, weighted=True, hist_count=False, hist_weight=True)
        os.environ["PGX_QUORUM_PATH"] = "scan"
        a.hist_ordered_growth(cov, thr, weighted=True, hist_count=True, hist_weight=True)
        os.environ.pop("PGX_QUORUM_PATH")
        orders = synth.random_orders(4, G, seed=3)
        a.permuted_growth(orders, cov, thr, weighted=False)   # coverage-sorted copy (4 orders)
        a.permuted_growth(orders, cov, thr, weighted=True)    # weight-sorted copy
        a.permuted_growth(orders, [1, 2], None, weighted=False)
        os.environ["PGX_SIM"] = "mma"
        a.similarity(weighted=False)
        os.environ.pop("PGX_SIM")
        a.similarity(weighted=True)
        info = a.last_launch_info()
    print(N, G, "ok", info, flush=True)
print("done")
