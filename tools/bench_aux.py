#!/usr/bin/env python3
"""Timings of the secondary hot-path kernels on one B200 (BASELINE.json configs 3 and 4 shapes):
general-quorum ordered growth, bp-weighted pass, permuted growth (group-major), similarity, transpose.
Usage: python tools/bench_aux.py [--quick]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth

quick = "--quick" in sys.argv
PEAK = 6541.5

def ev_time(fn, reps=5, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3

def report(name, ms, **kw):
    print(json.dumps({"what": name, "ms": round(ms, 4), **kw}), flush=True)

def cutoffs(G, pairs):
    return [max(1, c) for c, _ in pairs], np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])

# ---- config 3 shape: 5M x 512, pairs (1,0) (2,0.5) (4,0.9) ----
N, G = (1_000_000, 512) if quick else (5_000_000, 512)
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 3)
torch.cuda.synchronize()
a = pb.DeviceAbacus(N, G)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
cov, thr = cutoffs(G, pairs)
alg = N * ((G + 63) // 64) * 8
for weighted in (False, True):
    ms = ev_time(lambda: a.ordered_growth(cov, thr, weighted=weighted))
    report("ordered_growth 3 pairs (1 fast + 2 quorum, 2 launches)", ms, N=N, G=G, weighted=weighted,
           gbps_per_pass=round(2 * alg / ms / 1e6, 1))
    ms = ev_time(lambda: a.ordered_growth(cov[1:], thr[1:], weighted=weighted))
    report("ordered_growth quorum-only 2 pairs (k_scan<quorum>)", ms, N=N, G=G, weighted=weighted, gbps=round(alg / ms / 1e6, 1))
    ms = ev_time(lambda: a.hist_ordered_growth([1], None, weighted=weighted, hist_count=not weighted, hist_weight=weighted))
    report("hist + growth c=1 q=0 (k_scan<fast>)", ms, N=N, G=G, weighted=weighted, gbps=round((alg + (4 * N if weighted else 0)) / ms / 1e6, 1))
P = 8 if quick else 100
orders = synth.random_orders(P, G, seed=synth.SEED_BASE + 3)
t0 = time.perf_counter(); a.permuted_growth(orders[:1], [1], None); torch.cuda.synchronize()
report("first permuted call (includes transpose + countable)", (time.perf_counter() - t0) * 1e3, N=N, G=G)
for weighted in (False, True):
    ms = ev_time(lambda: a.permuted_growth(orders, [1], None, weighted=weighted), reps=3)
    report(f"permuted_growth {P} orders, c=1 q=0", ms, N=N, G=G, weighted=weighted, ms_per_order=round(ms / P, 4),
           gbps=round(P * alg / ms / 1e6, 1), frac_hbm=round(P * alg / ms / 1e6 / PEAK, 3))
    ms = ev_time(lambda: a.permuted_growth(orders, cov, thr, weighted=weighted), reps=3)
    report(f"permuted_growth {P} orders, 3 pairs (1,0)(2,.5)(4,.9)", ms, N=N, G=G, weighted=weighted, ms_per_order=round(ms / P, 4),
           gbps=round(P * alg / ms / 1e6, 1), frac_hbm=round(P * alg / ms / 1e6 / PEAK, 3))
a.close(); del bitmap, weight
# ---- device-side abacus build from an ItemTable vs the reference's host passes ----
from oracle import oracle as po
Nb, Gb = (100_000, 256) if quick else (400_000, 512)
bits_b, bitmap_b, _ = synth.numpy_table(Nb, Gb, seed=9)
items, prefsum, op, og = po.bitmap_to_item_table(bitmap_b, Gb)
b = pb.DeviceAbacus(Nb, Gb)
pg = np.arange(Gb, dtype=np.int64)
b.build(items, prefsum, pg)
assert np.array_equal(b.download(), bitmap_b)
b.clear()
ms = ev_time(lambda: b.build(items, prefsum, pg), reps=3)
t0 = time.perf_counter(); po.abacus_by_total(Nb, items, prefsum, op, og); t_cov = time.perf_counter() - t0
t0 = time.perf_counter(); po.csr_build(Nb, items, prefsum, op, og); t_csr = time.perf_counter() - t0
report("pgx_abacus_build (H2D of the ItemTable + k_build)", ms, N=Nb, G=Gb, steps=int(items.size), steps_per_s=round(items.size / ms * 1e3 / 1e9, 2),
       unit="G steps/s", cpu_coverage_pass_ms=round(t_cov * 1e3, 1), cpu_csr_build_ms=round(t_csr * 1e3, 1))
b.close()
# ---- config 4 shape: similarity 10M x 1024 ----
N, G = (1_000_000, 1024) if quick else (10_000_000, 1024)
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 4)
torch.cuda.synchronize()
a = pb.DeviceAbacus(N, G)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
t0 = time.perf_counter(); a.similarity(weighted=False, row_begin=0, row_end=64); torch.cuda.synchronize()
report("first similarity call, 64 rows (includes transpose)", (time.perf_counter() - t0) * 1e3, N=N, G=G)
words = (N + 64) // 64
for rows in ((128, G) if not quick else (128, G)):
    ms = ev_time(lambda: a.similarity(weighted=False, row_begin=0, row_end=rows), reps=3)
    ops = rows * G * words
    report(f"similarity rows [0,{rows}) x {G}", ms, N=N, G=G, weighted=False, pair_words_per_s=round(ops / ms * 1e3 / 1e12, 3), unit="T and+popc(64b)/s")
ms = ev_time(lambda: a.similarity(weighted=True, row_begin=0, row_end=128), reps=2)
report("similarity bp-weighted rows [0,128)", ms, N=N, G=G, weighted=True)
a.close()
