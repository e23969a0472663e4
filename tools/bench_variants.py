#!/usr/bin/env python3
"""Old vs new kernels on one B200, same inputs, results compared bit for bit, timings side by side:
  * general-quorum growth: k_gm_growth<P,true> (PGX_GM_QUORUM=old) vs k_gm_quorum (default)  -- BASELINE config 3 shape
  * similarity, unweighted: plain AND + POPC (PGX_SIM=plain) vs carry-save pairs (default)    -- BASELINE config 4 shape
Usage: python tools/bench_variants.py [--quick]   (one JSON line per measurement)"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth

quick = "--quick" in sys.argv


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return out, float(np.median(ts)) * 1e3


def report(**kw):
    print(json.dumps(kw), flush=True)


def with_env(key, val, fn):
    old = os.environ.get(key)
    if val is None:
        os.environ.pop(key, None)
    else:
        os.environ[key] = val
    try:
        return fn()
    finally:
        if old is None:
            os.environ.pop(key, None)
        else:
            os.environ[key] = old


def cutoffs(G, pairs):
    return [max(1, c) for c, _ in pairs], np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])


# ---- config 3: 5M x 512, (1,0) (2,0.5) (4,0.9), 100 orders ----
N, G, P = (1_000_000, 512, 10) if quick else (5_000_000, 512, 100)
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 3)
torch.cuda.synchronize()
a = pb.DeviceAbacus(N, G)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
orders = synth.random_orders(P, G, seed=synth.SEED_BASE + 3)
for pairs in ([(1, 0.0), (2, 0.5), (4, 0.9)], [(2, 0.5)], [(1, 0.1), (2, 0.5), (4, 0.9), (1, 1.0)]):
    cov, thr = cutoffs(G, pairs)
    for weighted in (False, True):
        res = {}
        for name, val in (("old", "old"), ("table", None)):
            out, ms = with_env("PGX_GM_QUORUM", val, lambda: timed(lambda: a.permuted_growth(orders, cov, thr, weighted=weighted)))
            res[name] = (out, ms, a.last_launch_info())
        same = bool(np.array_equal(res["old"][0], res["table"][0]))
        report(what="permuted_growth", N=N, G=G, orders=P, pairs=pairs, weighted=weighted, old_ms=round(res["old"][1], 3),
               table_ms=round(res["table"][1], 3), speedup=round(res["old"][1] / res["table"][1], 2), identical=same,
               launch=res["table"][2])
# q = 0 only (HBM-bound k_gm_growth<.,false>): grid mapping, column-fastest (old) vs order-fastest (L2 reuse across orders)
alg = N * ((G + 63) // 64) * 8
for weighted in (False, True):
    res = {}
    for name, val in (("col", "col"), ("order", None)):
        out, ms = with_env("PGX_GM_GRID", val, lambda: timed(lambda: a.permuted_growth(orders, [1], None, weighted=weighted)))
        res[name] = (out, ms)
    report(what="permuted_growth q=0 grid mapping", N=N, G=G, orders=P, weighted=weighted, col_fastest_ms=round(res["col"][1], 3),
           order_fastest_ms=round(res["order"][1], 3), identical=bool(np.array_equal(res["col"][0], res["order"][0])),
           algorithmic_gbps_order_fastest=round(P * alg / res["order"][1] / 1e6, 1))
cov3, thr3 = cutoffs(G, [(1, 0.0), (2, 0.5), (4, 0.9)])
out_c, ms_c = with_env("PGX_GM_GRID", "col", lambda: timed(lambda: a.permuted_growth(orders, cov3, thr3)))
out_o, ms_o = timed(lambda: a.permuted_growth(orders, cov3, thr3))
report(what="permuted_growth 3 pairs grid mapping (k_gm_quorum)", N=N, G=G, orders=P, col_fastest_ms=round(ms_c, 3),
       order_fastest_ms=round(ms_o, 3), identical=bool(np.array_equal(out_c, out_o)))
# group order, one pass (what ordered-histgrowth with a quorum runs on large tables)
cov, thr = cutoffs(G, [(1, 0.0), (2, 0.5), (4, 0.9)])
for weighted in (False, True):
    res = {}
    for name, val in (("old", "old"), ("table", None)):
        out, ms = with_env("PGX_GM_QUORUM", val, lambda: timed(lambda: a.ordered_growth(cov, thr, weighted=weighted), reps=5))
        res[name] = (out, ms)
    scan, ms_scan = with_env("PGX_QUORUM_PATH", "scan", lambda: timed(lambda: a.ordered_growth(cov, thr, weighted=weighted), reps=3))
    report(what="ordered_growth 3 pairs", N=N, G=G, weighted=weighted, old_ms=round(res["old"][1], 3), table_ms=round(res["table"][1], 3),
           node_major_ms=round(ms_scan, 3), identical=bool(np.array_equal(res["old"][0], res["table"][0]) and np.array_equal(scan, res["table"][0])))
a.close(); del bitmap, weight
torch.cuda.empty_cache()

# ---- config 4: similarity 10M x 1024 ----
N, G = (1_000_000, 1024) if quick else (10_000_000, 1024)
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 4)
torch.cuda.synchronize()
a = pb.DeviceAbacus(N, G)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
a.similarity(weighted=False, row_begin=0, row_end=64)  # transpose
words = (N + 64) // 64
for rows in (128, G):
    res = {}
    for name, val in (("plain", "plain"), ("csa", None)):
        out, ms = with_env("PGX_SIM", val, lambda: timed(lambda: a.similarity(weighted=False, row_begin=0, row_end=rows)))
        res[name] = (out, ms, a.last_launch_info())
    same = bool(np.array_equal(res["plain"][0][0], res["csa"][0][0]) and np.array_equal(res["plain"][0][1], res["csa"][0][1]))
    report(what="similarity", N=N, G=G, rows=rows, plain_ms=round(res["plain"][1], 3), csa_ms=round(res["csa"][1], 3),
           speedup=round(res["plain"][1] / res["csa"][1], 2), identical=same, launch=res["csa"][2],
           csa_pair_words_per_s=round(rows * G * words / res["csa"][1] * 1e3 / 1e12, 3))
a.close()
