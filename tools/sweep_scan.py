#!/usr/bin/env python3
"""Tuning sweep for k_scan: tile size / stages / CTAs per SM via the PGX_SCAN_* overrides.
Usage (GPU box): python tools/sweep_scan.py [n_items] [n_groups]"""
import itertools
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
weighted = len(sys.argv) > 3 and sys.argv[3] == "bp"
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 1)
torch.cuda.synchronize()
a = pb.DeviceAbacus(N, G)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr() if weighted else None, keepalive=(bitmap, weight))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
a.set_stream(stream.cuda_stream)
out = torch.zeros(a.fused_out_words(1), dtype=torch.int64, device="cuda")
W = (G + 63) // 64
alg = N * W * 8 + (4 * N if weighted else 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if alg < (256 << 20) else None
ref = None
rowbytes = pb.row_words(G) * 8
tiles = sorted({t for t in (128, 256, 512, 768, 1024, 2048, 3072, 4096, 8192) if 8192 <= t * rowbytes <= 98304})
for tile, ctas, stages in itertools.product([0] + tiles, [0, 1, 2], [0, 2, 3, 4, 6]):
    os.environ["PGX_SCAN_TILE"] = str(tile)
    os.environ["PGX_SCAN_CTAS"] = str(ctas)
    os.environ["PGX_SCAN_STAGES"] = str(stages)
    try:
        for _ in range(3):
            a.fused_pass_async(out.data_ptr(), [1], None, weighted=weighted, hist_count=True, hist_weight=weighted)
    except pb.PgxError as e:
        continue
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a.fused_pass_async(out.data_ptr(), [1], None, weighted=weighted, hist_count=True, hist_weight=weighted)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = out.cpu().numpy()
    if ref is None:
        ref = res.copy()
    ok = bool(np.array_equal(ref, res))
    ms = float(np.median(ts))
    print(json.dumps({"tile": tile, "ctas": ctas, "stages": stages, "ms": round(ms, 5), "gbps": round(alg / ms / 1e6, 1),
                      "ok": ok, "launch": a.last_launch_info()}), flush=True)
