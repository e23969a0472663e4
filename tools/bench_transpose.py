#!/usr/bin/env python3
"""k_transpose (node-major -> group-major copy) alone: CUDA-event time via the library's kernel timing, per shape."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth
peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6552.6
dev = torch.device("cuda", 0)
for N, G in [(10_000_000, 1024), (5_000_000, 512), (10_000_000, 256), (10_000_000, 64), (3_760_000, 44)]:
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 4, device=dev)
    ts = []
    for it in range(6):
        a = pb.DeviceAbacus(N, G, device=0)
        a.adopt_device(bitmap.data_ptr(), None, keepalive=bitmap)
        a.set_timing(True)
        inter, ln = a.similarity(row_begin=0, row_end=0)   # derives the group-major copy, computes len[] only
        ms, n = a.kernel_time_ms()
        ts.append(ms)
        chk = int(ln.sum())
        a.close()
    W = (G + 63) // 64
    bytes_ = 2 * N * W * 8
    us = float(np.median(ts[1:])) * 1e3
    print(json.dumps({"N": N, "G": G, "transpose_us": round(us, 1), "gbps_read_plus_write": round(bytes_ / us / 1e3, 1),
                      "frac_of_hbm": round(bytes_ / us / 1e3 / peak, 3), "len_checksum": chk}), flush=True)
    del bitmap, weight
