#!/usr/bin/env python3
"""Fixed-cost probe: k_scan time vs item count at G=256 / 1024, with L2 write-flush, read-flush, no flush."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for G in (256, 1024):
    bitmap, _ = synth.torch_table(4_000_000, G, seed=1)
    for N in (1000, 100_000, 1_000_000, 2_000_000, 4_000_000):
        a = pb.DeviceAbacus(N, G); a.adopt_device(bitmap.data_ptr(), None, keepalive=bitmap); a.set_stream(stream.cuda_stream)
        out = torch.zeros(a.fused_out_words(1), dtype=torch.int64, device="cuda")
        res = {}
        for mode in ("write_flush", "read_flush", "none"):
            ts = []
            for _ in range(12):
                if mode == "write_flush": flush.fill_(1)
                elif mode == "read_flush": flush.sum()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); a.fused_pass_async(out.data_ptr(), [1], None); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[mode] = round(float(np.median(ts[2:])) * 1e3, 2)
        print(json.dumps({"G": G, "N": N, "us": res, "launch": a.last_launch_info()}), flush=True)
        a.close()
