#!/usr/bin/env python3
"""k_scan<fast> time at small G with / without the joint histogram (PGX_SCAN_JOINT=2 disables it)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
N = 10_000_000
for G in (44, 64, 100, 128, 256):
    bitmap, weight = synth.torch_table(N, G, seed=3)
    for weighted in (False, True):
        a = pb.DeviceAbacus(N, G); a.adopt_device(bitmap.data_ptr(), weight.data_ptr() if weighted else None, keepalive=(bitmap, weight)); a.set_stream(stream.cuda_stream)
        out = torch.zeros(a.fused_out_words(3), dtype=torch.int64, device="cuda")
        res = {}
        for T in (1, 3):
            for mode in ("joint", "plain"):
                os.environ["PGX_SCAN_JOINT"] = "2" if mode == "plain" else "0"
                cov = [1, 2, 3][:T]
                ts = []
                for _ in range(12):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); a.fused_pass_async(out.data_ptr(), cov, None, weighted=weighted, hist_count=not weighted, hist_weight=weighted); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                res[f"T{T}_{mode}_us"] = round(float(np.median(ts[2:])) * 1e3, 1)
                res[f"T{T}_{mode}_chk"] = int(out.sum().item())
        alg = N * ((G + 63) // 64) * 8 + (4 * N if weighted else 0)
        res["gbps_T1_joint"] = round(alg / res["T1_joint_us"] / 1e3, 0)
        print(json.dumps({"G": G, "weighted": weighted, **res, "launch": a.last_launch_info()[:90]}), flush=True)
        a.close()
