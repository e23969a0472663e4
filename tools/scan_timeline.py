#!/usr/bin/env python3
"""Phase timeline of single k_scan* launches (PGX_SCAN_TS=1: globaltimer stamps per CTA, printed by the library on stderr)
for the shapes where the fixed cost dominates.  One warm-up launch, then one stamped launch per case after an L2 flush."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for N, G in [(1_000, 256), (1_000_000, 256), (1_000_000, 44), (3_760_000, 44), (10_000_000, 44), (10_000_000, 1024)]:
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 7, device=dev)
    a = pb.DeviceAbacus(N, G, device=0)
    a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
    out = torch.zeros(a.fused_out_words(1), dtype=torch.int64, device=dev)
    for variant, env in (("default", {}), ("novert", {"PGX_SCAN_VERT": "2"}), ("atomics", {"PGX_SCAN_PRIV": "2"})):
        for k in ("PGX_SCAN_VERT", "PGX_SCAN_PRIV"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for flush_mode in ("write", "read"):
            a.fused_pass_async(out.data_ptr(), [1], None, weighted=False, hist_count=True, hist_weight=False)
            torch.cuda.synchronize()
            if flush_mode == "write":
                flush.fill_(1)
            else:
                flush.sum()
            torch.cuda.synchronize()
            os.environ["PGX_SCAN_TS"] = "1"
            sys.stderr.write(f"== {N} x {G} {variant} after {flush_mode}-flush: ")
            sys.stderr.flush()
            a.fused_pass_async(out.data_ptr(), [1], None, weighted=False, hist_count=True, hist_weight=False)
            torch.cuda.synchronize()
            os.environ.pop("PGX_SCAN_TS", None)
            sys.stderr.write("   " + a.last_launch_info() + "\n")
    a.close()
    del bitmap, weight
