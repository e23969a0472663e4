#!/usr/bin/env python3
"""tensor-core similarity (PGX_SIM=mma) against the AND / POPC kernel (PGX_SIM=csa): bit-exact on several shapes, then timing."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth

def run(a, mode, **kw):
    os.environ["PGX_SIM"] = mode
    try:
        return a.similarity(**kw)
    finally:
        os.environ.pop("PGX_SIM", None)

ok = True
for N, G in [(3000, 256), (70_000, 300), (200_001, 1024), (5000, 700), (127, 256), (100_000, 257), (64 * 7 + 1, 512)]:
    bits, bitmap, weight = synth.numpy_table(N, G, seed=N + G)
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weight)
        i0, l0 = run(a, "csa")
        i1, l1 = run(a, "mma")
        same = bool(np.array_equal(i0, i1) and np.array_equal(l0, l1))
        p0, _ = run(a, "csa", row_begin=64, row_end=min(G, 300))
        p1, _ = run(a, "mma", row_begin=64, row_end=min(G, 300))
        u0, _ = run(a, "csa", row_begin=128, row_end=min(G, 256), upper=True)
        u1, _ = run(a, "mma", row_begin=128, row_end=min(G, 256), upper=True)
        same_rows = bool(np.array_equal(p0, p1))
        same_upper = bool(np.array_equal(np.triu(np.pad(u0, ((128, 0), (0, 0))))[128:], np.triu(np.pad(u1, ((128, 0), (0, 0))))[128:]))
        print(json.dumps({"N": N, "G": G, "full": same, "rows": same_rows, "upper": same_upper, "launch": a.last_launch_info(),
                          "max_abs_diff": int(np.abs(i0.astype(np.int64) - i1.astype(np.int64)).max())}), flush=True)
        ok = ok and same and same_rows and same_upper
if "--time" in sys.argv:
    dev = torch.device("cuda", 0)
    for N, G in [(10_000_000, 1024), (5_000_000, 512), (10_000_000, 256)]:
        bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 4, device=dev)
        a = pb.DeviceAbacus(N, G, device=0)
        a.adopt_device(bitmap.data_ptr(), None, keepalive=bitmap)
        a.set_timing(True)
        res = {}
        for mode in ("csa", "mma"):
            run(a, mode)
            a.kernel_time_ms()
            ts = []
            for _ in range(5):
                out = run(a, mode)
                ts.append(a.kernel_time_ms()[0])
            res[mode] = (float(np.median(ts)), int(out[0].sum() % (1 << 61)))
        print(json.dumps({"N": N, "G": G, "csa_ms": round(res["csa"][0], 3), "mma_ms": round(res["mma"][0], 3),
                          "same_checksum": res["csa"][1] == res["mma"][1]}), flush=True)
        a.close()
        del bitmap, weight
print("ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
