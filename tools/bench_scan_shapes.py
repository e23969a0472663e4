#!/usr/bin/env python3
"""k_scan / k_scan_priv over a list of shapes: median CUDA-event time of the fused pass (hist + q = 0 growth), L2 flushed
between launches for inputs < 256 MB, for the lane-private-counter kernel (default) and the shared-atomics kernel
(PGX_SCAN_PRIV=2).  One JSON line per case.   python tools/bench_scan_shapes.py [--quick]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import synth

peak = 6552.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
quick = "--quick" in sys.argv
shapes = [(10_000_000, 44), (10_000_000, 64), (10_000_000, 100), (10_000_000, 128), (10_000_000, 256), (1_000_000, 256),
          (1_000_000, 44), (3_760_000, 44), (10_000_000, 300), (10_000_000, 512)]
if quick:
    shapes = shapes[:2] + shapes[4:6]
for N, G in shapes:
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 7, device=dev)
    Wp = pb.row_words(G)
    a = pb.DeviceAbacus(N, G, device=0)
    a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
    a.set_stream(stream.cuda_stream)
    for mode, kw in (("count T=1", dict(cov=[1], weighted=False, hist_count=True, hist_weight=False)),
                     ("count T=3", dict(cov=[1, 2, 4], weighted=False, hist_count=True, hist_weight=False)),
                     ("count hist only", dict(cov=[], weighted=False, hist_count=True, hist_weight=False)),
                     ("bp T=1", dict(cov=[1], weighted=True, hist_count=False, hist_weight=True)),
                     ("bp hist only", dict(cov=[], weighted=False, hist_count=False, hist_weight=True))):
        T = len(kw["cov"])
        out = torch.zeros(a.fused_out_words(max(T, 1)), dtype=torch.int64, device=dev)
        res = {}
        for name, env in (("priv", None), ("atomics", "2"), ("novert", "v")):
            os.environ.pop("PGX_SCAN_PRIV", None)
            os.environ.pop("PGX_SCAN_VERT", None)
            if env == "v":
                os.environ["PGX_SCAN_VERT"] = "2"  # the library's choice without the vertical-counter kernel
            elif env:
                os.environ["PGX_SCAN_PRIV"] = env
            ts = []
            for it in range(14):
                if (N + 1) * Wp * 8 < 256 * 1024 * 1024:
                    flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                a.fused_pass_async(out.data_ptr(), kw["cov"], None, weighted=kw["weighted"], hist_count=kw["hist_count"],
                                   hist_weight=kw["hist_weight"])
                e1.record()
                torch.cuda.synchronize()
                if it >= 4:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            res[name] = (float(np.median(ts)), int(out.cpu().numpy().view(np.uint64).sum() % (1 << 61)), a.last_launch_info())
        os.environ.pop("PGX_SCAN_PRIV", None)
        os.environ.pop("PGX_SCAN_VERT", None)
        bytes_ = N * ((G + 63) // 64) * 8 + (4 * N if (kw["weighted"] or kw["hist_weight"]) else 0)
        us = res["priv"][0]
        print(json.dumps({"N": N, "G": G, "mode": mode, "priv_us": round(us, 2), "atomics_us": round(res["atomics"][0], 2), "novert_us": round(res["novert"][0], 2),
                          "gbps": round(bytes_ / us / 1e3, 1), "frac_of_hbm": round(bytes_ / us / 1e3 / peak, 3),
                          "same_result": res["priv"][1] == res["atomics"][1] == res["novert"][1], "launch": res["priv"][2]}), flush=True)
    a.close()
    del bitmap, weight
