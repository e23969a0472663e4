#!/usr/bin/env python3
"""bench.py -- ordered-histgrowth throughput (item x group cells / s) of the fused B200 pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload target|c2]

A "step" is one fused pass (coverage histogram + ordered growth, count = node, coverage 1, quorum 0)
over one resident abacus bitmap:

  target  10,000,000 items x 1024 groups  (BASELINE.json north_star target shape; 1.28 GB > L2)   [default]
  c2       1,000,000 items x  256 groups  (BASELINE.json configs[1]; 32 MB, L2 flushed between steps)

N > 1 (torchrun, one rank per GPU): every rank holds its own item-range shard of the same shape
(weak scaling) and each step ends with the path's one exchange: an all-reduce of the KB-sized
u64 result vector (NCCL, integer sum -> order independent, bit exact).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
algorithm (oracle/, the Rust reference cannot be built in this image) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "target": dict(n_items=10_000_000, n_groups=1024,
                   name="ordered-histgrowth synthetic 10M items x 1024 groups, count=node, coverage=1 quorum=0 "
                        "(north_star target shape)"),
    "c2": dict(n_items=1_000_000, n_groups=256,
               name="ordered-histgrowth synthetic 1M items x 256 groups, count=node, coverage=1 quorum=0 "
                    "(BASELINE.json configs[1])"),
}
METRIC = "ordered-histgrowth item x group cells per second"
UNIT = "cells/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the scan kernel from the committed ncu summary, if any."""
    path = os.path.join(ROOT, "profiles", "scan_ncu_summary.json")
    try:
        d = json.load(open(path))
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the reference algorithm (oracle port) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------

def cpu_sample_table(n_items, n_groups, seed):
    from panacus_b200 import synth
    bits, bitmap, weight = synth.numpy_table(n_items, n_groups, seed=seed)
    return bitmap, weight


def cpu_reference_pass(tables, n_groups):
    """One pass of the reference hot path at the seam ItemTable -> (hist, ordered growth curve):
    AbacusByTotal::coverage + construct_hist, AbacusByGroup CSR build + calc_growth (c=1, q=0)."""
    from oracle import oracle as po
    items, prefsum, op, og, n_items = tables
    t0 = time.perf_counter()
    countable = po.abacus_by_total(n_items, items, prefsum, op, og)
    hist = po.construct_hist(countable, n_groups)
    t1 = time.perf_counter()
    r, c, v = po.csr_build(n_items, items, prefsum, op, og)
    t2 = time.perf_counter()
    curve = po.calc_growth(r, c, n_groups, po.absolute(1), po.relative(0.0))
    t3 = time.perf_counter()
    return dict(hist_s=t1 - t0, csr_s=t2 - t1, growth_s=t3 - t2, total_s=t3 - t0, hist=hist, curve=curve)


def make_cpu_tables(n_items, n_groups, seed):
    from oracle import oracle as po
    bitmap, weight = cpu_sample_table(n_items, n_groups, seed)
    items, prefsum, op, og = po.bitmap_to_item_table(bitmap, n_groups)
    return (items, prefsum, op, og, n_items), bitmap


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    G = wl["n_groups"]
    sample_items = args.cpu_sample_items or (100_000 if G >= 1024 else 400_000)
    tables, _ = make_cpu_tables(sample_items, G, seed=0x5EED0001)
    for _ in range(args.warmup):
        cpu_reference_pass(tables, G)
    t0 = time.perf_counter()
    parts = [cpu_reference_pass(tables, G) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    cells = float(sample_items) * G * args.steps
    value = cells / dt
    sample = (f"{sample_items} items x {G} groups per step (same generator as the GPU workload); seam = ItemTable -> "
              f"coverage+hist ({np.mean([p['hist_s'] for p in parts]):.3f}s) + CSR build "
              f"({np.mean([p['csr_s'] for p in parts]):.3f}s) + calc_growth c=1 q=0 "
              f"({np.mean([p['growth_s'] for p in parts]):.3f}s); reference parallelises only across threshold "
              f"pairs (1 pair -> 1 thread)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["name"], "n_items": wl["n_items"], "n_groups": G, "count": "node",
                   "coverage": 1, "quorum": 0},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    import panacus_b200 as pb
    from panacus_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    N, G = args.items or wl["n_items"], args.groups or wl["n_groups"]
    Wp = pb.row_words(G)
    W = (G + 63) // 64

    # synthetic shard, generated in HBM (each rank: its own item range, different seed)
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 1 + rank, device=dev)
    del weight  # count = node: unit weights, no weight traffic
    torch.cuda.synchronize()

    a = pb.DeviceAbacus(N, G, device=local_rank)
    a.adopt_device(bitmap.data_ptr(), None, keepalive=bitmap)
    # one explicit (non-default) stream carries the kernels, the events that time them and the NCCL exchange
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    a.set_stream(stream.cuda_stream)
    T = 1
    fused = world > 1 and args.exchange == "fused"
    if fused:  # in-kernel all-reduce over NVLink peer memory: the step stays ONE kernel launch
        from panacus_b200 import sharding
        sharding.connect_fused_exchange(a)
    nccl_exchange = world > 1 and not fused
    out = torch.zeros(a.fused_out_words(T), dtype=torch.int64, device=dev)
    cov = [1]

    flush = None
    bitmap_bytes = (N + 1) * Wp * 8
    if bitmap_bytes < 256 * 1024 * 1024:  # smaller than 2x L2: flush L2 between steps
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        if flush is not None:
            flush.fill_(1)
        a.fused_pass_async(out.data_ptr(), cov, None, weighted=False, hist_count=True, hist_weight=False)
        if nccl_exchange:
            dist.all_reduce(out)  # the path's one exchange: sum of per-shard integer results

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, device time via CUDA events on the launching stream ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = a.launch_count
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()  # all ranks enter the timed region together (the passes are collective)
        torch.cuda.synchronize()
    profiling = os.environ.get("PGX_PROFILE_RANGE") == "1"
    if profiling:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        k_ev[i][0].record()
        a.fused_pass_async(out.data_ptr(), cov, None, weighted=False, hist_count=True, hist_weight=False)
        k_ev[i][1].record()
        if nccl_exchange:
            dist.all_reduce(out)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if profiling:
        torch.cuda.cudart().cudaProfilerStop()
    launches = a.launch_count - launches0
    res = out.cpu().numpy().view(np.uint64).copy()
    kernel_ms = [s.elapsed_time(e) for s, e in k_ev]
    if os.environ.get("PGX_DEBUG_TIMES") == "1":
        starts = [e0.elapsed_time(s) for s, _ in k_ev]
        print(f"[rank {rank}] kernel_ms", [round(x, 3) for x in kernel_ms[:12]], "start offsets",
              [round(x, 3) for x in starts[:12]], file=sys.stderr, flush=True)
    if flush is not None:
        total_ms = float(np.sum(kernel_ms))  # L2-flush writes are not part of the step
    else:
        total_ms = e0.elapsed_time(e1)
    t_max = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms = float(t_max.item())
    # keep every GPU under the same load a little longer so the clock sampler sees it (same count on all ranks:
    # with the fused exchange every pass is collective)
    n_extra = max(1, min(5000, int(250.0 / max(total_ms / args.steps, 1e-3))))
    for _ in range(n_extra):
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps
    cells_per_step = float(N) * G * world
    value = cells_per_step / (ms_per_step * 1e-3)

    # sanity: the result must be a real growth curve (guards against timing a no-op)
    hc, _, curve = pb.curve_from_fused(res, G, T)
    tot_items = N * world if world > 1 else N
    assert int(hc.sum()) == tot_items, (int(hc.sum()), tot_items)
    assert int(curve[0, -1]) == tot_items - int(hc[0])

    # ---- roofline of the dominant kernel (k_scan): algorithmic bytes / mean launch duration ----
    peak, peak_src = measured_peaks()
    alg_bytes = float(N) * W * 8  # N * ceil(G/64) * 8 (unweighted)
    mean_kernel_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (mean_kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload), "peak_source": peak_src, "kernel": "k_scan<fast>",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_mean": mean_kernel_ms,
                "kernel_ms_min": float(np.min(kernel_ms)), "launch": a.last_launch_info()}

    # ---- e2e: host bitmap (pinned) -> H2D -> fused pass -> D2H of the result, every step ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty((N + 1, Wp), dtype=torch.int64, pin_memory=True)
        host.copy_(bitmap)
        torch.cuda.synchronize()
        host_np = host.numpy().view(np.uint64)
        b = pb.DeviceAbacus(N, G, device=local_rank)
        if fused:
            sharding.connect_fused_exchange(b)
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        b.upload(host_np)  # warm-up (allocations, first-touch)
        b.hist_ordered_growth(cov, None, weighted=False)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            b.upload(host_np)
            hc2, _, cv2 = b.hist_ordered_growth(cov, None, weighted=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dt = float(t_e.item())
        assert np.array_equal(cv2[0], np.cumsum(res[2 * (G + 1):2 * (G + 1) + G], dtype=np.uint64)) or world > 1
        e2e = {"value": cells_per_step * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": int(bitmap_bytes),
               "d2h_bytes_per_step": int((2 * (G + 1) + G) * 8), "steps": n_e2e,
               "api": "DeviceAbacus.upload(host bitmap) + hist_ordered_growth -> pgx_abacus_upload + "
                      "pgx_hist_ordered_growth"}
        b.close()
        del host

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_items = args.cpu_sample_items or (200_000 if G >= 1024 else 800_000)
        tables, sbitmap = make_cpu_tables(sample_items, G, seed=0x5EED0001)
        r = cpu_reference_pass(tables, G)
        # the GPU path must reproduce the CPU result on this very sample
        with pb.DeviceAbacus(sample_items, G, device=local_rank) as c:
            c.upload(sbitmap)
            hc3, _, cv3 = c.hist_ordered_growth([1], None, weighted=False)
        parity = bool(np.array_equal(hc3, r["hist"]) and np.array_equal(cv3[0].astype(np.float64), r["curve"]))
        cpu = {"value": float(sample_items) * G / r["total_s"], "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{sample_items} items x {G} groups, one pass: coverage+hist {r['hist_s']:.2f}s + CSR build "
                         f"{r['csr_s']:.2f}s + calc_growth {r['growth_s']:.2f}s (oracle/ C restatement; the reference "
                         f"runs this path on one thread for a single threshold pair)",
               "growth_only_value": float(sample_items) * G / r["growth_s"], "gpu_matches_cpu_on_sample": parity}
        assert parity, "GPU result differs from the CPU oracle on the baseline sample"

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["name"], "n_items_per_gpu": N, "n_groups": G, "count": "node", "coverage": 1,
                       "quorum": 0, "sharding": "item ranges, one shard per GPU" if world > 1 else "single GPU",
                       "exchange": ("none" if world == 1 else
                                    "in-kernel all-reduce over NVLink peer memory (last CTA pushes %d u64 words to every rank)"
                                    % out.numel() if fused else "ncclAllReduce(int64 sum) of %d words per step" % out.numel()),
                       "cache": "L2 flushed between steps (256 MB write)" if flush is not None
                                else "input %.2f GB per GPU >> 126 MB L2, no flush" % (bitmap_bytes / 1e9)},
            "gbps_per_gpu": alg_bytes / (ms_per_step * 1e-3) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    a.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="target")
    ap.add_argument("--items", type=int, default=0, help="override items per GPU")
    ap.add_argument("--groups", type=int, default=0)
    ap.add_argument("--cpu-sample-items", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--exchange", choices=["fused", "nccl"], default="fused",
                    help="N > 1: how the per-shard results are summed (fused = inside k_scan over NVLink peer memory)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
