#!/usr/bin/env python3
"""bench.py -- throughput (item x group cells / s) of the panacus counting hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W] [--scaling weak|strong]

Workloads (BASELINE.json; `config.workload` names the one that ran):

  target  ordered-histgrowth, 10,000,000 items x 1024 groups, count=node, coverage 1 quorum 0   [default]
          (north_star target shape; 1.28 GB >> L2).  A step = one fused pass (coverage histogram + ordered growth)
          over the resident abacus bitmap.  N > 1: item-range shards, the path's one exchange (sum of the KB-sized
          result vectors) fused into the scan kernel over NVLink peer memory (--exchange nccl: ncclAllReduce).
          --scaling weak (default): one 10M x 1024 shard per GPU; --scaling strong: one graph cut into N item ranges.
  c2      configs[1]: 1M x 256, same pass; 32 MB, L2 flushed between steps.
  c3      configs[2]: growth under 100 random group orders, 5M x 512, (coverage, quorum) = (1,0) (2,0.5) (4,0.9);
          order p -> rank p % N, bitmap replicated, pgx_permuted_growth_sharded (ncclAllGather of device-resident curves).
  c4      configs[3]: all-pairs similarity, 10M x 1024; upper-triangle row blocks per rank, pgx_similarity_sharded.
  c5      configs[4] stand-in: `panacus histgrowth -c bp -S -q 0,0.5,1 -l 0,1,2` on a generated chr22-shaped GFA
          (the HPRC file is not available offline), wall clock with a phase breakdown.

`value`  whole-job throughput with the inputs resident in HBM (device time, CUDA events, max over ranks).
`e2e`    the same metric through the C ABI from HOST buffers.  target / c2: the seam the CPU reference is timed on --
         ItemTable (u32 ids, page-locked) -> pgx_abacus_build_u32 -> pgx_hist_ordered_growth -> results on the host;
         `e2e_bitmap_seam` keeps the packed-bitmap upload variant.  c3 / c4: bitmap upload (+ NVLink broadcast) + call.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/; the Rust reference cannot be built in
this image) on a bounded sample of the same workload and prints the same `config`.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "target": dict(n_items=10_000_000, n_groups=1024, kind="ordered",
                   name="ordered-histgrowth synthetic 10M items x 1024 groups, count=node, coverage=1 quorum=0 "
                        "(north_star target shape)"),
    "c2": dict(n_items=1_000_000, n_groups=256, kind="ordered",
               name="ordered-histgrowth synthetic 1M items x 256 groups, count=node, coverage=1 quorum=0 "
                    "(BASELINE.json configs[1])"),
    "c3": dict(n_items=5_000_000, n_groups=512, n_orders=100, kind="permuted",
               name="growth under 100 random group orders, synthetic 5M items x 512 groups, (coverage,quorum) = "
                    "(1,0) (2,0.5) (4,0.9), count=node (BASELINE.json configs[2])"),
    "c4": dict(n_items=10_000_000, n_groups=1024, kind="similarity",
               name="similarity (all-pairs intersections + Jaccard), synthetic 10M items x 1024 groups, count=node "
                    "(BASELINE.json configs[3])"),
    "c5": dict(n_items=3_759_736, n_groups=44, kind="cli",
               name="panacus histgrowth -c bp -S -q 0,0.5,1 -l 0,1,2 on a generated chr22-shaped GFA (3.76M segments, "
                    "44 sample groups, U-shaped coverage; stand-in for BASELINE.json configs[4], file not available offline)"),
}
METRIC = "ordered-histgrowth item x group cells per second"
UNIT = "cells/s"
C3_PAIRS = [(1, 0.0), (2, 0.5), (4, 0.9)]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """(dram bytes per launch of the dominant kernel, source) from the committed ncu summary of the same command."""
    path = os.path.join(ROOT, "profiles", "scan_ncu_summary.json")
    try:
        d = json.load(open(path)).get(workload, {})
        return d.get("dram_bytes_per_launch"), "profiles/scan_ncu_summary.json (ncu --set full of this command; not measured in this run)"
    except Exception:
        return None, None


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


def shape_of(args):
    wl = WORKLOADS[args.workload]
    return wl, args.items or wl["n_items"], args.groups or wl["n_groups"]


def workload_config(args, world):
    """The `config` object: identical for both arms (the reference arm describes the same workload)."""
    wl, N, G = shape_of(args)
    kind = wl["kind"]
    W = (G + 63) // 64
    cfg = {"workload": wl["name"], "n_groups": G, "count": "bp" if kind == "cli" else "node", "n_gpus": world}
    if kind == "ordered":
        strong = args.scaling == "strong"
        cfg.update({"n_items_per_gpu": N // world if strong else N, "n_items_total": N if strong else N * world,
                    "coverage": 1, "quorum": 0, "scaling": "strong" if strong else "weak",
                    "sharding": "single GPU" if world == 1 else "item ranges, one shard per GPU",
                    "exchange": "none" if world == 1 else
                                ("in-kernel all-reduce over NVLink peer memory (last CTA of k_scan pushes the result vector "
                                 "to every rank)" if args.exchange == "fused" else "ncclAllReduce(u64 sum) of the result vector"),
                    "cache": "L2 flushed between steps (256 MB write)" if (N // world if strong else N) * W * 8 < 256 * 2**20
                             else "input per GPU >> 126 MB L2, no flush"})
    elif kind == "permuted":
        cfg.update({"n_items_total": N, "n_orders": args.orders or wl["n_orders"],
                    "thresholds": [list(p) for p in C3_PAIRS], "scaling": "strong",
                    "sharding": "single GPU" if world == 1 else "bitmap replicated; order p on rank p % n_gpus; ncclAllGather of the curves",
                    "cache": "input 0.32 GB >> L2, no flush"})
    elif kind == "similarity":
        cfg.update({"n_items_total": N, "scaling": "strong",
                    "sharding": "single GPU (upper triangle + mirror)" if world == 1 else
                                "bitmap replicated; two folded upper-triangle row blocks per rank; ncclAllGather + device-side assembly",
                    "cache": "input 1.28 GB >> L2, no flush"})
    else:
        cfg.update({"n_items_total": N, "scaling": "weak", "sharding": "single GPU", "thresholds": [[0, 0.0], [1, 0.5], [2, 1.0]]})
    return cfg


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference algorithm (oracle port) on bounded samples
# ---------------------------------------------------------------------------------------------------

def make_cpu_tables(n_items, n_groups, seed):
    from oracle import oracle as po
    from panacus_b200 import synth
    bits, bitmap, weight = synth.numpy_table(n_items, n_groups, seed=seed)
    items, prefsum, op, og = po.bitmap_to_item_table(bitmap, n_groups)
    return (items, prefsum, op, og, n_items), bitmap, weight


def cpu_reference_pass(tables, n_groups, pairs=((1, 0.0),)):
    """One pass of the reference hot path at the seam ItemTable -> (hist, ordered growth curves):
    AbacusByTotal::coverage + construct_hist, AbacusByGroup CSR build + calc_growth per pair (serial here; the
    reference runs the pairs on rayon threads, src/analyses/ordered_histgrowth.rs:174-188)."""
    from oracle import oracle as po
    items, prefsum, op, og, n_items = tables
    t0 = time.perf_counter()
    countable = po.abacus_by_total(n_items, items, prefsum, op, og)
    hist = po.construct_hist(countable, n_groups)
    t1 = time.perf_counter()
    r, c, v = po.csr_build(n_items, items, prefsum, op, og)
    t2 = time.perf_counter()
    curves = [po.calc_growth(r, c, n_groups, po.absolute(cv), po.relative(q)) for cv, q in pairs]
    t3 = time.perf_counter()
    return dict(hist_s=t1 - t0, csr_s=t2 - t1, growth_s=t3 - t2, total_s=t3 - t0, hist=hist, curve=curves[0], curves=curves)


def reference_line(args, value, ms_per_step, sample, cores=1):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": workload_config(args, world)["scaling"], "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    wl, N, G = shape_of(args)
    kind = wl["kind"]
    if kind == "cli":
        return run_c5(args, reference=True)
    from oracle import oracle as po
    if kind == "ordered":
        sample_items = args.cpu_sample_items or (100_000 if G >= 1024 else 400_000)
        tables, _, _ = make_cpu_tables(sample_items, G, seed=0x5EED0001)
        for _ in range(args.warmup):
            cpu_reference_pass(tables, G)
        t0 = time.perf_counter()
        parts = [cpu_reference_pass(tables, G) for _ in range(args.steps)]
        dt = time.perf_counter() - t0
        value = float(sample_items) * G * args.steps / dt
        sample = (f"{sample_items} items x {G} groups per step (same generator as the GPU workload); seam = ItemTable -> "
                  f"coverage+hist ({np.mean([p['hist_s'] for p in parts]):.3f}s) + CSR build "
                  f"({np.mean([p['csr_s'] for p in parts]):.3f}s) + calc_growth c=1 q=0 "
                  f"({np.mean([p['growth_s'] for p in parts]):.3f}s); the reference parallelises only across threshold "
                  f"pairs (1 pair -> 1 thread)")
        print(json.dumps(reference_line(args, value, dt / args.steps * 1e3, sample)))
        return 0
    if kind == "permuted":
        # the reference computes one order per `--order` run: rebuild the abacus under the order, then calc_growth per pair
        sample_items, k_orders = args.cpu_sample_items or 100_000, 2
        from panacus_b200 import pack_bits, synth
        bits, _, _ = synth.numpy_table(sample_items, G, seed=0x5EED0003)
        orders = synth.random_orders(k_orders, G, seed=synth.SEED_BASE + 3)

        def one_step():
            t0 = time.perf_counter()
            for o in orders:
                items, prefsum, op, og = po.bitmap_to_item_table(pack_bits(bits[:, o]), G)  # not timed as reference work
                t1 = time.perf_counter()
                r, c, _ = po.csr_build(sample_items, items, prefsum, op, og)
                for cv, q in C3_PAIRS:
                    po.calc_growth(r, c, G, po.absolute(cv), po.relative(q))
                one_step.busy += time.perf_counter() - t1
            return time.perf_counter() - t0
        one_step.busy = 0.0
        for _ in range(min(args.warmup, 1)):
            one_step()
        one_step.busy = 0.0
        steps = max(1, min(args.steps, 3))
        for _ in range(steps):
            one_step()
        per_order = one_step.busy / (steps * k_orders)
        value = float(sample_items) * G / per_order  # cells of one order per second == cells x orders / (orders x per_order)
        sample = (f"{k_orders} orders x {len(C3_PAIRS)} pairs on {sample_items} items x {G} groups per step: CSR rebuild under the "
                  f"order + calc_growth per pair = {per_order:.3f} s per order, 1 thread (the reference would use up to 3 threads "
                  f"for the 3 pairs); throughput is per order, so it is comparable with cells x orders / time")
        line = reference_line(args, value, per_order * k_orders * 1e3, sample)
        line["steps"] = args.steps
        print(json.dumps(line))
        return 0
    if kind == "similarity":
        # Similarity::set_table is a hash-map update per ordered group pair per item: sum of deg^2 -- timed at a small shape
        sample_items, Gs = args.cpu_sample_items or 20_000, min(G, 256)
        tables, _, _ = make_cpu_tables(sample_items, Gs, seed=0x5EED0004)
        items, prefsum, op, og, n_items = tables
        r, c, _ = po.csr_build(n_items, items, prefsum, op, og)
        steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(steps):
            po.similarity(r, c, Gs)
        dt = (time.perf_counter() - t0) / steps
        value = float(sample_items) * Gs / dt
        sample = (f"{sample_items} items x {Gs} groups per step (the reference's per-item pair updates are sum(deg^2): the full "
                  f"shape is ~10^12 hash updates and is not run; cells/s at this smaller shape OVERSTATES the CPU at {G} groups, "
                  f"where the work per cell is {G // Gs}x larger); 1 thread (serial in the reference)")
        print(json.dumps(reference_line(args, value, dt * 1e3, sample)))
        return 0
    raise SystemExit("unknown workload kind")


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------

class Ctx:
    pass


def bind_to_gpu_numa_node(torch, local):
    """N > 1: run this rank (and allocate its page-locked tables) on the NUMA node its GPU hangs off, so that eight ranks'
    uploads do not cross the inter-socket link.  Best effort: returns the node, or None when it cannot be determined / the
    node's CPUs are not in this process' affinity mask / PGX_BENCH_NUMA=0."""
    if os.environ.get("PGX_BENCH_NUMA") == "0":
        return None
    try:
        prop = torch.cuda.get_device_properties(local)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def setup_gpu(args):
    import torch
    import torch.distributed as dist
    c = Ctx()
    c.torch, c.dist = torch, dist
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    c.numa_node = bind_to_gpu_numa_node(torch, c.local) if c.world > 1 else None
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)
    # one explicit (non-default) stream carries the kernels, the events that time them and the exchange
    c.stream = torch.cuda.Stream(device=c.dev)
    torch.cuda.set_stream(c.stream)
    assert c.stream.cuda_stream != 0
    return c


def max_over_ranks(c, x):
    t = c.torch.tensor([x], dtype=c.torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return float(t.item())


def barrier(c):
    c.torch.cuda.synchronize()
    if c.world > 1:
        c.dist.barrier()
        c.torch.cuda.synchronize()


def host_mem_available():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return 0


def gpu_item_table(c, bitmap, N, G):
    """ItemTable of the resident bitmap, one path per group (the table a GFA with P lines `g: id+,id+,...` parses to,
    src/util.rs:80-93): (items u32 in page-locked host memory, id_prefsum u64[G+1], path_group i64[G])."""
    import panacus_b200 as pb
    torch = c.torch
    counts = np.zeros(G, dtype=np.int64)
    for w in range((G + 63) // 64):
        word = bitmap[:, w].contiguous()
        for b in range(min(64, G - w * 64)):
            counts[w * 64 + b] = int(((word >> b) & 1).sum().item())
    prefsum = np.zeros(G + 1, dtype=np.uint64)
    prefsum[1:] = np.cumsum(counts).astype(np.uint64)
    total = int(prefsum[-1])
    items = pb.pinned_empty(total, np.uint32)
    view = torch.from_numpy(items.view(np.int32))
    for w in range((G + 63) // 64):
        word = bitmap[:, w].contiguous()
        for b in range(min(64, G - w * 64)):
            g = w * 64 + b
            ids = torch.nonzero((word >> b) & 1).view(-1).to(torch.int32)
            view[int(prefsum[g]):int(prefsum[g + 1])].copy_(ids)
    torch.cuda.synchronize()
    return items, prefsum, np.arange(G, dtype=np.int64)


def finish_line(c, args, line):
    if c.rank == 0:
        print(json.dumps(line))
    if c.world > 1:
        c.dist.barrier()
        c.dist.destroy_process_group()
    return 0


def run_ordered(args):
    import panacus_b200 as pb
    from panacus_b200 import sharding, synth
    c = setup_gpu(args)
    torch, dist, world, rank, dev = c.torch, c.dist, c.world, c.rank, c.dev
    wl, N_total, G = shape_of(args)
    strong = args.scaling == "strong"
    N = N_total // world if strong else N_total  # items per GPU
    Wp, W = pb.row_words(G), (G + 63) // 64

    # synthetic shard, generated in HBM (each rank: its own item range, different seed)
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 1 + rank, device=dev)
    del weight  # count = node: unit weights, no weight traffic
    torch.cuda.synchronize()

    a = pb.DeviceAbacus(N, G, device=c.local)
    a.adopt_device(bitmap.data_ptr(), None, keepalive=bitmap)
    a.set_stream(c.stream.cuda_stream)
    T = 1
    fused = world > 1 and args.exchange == "fused"
    if fused:  # in-kernel all-reduce over NVLink peer memory: the step stays ONE kernel launch
        sharding.connect_fused_exchange(a)
    nccl_exchange = world > 1 and not fused
    out = torch.zeros(a.fused_out_words(T), dtype=torch.int64, device=dev)
    cov = [1]

    flush = None
    bitmap_bytes = (N + 1) * Wp * 8
    if bitmap_bytes < 256 * 1024 * 1024:  # smaller than 2x L2: flush L2 between steps
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush_l2():
        # 256 MB write evicts the table from L2; the spin kernel after it (~100 us) keeps the stream busy while the host
        # enqueues the event and the pass, so that no host-side launch gap lands between the two events of a step
        flush.fill_(1)
        torch.cuda._sleep(200_000)

    def step():
        if flush is not None:
            flush_l2()
        a.fused_pass_async(out.data_ptr(), cov, None, weighted=False, hist_count=True, hist_weight=False)
        if nccl_exchange:
            dist.all_reduce(out)  # the path's one exchange: sum of per-shard integer results

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, device time via CUDA events on the launching stream ----
    sampler = ClockSampler(c.local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = a.launch_count
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(c)  # all ranks enter the timed region together (the passes are collective)
    profiling = os.environ.get("PGX_PROFILE_RANGE") == "1"
    if profiling:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for i in range(args.steps):
        if flush is not None:
            flush_l2()
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
    if fused:
        a.exchange_status()  # raises if the in-kernel watchdog saw a peer missing
    res = out.cpu().numpy().view(np.uint64).copy()
    kernel_ms = [s.elapsed_time(e) for s, e in k_ev]
    if os.environ.get("PGX_DEBUG_TIMES") == "1":
        print(f"[rank {rank}] kernel_ms", [round(x, 3) for x in kernel_ms[:12]], file=sys.stderr, flush=True)
    total_ms = float(np.sum(kernel_ms)) if flush is not None else e0.elapsed_time(e1)  # L2-flush writes are not part of the step
    total_ms = max_over_ranks(c, total_ms)
    # keep every GPU under the same load a little longer so the clock sampler sees it (same count on all ranks)
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
    tot_items = N * world
    assert int(hc.sum()) == tot_items, (int(hc.sum()), tot_items)
    assert int(curve[0, -1]) == tot_items - int(hc[0])

    # ---- roofline of the dominant kernel (k_scan): algorithmic bytes / mean launch duration ----
    peak, peak_src = measured_peaks()
    alg_bytes = float(N) * W * 8  # N * ceil(G/64) * 8 (unweighted)
    mean_kernel_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (mean_kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.workload) if not strong and not args.items else (None, None)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "k_scan<fast>",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_mean": mean_kernel_ms,
                "kernel_ms_min": float(np.min(kernel_ms)), "launch": a.last_launch_info()}

    # ---- e2e at the reference's seam: ItemTable (host) -> build -> fused pass -> results on the host ----
    e2e = e2e_bitmap = None
    if not args.no_e2e:
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        # (a) bitmap seam: packed host bitmap (pinned) -> H2D -> pass -> D2H
        host = torch.empty((N + 1, Wp), dtype=torch.int64, pin_memory=True)
        host.copy_(bitmap)
        torch.cuda.synchronize()
        host_np = host.numpy().view(np.uint64)
        b = pb.DeviceAbacus(N, G, device=c.local)
        if fused:
            sharding.connect_fused_exchange(b)
        b.upload(host_np)  # warm-up (allocations, first touch)
        b.hist_ordered_growth(cov, None, weighted=False)
        barrier(c)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            b.upload(host_np)
            hc2, _, cv2 = b.hist_ordered_growth(cov, None, weighted=False)
        torch.cuda.synchronize()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        want_curve = np.cumsum(res[2 * (G + 1):2 * (G + 1) + G], dtype=np.uint64)
        assert np.array_equal(cv2[0], want_curve)
        e2e_bitmap = {"value": cells_per_step * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": int(bitmap_bytes),
                      "d2h_bytes_per_step": int((2 * (G + 1) + G) * 8), "steps": n_e2e,
                      "api": "pgx_abacus_upload(packed host bitmap, pinned) + pgx_hist_ordered_growth"}
        del host, host_np
        # (b) ItemTable seam (what the CPU reference arm is timed on, abacus.rs:539-586, 859-1032)
        hc_local = a.hist()[0] if world == 1 else None
        steps_total = int((np.arange(G + 1, dtype=np.float64) * hc_local).sum()) if world == 1 else None
        need = steps_total * 4 if steps_total else (int(N) * G // 2) * 4  # the pinned table: u32 per step
        avail = host_mem_available()
        if avail and avail < 1.5 * need * world:  # every rank of this node pins its own table
            e2e = {"skipped": f"host memory: {avail >> 30} GiB available, ItemTable needs ~{need >> 30} GiB pinned per rank"}
        else:
            items, prefsum, path_group = gpu_item_table(c, bitmap, N, G)
            if steps_total is not None:
                assert items.size == steps_total, (items.size, steps_total)
            b.clear()
            b.build(items, prefsum, path_group)  # warm-up: staging buffers, copy stream
            hc3, _, cv3 = b.hist_ordered_growth(cov, None, weighted=False)
            assert np.array_equal(cv3[0], want_curve) and (world > 1 or np.array_equal(hc3, hc)), "ItemTable build differs"
            barrier(c)
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                b.clear()
                b.build(items, prefsum, path_group)
                hc3, _, cv3 = b.hist_ordered_growth(cov, None, weighted=False)
            torch.cuda.synchronize()
            dt = max_over_ranks(c, time.perf_counter() - t0)
            # the build alone (H2D overlapped with the scatter kernel)
            t0 = time.perf_counter()
            b.clear()
            b.build(items, prefsum, path_group)
            dt_build = max_over_ranks(c, time.perf_counter() - t0)
            build_info = b.last_launch_info()
            hc3, _, cv3 = b.hist_ordered_growth(cov, None, weighted=False)  # (collective with the fused exchange: every rank)
            e2e = {"value": cells_per_step * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": int(items.nbytes + prefsum.nbytes + path_group.nbytes),
                   "d2h_bytes_per_step": int((2 * (G + 1) + G) * 8), "steps": n_e2e, "seam": "ItemTable (same as the reference arm)",
                   "numa_node_of_rank0": c.numa_node,
                   "item_table_steps": int(items.size), "ids": "u32, page-locked host memory",
                   "build_steps_per_s": float(items.size) / dt_build, "build": build_info,
                   "api": "pgx_abacus_clear + pgx_abacus_build_u32 (chunked H2D overlapped with k_build) + pgx_hist_ordered_growth"}
            del items
        b.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_items = args.cpu_sample_items or (200_000 if G >= 1024 else 800_000)
        tables, sbitmap, _ = make_cpu_tables(sample_items, G, seed=0x5EED0001)
        r = cpu_reference_pass(tables, G)
        # the GPU path must reproduce the CPU result on this very sample, from the same ItemTable
        with pb.DeviceAbacus(sample_items, G, device=c.local) as s:
            s.build(tables[0], tables[1], np.arange(G, dtype=np.int64))
            hc3, _, cv3 = s.hist_ordered_growth([1], None, weighted=False)
        parity = bool(np.array_equal(hc3, r["hist"]) and np.array_equal(cv3[0].astype(np.float64), r["curve"]))
        cpu = {"value": float(sample_items) * G / r["total_s"], "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{sample_items} items x {G} groups, one pass: coverage+hist {r['hist_s']:.2f}s + CSR build "
                         f"{r['csr_s']:.2f}s + calc_growth {r['growth_s']:.2f}s (oracle/ C restatement; the reference "
                         f"runs this path on one thread for a single threshold pair)",
               "growth_only_value": float(sample_items) * G / r["growth_s"], "gpu_matches_cpu_on_sample": parity}
        assert parity, "GPU result differs from the CPU oracle on the baseline sample"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, world),
        "gbps_per_gpu": alg_bytes / (ms_per_step * 1e-3) / 1e9,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e if e2e and "value" in e2e else e2e_bitmap,
        "e2e_bitmap_seam": e2e_bitmap, "e2e_note": None if e2e and "value" in e2e else (e2e or {}).get("skipped"),
        "gpu_launches": int(launches), "clocks": clocks,
    }
    a.close()
    return finish_line(c, args, line)


def run_sharded(args):
    """c3 (permuted growth) and c4 (similarity): work-item sharding behind the C ABI, strong scaling."""
    import panacus_b200 as pb
    from panacus_b200 import synth
    c = setup_gpu(args)
    torch, world, rank, dev = c.torch, c.world, c.rank, c.dev
    wl, N, G = shape_of(args)
    kind = wl["kind"]
    Wp, W = pb.row_words(G), (G + 63) // 64
    seed = synth.SEED_BASE + (3 if kind == "permuted" else 4)
    bitmap, weight = synth.torch_table(N, G, seed=seed, device=dev)  # same seed on every rank: replicated bitmap
    del weight
    torch.cuda.synchronize()
    a = pb.DeviceAbacus(N, G, device=c.local)
    a.adopt_device(bitmap.data_ptr(), None, keepalive=bitmap)
    a.set_stream(c.stream.cuda_stream)
    comm = pb.Comm.from_torch_distributed(c.local) if world > 1 else None
    bitmap_bytes = (N + 1) * Wp * 8

    if kind == "permuted":
        P = args.orders or wl["n_orders"]
        cov = [cv for cv, _ in C3_PAIRS]
        thr = np.stack([pb.quorum_thresholds(G, q) for _, q in C3_PAIRS])
        orders = synth.random_orders(P, G, seed=seed)
        result = pb.pinned_empty((P, len(cov), G), np.uint64)
        if comm is not None:
            def call(h):
                return h.permuted_growth_sharded(comm, orders, cov, thr, weighted=False, out=result)
        else:
            def call(h):
                return h.permuted_growth(orders, cov, thr, weighted=False, out=result)
        cells_per_step = float(N) * G * P
        n_local = len(range(rank, P, world))
        alg_bytes = float(n_local) * N * W * 8  # one bitmap pass per order
        kernel_name = "k_gm_quorum (+ k_gm_union rider)"
        d2h = int(result.nbytes)
        api = "pgx_permuted_growth_sharded" if comm is not None else "pgx_permuted_growth"
    else:
        result = pb.pinned_empty((G, G), np.uint64)
        if comm is not None:
            def call(h):
                return h.similarity_sharded(comm, weighted=False, out_inter=result)
        else:
            def call(h):
                return h.similarity(weighted=False, out_inter=result)
        cells_per_step = float(N) * G
        alg_bytes = float(N) * W * 8  # the bitmap read once (the kernel is POPC / LOP3 issue bound, not HBM bound)
        kernel_name = "k_gm_similarity<csa>"
        d2h = int(result.nbytes + G * 8)
        api = "pgx_similarity_sharded" if comm is not None else "pgx_similarity"

    for _ in range(max(args.warmup, 3)):  # the first call also derives the group-major copy (not part of a step)
        call(a)
    torch.cuda.synchronize()
    a.set_timing(True)
    a.kernel_time_ms()
    sampler = ClockSampler(c.local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = a.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(c)
    e0.record()
    for _ in range(args.steps):
        out = call(a)
    e1.record()
    torch.cuda.synchronize()
    total_ms = max_over_ranks(c, e0.elapsed_time(e1))
    kernel_ms, n_sections = a.kernel_time_ms()
    launch_info = a.last_launch_info()  # (with timing on, the sharded similarity reports its phase times here)
    a.set_timing(False)
    launches = a.launch_count - launches0
    n_extra = max(1, min(200, int(250.0 / max(total_ms / args.steps, 1e-3))))
    for _ in range(n_extra):
        call(a)
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps
    value = cells_per_step / (ms_per_step * 1e-3)

    # sanity + a checksum that must not depend on the number of GPUs
    if kind == "permuted":
        curves = out
        assert np.all(curves[:, 0, -1] == curves[0, 0, -1]) and int(curves[0, 0, -1]) > 0  # union of all groups: order independent
        checksum = int(curves.astype(np.uint64).sum() % (1 << 61))
    else:
        inter, ln = out
        assert np.array_equal(np.diag(inter), ln) and np.array_equal(inter, inter.T)
        checksum = int(inter.sum() % (1 << 61))

    peak, peak_src = measured_peaks()
    kms = kernel_ms / args.steps
    achieved = alg_bytes / (kms * 1e-3) / 1e9 if kms > 0 else None
    tensor = None
    if kind != "permuted" and "k_sim_mma" in launch_info:
        # the similarity contraction runs on the tensor cores (tcgen05.mma kind::i8): 128 x 256 tiles that touch the upper
        # triangle, 2 ops per (row, column, item); nominal i8 rate = 2 x bf16, so the reference figure is 2 x the measured
        # bf16 throughput of MEASURED_PEAKS.json (sustained: the kernel runs for milliseconds)
        kernel_name = "k_sim_mma"
        tiles = sum(1 for by in range((G + 127) // 128) for bx in range((G + 255) // 256) if bx * 256 + 256 > by * 128)
        ops = 2.0 * tiles * 128 * 256 * (N + 1) / world
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            tpeak = 2.0 * float(mp.get("bf16_tflops_sustained") or mp["bf16_tflops"])
            tsrc = "2 x measured bf16 (MEASURED_PEAKS.json, sustained)"
        except Exception:
            tpeak, tsrc = 2.0 * 2250.0, "2 x nominal dense bf16 (fallback)"
        tach = ops / (kms * 1e-3) / 1e12 if kms > 0 else None
        tensor = {"bound": "tensor", "achieved": tach, "peak": tpeak, "unit": "TOP/s (i8)", "frac": tach / tpeak if tach else None,
                  "peak_source": tsrc, "ops_per_launch": ops,
                  "note": "the tensor pipe is fed by a bit -> u8 expansion on the integer pipes, which is the larger half of the kernel"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                "traffic": None, "peak_source": peak_src, "kernel": kernel_name, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms_mean": kms, "kernel_sections_per_step": n_sections / args.steps, "launch": launch_info,
                "note": ("integer-ALU bound (bit-sliced rank counters), DRAM sees about one pass for all orders of a launch (L2 reuse)"
                         if kind == "permuted" else "not HBM bound (k_sim_mma: tensor cores + bit expansion; k_gm_similarity: POPC / LOP3 issue); "
                                                    "the HBM figure is reported because the contract asks for it")}
    if tensor:
        roofline["tensor"] = tensor

    # ---- e2e: packed bitmap in pinned host memory on rank 0 -> H2D (+ NVLink broadcast) -> transpose -> call -> host ----
    e2e = None
    if not args.no_e2e:
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        host_np = None
        if rank == 0:
            host = torch.empty((N + 1, Wp), dtype=torch.int64, pin_memory=True)
            host.copy_(bitmap)
            torch.cuda.synchronize()
            host_np = host.numpy().view(np.uint64)
        b = pb.DeviceAbacus(N, G, device=c.local)

        def e2e_step():
            if rank == 0:
                b.upload(host_np)
            if comm is not None:
                b.broadcast(comm, root=0)
            return call(b)
        e2e_step()
        barrier(c)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            out2 = e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        chk2 = int((out2 if kind == "permuted" else out2[0]).astype(np.uint64).sum() % (1 << 61))
        assert chk2 == checksum, "e2e result differs"
        e2e = {"value": cells_per_step * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": int(bitmap_bytes), "d2h_bytes_per_step": d2h,
               "steps": n_e2e, "api": "pgx_abacus_upload (rank 0, pinned host bitmap)" + (" + pgx_abacus_broadcast (NVLink)" if comm else "") +
                                      " + group-major transpose + " + api}
        b.close()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": None, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks, "checksum": checksum, "api": api,
    }
    a.close()
    if comm is not None:
        comm.close()
    return finish_line(c, args, line)


# ---------------------------------------------------------------------------------------------------
# c5: the CLI on a chr22-shaped GFA
# ---------------------------------------------------------------------------------------------------

def run_c5(args, reference=False):
    from tools import chr22_shape
    return chr22_shape.bench(args, reference=reference, metric=METRIC, unit=UNIT, config=workload_config(args, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="target")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak", help="target / c2 with N > 1")
    ap.add_argument("--items", type=int, default=0, help="override the item count")
    ap.add_argument("--groups", type=int, default=0)
    ap.add_argument("--orders", type=int, default=0, help="c3: number of group orders")
    ap.add_argument("--cpu-sample-items", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--exchange", choices=["fused", "nccl"], default="fused",
                    help="N > 1: how the per-shard results are summed (fused = inside k_scan over NVLink peer memory)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    kind = WORKLOADS[args.workload]["kind"]
    if kind == "ordered":
        return run_ordered(args)
    if kind == "cli":
        return run_c5(args)
    return run_sharded(args)


if __name__ == "__main__":
    sys.exit(main())
