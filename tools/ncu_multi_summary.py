#!/usr/bin/env python3
"""Summarise every launch of an ncu --set full report (run here, no GPU needed):
    python tools/ncu_multi_summary.py gpurun_out/r1_secondary.ncu-rep profiles/r1_secondary_ncu_summary.txt
Per launch: duration, DRAM bytes, issue / pipe utilisation, occupancy, and the top warp-stall reasons."""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
lines = []
for r in rows[2:]:
    lines.append("--- " + r[hdr.index("Kernel Name")])
    for k in WANT:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"  {k} = {r[i]} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
           (h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")):
            try:
                stalls.append((float(r[i].replace(",", "")), h))
            except ValueError:
                pass
    for v, h in sorted(stalls, reverse=True)[:6]:
        lines.append(f"  stall {h.split('stalled_')[1].split('.')[0].replace('_per_issue_active', '')} = {v:.2f} warps per issue")
    pipes = []
    for i, h in enumerate(hdr):
        if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum"):
            try:
                pipes.append((float(r[i].replace(",", "")), h))
            except ValueError:
                pass
    if pipes:
        lines.append("  warp instructions by pipe: " + ", ".join(f"{h[len('sm__inst_executed_pipe_'):-4]} {v:.3g}" for v, h in sorted(pipes, reverse=True)[:8]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
