#!/usr/bin/env python3
"""Summarise an ncu report (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof_scan.ncu-rep profiles/r1_scan target
writes <prefix>_ncu_summary.txt (key metrics per captured launch + hottest SASS lines) and updates
profiles/scan_ncu_summary.json[workload] = {dram_bytes_per_launch, ...} (read by bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys

rep, prefix, workload = sys.argv[1], sys.argv[2], sys.argv[3]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines, launches = [], []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    lines.append(f"--- {name}")
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"  {k} = {r[i]} {units[i]}")
            d[k] = (r[i], units[i])
    launches.append(d)

def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    h = srows[1]
    isrc, iex, ist = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    body = []
    for r in srows[2:]:
        if len(r) < 10 or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    tot_ex = sum(int(r[iex]) for r in body)
    tot_st = sum(int(r[ist]) for r in body)
    lines.append(f"\nSASS (first captured launch): {len(body)} instructions, {tot_ex} warp-instructions executed, {tot_st} stall samples")
    lines.append("top stall-sample instructions:")
    for r in sorted(body, key=lambda r: -int(r[ist]))[:14]:
        lines.append(f"  {int(r[ist]):6d} samples  {int(r[iex]):9d} exec  {r[isrc].strip()[:80]}")
open(prefix + "_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if launches:
    rd = [to_bytes(*d["dram__bytes_read.sum"]) for d in launches]
    wr = [to_bytes(*d["dram__bytes_write.sum"]) for d in launches]
    path = os.path.join(os.path.dirname(prefix) or ".", "scan_ncu_summary.json")
    allw = json.load(open(path)) if os.path.exists(path) else {}
    allw[workload] = {"dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd), "dram_read": sum(rd) / len(rd),
                      "dram_write": sum(wr) / len(wr), "launches_captured": len(rd),
                      "duration_us_under_ncu": [d["gpu__time_duration.sum"][0] for d in launches], "source": os.path.basename(rep)}
    json.dump(allw, open(path, "w"), indent=1)
