#!/bin/bash
# full GPU parity suite + c3 (plane skipping) + vert timeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/r2_run22_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_run22_pytest.log
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_run22_c3.json 2> gpurun_out/r2_run22_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_run22_c3.json").read().strip().splitlines()[-1]); print("c3", d["ms_per_step"], d["roofline"].get("kernel_ms_mean"), d["roofline"].get("launch"), d.get("checksum"))
PY
timeout 300 python tools/scan_timeline.py > gpurun_out/r2_run22_timeline.out 2> gpurun_out/r2_run22_timeline.txt; echo "timeline rc=$?"; grep -A1 "x 44 default after write" gpurun_out/r2_run22_timeline.txt | cut -c1-420
timeout 300 python tools/bench_scan_shapes.py --quick > gpurun_out/r2_run22_scan_shapes.jsonl 2> gpurun_out/r2_run22_scan_shapes.err; echo "shapes rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2_run22_scan_shapes.jsonl"):
    d=json.loads(l)
    if d["G"] <= 64: print(d["N"],d["G"],d["mode"].ljust(16),"default",d["priv_us"],"atomics",d["atomics_us"],"novert",d["novert_us"],"frac",d["frac_of_hbm"],d["same_result"],d["launch"][:28])
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_run22_smoke.log 2>&1; tail -1 gpurun_out/r2_run22_smoke.log
