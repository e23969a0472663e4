#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vertical or lane_private or random_tables or garbage or chrM or witness or small_fixtures" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
timeout 300 python tools/scan_timeline.py > gpurun_out/r2k_timeline.out 2> gpurun_out/r2k_timeline.txt; echo "timeline rc=$?"; grep -A1 "x 44 default after write" gpurun_out/r2k_timeline.txt | cut -c1-520
timeout 400 python tools/bench_scan_shapes.py > gpurun_out/r2k_scan_shapes.jsonl 2> gpurun_out/r2k_scan_shapes.err; echo "shapes rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2k_scan_shapes.jsonl"):
    d=json.loads(l)
    if d["G"] <= 128 and d["mode"].startswith("count"): print(d["N"],d["G"],d["mode"].ljust(16),"default",d["priv_us"],"atomics",d["atomics_us"],"novert",d["novert_us"],"frac",d["frac_of_hbm"],d["same_result"],d["launch"][:28])
PY
