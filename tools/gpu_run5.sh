#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_private or random_tables or small_fixtures or chrM or extreme or garbage or many_thresholds or dense or config2" > gpurun_out/r2_run5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_run5_pytest.log
timeout 900 python tools/bench_scan_shapes.py > gpurun_out/r2_run5_scan_shapes.jsonl 2> gpurun_out/r2_run5_scan_shapes.err; echo "shapes rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2_run5_scan_shapes.jsonl"):
    d=json.loads(l); print(d["N"],d["G"],d["mode"].ljust(16),"priv",d["priv_us"],"atomics",d["atomics_us"],"frac",d["frac_of_hbm"],d["same_result"],d["launch"][:40], d["launch"].split("smem=")[1][:40])
PY
tail -3 gpurun_out/r2_run5_scan_shapes.err
