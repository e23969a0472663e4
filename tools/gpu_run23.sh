#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vertical or cutoff or similarity or large_permuted or lane_private" > gpurun_out/r2_run23_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_run23_pytest.log
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 > gpurun_out/r2_run23_c4.json 2> gpurun_out/r2_run23_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_run23_c4.json").read().strip().splitlines()[-1]); print("c4", d["ms_per_step"], d["roofline"].get("kernel_ms_mean"), d["roofline"].get("launch"), d.get("checksum"))
PY
timeout 300 python tools/scan_timeline.py > gpurun_out/r2_run23_timeline.out 2> gpurun_out/r2_run23_timeline.txt; echo "timeline rc=$?"; grep -A1 "x 44 default after write" gpurun_out/r2_run23_timeline.txt | cut -c1-520
