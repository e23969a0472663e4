#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli.py -m gpu -x -q > gpurun_out/r2s_pytest_cli.log 2>&1; echo "pytest cli rc=$?"; tail -3 gpurun_out/r2s_pytest_cli.log
timeout 600 python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/r2s_bench_c5.json 2> gpurun_out/r2s_bench_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2s_bench_c5.json").read().strip().splitlines()[-1])
print("c5 ms/step", d.get("ms_per_step"), d.get("phases_ms"), "cpu", d.get("cpu_baseline"))
PY
