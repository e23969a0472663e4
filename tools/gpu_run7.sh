#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run7_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_run7_pytest.log
timeout 600 python tools/bench_transpose.py > gpurun_out/r2_run7_transpose.jsonl 2> gpurun_out/r2_run7_transpose.err; echo "transpose rc=$?"; cat gpurun_out/r2_run7_transpose.jsonl; tail -3 gpurun_out/r2_run7_transpose.err
timeout 600 python bench.py --workload c5 --steps 3 > gpurun_out/r2_run7_bench_c5.json 2> gpurun_out/r2_run7_bench_c5.err; echo "c5 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_run7_bench_c5.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['phases_ms'], d['cpu_baseline'])"; tail -3 gpurun_out/r2_run7_bench_c5.err
