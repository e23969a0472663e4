#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_run6_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_run6_pytest.log
timeout 600 python tools/bench_transpose.py > gpurun_out/r2_run6_transpose.jsonl 2> gpurun_out/r2_run6_transpose.err; echo "transpose rc=$?"; cat gpurun_out/r2_run6_transpose.jsonl; tail -3 gpurun_out/r2_run6_transpose.err
timeout 600 python tools/bench_scan_shapes.py --quick > gpurun_out/r2_run6_scan_shapes.jsonl 2> gpurun_out/r2_run6_scan_shapes.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_run6_scan_shapes.jsonl"):
    d=json.loads(l); print(d["N"],d["G"],d["mode"].ljust(16),"priv",d["priv_us"],"atomics",d["atomics_us"],"frac",d["frac_of_hbm"],d["same_result"],d["launch"][:24], d["launch"].split("grid=")[1][:60])
PY
timeout 600 python bench.py --workload c5 --steps 3 > gpurun_out/r2_run6_bench_c5.json 2> gpurun_out/r2_run6_bench_c5.err; echo "c5 rc=$?"; cut -c1-1800 gpurun_out/r2_run6_bench_c5.json; tail -3 gpurun_out/r2_run6_bench_c5.err
timeout 600 python bench.py --impl reference --workload c5 --steps 1 --warmup 0 > gpurun_out/r2_run6_bench_c5_ref.json 2> gpurun_out/r2_run6_bench_c5_ref.err; echo "c5 ref rc=$?"; cut -c1-400 gpurun_out/r2_run6_bench_c5_ref.json
