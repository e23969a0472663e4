#!/bin/bash
# in-register transpose + vert read-out: parity, transpose bench, scan phase timelines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transpose or vertical or similarity or permuted or smoke" > gpurun_out/r2_run19_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_run19_pytest.log
timeout 300 python tools/bench_transpose.py > gpurun_out/r2_run19_transpose.jsonl 2> gpurun_out/r2_run19_transpose.err; echo "transpose rc=$?"; cat gpurun_out/r2_run19_transpose.jsonl; tail -3 gpurun_out/r2_run19_transpose.err
PGX_TRANSPOSE=shfl timeout 300 python tools/bench_transpose.py > gpurun_out/r2_run19_transpose_shfl.jsonl 2> gpurun_out/r2_run19_transpose_shfl.err; echo "transpose shfl rc=$?"; cat gpurun_out/r2_run19_transpose_shfl.jsonl
timeout 300 python tools/scan_timeline.py > gpurun_out/r2_run19_timeline.out 2> gpurun_out/r2_run19_timeline.txt; echo "timeline rc=$?"; cat gpurun_out/r2_run19_timeline.txt
