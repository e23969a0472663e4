#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "permuted or similarity or large or full_size or config3" > gpurun_out/r2_run8_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_run8_pytest.log
TR_VARIANTS=1,2,3,4,6 timeout 600 python tools/bench_transpose.py > gpurun_out/r2_run8_transpose.jsonl 2> gpurun_out/r2_run8_transpose.err; echo "transpose rc=$?"; cat gpurun_out/r2_run8_transpose.jsonl; tail -3 gpurun_out/r2_run8_transpose.err
