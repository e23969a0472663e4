#!/bin/bash
# compute-sanitizer on the round-2 kernels + parity of the last scan change
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_private or random_tables" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest.log
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_targets.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" gpurun_out/r2_sanitizer_$tool.txt | tail -3
done
