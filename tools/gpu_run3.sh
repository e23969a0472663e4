#!/bin/bash
# round-2 GPU session 3 (1 GPU): the lane-private-counter scan kernel -- parity, then timings over shapes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_private or random_tables or small_fixtures or chrM or extreme or garbage or many_thresholds or dense" > gpurun_out/r2_run3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_run3_pytest.log
tail -25 gpurun_out/r2_run3_pytest.log
timeout 900 python tools/bench_scan_shapes.py > gpurun_out/r2_run3_scan_shapes.jsonl 2> gpurun_out/r2_run3_scan_shapes.err; echo "shapes rc=$?"
cat gpurun_out/r2_run3_scan_shapes.jsonl | cut -c1-420; tail -5 gpurun_out/r2_run3_scan_shapes.err
