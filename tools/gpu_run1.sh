#!/bin/bash
# round-2 GPU session 1 (1 GPU): full parity suite + the new bench workloads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_run1_env.txt 2>&1
free -g >> gpurun_out/r2_run1_env.txt; nproc >> gpurun_out/r2_run1_env.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_run1_pytest.log
tail -5 gpurun_out/r2_run1_pytest.log
for wl in target c2 c3 c4; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/r2_run1_bench_$wl.json 2> gpurun_out/r2_run1_bench_$wl.err; echo "bench $wl rc=$?"
  tail -c 1500 gpurun_out/r2_run1_bench_$wl.json; tail -3 gpurun_out/r2_run1_bench_$wl.err
done
