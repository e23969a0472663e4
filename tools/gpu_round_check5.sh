#!/bin/bash
# 2-GPU call: sharded parity (tools/check_multigpu.py) and the sharded config 3 / 4 timings with the new kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
START=$(date +%s)
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multigpu.py > gpurun_out/check_multigpu_2gpu.log 2>&1
echo "check rc=$? at $(( $(date +%s) - START )) s"; tail -2 gpurun_out/check_multigpu_2gpu.log
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/bench_sharded.py > gpurun_out/bench_sharded_2gpu.txt 2> gpurun_out/bench_sharded_2gpu.err
echo "sharded rc=$? at $(( $(date +%s) - START )) s"; cat gpurun_out/bench_sharded_2gpu.txt; tail -2 gpurun_out/bench_sharded_2gpu.err
