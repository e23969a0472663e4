#!/bin/bash
# 2-GPU call (gpurun --gpus 2 --timeout 100 -- 'bash tools/gpu_multi_check.sh'): similarity tests on one GPU, sharded
# parity (tools/check_multigpu.py) and the sharded config 3 / 4 timings (tools/bench_sharded.py)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
START=$(date +%s)
CUDA_VISIBLE_DEVICES=0 timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "test_similarity" > gpurun_out/pytest_similarity.log 2>&1 &
timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multigpu.py > gpurun_out/check_multigpu_2gpu.log 2>&1
echo "check rc=$? at $(( $(date +%s) - START )) s"; tail -2 gpurun_out/check_multigpu_2gpu.log
wait
tail -3 gpurun_out/pytest_similarity.log
timeout 30 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/bench_sharded.py > gpurun_out/bench_sharded_2gpu.txt 2> gpurun_out/bench_sharded_2gpu.err
echo "sharded rc=$? at $(( $(date +%s) - START )) s"; cat gpurun_out/bench_sharded_2gpu.txt; tail -2 gpurun_out/bench_sharded_2gpu.err
