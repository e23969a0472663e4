#!/bin/bash
# round-2 GPU session 2 (N GPUs): multi-GPU parity tests + sharded bench workloads
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_run25_env.txt 2>&1; free -g >> gpurun_out/r2_run25_env.txt; nproc >> gpurun_out/r2_run25_env.txt
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r2_run25_pytest_multigpu_${N}gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_run25_pytest_multigpu_${N}gpu.log
tail -15 gpurun_out/r2_run25_pytest_multigpu_${N}gpu.log
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r2_run25_bench_${name}_${N}gpu.json 2> gpurun_out/r2_run25_bench_${name}_${N}gpu.err
  echo "bench $name N=$N rc=$?"; tail -c 900 gpurun_out/r2_run25_bench_${name}_${N}gpu.json; echo; grep -v "^W10\|^\*\*\*\|OMP_NUM" gpurun_out/r2_run25_bench_${name}_${N}gpu.err | tail -5
}
run c3 --workload c3 --steps 20 --warmup 5
run c4 --workload c4 --steps 20 --warmup 5
run target_weak --workload target --steps 50 --warmup 5
run target_strong --workload target --scaling strong --steps 50 --warmup 5 --no-e2e
