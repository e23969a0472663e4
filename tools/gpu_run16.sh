#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r2_run16_pytest_multigpu_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_run16_pytest_multigpu_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c4 --steps 20 --warmup 5 > gpurun_out/r2_run16_c4_${N}gpu.json 2> gpurun_out/r2_run16_c4.err; echo "c4 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_run16_c4_${N}gpu.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['launch'], d['checksum'], d['e2e'])"
