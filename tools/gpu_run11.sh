#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c4 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_run11_c4_${N}gpu.json 2> gpurun_out/r2_run11_c4.err; echo "c4 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_run11_c4_2gpu.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['launch'])"
timeout 600 python bench.py --workload c2 --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2_run11_c2.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_run11_c2.json').read().strip().splitlines()[-1]); print('c2', d['ms_per_step'], d['roofline']['kernel_ms_mean'], d['roofline']['kernel_ms_min'])"
