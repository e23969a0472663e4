#!/bin/bash
# end-of-round 8-GPU session (trimmed: no e2e legs): multi-GPU parity + sharded c3 / c4 + weak / strong scaling of the fused pass
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r2b_pytest_multigpu_${N}gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest_multigpu_${N}gpu.log
tail -4 gpurun_out/r2b_pytest_multigpu_${N}gpu.log
run() { # name, args...
  name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r2b_bench_${name}_${N}gpu.json 2> gpurun_out/r2b_bench_${name}_${N}gpu.err
  echo "bench $name N=$N rc=$?"; grep -v "^W10\|^\*\*\*\|OMP_NUM" gpurun_out/r2b_bench_${name}_${N}gpu.err | tail -3
}
run c3 --workload c3 --steps 20 --warmup 5 --no-e2e
run c4 --workload c4 --steps 20 --warmup 5 --no-e2e
run target_weak --workload target --steps 50 --warmup 5 --no-e2e --no-cpu
run target_strong --workload target --scaling strong --steps 50 --warmup 5 --no-e2e --no-cpu
python - <<PY
import json
for f in ("c3","c4","target_weak","target_strong"):
    try:
        d=json.loads(open(f"gpurun_out/r2b_bench_{f}_${N}gpu.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "value", d["value"], "kernel_ms", d["roofline"].get("kernel_ms_mean"), "checksum", d.get("checksum"), d["roofline"].get("launch","")[:110])
    except Exception as e: print(f, "failed", e)
PY
