#!/bin/bash
# same 8-GPU box: fused pass at N = 1 and N = 8 (weak), dynamic vs static tiles -- the weak-scaling efficiency without box-to-box spread
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="--workload target --steps 50 --warmup 5 --no-e2e --no-cpu"
timeout 200 python bench.py $B > gpurun_out/r2f_weak_1gpu_dynamic.json 2> gpurun_out/r2f_weak_1gpu.err; echo "N=1 rc=$?"
for mode in dynamic static; do
  if [ $mode = static ]; then export PGX_SCAN_STATIC=1; else unset PGX_SCAN_STATIC; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 $B > gpurun_out/r2f_weak_8gpu_$mode.json 2> gpurun_out/r2f_weak_8gpu_$mode.err; echo "N=8 $mode rc=$?"
done
unset PGX_SCAN_STATIC
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 $B > gpurun_out/r2f_weak_4gpu_dynamic.json 2> gpurun_out/r2f_weak_4gpu.err; echo "N=4 rc=$?"
python - <<'PY'
import json
for f in ("1gpu_dynamic","4gpu_dynamic","8gpu_dynamic","8gpu_static"):
    try:
        d=json.loads(open(f"gpurun_out/r2f_weak_{f}.json").read().strip().splitlines()[-1]); print(f, "ms/step", round(d["ms_per_step"],5), "value", d["value"], "kernel_ms", d["roofline"].get("kernel_ms_mean"))
    except Exception as e: print(f, "failed", e)
PY
