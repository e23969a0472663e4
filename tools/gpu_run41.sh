#!/bin/bash
# 8 GPUs, ItemTable-seam e2e of the fused pass with / without binding every rank to its GPU's NUMA node
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2n_topo.txt 2>&1; lscpu | grep -i -E "numa|socket|model name" >> gpurun_out/r2n_topo.txt; nproc >> gpurun_out/r2n_topo.txt
for mode in bind nobind; do
  if [ $mode = nobind ]; then export PGX_BENCH_NUMA=0; else unset PGX_BENCH_NUMA; fi
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload target --steps 20 --warmup 5 --e2e-steps 2 --no-cpu > gpurun_out/r2n_e2e_8gpu_$mode.json 2> gpurun_out/r2n_e2e_8gpu_$mode.err; echo "$mode rc=$?"
done
python - <<'PY'
import json
for m in ("bind","nobind"):
    try:
        d=json.loads(open(f"gpurun_out/r2n_e2e_8gpu_{m}.json").read().strip().splitlines()[-1]); e=d.get("e2e") or {}
        print(m, "ms/step", round(d["ms_per_step"],4), "e2e", e.get("value"), "build_steps_per_s", e.get("build_steps_per_s"), "numa", e.get("numa_node_of_rank0"), e.get("skipped"))
    except Exception as ex: print(m, "failed", ex)
PY
tail -12 gpurun_out/r2n_topo.txt
