#!/bin/bash
# init-cost probe + full GPU parity suite + c4 bench on one GPU (tensor-core similarity)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/probe_init.py > gpurun_out/r2_run17_probe_init.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2_run17_probe_init.txt
timeout 600 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/r2_run17_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_run17_pytest.log
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 > gpurun_out/r2_run17_c4_1gpu.json 2> gpurun_out/r2_run17_c4.err; echo "c4 rc=$?"; tail -c 600 gpurun_out/r2_run17_c4_1gpu.json
timeout 400 python tools/bench_scan_shapes.py > gpurun_out/r2_run17_scan_shapes.jsonl 2> gpurun_out/r2_run17_scan_shapes.err; echo "shapes rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2_run17_scan_shapes.jsonl"):
    d=json.loads(l); print(d["N"],d["G"],d["mode"].ljust(16),"default",d["priv_us"],"atomics",d["atomics_us"],"frac",d["frac_of_hbm"],d["same_result"],d["launch"][:18])
PY
timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_run17_target.json 2> gpurun_out/r2_run17_target.err; echo "target rc=$?"; tail -c 900 gpurun_out/r2_run17_target.json
