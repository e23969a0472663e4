#!/bin/bash
# direct epilogue + k_scan_vert: parity of the scan kernels, shape sweep, target / c2 bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vertical or ticket or lane_private or random_tables or small_fixtures or chrM or extreme or garbage or many_thresholds or dense or config2 or witness" > gpurun_out/r2_run18_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_run18_pytest.log
timeout 400 python tools/bench_scan_shapes.py > gpurun_out/r2_run18_scan_shapes.jsonl 2> gpurun_out/r2_run18_scan_shapes.err; echo "shapes rc=$?"; tail -3 gpurun_out/r2_run18_scan_shapes.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_run18_scan_shapes.jsonl"):
    d=json.loads(l); print(d["N"],d["G"],d["mode"].ljust(16),"default",d["priv_us"],"atomics",d["atomics_us"],"novert",d["novert_us"],"frac",d["frac_of_hbm"],d["same_result"],d["launch"][:28])
PY
timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_run18_target.json 2> gpurun_out/r2_run18_target.err; echo "target rc=$?"
timeout 200 python bench.py --workload c2 --steps 50 --warmup 5 > gpurun_out/r2_run18_c2.json 2> gpurun_out/r2_run18_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ("target","c2"):
    d=json.loads(open(f"gpurun_out/r2_run18_{f}.json").read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_mean"], d["roofline"]["launch"])
PY
