#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "permuted or config3 or large_permuted or cutoff" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2j_pytest.log
timeout 400 python tools/time_bp_quorum.py > gpurun_out/r2j_bp_quorum.json 2> gpurun_out/r2j_bp_quorum.err; echo "bp rc=$?"; cat gpurun_out/r2j_bp_quorum.json; tail -2 gpurun_out/r2j_bp_quorum.err
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2j_c3.json 2> gpurun_out/r2j_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_c3.json").read().strip().splitlines()[-1]); print("c3", d["ms_per_step"], d["roofline"].get("kernel_ms_mean"), d["roofline"].get("launch"), d.get("checksum"))
PY
