#!/bin/bash
# last verification of the round (final library): full GPU suite, smoke, default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; tail -1 gpurun_out/r2final_smoke.log
timeout 600 python bench.py > gpurun_out/r2final_bench_default.json 2> gpurun_out/r2final_bench_default.err; echo "bench default rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2final_bench_default.json").read().strip().splitlines()[-1])
print("default value", d.get("value"), "ms/step", d.get("ms_per_step"), "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
PY
