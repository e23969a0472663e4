#!/bin/bash
# end-of-round verification: full GPU suite, smoke, default bench line + reference arm, c4 / c5 lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -1 gpurun_out/r2c_smoke.log
timeout 600 python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; echo "bench reference rc=$?"
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
for f in ("default","reference","c4","c5"):
    try:
        d=json.loads(open(f"gpurun_out/r2c_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", d.get("value"), "ms/step", d.get("ms_per_step"), "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("tensor"))
    except Exception as e: print(f, "failed", e)
PY
