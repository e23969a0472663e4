#!/bin/bash
# last verification of the round: full GPU suite, smoke, default bench line, ncu summary of the two-word k_scan_vert
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out /tmp/ncu
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -1 gpurun_out/r2h_smoke.log
timeout 600 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; echo "bench default rc=$?"
timeout 300 python bench.py --workload c2 --steps 50 --warmup 5 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err; echo "c2 rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_scan_vert -s 2 -c 1 -o /tmp/ncu/vert2 -f python tools/one_scan.py 10000000 100 count 0 > gpurun_out/r2h_prof_vert2.log 2>&1; echo "full vert2 rc=$?"
python tools/ncu_multi_summary.py /tmp/ncu/vert2.ncu-rep gpurun_out/r2h_vert_two_words_ncu_summary.txt > /dev/null 2>&1
python - <<'PY'
import json
for f in ("default","c2"):
    d=json.loads(open(f"gpurun_out/r2h_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value", d.get("value"), "ms/step", d.get("ms_per_step"), "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"), d.get("clocks"))
PY
head -12 gpurun_out/r2h_vert_two_words_ncu_summary.txt
