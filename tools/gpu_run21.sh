#!/bin/bash
# vert fold v3 + coverage-sorted quorum (3 CTAs/SM): parity, c3 A/B, vert timeline, quick shapes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vertical or permuted or config3 or large_permuted or random_tables" > gpurun_out/r2_run21_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_run21_pytest.log
for mode in covsort natural; do
  if [ $mode = natural ]; then export PGX_GM_COVSORT=0; else unset PGX_GM_COVSORT; fi
  timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_run21_c3_$mode.json 2> gpurun_out/r2_run21_c3_$mode.err; echo "c3 $mode rc=$?"
done
unset PGX_GM_COVSORT
python - <<'PY'
import json
for m in ("covsort","natural"):
    d=json.loads(open(f"gpurun_out/r2_run21_c3_{m}.json").read().strip().splitlines()[-1]); print(m, d["ms_per_step"], d["roofline"].get("kernel_ms_mean"), d["roofline"].get("launch"), d.get("checksum"))
PY
timeout 300 python tools/scan_timeline.py > gpurun_out/r2_run21_timeline.out 2> gpurun_out/r2_run21_timeline.txt; echo "timeline rc=$?"; grep -A1 "x 44 default after write" gpurun_out/r2_run21_timeline.txt | cut -c1-420
timeout 300 python tools/bench_scan_shapes.py --quick > gpurun_out/r2_run21_scan_shapes.jsonl 2> gpurun_out/r2_run21_scan_shapes.err; echo "shapes rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2_run21_scan_shapes.jsonl"):
    d=json.loads(l); print(d["N"],d["G"],d["mode"].ljust(16),"default",d["priv_us"],"atomics",d["atomics_us"],"novert",d["novert_us"],"frac",d["frac_of_hbm"],d["same_result"],d["launch"][:28])
PY
