#!/bin/bash
# dynamic tile scheduler + vert fold: parity, shapes, target / c2 (dynamic vs static), timeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vertical or ticket or lane_private or random_tables or small_fixtures or chrM or extreme or garbage or many_thresholds or dense or config2 or witness or full_size" > gpurun_out/r2_run20_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_run20_pytest.log
timeout 400 python tools/bench_scan_shapes.py > gpurun_out/r2_run20_scan_shapes.jsonl 2> gpurun_out/r2_run20_scan_shapes.err; echo "shapes rc=$?"; tail -3 gpurun_out/r2_run20_scan_shapes.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_run20_scan_shapes.jsonl"):
    d=json.loads(l); print(d["N"],d["G"],d["mode"].ljust(16),"default",d["priv_us"],"atomics",d["atomics_us"],"novert",d["novert_us"],"frac",d["frac_of_hbm"],d["same_result"],d["launch"][:28])
PY
for mode in dynamic static; do
  if [ $mode = static ]; then export PGX_SCAN_STATIC=1; else unset PGX_SCAN_STATIC; fi
  timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_run20_target_$mode.json 2> gpurun_out/r2_run20_target_$mode.err; echo "target $mode rc=$?"
  timeout 200 python bench.py --workload c2 --steps 50 --warmup 5 > gpurun_out/r2_run20_c2_$mode.json 2> gpurun_out/r2_run20_c2_$mode.err; echo "c2 $mode rc=$?"
done
unset PGX_SCAN_STATIC
python - <<'PY'
import json
for m in ("dynamic","static"):
  for f in ("target","c2"):
    d=json.loads(open(f"gpurun_out/r2_run20_{f}_{m}.json").read().strip().splitlines()[-1]); print(m, f, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_mean"], d["roofline"]["launch"])
PY
timeout 300 python tools/scan_timeline.py > gpurun_out/r2_run20_timeline.out 2> gpurun_out/r2_run20_timeline.txt; echo "timeline rc=$?"; grep -A1 "default after write" gpurun_out/r2_run20_timeline.txt | cut -c1-420
