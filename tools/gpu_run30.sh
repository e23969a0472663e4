#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_private or random_tables or vertical or ticket or config2 or full_size or many_thresholds" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2d_pytest.log
for mode in hybrid off; do
  if [ $mode = off ]; then export PGX_SCAN_HYBRID=2; else unset PGX_SCAN_HYBRID; fi
  timeout 400 python tools/bench_scan_shapes.py > gpurun_out/r2d_scan_shapes_$mode.jsonl 2> gpurun_out/r2d_scan_shapes_$mode.err; echo "shapes $mode rc=$?"
  timeout 200 python bench.py --workload c2 --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2d_c2_$mode.json 2> gpurun_out/r2d_c2_$mode.err
done
unset PGX_SCAN_HYBRID
python - <<'PY'
import json
a={}
for l in open("gpurun_out/r2d_scan_shapes_off.jsonl"):
    d=json.loads(l); a[(d["N"],d["G"],d["mode"])]=d
for l in open("gpurun_out/r2d_scan_shapes_hybrid.jsonl"):
    d=json.loads(l); o=a.get((d["N"],d["G"],d["mode"]))
    if o and d["G"]>=100: print(d["N"],d["G"],d["mode"].ljust(16),"hybrid-build default",d["priv_us"],"| off default",o["priv_us"],"atomics",o["atomics_us"],d["same_result"],d["launch"][:30])
for m in ("hybrid","off"):
    d=json.loads(open(f"gpurun_out/r2d_c2_{m}.json").read().strip().splitlines()[-1]); print("c2",m,d["ms_per_step"],d["roofline"]["launch"][:40])
PY
