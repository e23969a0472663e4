#!/bin/bash
# c2 (1M x 256) with smaller tiles / deeper rings (PGX_SCAN_TILE / PGX_SCAN_STAGES): does a shorter first tile + finer dynamic scheduling pay?
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
: > gpurun_out/r2m_c2_tiles.txt
for cfg in "0 0" "512 3" "512 4" "384 4" "256 4" "256 6" "128 8"; do
  set -- $cfg
  if [ "$1" != 0 ]; then export PGX_SCAN_TILE=$1 PGX_SCAN_STAGES=$2; else unset PGX_SCAN_TILE PGX_SCAN_STAGES; fi
  timeout 200 python bench.py --workload c2 --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2m_c2.json 2> gpurun_out/r2m_c2.err
  python - >> gpurun_out/r2m_c2_tiles.txt <<PY
import json
d=json.loads(open("gpurun_out/r2m_c2.json").read().strip().splitlines()[-1]); print("tile=$1 stages=$2", round(d["ms_per_step"]*1e3,2), "us", d["roofline"]["launch"])
PY
done
cat gpurun_out/r2m_c2_tiles.txt
