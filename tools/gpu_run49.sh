#!/bin/bash
# c5 bench line with the final tool (median run with its own phases, every run listed)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/r2v_bench_c5.json 2> gpurun_out/r2v_bench_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_bench_c5.json").read().strip().splitlines()[-1])
print("median wall", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in d["phases_ms"].items()}); print(d["runs"]); print(d["cpu_baseline"])
PY
