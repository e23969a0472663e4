#!/bin/bash
# ncu evidence for the end-of-round kernels: launch lists of the bench commands + one --set full capture each
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="--no-e2e --no-cpu"
PGX_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_target.csv python bench.py --steps 5 --warmup 3 $B > gpurun_out/r2b_launches_target.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 4 -c 2 -o gpurun_out/r2b_prof_scan -f python bench.py --steps 5 --warmup 3 $B > gpurun_out/r2b_prof_scan.log 2>&1; echo "full scan rc=$?"
PGX_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_c2.csv python bench.py --workload c2 --steps 5 --warmup 3 $B > gpurun_out/r2b_launches_c2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 4 -c 2 -o gpurun_out/r2b_prof_scan_c2 -f python bench.py --workload c2 --steps 5 --warmup 3 $B > gpurun_out/r2b_prof_scan_c2.log 2>&1; echo "full c2 rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_transpose_reg -s 1 -c 1 -o gpurun_out/r2b_prof_transpose -f python tools/one_transpose.py 10000000 1024 > gpurun_out/r2b_prof_transpose.log 2>&1; echo "full transpose rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_scan_vert -s 2 -c 1 -o gpurun_out/r2b_prof_vert -f python tools/one_scan.py 10000000 44 count 1 > gpurun_out/r2b_prof_vert.log 2>&1; echo "full vert rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_gm_quorum -c 2 -o gpurun_out/r2b_prof_quorum -f python tools/ncu_targets.py > gpurun_out/r2b_prof_quorum.log 2>&1; echo "full quorum rc=$?"
for wl in target c2 c3 c4; do timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/r2b_bench_$wl.json 2> gpurun_out/r2b_bench_$wl.err; echo "bench $wl rc=$?"; done
python - <<'PY'
import json
for f in ("target","c2","c3","c4"):
    d=json.loads(open(f"gpurun_out/r2b_bench_{f}.json").read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("kernel_ms_mean"), (d.get("e2e") or {}).get("value"))
PY
ls -la gpurun_out | grep r2b_prof
