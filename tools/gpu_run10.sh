#!/bin/bash
# round-2 profile evidence for the bench command: ncu launch list of the timed region + one full capture of k_scan
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
PGX_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_target.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_launches_target.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 4 -c 2 -o gpurun_out/r2_prof_scan -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_prof_scan.log 2>&1; echo "full rc=$?"
PGX_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_launches_c2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 4 -c 2 -o gpurun_out/r2_prof_scan_c2 -f python bench.py --workload c2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_prof_scan_c2.log 2>&1
for wl in target c2; do timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/r2_run10_bench_$wl.json 2> gpurun_out/r2_run10_bench_$wl.err; echo "bench $wl rc=$?"; done
grep -c k_scan gpurun_out/r2_launches_target.csv; ls -la gpurun_out | grep r2_prof
