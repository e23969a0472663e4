#!/bin/bash
# ncu evidence for the end-of-round kernels: launch lists of the bench commands + one --set full capture each, summarised
# ON THE BOX (the reports themselves are too large to bring back together: 64 MiB limit)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out /tmp/ncu
B="--no-e2e --no-cpu"
PGX_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_target.csv python bench.py --steps 5 --warmup 3 $B > gpurun_out/r2b_launches_target.log 2>&1; echo "launch list rc=$?"
PGX_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_c2.csv python bench.py --workload c2 --steps 5 --warmup 3 $B > gpurun_out/r2b_launches_c2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 4 -c 2 -o /tmp/ncu/scan -f python bench.py --steps 5 --warmup 3 $B > gpurun_out/r2b_prof_scan.log 2>&1; echo "full scan rc=$?"
python tools/ncu_summary.py /tmp/ncu/scan.ncu-rep gpurun_out/r2b_scan target > /dev/null 2>&1; ls -la /tmp/ncu
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 4 -c 2 -o /tmp/ncu/scan_c2 -f python bench.py --workload c2 --steps 5 --warmup 3 $B > gpurun_out/r2b_prof_scan_c2.log 2>&1; echo "full c2 rc=$?"
python tools/ncu_summary.py /tmp/ncu/scan_c2.ncu-rep gpurun_out/r2b_scan_c2 c2 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_transpose_reg -s 1 -c 1 -o /tmp/ncu/transpose -f python tools/one_transpose.py 10000000 1024 > gpurun_out/r2b_prof_transpose.log 2>&1; echo "full transpose rc=$?"
python tools/ncu_multi_summary.py /tmp/ncu/transpose.ncu-rep gpurun_out/r2b_transpose_ncu_summary.txt > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_scan_vert -s 2 -c 1 -o /tmp/ncu/vert -f python tools/one_scan.py 10000000 44 count 1 > gpurun_out/r2b_prof_vert.log 2>&1; echo "full vert rc=$?"
python tools/ncu_multi_summary.py /tmp/ncu/vert.ncu-rep gpurun_out/r2b_vert_ncu_summary.txt > /dev/null 2>&1
python tools/ncu_summary.py /tmp/ncu/vert.ncu-rep gpurun_out/r2b_vert_sass vert_10Mx44 > /dev/null 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_gm_quorum -c 2 -o /tmp/ncu/quorum -f python tools/ncu_targets.py > gpurun_out/r2b_prof_quorum.log 2>&1; echo "full quorum rc=$?"
python tools/ncu_multi_summary.py /tmp/ncu/quorum.ncu-rep gpurun_out/r2b_quorum_ncu_summary.txt > /dev/null 2>&1
cp /tmp/ncu/scan.ncu-rep gpurun_out/r2b_prof_scan.ncu-rep   # the headline kernel's report comes back whole (19 MB)
du -sh gpurun_out; ls -la gpurun_out
