#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_private" > gpurun_out/r2_run4_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_run4_pytest.log
M="gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_fma.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warp_latency_issue_stalled_mio_throttle.ratio,smsp__average_warp_latency_issue_stalled_lg_throttle.ratio,smsp__average_warp_latency_issue_stalled_membar.ratio,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio,smsp__average_warp_latency_issue_stalled_branch_resolving.ratio,smsp__average_warp_latency_issue_stalled_sleeping.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio,smsp__average_warp_latency_issue_stalled_selected.ratio"
for v in priv atomics; do
  if [ $v = atomics ]; then export PGX_SCAN_PRIV=2; else unset PGX_SCAN_PRIV; fi
  timeout 300 ncu --metrics $M --clock-control none -k regex:k_scan -c 3 --csv --log-file gpurun_out/r2_run4_ncu_44_$v.csv python tools/one_scan.py 10000000 44 count 1 > gpurun_out/r2_run4_ncu_44_$v.log 2>&1
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 2 -c 1 -o gpurun_out/r2_run4_full_44_$v -f python tools/one_scan.py 10000000 44 count 1 >> gpurun_out/r2_run4_ncu_44_$v.log 2>&1
  echo "ncu $v rc=$?"
done
unset PGX_SCAN_PRIV
ls -la gpurun_out | tail -8
