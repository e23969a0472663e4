#!/bin/bash
# One gpurun call (gpurun --timeout 780 -- 'BUDGET_S=640 bash tools/gpu_round_check.sh'): GPU parity suite, old-vs-new
# kernel comparison, the bench line, smoke, secondary timings, ncu --set full of the secondary kernels, the c2 bench line.  Everything lands in gpurun_out/ (merged back by gpurun).  The steps share one deadline
# (BUDGET_S seconds from the start, default 600) so that the call ends on its own, most important step first.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
START=$(date +%s)
BUDGET_S=${BUDGET_S:-600}
left() { echo $(( START + BUDGET_S - $(date +%s) )); }
step() {  # step <name> <max seconds> <command...>
    local name=$1 max=$2; shift 2
    local l; l=$(left)
    if [ "$l" -lt 20 ]; then echo "== $name: skipped (deadline)"; return; fi
    [ "$l" -lt "$max" ] && max=$l
    local t0; t0=$(date +%s)
    timeout "$max" "$@"
    echo "== $name rc=$? in $(( $(date +%s) - t0 )) s (limit $max)"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
step pytest 360 bash -c 'python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log'
tail -4 gpurun_out/pytest_gpu.log
step variants 180 bash -c 'python tools/bench_variants.py > gpurun_out/bench_variants.txt 2> gpurun_out/bench_variants.err'
cat gpurun_out/bench_variants.txt; tail -3 gpurun_out/bench_variants.err
step bench 180 bash -c 'python bench.py --steps 50 --warmup 5 > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err'
tail -c 400 gpurun_out/bench_target.json
step smoke 90 bash -c 'python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1'
tail -1 gpurun_out/smoke.log
step aux 200 bash -c 'python tools/bench_aux.py > gpurun_out/bench_aux.txt 2> gpurun_out/bench_aux.err'
step ncu_full 200 bash -c 'ncu --set full --clock-control none --import-source on -k regex:k_gm_quorum\|k_gm_similarity\|k_gm_union -c 8 -f -o gpurun_out/r1_secondary python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1'
step bench_c2 120 bash -c 'python bench.py --workload c2 --steps 50 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err'
echo "total $(( $(date +%s) - START )) s"
