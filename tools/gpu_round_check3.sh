#!/bin/bash
# Third GPU call: parity suite + variants + ncu of the union / quorum / similarity kernels after the loop rework.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
START=$(date +%s)
BUDGET_S=${BUDGET_S:-330}
left() { echo $(( START + BUDGET_S - $(date +%s) )); }
step() {
    local name=$1 max=$2; shift 2
    local l; l=$(left)
    if [ "$l" -lt 20 ]; then echo "== $name: skipped (deadline)"; return; fi
    [ "$l" -lt "$max" ] && max=$l
    local t0; t0=$(date +%s)
    timeout "$max" "$@"
    echo "== $name rc=$? in $(( $(date +%s) - t0 )) s (limit $max)"
}
step pytest 200 bash -c 'python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log'
tail -12 gpurun_out/pytest_gpu.log
step variants 90 bash -c 'python tools/bench_variants.py > gpurun_out/bench_variants.txt 2> gpurun_out/bench_variants.err'
cat gpurun_out/bench_variants.txt; tail -3 gpurun_out/bench_variants.err
step aux 90 bash -c 'python tools/bench_aux.py > gpurun_out/bench_aux.txt 2> gpurun_out/bench_aux.err'
grep -E "permuted|similarity rows|ordered_growth 3" gpurun_out/bench_aux.txt
step ncu_full 150 bash -c 'ncu --set full --clock-control none --import-source on -k regex:k_gm_quorum\|k_gm_similarity\|k_gm_union -c 8 -f -o gpurun_out/r1_secondary_v2 python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1'
tail -2 gpurun_out/ncu_full.log
echo "total $(( $(date +%s) - START )) s"
