#!/bin/bash
# Second GPU call of the round: parity suite, variants, ncu --set full of the secondary kernels, c2 bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
START=$(date +%s)
BUDGET_S=${BUDGET_S:-420}
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
step pytest 240 bash -c 'python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log'
tail -4 gpurun_out/pytest_gpu.log
step variants 120 bash -c 'python tools/bench_variants.py > gpurun_out/bench_variants.txt 2> gpurun_out/bench_variants.err'
cat gpurun_out/bench_variants.txt; tail -3 gpurun_out/bench_variants.err
step ncu_full 200 bash -c 'ncu --set full --clock-control none --import-source on -k regex:k_gm_quorum\|k_gm_similarity -c 8 -f -o gpurun_out/r1_secondary python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1'
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
step aux 120 bash -c 'python tools/bench_aux.py > gpurun_out/bench_aux.txt 2> gpurun_out/bench_aux.err'
step bench_c2 120 bash -c 'python bench.py --workload c2 --steps 50 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err'
tail -c 300 gpurun_out/bench_c2.json
echo "total $(( $(date +%s) - START )) s"
