#!/bin/bash
# Final validation call of the round: parity suite, variants, the bench line and smoke().
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
START=$(date +%s)
BUDGET_S=${BUDGET_S:-200}
left() { echo $(( START + BUDGET_S - $(date +%s) )); }
step() {
    local name=$1 max=$2; shift 2
    local l; l=$(left)
    if [ "$l" -lt 15 ]; then echo "== $name: skipped (deadline)"; return; fi
    [ "$l" -lt "$max" ] && max=$l
    local t0; t0=$(date +%s)
    timeout "$max" "$@"
    echo "== $name rc=$? in $(( $(date +%s) - t0 )) s (limit $max)"
}
step pytest 150 bash -c 'python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log'
tail -12 gpurun_out/pytest_gpu.log
step variants 60 bash -c 'python tools/bench_variants.py > gpurun_out/bench_variants.txt 2> gpurun_out/bench_variants.err'
cat gpurun_out/bench_variants.txt; tail -3 gpurun_out/bench_variants.err
step bench 60 bash -c 'python bench.py --steps 50 --warmup 5 > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err'
tail -c 300 gpurun_out/bench_target.json
step smoke 30 bash -c 'python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1'
tail -1 gpurun_out/smoke.log
echo "total $(( $(date +%s) - START )) s"
