#!/bin/bash
# One gpurun call: GPU parity suite, old-vs-new kernel comparison, the bench line and an ncu launch list.
# Everything lands in gpurun_out/ (merged back by gpurun).  Each step has its own timeout so that a slow
# step cannot starve the ones after it.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 420 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time timeout 240 python tools/bench_variants.py ) > gpurun_out/bench_variants.txt 2> gpurun_out/bench_variants.err
echo "variants rc=$?"; cat gpurun_out/bench_variants.txt
( time timeout 200 python bench.py --steps 50 --warmup 5 ) > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_target.json
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 240 python tools/bench_aux.py ) > gpurun_out/bench_aux.txt 2> gpurun_out/bench_aux.err
echo "aux rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_variants.csv python tools/bench_variants.py --quick > gpurun_out/ncu_variants.log 2>&1
echo "ncu rc=$?"
