#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_transpose -s 1 -c 1 -o gpurun_out/r2_run9_transpose_full -f python tools/one_transpose.py 10000000 1024 > gpurun_out/r2_run9_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_run9_transpose_full.ncu-rep --page raw --csv > gpurun_out/r2_run9_transpose_raw.csv 2>/dev/null
ls -la gpurun_out/r2_run9*
