#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 240 python tools/check_simmma.py --time > gpurun_out/r2_run14_simmma.log 2>&1; echo "simmma rc=$?"; tail -12 gpurun_out/r2_run14_simmma.log | cut -c1-250
timeout 200 python tools/time_simmma.py 2>&1 | tee gpurun_out/r2_run14_simmma_debug.log | tail -4
