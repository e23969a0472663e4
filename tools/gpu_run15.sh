#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/time_simmma.py 2>&1 | tee gpurun_out/r2_run15_simmma_debug.log | tail -6
