#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 240 python tools/check_simmma.py --time > gpurun_out/r2_run12_simmma.log 2>&1; echo "simmma rc=$?"; tail -20 gpurun_out/r2_run12_simmma.log | cut -c1-300
nvidia-smi --query-gpu=name,memory.used --format=csv | tail -1
