#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "permuted or similarity or config3 or large_permuted or cutoff or transpose" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2i_pytest.log
timeout 400 python tools/time_bp_quorum.py > gpurun_out/r2i_bp_quorum.json 2> gpurun_out/r2i_bp_quorum.err; echo "bp rc=$?"; cat gpurun_out/r2i_bp_quorum.json; tail -2 gpurun_out/r2i_bp_quorum.err
