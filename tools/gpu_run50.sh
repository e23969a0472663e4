#!/bin/bash
# CLI GPU tests with the final host binary (lean parse with subset lists, BGZF input)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_cli.py -m gpu -x -q > gpurun_out/r2w_pytest_cli.log 2>&1; echo "pytest cli rc=$?"; tail -3 gpurun_out/r2w_pytest_cli.log
