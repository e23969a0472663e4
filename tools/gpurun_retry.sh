#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box / slots draining): tools/gpurun_retry.sh <timeout_s> <command...>
T=$1; shift
for i in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
    if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
    echo "$out"; exit 0
done
echo "gave up after 40 transient answers"; exit 3
