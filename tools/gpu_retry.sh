#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod answers "transient"/busy
t=$1; shift
for i in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun --timeout "$t" "$@" 2>&1)
    if echo "$out" | grep -q "status=transient\|rc=3\|no box or slot"; then sleep 45; continue; fi
    echo "$out"; exit 0
done
echo "gave up after 40 tries"; exit 3
