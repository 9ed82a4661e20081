#!/bin/bash
# gpurun with retries while the pod answers "busy" (status transient): usage  tools/gpurun_retry.sh <timeout-s> <script>
T=$1; shift
for i in $(seq 1 20); do
    gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
    if grep -q "status=transient" /tmp/gpurun_last.log; then sleep 60; continue; fi
    break
done
cat /tmp/gpurun_last.log
