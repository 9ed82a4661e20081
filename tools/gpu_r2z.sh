#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for tool in memcheck racecheck synccheck; do
    timeout 500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > $O/r2z_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; tail -2 $O/r2z_sanitizer_$tool.log
done
