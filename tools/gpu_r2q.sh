#!/bin/bash
# ncu --set full of the packed AO stream on config 5
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o gpurun_out/r2q_rays_packed python tools/profile_run.py --skip-ppll > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2q_ncu.log
