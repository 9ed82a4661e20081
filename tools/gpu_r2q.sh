#!/bin/bash
# ncu --set full of the shipped AO stream on config 5 (end of round 2)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o gpurun_out/r2z_rays_final python tools/profile_run.py --skip-ppll > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2z_ncu.log
