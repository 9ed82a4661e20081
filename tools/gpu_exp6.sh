#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o gpurun_out/e6_k_rtao_rays_q python tools/profile_run.py --skip-ppll --opt b200_ao_queue=true b200_ao_min_blocks=8 > gpurun_out/e6_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/e6_ncu.log
