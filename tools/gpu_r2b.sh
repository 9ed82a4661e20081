#!/bin/bash
# ncu --set full of the object-order PPLL gather and the in-register resolve on config 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ppll_gather_raster|k_ppll_resolve" -s 2 -c 2 -f -o gpurun_out/r2b_ppll_raster \
    python tools/profile_run.py --skip-tubes --opt b200_ppll_gather_mode=raster b200_ppll_reg_sort=true > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2b_ncu.log
