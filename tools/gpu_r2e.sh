#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o gpurun_out/r2e_rays_wide python tools/profile_run.py --skip-ppll --opt b200_ao_wide=true > gpurun_out/r2e_ncu_wide.log 2>&1; echo "ncu wide rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o gpurun_out/r2e_rays_wide_top85 python tools/profile_run.py --skip-ppll --opt b200_ao_wide=true b200_ao_wide_top=85 > gpurun_out/r2e_ncu_wide85.log 2>&1; echo "ncu wide85 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ppll_gather_raster -s 1 -c 1 -f -o gpurun_out/r2e_raster_v2 python tools/profile_run.py --skip-tubes --opt b200_ppll_raster_min_blocks=6 > gpurun_out/r2e_ncu_raster.log 2>&1; echo "ncu raster rc=$?"
for o in "b200_ao_wide=true b200_ao_wide_top=85" "b200_ao_wide=true b200_ao_wide_top=341" "b200_ao_wide=true b200_ao_min_blocks=9"; do
  opts=(); for k in $o; do opts+=(--opt $k); done
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none "${opts[@]}" 2>gpurun_out/r2e_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$o', 'frame %.2f ao %.2f T %.2f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['T_per_ao_ray']))"
done
