#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ab.sh r2j config5 "-" "b200_bvh_builder=ploc" "b200_bvh_builder=ploc b200_bvh_ploc_radius=8" "b200_bvh_builder=ploc b200_bvh_ploc_radius=32"
for f in gpurun_out/r2j_*.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); c=d['config']; print('$f', 'build ms', c['bvh_build_ms'], 'first', c['bvh_build_ms_first_in_process'], 'prim steps', c['primary_packet_steps_per_ray'])"; done
bash tools/gpu_ab.sh r2j3 config3 "-" "b200_bvh_builder=ploc"
bash tools/gpu_ab.sh r2j4 config4 "-" "b200_bvh_builder=ploc b200_ppll_gather_mode=raycast" "b200_ppll_gather_mode=raycast"
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "crop" 2>&1 | tail -3
