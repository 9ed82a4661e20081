#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ab.sh r2h config5 "-" "b200_ao_wide=true" "b200_ao_wide=true b200_ao_raybuf=true" "b200_ao_wide=true b200_ao_raybuf=true b200_ao_refill_below=28" "b200_ao_wide=true b200_ao_raybuf=true b200_ao_refill_below=30" "b200_ao_wide=true b200_ao_raybuf=true b200_ao_refill_below=32" "b200_ao_raybuf=true" "b200_ao_raybuf=true b200_ao_refill_below=30" "b200_ao_wide=true b200_ao_raybuf=true b200_ao_refill_below=30 b200_ao_leaf_vote=16" "b200_ao_wide=true b200_ao_raybuf=true b200_ao_refill_below=30 b200_ao_leaf_vote=20"
bash tools/gpu_ab.sh r2h3 config3 "-" "b200_ao_wide=true b200_ao_raybuf=true b200_ao_refill_below=30"
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
