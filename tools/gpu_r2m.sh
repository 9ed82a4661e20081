#!/bin/bash
mkdir -p gpurun_out
for v in default flush8 flush16 flush24 flush32; do
  if [ "$v" != "default" ]; then export LINEVIS_B200_LIB=build/liblinevis_b200_$v.so; else unset LINEVIS_B200_LIB; fi
  echo "== $v"
  bash tools/gpu_ab.sh r2m_$v config5 "-"
  bash tools/gpu_ab.sh r2m3_$v config3 "-"
done
unset LINEVIS_B200_LIB
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
