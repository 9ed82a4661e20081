#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "stack_layouts" > gpurun_out/e2_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/e2_tests.log
timeout 900 python tools/sweep.py --leaf 1 --refill 24 --vote 12 --minb 7 8 9 10 --stack 1 8 12 16 24 32 > gpurun_out/e2_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/e2_sweep.log
