#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $O/r2y_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2y_pytest_gpu.log
( time timeout 900 python bench.py ) > $O/r2y_bench.json 2> $O/r2y_bench.err; echo "bench rc=$?"; tail -4 $O/r2y_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o gpurun_out/r2y_rays_final python tools/profile_run.py --skip-ppll > gpurun_out/r2y_ncu.log 2>&1; echo "ncu rc=$?"
