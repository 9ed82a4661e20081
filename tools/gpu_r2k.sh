#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2k_pytest_gpu.log
bash tools/gpu_ab.sh r2k config5 "-"
bash tools/gpu_ab.sh r2k3 config3 "-"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|DeviceRadixSort|DeviceScan" -c 60 --csv --log-file $O/r2k_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2k_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 python tools/tri_timing.py 1000000 > $O/r2k_tri_timing.log 2>&1; cat $O/r2k_tri_timing.log
