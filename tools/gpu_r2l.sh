#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|DeviceRadixSort" -c 400 --csv --log-file $O/r2z_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2z_launches_bench.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py $O/r2z_launches_bench.csv | tail -30
