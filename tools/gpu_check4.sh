#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary4.txt; : > $S
t0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c4_tests.log 2>&1; echo "all_tests rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c4_bench_n1.json 2> gpurun_out/c4_bench_n1.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/c4_bench_ref.json 2> gpurun_out/c4_bench_ref.err; echo "ref rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/c4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c4_ncu_bench.log 2>&1; echo "ncu_launches rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
cat $S; tail -3 gpurun_out/c4_tests.log; head -c 300 gpurun_out/c4_bench_n1.json
