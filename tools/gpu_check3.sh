#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary3.txt; : > $S
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_tests.log 2>&1; echo "all_tests rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 600 python tools/balance.py --world 8 --tile 64 32 16 > gpurun_out/c3_balance.log 2>&1; echo "balance rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
cat $S; tail -3 gpurun_out/c3_tests.log; cat gpurun_out/c3_balance.log; head -c 300 gpurun_out/c3_bench_n1.json
