#!/bin/bash
# One GPU-box visit: new-feature parity first, then the benchmark line, then the full -m gpu suite, then the reference arm.
# Every step has its own timeout and log under gpurun_out/ so that a late failure cannot hide an early result.
mkdir -p gpurun_out
S=gpurun_out/summary.txt
: > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $S 2>&1
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_prebaker.py tests/test_host_adapter.py -m gpu -q > gpurun_out/t_new.log 2>&1; echo "new_tests rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 300 python tools/prebake_bench.py > gpurun_out/prebake.json 2> gpurun_out/prebake.err; echo "prebake rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_prebaker.py > gpurun_out/t_all.log 2>&1; echo "all_tests rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
cat $S; tail -5 gpurun_out/t_new.log; tail -3 gpurun_out/t_all.log; head -c 600 gpurun_out/prebake.json
