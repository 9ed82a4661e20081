#!/bin/bash
# full validation of the new default AO kernel configuration + fresh profiles
mkdir -p gpurun_out
S=gpurun_out/summary2.txt; : > $S
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_tests.log 2>&1; echo "all_tests rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c2_bench_n1.json 2> gpurun_out/c2_bench_n1.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c2_bench_ref.json 2> gpurun_out/c2_bench_ref.err; echo "ref rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c2_ncu_bench.log 2>&1; echo "ncu_launches rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o gpurun_out/c2_k_rtao_rays python tools/profile_run.py --skip-ppll > gpurun_out/c2_ncu_full.log 2>&1; echo "ncu_full rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 300 python tools/prebake_bench.py > gpurun_out/c2_prebake.json 2> gpurun_out/c2_prebake.err; echo "prebake rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
cat $S; tail -3 gpurun_out/c2_tests.log; head -c 400 gpurun_out/c2_bench_n1.json
