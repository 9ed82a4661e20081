#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/sweep.py --leaf 1 --combo 9:6:24:12 9:8:24:12 9:9:24:12 9:10:24:12 9:11:24:12 9:12:24:12 9:13:24:12 10:9:24:12 10:12:24:12 8:11:24:12 8:15:24:12 \
   9:12:16:12 9:12:20:12 9:12:28:12 9:12:32:12 9:12:24:8 9:12:24:16 9:12:24:20 9:10:20:16 9:10:28:8 > gpurun_out/e3_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/e3_sweep.log
