#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python tools/ao_ab.py --workload config5 --variant "" --variant "b200_ao_wide_reps=2" --variant "b200_ao_wide_reps=2,b200_ao_refill_below=26" --variant "b200_ao_wide_reps=3" > $O/r2p_ab5.log 2>&1; echo "ab5 rc=$?"; cat $O/r2p_ab5.log | tail -5
