#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python tools/ao_ab.py --workload config5 --variant "" --variant "b200_ao_direct_queue=true" --variant "b200_ao_direct_queue=true,b200_ao_refill_below=30" --variant "b200_ao_direct_queue=true,b200_ao_refill_below=26" --variant "b200_ao_direct_queue=true,b200_ao_leaf_vote=20"  > $O/r2p_ab5.log 2>&1; echo "ab5 rc=$?"; cat $O/r2p_ab5.log | tail -8
python tools/ao_ab.py --workload config3 --variant "" --variant "b200_ao_direct_queue=true"  > $O/r2p_ab3.log 2>&1; echo "ab3 rc=$?"; cat $O/r2p_ab3.log | tail -3
