#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python tools/ao_ab.py --workload config5 --variant "b200_ao_packed=false" --variant "" --variant "b200_ao_refill_below=29" > $O/r2p_ab5.log 2>&1; echo "ab5 rc=$?"; cat $O/r2p_ab5.log | tail -4
python tools/ao_ab.py --workload config3 --variant "" > $O/r2p_ab3.log 2>&1; echo "ab3 rc=$?"; cat $O/r2p_ab3.log | tail -2
