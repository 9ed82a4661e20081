#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python tools/ao_ab.py --workload config5 --variant "b200_tube_prepass=false" --variant "" > $O/r2p_seed5.log 2>&1; echo "seed5 rc=$?"; cat $O/r2p_seed5.log | tail -3
python tools/ao_ab.py --workload config3 --variant "b200_tube_prepass=false" --variant "" > $O/r2p_seed3.log 2>&1; echo "seed3 rc=$?"; cat $O/r2p_seed3.log | tail -3
