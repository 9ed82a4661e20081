#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python tools/ao_ab.py --workload config5 --variant "" --variant "b200_bvh_builder=sah"  > $O/r2p_sah5.log 2>&1; echo "sah5 rc=$?"; cat $O/r2p_sah5.log | tail -4
python tools/ao_ab.py --workload config3 --variant "" --variant "b200_bvh_builder=sah"  > $O/r2p_sah3.log 2>&1; echo "sah3 rc=$?"; cat $O/r2p_sah3.log | tail -3
