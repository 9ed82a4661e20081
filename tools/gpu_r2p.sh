#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python tools/ao_ab.py --workload config5 --variant "b200_packet_wide=false" --variant "" > $O/r2p_pw5.log 2>&1; echo "pw5 rc=$?"; cat $O/r2p_pw5.log | tail -3
python tools/ao_ab.py --workload config3 --variant "b200_packet_wide=false" --variant "" > $O/r2p_pw3.log 2>&1; echo "pw3 rc=$?"; cat $O/r2p_pw3.log | tail -3
