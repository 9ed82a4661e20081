#!/bin/bash
mkdir -p gpurun_out
python tools/shard_emul.py --world 2 8 --balance contiguous > gpurun_out/r2s_shard_cont.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s_shard_cont.log
python tools/shard_emul.py --world 8 --balance contiguous --tile 32 > gpurun_out/r2s_shard_cont32.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s_shard_cont32.log
