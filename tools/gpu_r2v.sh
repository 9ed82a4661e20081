#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "sample_batch or packed or rtao or config5_crop" > $O/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2v_pytest.log
for sh in tiles samples; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none --shard $sh > $O/r2v_n2_$sh.json 2> $O/r2v_n2_$sh.err; echo "n2 $sh rc=$?"; tail -2 $O/r2v_n2_$sh.err
python - $sh <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2v_n2_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['config']['k_rtao_rays_ms_per_rank'], d['e2e']['ms_per_step'])
PY
done
