#!/bin/bash
# the driver's own 8-GPU command, default flags
mkdir -p gpurun_out
O=gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 ) > $O/r2z_n8_default.json 2> $O/r2z_n8_default.err; echo "n8 default rc=$?"; tail -4 $O/r2z_n8_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_n8_default.json').read().strip().splitlines()[0])
print(round(d['value'],1), round(d['ms_per_step'],3), d['config'].get('ms_one_frame_in_flight'), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e'].get('frames_complete_and_equal'), 'ppll4', round(d['ppll_config4']['ms_per_step'],3) if 'ppll_config4' in d else None)
print(d['config']['parallelism'])
PY
