#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none > $O/r2x_n1.json 2> $O/r2x_n1.err; echo "n1 rc=$?"; tail -3 $O/r2x_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none > $O/r2x_n2.json 2> $O/r2x_n2.err; echo "n2 rc=$?"; tail -3 $O/r2x_n2.err
python - <<'PY'
import json
for n in ("n1","n2"):
    d=json.loads(open('gpurun_out/r2x_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['value'],1), round(d['ms_per_step'],3), d['config'].get('ms_one_frame_in_flight'), d['config'].get('frames_in_flight_equal_to_single'), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e'].get('frames_complete_and_equal'), d['e2e'].get('frames_in_flight'))
PY
