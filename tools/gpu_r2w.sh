#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none "$@" > $O/r2w_n8_$name.json 2> $O/r2w_n8_$name.err; echo "n8 $name rc=$?"
python - $name <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2w_n8_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), [round(x,2) for x in d['config']['k_rtao_rays_ms_per_rank']], round(d['e2e']['ms_per_step'],3), d['config'].get('ms_one_frame_in_flight'))
PY
}
run tiles --shard tiles
run samples --shard samples
run fif2 --shard tiles --frames-in-flight 2
