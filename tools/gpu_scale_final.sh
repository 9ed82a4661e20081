#!/bin/bash
# final scaling table of the round on ONE 8-GPU box: config 5 at N = 1, 2, 4, 8 (two frames in flight), config 4 at 1 and 8
mkdir -p gpurun_out
O=gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n"; fi
  timeout 300 $cmd --steps 10 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none > $O/r2f_c5_n$n.json 2> $O/r2f_c5_n$n.err; echo "c5 n$n rc=$?"
done
for n in 1 8; do
  if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) bench.py --gpus $n"; fi
  timeout 300 $cmd --workload config4 --steps 10 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2f_c4_n$n.json 2> $O/r2f_c4_n$n.err; echo "c4 n$n rc=$?"
done
python - <<'PY'
import json
for f in ("c5_n1","c5_n2","c5_n4","c5_n8","c4_n1","c4_n8"):
    try:
        d=json.loads(open('gpurun_out/r2f_%s.json'%f).read().strip().splitlines()[0])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'one-in-flight', d['config'].get('ms_one_frame_in_flight'), 'e2e', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
