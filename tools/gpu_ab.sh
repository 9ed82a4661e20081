#!/bin/bash
# generic A/B: tools/gpu_ab.sh <tag> <workload> "<opts a>" "<opts b>" ...   (each opts string: space-separated key=value, or "-" for none)
mkdir -p gpurun_out
TAG=$1; shift; WL=$1; shift
i=0
for o in "$@"; do
  opts=(); if [ "$o" != "-" ]; then for k in $o; do opts+=(--opt $k); done; fi
  timeout 400 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none "${opts[@]}" 2>gpurun_out/${TAG}_$i.err > gpurun_out/${TAG}_$i.json
  python - "$o" gpurun_out/${TAG}_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); c=d['config']
    if 'ms_gather' in c: print('%-60s frame %.2f gather %.2f resolve %.2f frags %d' % (sys.argv[1], d['ms_per_step'], c['ms_gather'], c['ms_resolve'], c['frags_sorted']))
    else: print('%-60s frame %.2f ao %.2f T %.2f I %.2f  %.0f Mrays/s' % (sys.argv[1], d['ms_per_step'], d['roofline']['kernel_ms'], c['T_per_ao_ray'], c['I_per_ao_ray'], d['value']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
  i=$((i+1))
done
