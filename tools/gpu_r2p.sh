#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for v in "--steps 5" "--steps 5 --frames-in-flight 1" "--steps 10" "--steps 5"; do
timeout 600 python bench.py $v --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none > $O/r2p_e2e.json 2> $O/r2p_e2e.err; echo "[$v] rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_e2e.json').read().strip().splitlines()[0])
print(round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2e32f', round(d['e2e_rgba32f']['ms_per_step'],3), d['e2e'].get('frames_in_flight'))
PY
done
