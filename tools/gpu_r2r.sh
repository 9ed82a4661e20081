#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for v in "" "--opt b200_packet_carveout=100"; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none $v > $O/r2r_fif2.json 2> $O/r2r_fif2.err; echo "fif2 [$v] rc=$?"; tail -3 $O/r2r_fif2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_fif2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k:d['config'].get(k) for k in ('frames_in_flight','ms_one_frame_in_flight','frames_in_flight_equal_to_single')}, d['e2e']['ms_per_step'])
PY
done
