#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2p_c4.json 2> $O/r2p_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2p_c2.json 2> $O/r2p_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ("c4","c2"):
    d=json.loads(open('gpurun_out/r2p_%s.json'%f).read().strip().splitlines()[0])
    print(f, round(d['value'],1), round(d['ms_per_step'],3), {k:d['config'][k] for k in ('ms_gather','ms_resolve')})
PY
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "ppll or tubes or config4 or config3" 2>&1 | tail -2
