#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $O/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed" $O/r2z_pytest_gpu.log | tail -1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2z_smoke.log
( time timeout 900 python bench.py ) > $O/r2z_bench.json 2> $O/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[0])
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), round(d['roofline'].get('frac_ref_tree') or 0,3), d['roofline']['kernel'], round(d['roofline']['kernel_ms'],2), 'parity', d.get('parity_max_abs_delta'), round(d['cpu_baseline']['value'],2), d['e2e'].get('frames_in_flight'))
PY
