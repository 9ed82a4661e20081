#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $O/r2y_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2y_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2y_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2y_smoke.log
( time timeout 900 python bench.py --workload config4 ) > $O/r2y_bench_c4.json 2> $O/r2y_bench_c4.err; echo "bench c4 rc=$?"; tail -3 $O/r2y_bench_c4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload config4 --steps 5 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2y_c4_n2.json 2> $O/r2y_c4_n2.err; echo "c4 n2 rc=$?"; tail -2 $O/r2y_c4_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2y_c5_n2.json 2> $O/r2y_c5_n2.err; echo "c5 n2 rc=$?"; tail -2 $O/r2y_c5_n2.err
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > $O/r2y_ref_n2.json 2> $O/r2y_ref_n2.err; echo "ref n2 rc=$?"; tail -3 $O/r2y_ref_n2.err
python - <<'PY'
import json
for f in ("r2y_bench_c4","r2y_c4_n2","r2y_c5_n2","r2y_ref_n2"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[0])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3) if 'ms_per_step' in d.get('e2e',{}) else d.get('e2e'), d.get('parity_max_abs_delta'))
    except Exception as e: print(f, 'ERR', e)
PY
