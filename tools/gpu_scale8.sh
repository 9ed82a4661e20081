#!/bin/bash
# 8-GPU scaling check: tile sizes of the image shards, value and e2e; one run each (N = 8), plus N = 1 on the same box.
mkdir -p gpurun_out
O=gpurun_out
run() {  # name, nproc, extra args...
    local name=$1; shift; local n=$1; shift
    if [ "$n" = "1" ]; then timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --ppll-workload none --no-cpu-baseline --no-ncu "$@" > $O/s8_$name.json 2> $O/s8_$name.err
    else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 5 --warmup 3 --ppll-workload none "$@" > $O/s8_$name.json 2> $O/s8_$name.err; fi
    python - $name $O/s8_$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); c=d["config"]
    print("%-14s n=%d frame %.3f ms  %.0f Mrays/s  e2e %.3f ms (%.0f)  rays/rank ms %s  assemble %s" % (sys.argv[1], d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], ["%.2f" % x for x in c["k_rtao_rays_ms_per_rank"]], c.get("assemble_ms")))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
}
run n1 1
run n8_t64 8 --tile 64
run n8_t32 8 --tile 32
run n8_t16 8 --tile 16
run n2_t32 2 --tile 32
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/s8_ref_n2.json 2> $O/s8_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-300 $O/s8_ref_n2.json
