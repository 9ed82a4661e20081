#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() {  # name, nproc, extra args...
    local name=$1; shift; local n=$1; shift
    if [ "$n" = "1" ]; then timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-ncu "$@" > $O/s8b_$name.json 2> $O/s8b_$name.err
    else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 5 --warmup 3 "$@" > $O/s8b_$name.json 2> $O/s8b_$name.err; fi
    python - $name $O/s8b_$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); c=d["config"]
    if "ms_gather" in c: print("%-14s n=%d frame %.3f ms  %.0f Mfrags/s  e2e %.3f ms (%.0f) gather %.2f resolve %.2f" % (sys.argv[1], d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], c["ms_gather"], c["ms_resolve"]))
    else: print("%-14s n=%d frame %.3f ms  %.0f Mrays/s  e2e %.3f ms (%.0f)  rays/rank ms %s  assemble %s" % (sys.argv[1], d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], ["%.2f" % x for x in c["k_rtao_rays_ms_per_rank"]], c.get("assemble_ms")))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
}
run n1 1 --ppll-workload none
run n8 8 --ppll-workload none
run n8_noprepass 8 --ppll-workload none --opt b200_tube_prepass=false
run n4 4 --ppll-workload none
run c4_n1 1 --workload config4
run c4_n8 8 --workload config4
