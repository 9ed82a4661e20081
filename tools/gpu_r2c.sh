#!/bin/bash
# raster gather v2 (ray table, hit queue, per-segment hoisting): parity + A/B on config 4 / config 2 through the PPLL headline bench
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2c_pytest_gpu.log
ab() {   # name, workload, options...
    local name=$1; shift; local wl=$1; shift
    local opts=(); for o in "$@"; do opts+=(--opt "$o"); done
    timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-ncu "${opts[@]}" > $O/r2c_ab_$name.json 2> $O/r2c_ab_$name.err; echo "ab $name rc=$?"
}
ab c4_default config4
ab c4_slack1 config4 b200_ppll_raster_slack=1.0
ab c4_slack025 config4 b200_ppll_raster_slack=0.25
ab c4_mb5 config4 b200_ppll_raster_min_blocks=5
ab c4_mb6 config4 b200_ppll_raster_min_blocks=6
ab c4_tile512 config4 b200_ppll_resolve_tile=512
ab c4_raycast config4 b200_ppll_gather_mode=raycast
ab c4_raycast_tile512 config4 b200_ppll_gather_mode=raycast b200_ppll_resolve_tile=512
ab c2_default config2
ab c2_mb5 config2 b200_ppll_raster_min_blocks=5
ab c2_raycast config2 b200_ppll_gather_mode=raycast
# the full headline line for config 4 (CPU baseline + parity + live ncu)
timeout 600 python bench.py --workload config4 --steps 5 --warmup 3 > $O/r2c_bench_config4.json 2> $O/r2c_bench_config4.err; echo "bench config4 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    c = d["config"]
    print("%-36s frame %.2f ms  clear %.3f gather %.2f resolve %.2f  e2e %.2f ms  frags %d  parity %s" % (
        f.split("/")[-1], d["ms_per_step"], c["ms_clear"], c["ms_gather"], c["ms_resolve"], d["e2e"]["ms_per_step"], c["frags_sorted"], d.get("parity_max_abs_delta")))
PY
