#!/bin/bash
# wide quantised tree for the AO stream + raster gather with the float-error slack: parity, A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2d_pytest_gpu.log
ab() {   # name, workload, options...
    local name=$1; shift; local wl=$1; shift
    local opts=(); for o in "$@"; do opts+=(--opt "$o"); done
    timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-ncu --ppll-workload none "${opts[@]}" > $O/r2d_ab_$name.json 2> $O/r2d_ab_$name.err; echo "ab $name rc=$?"
}
ab c5_default config5
ab c5_wide config5 b200_ao_wide=true
ab c5_wide_mb7 config5 b200_ao_wide=true b200_ao_min_blocks=7
ab c5_wide_mb9 config5 b200_ao_wide=true b200_ao_min_blocks=9
ab c3_default config3
ab c3_wide config3 b200_ao_wide=true
ab c4_default config4
ab c4_mb6 config4 b200_ppll_raster_min_blocks=6
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    c = d["config"]
    if "ms_gather" in c:
        print("%-36s frame %.2f ms  gather %.2f resolve %.2f frags %d" % (f.split("/")[-1], d["ms_per_step"], c["ms_gather"], c["ms_resolve"], c["frags_sorted"]))
    else:
        print("%-36s frame %.2f ms  ao_stream %.2f  T/ray %.2f I/ray %.2f  %.0f Mrays/s" % (f.split("/")[-1], d["ms_per_step"], d["roofline"]["kernel_ms"], c["T_per_ao_ray"], c["I_per_ao_ray"], d["value"]))
PY
