#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2g_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --opt b200_ao_wide=true > $O/r2g_bench_wide.json 2> $O/r2g_bench_wide.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2g_bench_wide.json").read().strip().splitlines()[-1])
print("frame %.2f ms value %.0f  e2e %.2f ms (%.0f)  e2e32f %.2f ms  parity %s frac %.3f frac_ref %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e_rgba32f"]["ms_per_step"], d.get("parity"), d["roofline"]["frac"], d["roofline"].get("frac_ref_tree")))
print(d["roofline"].get("limiter")); print(d["roofline"].get("traffic"), d["roofline"].get("traffic_source")); print(d["e2e"])
PY
bash tools/gpu_ab.sh r2g config5 "b200_ao_wide=true b200_ao_wide_reps=2" "b200_ao_wide=true b200_ao_wide_reps=3" "b200_ao_wide=true b200_ao_refill_below=28" "b200_ao_wide=true b200_ao_refill_below=20" "b200_ao_wide=true b200_ao_leaf_vote=8" "b200_ao_wide=true b200_ao_leaf_vote=16" "b200_ao_wide=true b200_ao_stack=8" "b200_ao_wide=true b200_ao_wide_reps=2 b200_ao_min_blocks=9"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o gpurun_out/r2g_rays_wide python tools/profile_run.py --skip-ppll --opt b200_ao_wide=true > gpurun_out/r2g_ncu_wide.log 2>&1; echo "ncu wide rc=$?"
