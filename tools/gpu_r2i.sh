#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2i_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2i_bench_default.json 2> $O/r2i_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2i_bench_default.json").read().strip().splitlines()[-1])
print("frame %.2f ms value %.0f  e2e %.2f ms (%.0f)  parity %s frac %.3f frac_ref %s build %.1f ms" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("parity_max_abs_delta"), d["roofline"]["frac"], d["roofline"].get("frac_ref_tree"), d["config"]["bvh_build_ms"]))
print(d["roofline"].get("limiter")); print(d["roofline"].get("ref_tree")); print(d["cpu_baseline"])
for k in ("ppll","ppll_config4"): print(k, d[k]["ms_gather"], d[k]["ms_resolve"], d[k]["gather_mode"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2i_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tubes|k_rtao_primary" -s 2 -c 2 -f -o $O/r2i_primary_tubes python tools/profile_run.py --skip-ppll > $O/r2i_ncu_pt.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rtao_rays -s 1 -c 1 -f -o $O/r2i_rays_default python tools/profile_run.py --skip-ppll > $O/r2i_ncu_rays.log 2>&1; echo "ncu rc=$?"
