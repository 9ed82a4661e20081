#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
# estimate for the AO-sample-batch shard axis at N = 8: every rank would trace ALL hit pixels with 64 / 8 = 8 samples -- the work of a full frame at 8 spp
bash tools/gpu_ab.sh r2n config5 "-" "ambient_occlusion_samples_per_frame=8" "ambient_occlusion_samples_per_frame=16" "ambient_occlusion_samples_per_frame=32"
timeout 600 python bench.py --workload config4 --steps 5 --warmup 3 > $O/r2n_bench_config4.json 2> $O/r2n_bench_config4.err; echo "bench config4 rc=$?"
timeout 600 python bench.py --workload config2 --steps 5 --warmup 3 > $O/r2n_bench_config2.json 2> $O/r2n_bench_config2.err; echo "bench config2 rc=$?"
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --ppll-workload none > $O/r2n_bench_config3.json 2> $O/r2n_bench_config3.err; echo "bench config3 rc=$?"
python - <<'PY'
import json
for n in ("config4","config2","config3"):
    d=json.loads(open("gpurun_out/r2n_bench_%s.json"%n).read().strip().splitlines()[-1])
    print(n, "value %.0f %s frame %.3f ms e2e %.3f ms (%.0f) parity %s cpu %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("parity"), d.get("cpu_baseline",{}).get("value")))
    print("   roofline", {k:v for k,v in d["roofline"].items() if k in ("kernel","kernel_ms","frac","frac_ref_tree","traffic","achieved")}, d["roofline"].get("limiter"))
PY
