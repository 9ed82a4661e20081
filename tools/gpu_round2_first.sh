#!/bin/bash
# First GPU call of round 2 (one box, ~12 GPU-minutes): everything that was written after round 1's last GPU minute, in the order
# that matters if the call is cut short.  Run as
#     gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
# Results land in gpurun_out/r2a_*; DESIGN.md section 8 says what each line decides.
mkdir -p gpurun_out
O=gpurun_out

# 1. parity: the whole -m gpu suite (new since the last GPU run: degenerate segments, triangle tubes, qnodes, resolve variants,
#    raster gather, peer frame, rgba8, full-size kernel-independence properties)
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2a_pytest_gpu.log

# 2. the default bench line (also the first measurement of e2e_rgba8)
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2a_bench_default.json 2> $O/r2a_bench_default.err; echo "bench rc=$?"

# 3. A/B lines: one option each, CPU baseline skipped.  Compare ms_per_step / k_rtao_rays_ms_per_rank (AO stream) and
#    ppll.ms_resolve / ppll.ms_gather / ppll_config4.* (PPLL) against the default line.
ab() {   # name, options...
    local name=$1; shift
    local opts=(); for o in "$@"; do opts+=(--opt "$o"); done
    timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "${opts[@]}" > $O/r2a_ab_$name.json 2> $O/r2a_ab_$name.err; echo "ab $name rc=$?"
}
ab qnodes b200_ao_qnodes=true
ab regsort b200_ppll_reg_sort=true
ab tile256 b200_ppll_resolve_tile=256
ab tile512 b200_ppll_resolve_tile=512
ab raster b200_ppll_gather_mode=raster
ab raster_contig b200_ppll_gather_mode=raster_contiguous
ab raster_contig_regsort b200_ppll_gather_mode=raster_contiguous b200_ppll_reg_sort=true

# 4. FMA contraction on (build/liblinevis_b200_fmad.so, tools/build_variant.py): speed and max|delta| vs the strict build / the oracle
LINEVIS_B200_LIB=build/liblinevis_b200_fmad.so timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2a_ab_fmad.json 2> $O/r2a_ab_fmad.err; echo "ab fmad rc=$?"
timeout 300 python tools/fmad_delta.py > $O/r2a_fmad_delta.log 2>&1; cat $O/r2a_fmad_delta.log

# 5. compute-sanitizer on small scenes through every default kernel and the main variants
for tool in memcheck racecheck synccheck; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > $O/r2a_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; tail -2 $O/r2a_sanitizer_$tool.log
done

python - <<'EOF'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2a_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    p2, p4 = d.get("ppll", {}), d.get("ppll_config4", {})
    print("%-46s frame %.2f ms  ao_stream %s  e2e %.0f  rgba8 %s | cfg2 gather %.3f resolve %.3f | cfg4 gather %.2f resolve %.2f" % (
        f.split("/")[-1], d["ms_per_step"], d["config"].get("k_rtao_rays_ms_per_rank"), d["e2e"]["value"],
        d.get("e2e_rgba8", {}).get("value", d.get("e2e_rgba8", {}).get("error")),
        p2.get("ms_gather", -1), p2.get("ms_resolve", -1), p4.get("ms_gather", -1), p4.get("ms_resolve", -1)))
EOF
