#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2o_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2o_pytest_gpu.log
bash tools/gpu_ab.sh r2o config5 "-" "b200_bvh_morton=per_axis"
bash tools/gpu_ab.sh r2o3 config3 "-" "b200_bvh_morton=per_axis"
bash tools/gpu_ab.sh r2o4 config4 "-" "b200_bvh_morton=per_axis"
python - <<'PY'
# per-rank PPLL times of an 8-way tile shard, emulated on one GPU (config 4)
import sys, numpy as np
sys.path.insert(0, ".")
import bench, linevis_b200 as lv
pw = bench.PPLL_WORKLOADS["config4"]
pos, attr, seg = bench.generate(pw["gen"])
cam = lv.make_camera(pw["W"], pw["H"])
import torch
frame = torch.zeros((pw["H"], pw["W"], 4), dtype=torch.float32, device="cuda")
ctx = lv.Context(0)
ctx.set_transfer_function(lv.scenes.standard_transfer_function(opacity=(0.1, 0.6)))
ctx.set_option("ambient_occlusion_strength", 0.0); ctx.set_option("b200_expected_avg_depth_complexity", 24)
sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
for world in (1, 8):
    g, r = [], []
    for rank in range(world):
        ctx.set_tile_shard(rank, world, 64)
        for _ in range(3):
            st = ctx.render_ppll(sc, cam, 256, "priority_queue", 0, out=frame)[1]
        g.append(st["ms_gather"]); r.append(st["ms_resolve"])
    print("world %d: gather per rank %s  resolve %s" % (world, ["%.2f" % x for x in g], ["%.2f" % x for x in r]))
PY
