#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/r2t_ppll_tiles.log 2>&1 <<'PY'
# per-rank PPLL times of an 8-way tile shard, emulated on one GPU (config 4), for several tile sizes
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
for world, tile in ((1, 64), (8, 64), (8, 128), (8, 256), (8, 512)):
    g, r, f = [], [], []
    for rank in range(world):
        ctx.set_tile_shard(rank, world, tile)
        for _ in range(3):
            st = ctx.render_ppll(sc, cam, 256, "priority_queue", 0, out=frame)[1]
        g.append(st["ms_gather"]); r.append(st["ms_resolve"]); f.append(st["frags_sorted"])
    print("world %d tile %d: gather max %.2f mean %.2f  resolve max %.2f  frags min %.1fM max %.1fM" % (world, tile, max(g), np.mean(g), max(r), min(f) / 1e6, max(f) / 1e6), flush=True)
PY
echo rc=$?; tail -6 gpurun_out/r2t_ppll_tiles.log
