"""Per-rank PPLL kernel times of an N-way tile shard, emulated on ONE GPU (config 4): what the slowest rank of an N-GPU frame would take."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, linevis_b200 as lv
pw = bench.PPLL_WORKLOADS["config4"]
pos, attr, seg = bench.generate(pw["gen"])
cam = lv.make_camera(pw["W"], pw["H"])
frame = torch.zeros((pw["H"], pw["W"], 4), dtype=torch.float32, device="cuda")
ctx = lv.Context(0)
ctx.set_transfer_function(lv.scenes.standard_transfer_function(opacity=(0.1, 0.6)))
ctx.set_option("ambient_occlusion_strength", 0.0); ctx.set_option("b200_expected_avg_depth_complexity", 24)
sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
for world in (1, 2, 8):
    g, r, f = [], [], []
    for rank in range(world):
        ctx.set_tile_shard(rank, world, 64)
        for _ in range(3):
            st = ctx.render_ppll(sc, cam, 256, "priority_queue", 0, out=frame)[1]
        g.append(st["ms_gather"]); r.append(st["ms_resolve"]); f.append(st["frags_sorted"])
    print("world %d: gather per rank %s  resolve %s  frags %d" % (world, ["%.2f" % x for x in g], ["%.2f" % x for x in r], sum(f)), flush=True)
