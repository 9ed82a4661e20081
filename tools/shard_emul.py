"""Per-rank frame times of an N-way tile shard, emulated on ONE GPU (each rank's tiles rendered in turn): what the slowest rank of an
N-GPU run would take, without the fence.  python tools/shard_emul.py --world 8 [--tile 64] [--opt k=v ...]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import linevis_b200 as lv

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--world", type=int, nargs="+", default=[1, 8])
ap.add_argument("--tile", type=int, default=64)
ap.add_argument("--balance", default="", choices=["", "lpt", "contiguous"], help="cost-balanced tile ownership (lv_set_tile_owners) from the first frame's per-tile hit counts")
ap.add_argument("--opt", action="append", default=[])
args = ap.parse_args()
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS[args.workload]
pos, attr, seg = bench.generate(wl["gen"], dev)
cam = lv.make_camera(wl["W"], wl["H"])
frame = torch.zeros((wl["H"], wl["W"], 4), dtype=torch.float32, device=dev)
ctx = lv.Context(0)
ctx.set_transfer_function(lv.scenes.standard_transfer_function())
ctx.set_new_settings({"depth_cue_strength": 0.0, "ambient_occlusion_strength": 1.0, "ambient_occlusion_gamma": 1.0,
                      "ambient_occlusion_samples_per_frame": wl["ao_spp"], "ambient_occlusion_iterations": 1, "ambient_occlusion_radius": 0.1,
                      "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True,
                      "num_samples_per_frame": 1, "num_accumulated_frames": 1, "use_deterministic_sampling": False})
for kv in args.opt:
    ctx.set_option(*kv.split("=", 1))
sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
for world in args.world:
    owners = None
    if args.balance and world > 1:
        from linevis_b200 import sharding
        costs = sharding.tile_costs_single_gpu(ctx, sc, cam, args.tile, frame)
        owners = (sharding.balance_tiles if args.balance == "lpt" else sharding.balance_tiles_contiguous)(costs, world, args.tile * args.tile, wl["ao_spp"])
    rows = []
    for rank in range(world):
        ctx.set_tile_shard(rank, world, args.tile)
        if owners is not None:
            ctx.set_tile_owners(wl["W"], wl["H"], owners)
        ts = []
        for _ in range(4):
            st = ctx.render_tubes(sc, cam, 0, out=frame, stats=True)[1]
            ts.append((st["ms_total"], st["ms_rtao_rays"], st["ms_rtao"] - st["ms_rtao_rays"], st["ms_trace"], st["rays_ao"]))
        rows.append(np.min(np.array(ts[1:]), axis=0))
    rows = np.array(rows)
    print("world %d tile %d%s: frame max %.3f ms (mean %.3f, min %.3f)  AO stream max %.3f mean %.3f  rest-of-RTAO mean %.3f  tubes mean %.3f  AO rays min %.2fM max %.2fM" %
          (world, args.tile, " balanced " + args.balance if owners is not None else "", rows[:, 0].max(), rows[:, 0].mean(), rows[:, 0].min(), rows[:, 1].max(), rows[:, 1].mean(), rows[:, 2].mean(), rows[:, 3].mean(),
           rows[:, 4].min() / 1e6, rows[:, 4].max() / 1e6), flush=True)
    if world > 1:
        print("   per rank frame ms:", " ".join("%.2f" % x for x in rows[:, 0]))
