"""Tile-shard load balance, emulated on ONE GPU: render every rank's tile subset in turn and compare the frame times.
max/mean over the ranks is the factor a real N-GPU frame loses to imbalance (plus the gather)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import linevis_b200 as lv

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--world", type=int, nargs="+", default=[8])
ap.add_argument("--tile", type=int, nargs="+", default=[64, 32, 16])
args = ap.parse_args()
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS[args.workload]
pos, attr, seg = bench.generate(wl["gen"], dev)
cam = lv.make_camera(wl["W"], wl["H"])
frame = torch.zeros((wl["H"], wl["W"], 4), dtype=torch.float32, device=dev)
ctx = lv.Context(0)
ctx.set_transfer_function(lv.scenes.standard_transfer_function())
ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": wl["ao_spp"], "ambient_occlusion_iterations": 1,
                      "num_samples_per_frame": 1, "num_accumulated_frames": 1})
sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
for _ in range(2):
    _, st = ctx.render_tubes(sc, cam, 0, out=frame)
full = st["ms_total"]
print("full frame %.2f ms: k_rtao_rays %.2f, rest of RTAO %.2f, tubes %.2f" % (full, st["ms_rtao_rays"], st["ms_rtao"] - st["ms_rtao_rays"], st["ms_trace"]), flush=True)
for world in args.world:
    for tile in args.tile:
        ts, parts = [], []
        for r in range(world):
            ctx.set_tile_shard(r, world, tile)
            for _ in range(3):
                _, st = ctx.render_tubes(sc, cam, 0, out=frame)
            ts.append(st["ms_total"])
            parts.append((st["ms_rtao_rays"], st["ms_rtao"] - st["ms_rtao_rays"], st["ms_trace"], st["rays_ao"] / 1e6))
        ts = np.array(ts)
        pm = np.mean(parts, axis=0)
        print("   mean per rank: k_rtao_rays %.2f ms, rest of RTAO pass %.2f ms, tube pass %.2f ms, AO rays %.2f M (min %.2f max %.2f)" %
              (pm[0], pm[1], pm[2], pm[3], min(p[3] for p in parts), max(p[3] for p in parts)), flush=True)
        print("world %d tile %3d: rank ms min %.2f mean %.2f max %.2f  max/mean %.3f  ideal %.2f  speed-up bound %.2fx" %
              (world, tile, ts.min(), ts.mean(), ts.max(), ts.max() / ts.mean(), full / world, full / ts.max()), flush=True)
