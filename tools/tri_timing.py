"""Timing of the triangle-tube mode of the AO passes (b200_rtao_geometry = triangles) against the analytic capsules on the helix set of
config 2 (100 k segments, 400 polylines) at 1920x1080 with 16-spp RTAO, and on N random 2-point polylines (config 3's segment soup as
polylines): mesh size, mesh + BVH build time, frame time, AO kernel time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linevis_b200 as lv
from linevis_b200 import scenes


def run(name, d, W, H, spp):
    cam = lv.make_camera(W, H)
    for geom in ("capsules", "triangles"):
        ctx = lv.Context(0)
        ctx.set_transfer_function(scenes.standard_transfer_function())
        ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": spp, "ambient_occlusion_iterations": 1,
                              "num_samples_per_frame": 1, "num_accumulated_frames": 1, "b200_rtao_geometry": geom})
        sc = ctx.create_scene(d["pos"], d["attr"], d["seg"], scenes.LINE_WIDTH)
        sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"])
        t0 = time.time()
        img, st = ctx.render_tubes(sc, cam, 0)          # first frame: builds the tube mesh + its BVH in triangle mode
        first = time.time() - t0
        ts = []
        for _ in range(4):
            img, st = ctx.render_tubes(sc, cam, 0)
            ts.append((st["ms_total"], st["ms_rtao_rays"]))
        tot, rays = min(a for a, _ in ts), min(b for _, b in ts)
        n = st["rays_primary"] + st["rays_ao"]
        print("%-10s %-9s segs %8d  frame %7.2f ms  AO stream %7.2f ms  %7.0f Mrays/s  AO rays %9d  T/ray %5.1f I/ray %5.1f  first frame %.2f s" % (
            name, geom, d["seg"].shape[0], tot, rays, n / tot / 1e3, st["rays_ao"], st["ao_traversal_steps"] / max(st["rays_ao"], 1),
            st["ao_intersections"] / max(st["rays_ao"], 1), first), flush=True)
        sc.close(); ctx.close()


run("helix100k", scenes.helix_polylines(), 1920, 1080, 16)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pos, attr, seg = scenes.random_segments(n, seed=2002)
d = scenes.polylines_with_frames(pos, attr, np.arange(n + 1) * 2)
run("random%dk" % (n // 1000), d, 1920, 1080, 16)
