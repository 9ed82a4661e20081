"""A/B of AO ray-stream variants on one scene: k_rtao_rays time, frame time and a bit-exact comparison of the frames.

    python tools/ao_ab.py --workload config5 --variant "" --variant b200_ao_packed=false --variant b200_ao_min_blocks=7
Each --variant is a comma-separated list of key=value options applied on top of the defaults (the empty string = defaults)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import linevis_b200 as lv

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--variant", action="append", default=[])
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS[args.workload]
pos, attr, seg = bench.generate(wl["gen"], dev)
cam = lv.make_camera(wl["W"], wl["H"])
frame = torch.zeros((wl["H"], wl["W"], 4), dtype=torch.float32, device=dev)
ref = None
for variant in (args.variant or [""]):
    ctx = lv.Context(0)
    ctx.set_transfer_function(lv.scenes.standard_transfer_function())
    ctx.set_new_settings({"depth_cue_strength": 0.0, "ambient_occlusion_strength": 1.0, "ambient_occlusion_gamma": 1.0,
                          "ambient_occlusion_samples_per_frame": wl["ao_spp"], "ambient_occlusion_iterations": 1, "ambient_occlusion_radius": 0.1,
                          "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True,
                          "num_samples_per_frame": 1, "num_accumulated_frames": 1, "use_deterministic_sampling": False})
    for kv in filter(None, variant.split(",")):
        ctx.set_option(*kv.split("=", 1))
    sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    ts = []
    for i in range(args.reps):
        _, st = ctx.render_tubes(sc, cam, 0, out=frame, stats=True)
        ts.append((st["ms_rtao_rays"], st["ms_total"]))
    torch.cuda.synchronize()
    img = frame.clone()
    same = "reference" if ref is None else ("bit-exact" if torch.equal(img.view(torch.int32), ref.view(torch.int32)) else "DIFFERENT (max |d| %.3g)" % float((img - ref).abs().max()))
    if ref is None:
        ref = img
    k, t = np.min([a for a, _ in ts[1:]]), np.min([b for _, b in ts[1:]])
    print("%-60s k_rtao_rays %.2f ms  frame %.2f ms  T/ray %.2f I/ray %.2f  rays_ao %d  build %.1f ms  %s" %
          (variant or "(defaults)", k, t, st["ao_traversal_steps"] / max(1, st["rays_ao"]), st["ao_intersections"] / max(1, st["rays_ao"]), st["rays_ao"],
           sc.info()["build_ms"], same), flush=True)
    sc.close(); ctx.close()
