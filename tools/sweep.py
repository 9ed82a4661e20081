"""Parameter sweep on the GPU box: BVH leaf size x AO refill threshold for the tube+RTAO frame, leaf size for PPLL."""
import argparse, os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import linevis_b200 as lv

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--leaf", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--refill", type=int, nargs="+", default=[16, 24, 28, 32])
ap.add_argument("--ppll-workload", default="none")
ap.add_argument("--vote", type=int, nargs="+", default=[12])
ap.add_argument("--minb", type=int, nargs="+", default=[10])
ap.add_argument("--stack", type=int, nargs="+", default=[0])
ap.add_argument("--opt", type=str, nargs="*", default=[], help="extra key=value options applied to every configuration")
ap.add_argument("--combo", type=str, nargs="*", default=[], help="explicit minb:stack:refill:vote combinations instead of the product")
args = ap.parse_args()
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS[args.workload]
pos, attr, seg = bench.generate(wl["gen"], dev)
cam = lv.make_camera(wl["W"], wl["H"])
frame = torch.zeros((wl["H"], wl["W"], 4), dtype=torch.float32, device=dev)
for leaf in args.leaf:
    ctx = lv.Context(0)
    ctx.set_transfer_function(lv.scenes.standard_transfer_function())
    ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": wl["ao_spp"], "ambient_occlusion_iterations": 1,
                          "num_samples_per_frame": 1, "num_accumulated_frames": 1, "b200_bvh_leaf_size": leaf})
    for kv in args.opt:
        ctx.set_option(*kv.split("=", 1))
    sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    # e.g. --opt b200_ao_qnodes=true (quantised nodes), b200_ao_queue=false (leaf-vote kernel), b200_rtao_geometry=triangles (needs set_lines)
    combos = [tuple(int(v) for v in (c.split(":")[2], c.split(":")[3], c.split(":")[0], c.split(":")[1])) for c in args.combo] \
        or list(itertools.product(args.refill, args.vote, args.minb, args.stack))
    for refill, vote, minb, stack in combos:
        ctx.set_option("b200_ao_min_blocks", minb)
        ctx.set_option("b200_ao_stack", stack)
        ctx.set_option("b200_ao_refill_below", refill)
        ctx.set_option("b200_ao_leaf_vote", vote)
        ts = []
        for i in range(4):
            _, st = ctx.render_tubes(sc, cam, 0, out=frame)
            ts.append((st["ms_rtao_rays"], st["ms_total"]))
        k, t = np.min([a for a, _ in ts[1:]]), np.min([b for _, b in ts[1:]])
        rays = st["rays_primary"] + st["rays_ao"]
        by = 64 * st["ao_traversal_steps"] + 32 * st["ao_intersections"] + 4 * st["rays_ao"]
        print("leaf %d refill %2d vote %2d minb %2d stack %d: k_rtao_rays %.2f ms  frame %.2f ms  %.0f Mrays/s  T/ray %.1f I/ray %.1f  algGB/s %.0f  build %.1f ms" %
              (leaf, refill, vote, minb, stack, k, t, rays / t / 1e3, st["ao_traversal_steps"] / st["rays_ao"], st["ao_intersections"] / st["rays_ao"], by / k / 1e6, sc.info()["build_ms"]), flush=True)
    sc.close(); ctx.close()
if args.ppll_workload != "none":
    pw = bench.PPLL_WORKLOADS[args.ppll_workload]
    pos, attr, seg = bench.generate(pw["gen"], dev)
    cam = lv.make_camera(pw["W"], pw["H"])
    frame = torch.zeros((pw["H"], pw["W"], 4), dtype=torch.float32, device=dev)
    for leaf in args.leaf:
        ctx = lv.Context(0)
        ctx.set_transfer_function(lv.scenes.standard_transfer_function(opacity=(0.1, 0.6)))
        ctx.set_option("b200_bvh_leaf_size", leaf)
        sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
        for i in range(3):
            _, st = ctx.render_ppll(sc, cam, pw["max_frags"], "priority_queue", 0, out=frame)
        print("ppll leaf %d: gather %.2f ms resolve %.3f ms frags %d T/ray %.1f I/ray %.1f" %
              (leaf, st["ms_gather"], st["ms_resolve"], st["frags_sorted"], st["traversal_steps"] / st["rays_primary"], st["intersections"] / st["rays_primary"]), flush=True)
        sc.close(); ctx.close()
