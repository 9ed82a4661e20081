"""Short driver for ncu captures: a couple of frames of each hot path at a named workload (no timing, no oracle)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import linevis_b200 as lv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--ppll-workload", default="config4")
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--skip-tubes", action="store_true")
ap.add_argument("--skip-ppll", action="store_true")
ap.add_argument("--opt", type=str, nargs="*", default=[], help="extra key=value options for the contexts")
args = ap.parse_args()

import torch
dev = torch.device("cuda", 0)
if not args.skip_tubes:
    wl = bench.WORKLOADS[args.workload]
    pos, attr, seg = bench.generate(wl["gen"], dev)
    ctx = lv.Context(0)
    ctx.set_transfer_function(lv.scenes.standard_transfer_function())
    ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": wl["ao_spp"],
                          "ambient_occlusion_iterations": 1, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    for kv in args.opt:
        ctx.set_option(*kv.split("=", 1))
    sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    cam = lv.make_camera(wl["W"], wl["H"])
    frame = torch.zeros((wl["H"], wl["W"], 4), dtype=torch.float32, device=dev)
    for i in range(args.frames):
        _, st = ctx.render_tubes(sc, cam, 0, out=frame)
    print("tubes", {k: v for k, v in st.items() if v})
if not args.skip_ppll:
    pw = bench.PPLL_WORKLOADS[args.ppll_workload]
    pos, attr, seg = bench.generate(pw["gen"], dev)
    ctx = lv.Context(0)
    ctx.set_transfer_function(lv.scenes.standard_transfer_function(opacity=(0.1, 0.6)))
    ctx.set_option("ambient_occlusion_strength", 0.0)
    if "avg_depth" in pw:
        ctx.set_option("b200_expected_avg_depth_complexity", pw["avg_depth"])
    for kv in args.opt:
        ctx.set_option(*kv.split("=", 1))
    sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    cam = lv.make_camera(pw["W"], pw["H"])
    frame = torch.zeros((pw["H"], pw["W"], 4), dtype=torch.float32, device=dev)
    for i in range(args.frames):
        _, st = ctx.render_ppll(sc, cam, pw["max_frags"], "priority_queue", 0, out=frame)
    print("ppll", {k: v for k, v in st.items() if v})
