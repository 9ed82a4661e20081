#!/bin/bash
mkdir -p gpurun_out
timeout 120 python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, linevis_b200 as lv
from linevis_b200 import scenes
ctx = lv.Context(0)
for name, data, width in (("random", scenes.random_segments(20000, 0.02, seed=7), 0.003), ("helix", scenes.helix_lines(60, 101), 0.004)):
    sc = ctx.create_scene(*data, width)
    cam = lv.make_camera(240, 160)
    for dist in (True, False):
        ctx.set_new_settings({"ambient_occlusion_samples_per_frame": 8, "ambient_occlusion_radius": 0.3, "ambient_occlusion_distance_based": dist})
        ctx.set_option("b200_ao_queue", False); a, sa = ctx.render_rtao(sc, cam, 0)
        ctx.set_option("b200_ao_queue", True); b, sb = ctx.render_rtao(sc, cam, 0)
        print(name, dist, "bit-exact:", np.array_equal(a.view(np.uint32), b.view(np.uint32)), "T", sa["ao_traversal_steps"], sb["ao_traversal_steps"], "I", sa["ao_intersections"], sb["ao_intersections"], "rays", sa["rays_ao"], sb["rays_ao"], flush=True)
PY
timeout 600 python tools/sweep.py --leaf 1 --combo 9:12:24:12 8:12:24:12 7:12:24:12 9:12:24:8 9:12:24:16 9:12:24:20 9:12:28:12 9:12:20:12 9:8:24:12 --opt b200_ao_queue=true > gpurun_out/e5_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/e5_sweep.log
