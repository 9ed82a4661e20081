"""max|delta| of the -fmad=true build against the strict-IEEE build and against the oracle (SURVEY 7 hard part 5): tubes + RTAO and
PPLL frames of a 100 k-segment random scene at 480x270.  Run on the GPU box after `python tools/build_variant.py fmad`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linevis_b200 as lv
from oracle import lvo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pos, attr, seg = lv.scenes.random_segments(n_seg=100_000, seed=2002)
cam = lv.make_camera(480, 270)
tf = lv.scenes.standard_transfer_function()
tfp = lv.scenes.standard_transfer_function(opacity=(0.1, 0.6))
out = {}
for name, path in (("strict", None), ("fmad", os.path.join(ROOT, "build", "liblinevis_b200_fmad.so"))):
    ctx = lv.Context(0, lib_path=path)
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 8, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    img, st = ctx.render_tubes(sc, cam)
    ctx.set_transfer_function(tfp)
    pp, st2 = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="priority_queue")
    out[name] = (np.array(img), np.array(pp), st["rays_ao"], st2["frags_sorted"])
    sc.close(); ctx.close()
o = lvo.Oracle("own"); o.set_num_threads()
osc = o.scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
opts = lvo.default_options(ao_strength=1.0, ao_spp=8)
ao, _ = osc.render_rtao(cam, opts)
ref, _ = osc.render_tubes(cam, opts, tf, ao_tex=ao)
for name in out:
    img, pp, ra, fs = out[name]
    d = np.abs(img - ref)
    print("%-6s tubes+RTAO vs oracle: max|d| %.3g  mean|d| %.3g  pixels>1e-3: %d of %d | AO rays %d  PPLL frags %d" %
          (name, d.max(), d.mean(), int((d.max(axis=-1) > 1e-3).sum()), d.shape[0] * d.shape[1], ra, fs))
d = np.abs(out["fmad"][0] - out["strict"][0]); dp = np.abs(out["fmad"][1] - out["strict"][1])
print("fmad vs strict: tubes max|d| %.3g (pixels>1e-3: %d)  ppll max|d| %.3g (pixels>1e-3: %d)" %
      (d.max(), int((d.max(axis=-1) > 1e-3).sum()), np.nanmax(dp), int((np.nan_to_num(dp).max(axis=-1) > 1e-3).sum())))
