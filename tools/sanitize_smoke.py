"""Small scenes through every default kernel (and the main variants) for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(SURVEY 5 "New").  No torch: numpy + the C ABI only, so the sanitizer instruments nothing but this library.
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py [fast]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linevis_b200 as lv

fast = len(sys.argv) > 1 and sys.argv[1] == "fast"
cam = lv.make_camera(64, 48)
tf = lv.scenes.standard_transfer_function(opacity=(0.3, 0.8))
variants = [{}] if fast else [{}, {"b200_ao_queue": False, "b200_bvh_leaf_size": 4}, {"b200_ao_qnodes": True, "b200_ao_wide": False, "b200_ao_raybuf": False},
                              {"b200_ao_wide": False, "b200_ao_raybuf": False, "b200_tube_prepass": False, "b200_ppll_gather_mode": "raycast", "b200_ppll_reg_sort": False},
                              {"b200_ao_wide_top": 85}, {"b200_ao_packed": False}, {"b200_ao_tq_bits": 4}, {"b200_bvh_builder": "ploc"}, {"b200_frame_format": "rgba8", "b200_async_delivery": True},
                              {"b200_ppll_gather_mode": "raster_contiguous"}, {"b200_ppll_binned_resolve": True}, {"depth_cue_strength": 0.8},
                              {"geometry_mode": "Triangle Mesh", "b200_rtao_geometry": "triangles"}]
d = lv.scenes.helix_polylines(12, 41)
pos, attr, seg = d["pos"], d["attr"], d["seg"]
for v in variants:
    ctx = lv.Context(0)
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 4, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    ctx.set_new_settings(v)
    sc = ctx.create_scene(pos, attr, seg, 0.006)
    sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"])
    rgba8 = v.get("b200_frame_format") == "rgba8"
    out = (lambda: np.zeros((cam.height, cam.width), np.uint32)) if rgba8 else (lambda: None)
    img, st = ctx.render_tubes(sc, cam, out=out())
    pp, st2 = ctx.render_ppll(sc, cam, max_frags=128, sort_mode="priority_queue", out=out())
    ctx.set_tile_shard(1, 2, 16)
    img2, _ = ctx.render_tubes(sc, cam, out=out())
    pp2, _ = ctx.render_ppll(sc, cam, max_frags=128, sort_mode="bitonic", out=out())
    if not v:
        # AO-sample-batch stages (one rank plays both sample batches) and an explicit tile-owner map
        costs = ctx.tile_costs(cam.width, cam.height)
        ctx.set_tile_owners(cam.width, cam.height, (np.arange(costs.size) % 2).astype(np.uint8))
        ptr, n = ctx.sao_primary(sc, cam, 0)
        occ = ctx.frame_alloc(cam.width, cam.height)       # device scratch: 2 parts x n x 2 floats fit a W x H x 4 frame
        for part in range(2):
            ctx.sao_trace(sc, cam, 0, ptr, n, 2 * part, 2, occ + part * n * 2 * 4)
        img3, _ = ctx.sao_finish(sc, cam, 0, occ, 2, np.zeros((cam.height, cam.width, 4), np.float32))
        ctx.frame_free(occ)
    ctx.synchronize()
    print(v, "rays", st["rays_primary"] + st["rays_ao"], "frags", st2["frags_sorted"], "finite", bool(rgba8 or np.isfinite(img).all()), flush=True)
    sc.close(); ctx.close()
print("sanitize_smoke done")
