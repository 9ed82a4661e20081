"""GPU parity of the object-space AO prebaker ("RTAO (Prebaker)") against the oracle, through the C ABI.

Bars: baked ambientOcclusionFactors bit-exact (every iteration of the running mean, distance-based and any-hit), ray
counts exact, the frame shaded with the prebaked lookup bit-exact, PPLL fragments shaded with it bit-exact as multisets."""
import numpy as np
import pytest

import linevis_b200 as lv
from linevis_b200 import scenes
from oracle import lvo

pytestmark = pytest.mark.gpu


def _pair(ctx, oracle, d, width):
    sc = ctx.create_scene(d["pos"], d["attr"], d["seg"], width)
    sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"])
    osc = oracle.scene(d["pos"], d["attr"], d["seg"], width)
    osc.set_lines(d["tangent"], d["normal"])
    return sc, osc


BAKE_KEYS = {"b200_prebaker_iterations": 3, "b200_prebaker_samples_per_frame": 4, "b200_prebaker_subdivisions": 8,
             "b200_prebaker_param_segment_length": 0.01, "b200_prebaker_radius": 0.05, "b200_prebaker_distance_based": True}


@pytest.fixture()
def bake_ctx(ctx):
    ctx.set_new_settings(BAKE_KEYS)
    ctx.set_new_settings({"ambient_occlusion_mode": "RTAO (Prebaker)", "ambient_occlusion_strength": 0.0, "ambient_occlusion_gamma": 1.0,
                          "depth_cue_strength": 0.0, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    yield ctx
    ctx.set_new_settings({"ambient_occlusion_mode": "RTAO (Screen Space)", "ambient_occlusion_strength": 0.0, "ambient_occlusion_gamma": 1.0})


@pytest.mark.parametrize("distance_based", [True, False])
def test_baked_factors_bit_exact_every_iteration(bake_ctx, oracle, distance_based):
    ctx = bake_ctx
    ctx.set_option("b200_prebaker_distance_based", distance_based)
    d = scenes.helix_polylines(40, 61)
    width = 0.006
    sc, osc = _pair(ctx, oracle, d, width)
    bw, sl = oracle.ao_parametrize(d["pos"], d["line_offsets"], 0.01)
    ref = None
    for it in range(3):
        st = sc.ao_bake(1)
        got = sc.ao_read()
        ref, ost = osc.ao_bake_iteration(sl, it, factors=ref, radius=0.05, n_subdiv=8, spp=4, use_distance=distance_based)
        assert got["iterations_done"] == it + 1 and got["n_param"] == len(sl) and got["n_subdiv"] == 8
        assert np.array_equal(got["sampling_locations"], sl) and np.array_equal(got["blending_weights"], bw)
        assert st["rays_ao"] == ost["rays"] == len(sl) * 8 * 4
        assert np.array_equal(got["factors"].reshape(-1).view(np.uint32), ref.view(np.uint32)), "iteration %d" % it
    assert ref.min() < 0.9 and ref.max() <= 1.0 and 0.2 < ref.mean() < 0.999   # the helix does occlude itself
    assert sc.ao_bake(0)["rays_ao"] == 0                     # all b200_prebaker_iterations done: nothing left to bake
    sc.ao_bake_reset()
    assert sc.ao_bake(0)["rays_ao"] == 3 * len(sl) * 8 * 4   # immediate mode: all iterations in one call
    assert np.array_equal(sc.ao_read()["factors"].reshape(-1).view(np.uint32), ref.view(np.uint32))


def test_frame_with_prebaked_ao_bit_exact(bake_ctx, oracle):
    ctx = bake_ctx
    d = scenes.helix_polylines(40, 61)
    width = 0.006
    sc, osc = _pair(ctx, oracle, d, width)
    bw, sl = oracle.ao_parametrize(d["pos"], d["line_offsets"], 0.01)
    ref_f = None
    for it in range(3):
        ref_f, _ = osc.ao_bake_iteration(sl, it, factors=ref_f, radius=0.05, n_subdiv=8, spp=4)
    osc.set_static_ao(ref_f, 8, bw)
    tf = scenes.standard_transfer_function(opacity=(0.4, 1.0))
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 0.9, "ambient_occlusion_gamma": 1.5})
    cam = lv.make_camera(200, 120)
    # ITERATIVE_UPDATE: each of the first three frames runs one baking iteration before shading
    for it in range(3):
        img, st = ctx.render_tubes(sc, cam)
        assert sc.ao_read()["iterations_done"] == it + 1
    img2, st2 = ctx.render_tubes(sc, cam)
    assert st2["rays_ao"] == 0 and sc.ao_read()["iterations_done"] == 3
    opts = lvo.default_options(ao_strength=0.9, ao_gamma=1.5, use_static_ao=1)
    ref, _ = osc.render_tubes(cam, opts, tf)
    assert np.array_equal(img.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(img2.view(np.uint32), ref.view(np.uint32))
    plain, _ = osc.render_tubes(cam, lvo.default_options(), tf)
    assert np.abs(plain - ref).max() > 0.05                  # the AO term is visible


def test_ppll_with_prebaked_ao_bit_exact(bake_ctx, oracle):
    ctx = bake_ctx
    d = scenes.helix_polylines(30, 51)
    width = 0.008
    sc, osc = _pair(ctx, oracle, d, width)
    bw, sl = oracle.ao_parametrize(d["pos"], d["line_offsets"], 0.01)
    sc.ao_bake(0)
    got = sc.ao_read()
    osc.set_static_ao(got["factors"].reshape(-1), 8, bw)
    tf = scenes.standard_transfer_function(opacity=(0.2, 0.7))
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 1.0})
    cam = lv.make_camera(96, 64)
    img, st = ctx.render_ppll(sc, cam, max_frags=64, sort_mode="priority_queue")
    opts = lvo.default_options(ao_strength=1.0, use_static_ao=1)
    g = osc.ppll_gather(cam, opts, tf)
    assert st["frags_generated"] == g["counter"] > 0
    mine = ctx.ppll_read()
    a = lvo.per_pixel_lists(mine["heads"], mine["nodes"], cam, opts, oracle)
    b = lvo.per_pixel_lists(g["heads"], g["nodes"], cam, opts, oracle)
    assert a == b
    refp, _ = lvo.ppll_resolve(oracle, cam, opts, g["heads"], g["nodes"], 64, 0, canonical=True)
    nan = np.isnan(refp)                                     # all-alpha-0 lists resolve to 0/0 in the reference and here alike
    assert np.array_equal(np.isnan(img), nan)
    assert np.array_equal(img[~nan].view(np.uint32), refp[~nan].view(np.uint32))


def test_prebaker_errors(bake_ctx):
    ctx = bake_ctx
    pos, attr, seg = scenes.helix_lines(6, 12)
    sc = ctx.create_scene(pos, attr, seg, 0.01)
    ctx.set_transfer_function(scenes.standard_transfer_function())
    ctx.set_option("ambient_occlusion_strength", 1.0)
    with pytest.raises(lv.LineVisError) as e:
        ctx.render_tubes(sc, lv.make_camera(32, 32))
    assert e.value.code == -6 and "lv_scene_set_lines" in str(e.value)
    with pytest.raises(lv.LineVisError):
        sc.ao_bake(1)
    d = scenes.helix_polylines(6, 12)
    with pytest.raises(lv.LineVisError):   # wrong point count
        sc.set_lines(d["pos"][:-1], d["tangent"][:-1], d["normal"][:-1], d["line_offsets"])
    with pytest.raises(lv.LineVisError):
        ctx.set_option("ambient_occlusion_mode", "SSAO")
    # strength 0: the prebaker mode costs nothing and needs nothing
    ctx.set_option("ambient_occlusion_strength", 0.0)
    img, st = ctx.render_tubes(sc, lv.make_camera(32, 32))
    assert st["rays_ao"] == 0 and np.isfinite(img).all()
