"""GPU parity of the triangle-tube mode of the AO passes (b200_rtao_geometry = triangles) against the oracle's
restatement of the reference's tube mesh + barycentric fetch, through the C ABI.  The same statements hold on the host under the
SIMT emulator (tests/test_emu_simt.py); here they run at larger sizes on the real kernels."""
import numpy as np
import pytest

import linevis_b200 as lv
from linevis_b200 import scenes
from oracle import lvo

pytestmark = pytest.mark.gpu


def _scene(ctx, oracle, n_lines, n_pts, width):
    d = scenes.helix_polylines(n_lines, n_pts)
    sc = ctx.create_scene(d["pos"], d["attr"], d["seg"], width)
    sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"])
    osc = oracle.scene(d["pos"], d["attr"], d["seg"], width)
    osc.set_lines(d["tangent"], d["normal"])
    return d, sc, osc


@pytest.fixture()
def tri_ctx(ctx):
    ctx.set_new_settings({"b200_rtao_geometry": "triangles", "ambient_occlusion_mode": "RTAO (Screen Space)", "ambient_occlusion_gamma": 1.0,
                          "depth_cue_strength": 0.0, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    yield ctx
    ctx.set_new_settings({"b200_rtao_geometry": "capsules", "ambient_occlusion_mode": "RTAO (Screen Space)", "ambient_occlusion_strength": 0.0,
                          "tube_num_subdivisions": 6, "ambient_occlusion_radius": 0.1, "use_jittered_primary_rays": True})


@pytest.mark.parametrize("use_distance,jitter,n_sub", [(True, True, 6), (False, False, 6), (True, True, 8)])
def test_triangle_tube_rtao_bit_exact(tri_ctx, oracle, use_distance, jitter, n_sub):
    ctx = tri_ctx
    width = 0.006
    d, sc, osc = _scene(ctx, oracle, 40, 61, width)
    tm = lvo.TubeMesh(oracle, d["pos"], d["line_offsets"], width, n_sub)
    cam = lv.make_camera(200, 120)
    ctx.set_new_settings({"tube_num_subdivisions": n_sub, "ambient_occlusion_samples_per_frame": 8, "ambient_occlusion_distance_based": use_distance,
                          "use_jittered_primary_rays": jitter, "ambient_occlusion_radius": 0.1})
    ao, st = ctx.render_rtao(sc, cam, 0)
    ao2, _ = ctx.render_rtao(sc, cam, 1, out=ao.copy())
    opts = lvo.default_options(ao_strength=1.0, ao_spp=8, ao_use_distance=int(use_distance), ao_jitter_primary=int(jitter), tube_num_subdivisions=n_sub)
    ref, ost = tm.render_rtao(cam, opts, 0)
    ref2, _ = tm.render_rtao(cam, opts, 1, ao=ref.copy())
    assert st["pixels_hit"] == ost["pixels_hit"] > 500 and st["rays_ao"] == ost["rays_ao"]
    assert np.array_equal(ao.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(ao2.view(np.uint32), ref2.view(np.uint32))


def test_triangle_tube_frame_and_prebaker_bit_exact(tri_ctx, oracle):
    ctx = tri_ctx
    width = 0.006
    d, sc, osc = _scene(ctx, oracle, 40, 61, width)
    tm = lvo.TubeMesh(oracle, d["pos"], d["line_offsets"], width, 6)
    cam = lv.make_camera(200, 120)
    tf = scenes.standard_transfer_function(opacity=(0.4, 1.0))
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"tube_num_subdivisions": 6, "ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 4,
                          "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True, "ambient_occlusion_radius": 0.1})
    img, _ = ctx.render_tubes(sc, cam, 0)
    opts = lvo.default_options(ao_strength=1.0, ao_spp=4)
    ao, _ = tm.render_rtao(cam, opts, 0)
    ref, _ = osc.render_tubes(cam, opts, tf, ao_tex=ao)
    assert np.array_equal(img.view(np.uint32), ref.view(np.uint32))
    ctx.set_new_settings({"ambient_occlusion_mode": "RTAO (Prebaker)", "b200_prebaker_iterations": 2, "b200_prebaker_samples_per_frame": 4,
                          "b200_prebaker_subdivisions": 8, "b200_prebaker_param_segment_length": 0.01, "b200_prebaker_radius": 0.05,
                          "b200_prebaker_distance_based": True})
    bw, sl = oracle.ao_parametrize(d["pos"], d["line_offsets"], 0.01)
    ref_f = None
    for it in range(2):
        st = sc.ao_bake(1)
        ref_f, ost = osc.ao_bake_iteration(sl, it, factors=ref_f, radius=0.05, n_subdiv=8, spp=4, tube_mesh=tm)
        assert st["rays_ao"] == ost["rays"]
        assert np.array_equal(sc.ao_read()["factors"].reshape(-1).view(np.uint32), ref_f.view(np.uint32)), it


@pytest.mark.parametrize("n_sub,ao,jitter", [(6, False, False), (8, True, False), (6, True, True)])
def test_triangle_geometry_mode_of_the_tube_pass(tri_ctx, oracle, n_sub, ao, jitter):
    """geometry_mode = "Triangle Mesh" (RayTracingGeometryMode::TRIANGLE_MESH): the tube pass traces the reference's triangulated tubes and
    shades with ClosestHitTubeTriangles (barycentric normal / tangent / attribute, cap flag) -- frame bit-exact against the oracle's
    restatement, with the RTAO texture traced against the same mesh, with jittered multi-sample frames and accumulation."""
    d, sc, osc, width = _scene(tri_ctx, oracle, 40, 61, 0.006) + (0.006,)
    tm = lvo.TubeMesh(oracle, d["pos"], d["line_offsets"], width, n_sub)
    cam = lv.make_camera(200, 120)
    tf = scenes.standard_transfer_function(opacity=(0.4, 1.0))
    ctx = tri_ctx
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"geometry_mode": "Triangle Mesh", "tube_num_subdivisions": n_sub, "b200_rtao_geometry": "triangles",
                          "ambient_occlusion_strength": 1.0 if ao else 0.0, "ambient_occlusion_samples_per_frame": 4, "ambient_occlusion_radius": 0.2,
                          "num_samples_per_frame": 2 if jitter else 1, "num_accumulated_frames": 2 if jitter else 1})
    try:
        assert ctx.get_option("use_analytic_intersections") == "false"
        opts = lvo.default_options(ao_strength=1.0 if ao else 0.0, ao_spp=4, ao_radius=0.2, tube_num_subdivisions=n_sub,
                                   num_samples_per_frame=2 if jitter else 1, use_jittered_rays=int(jitter))
        img, ref = None, None
        for frame in range(2 if jitter else 1):
            img, st = ctx.render_tubes(sc, cam, frame, out=img)
            rao = tm.render_rtao(cam, opts, frame, ao=rao if frame else None)[0] if ao else None
            ref, ost = tm.render_tubes(d["attr"], cam, opts, tf, ao_tex=rao, frame_number=frame, rgba=ref)
            assert st["pixels_hit"] > 0 and np.isfinite(img).all()
            assert np.array_equal(img.view(np.uint32), ref.view(np.uint32)), frame
        analytic, _ = osc.render_tubes(cam, lvo.default_options(), tf)
        assert not np.array_equal(analytic, ref)            # it really is the other geometry
        ctx.set_new_settings({"ambient_occlusion_mode": "RTAO (Prebaker)", "ambient_occlusion_strength": 1.0})
        with pytest.raises(lv.LineVisError):
            ctx.render_tubes(sc, cam, 0)
        with pytest.raises(lv.LineVisError):
            ctx.set_option("geometry_mode", "Linear Swept Spheres")
    finally:
        ctx.set_new_settings({"geometry_mode": "AABBs (analytic)", "b200_rtao_geometry": "capsules", "ambient_occlusion_mode": "RTAO (Screen Space)",
                              "ambient_occlusion_strength": 0.0, "ambient_occlusion_radius": 0.1, "tube_num_subdivisions": 6,
                              "num_samples_per_frame": 1, "num_accumulated_frames": 1})

