"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bars: closest-hit records bit-exact (t bits, segment index, hit kind); AO image and shaded RGBA within
1e-3 per channel (in practice bit-exact, see DESIGN.md "Float conventions"); PPLL fragment counter, per-pixel list
lengths and the per-pixel multiset of (depth bits, colour) bit-exact."""
import numpy as np
import pytest

import linevis_b200 as lv
from linevis_b200 import scenes
from oracle import lvo

pytestmark = pytest.mark.gpu
TOL = 1e-3  # per-channel float tolerance of BASELINE.json's north_star


def _scene_pair(ctx, oracle, data, width):
    pos, attr, seg = data
    return ctx.create_scene(pos, attr, seg, width), oracle.scene(pos, attr, seg, width)


DATASETS = {
    "helix": lambda: (scenes.helix_lines(60, 101), 0.004),
    "random": lambda: (scenes.random_segments(20000, 0.02, seed=7), 0.003),
    "single": lambda: ((np.array([[-0.2, 0.0, 0.0], [0.2, 0.05, 0.0]], np.float32), np.array([0.1, 0.9], np.float32),
                        np.array([[0, 1]], np.uint32)), 0.05),
}


@pytest.mark.parametrize("name", list(DATASETS))
def test_primary_hits_bit_exact(ctx, oracle, name):
    data, width = DATASETS[name]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(200, 120)
    hits, st = ctx.trace_primary(sc, cam)
    ref, ost = osc.trace_primary(cam)
    assert np.array_equal(hits["prim"], ref["prim"])
    assert np.array_equal(hits["kind"], ref["kind"])
    assert np.array_equal(hits["t"].view(np.uint32), ref["t"].view(np.uint32))
    assert st["rays_primary"] == 200 * 120
    assert st["pixels_hit"] == int((ref["prim"] != 0xFFFFFFFF).sum()) > 0


def test_primary_empty_scene(ctx):
    sc = ctx.create_scene(np.zeros((0, 3), np.float32), np.zeros(0, np.float32), np.zeros((0, 2), np.uint32), 0.01)
    cam = lv.make_camera(64, 48)
    hits, st = ctx.trace_primary(sc, cam)
    assert (hits["prim"] == 0xFFFFFFFF).all() and st["pixels_hit"] == 0


@pytest.mark.parametrize("leaf", [1, 2, 4, 8])
def test_bvh_leaf_sizes_same_hits(ctx, oracle, leaf):
    data, width = DATASETS["random"]()
    ctx.set_option("b200_bvh_leaf_size", leaf)
    try:
        sc, osc = _scene_pair(ctx, oracle, data, width)
    finally:
        ctx.set_option("b200_bvh_leaf_size", 1)
    cam = lv.make_camera(160, 96)
    hits, _ = ctx.trace_primary(sc, cam)
    ref, _ = osc.trace_primary(cam)
    assert np.array_equal(hits["t"].view(np.uint32), ref["t"].view(np.uint32)) and np.array_equal(hits["prim"], ref["prim"])


def test_bvh_encloses_all_segments(ctx):
    (pos, attr, seg), width = DATASETS["random"]()
    sc = ctx.create_scene(pos, attr, seg, width)
    nodes = sc.bvh_nodes()
    info = sc.info()
    assert info["n_nodes"] == len(nodes)
    # walk from the root; every segment must be referenced exactly once and lie inside its leaf box
    seen = 0
    stack = [0]
    r = width * 0.5
    lo = np.minimum(pos[seg[:, 0]], pos[seg[:, 1]]) - r
    hi = np.maximum(pos[seg[:, 0]], pos[seg[:, 1]]) + r
    glo, ghi = lo.min(0), hi.max(0)
    assert np.allclose(info["aabb"][:3], glo, atol=1e-6) and np.allclose(info["aabb"][3:], ghi, atol=1e-6)
    while stack:
        n = nodes[stack.pop()]
        for side in ("l", "r"):
            mn, mx, ref, cnt = n[side + "min"], n[side + "max"], int(n[side + "ref"]), int(n[side + "count"])
            if not np.isfinite(mn).all():
                continue                                   # absent child: point box at +inf
            assert (mn <= mx).all() and (mn >= glo - 1e-6).all() and (mx <= ghi + 1e-6).all()
            if cnt:
                assert ref >> 31 == 1 and ((ref >> 27) & 15) + 1 == cnt
                first = ref & 0x07FFFFFF
                seen += cnt
            else:
                stack.append(ref)
    assert seen == seg.shape[0]


@pytest.mark.parametrize("jitter,use_distance,spp", [(False, True, 8), (True, True, 4), (False, False, 16), (True, True, 64), (False, True, 5)])
def test_rtao_parity(ctx, oracle, jitter, use_distance, spp):
    data, width = DATASETS["helix"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(120, 80)
    ctx.set_new_settings({"ambient_occlusion_samples_per_frame": spp, "ambient_occlusion_distance_based": use_distance,
                          "use_jittered_primary_rays": jitter, "ambient_occlusion_radius": 0.1})
    opts = lvo.default_options(ao_strength=1.0, ao_spp=spp, ao_use_distance=int(use_distance), ao_jitter_primary=int(jitter))
    ao, st = ctx.render_rtao(sc, cam, 0)
    ref, ost = osc.render_rtao(cam, opts, 0)
    assert np.abs(ao - ref).max() <= TOL
    assert np.array_equal(ao.view(np.uint32), ref.view(np.uint32)), "AO image expected bit-exact"
    assert st["rays_ao"] == ost["rays_ao"] == ost["pixels_hit"] * spp
    # second accumulated frame (running mean, VulkanRayTracedAmbientOcclusion.glsl:313-317)
    ao2, _ = ctx.render_rtao(sc, cam, 1, out=ao.copy())
    ref2, _ = osc.render_rtao(cam, opts, 1, ao=ref.copy())
    assert np.array_equal(ao2.view(np.uint32), ref2.view(np.uint32))


@pytest.mark.parametrize("queue,stack,minb", [(False, 0, 10), (False, 1, 9), (False, 8, 8), (False, 12, 0), (False, 16, 9),
                                              (True, 1, 9), (True, 8, 8), (True, 12, 9)])
@pytest.mark.parametrize("use_distance", [True, False])
def test_rtao_kernel_variants_bit_exact(ctx, oracle, queue, stack, minb, use_distance):
    """Every variant of the AO ray kernel -- leaf-vote (k_rtao_rays) or leaf-queue (k_rtao_rays_q), local / packed / shared-memory
    traversal stack, register budget -- gives the oracle's AO image bit for bit (the default, queue + 12 shared entries, is what
    all other tests run).  A long AO radius makes the stacks deep enough to spill out of their shared part."""
    data, width = DATASETS["random"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(120, 80)
    ctx.set_new_settings({"ambient_occlusion_samples_per_frame": 8, "ambient_occlusion_distance_based": use_distance,
                          "use_jittered_primary_rays": True, "ambient_occlusion_radius": 0.3,
                          "b200_ao_stack": stack, "b200_ao_queue": queue, "b200_ao_min_blocks": minb})
    try:
        ao, st = ctx.render_rtao(sc, cam, 0)
    finally:
        ctx.set_new_settings({"b200_ao_stack": 12, "b200_ao_queue": True, "b200_ao_min_blocks": 0, "ambient_occlusion_radius": 0.1})
    opts = lvo.default_options(ao_strength=1.0, ao_spp=8, ao_use_distance=int(use_distance), ao_jitter_primary=1, ao_radius=0.3)
    ref, ost = osc.render_rtao(cam, opts, 0)
    assert np.array_equal(ao.view(np.uint32), ref.view(np.uint32))
    assert st["rays_ao"] == ost["rays_ao"] > 0


@pytest.mark.parametrize("packed,tq_bits", [(True, 0), (True, 4), (False, 0)])
@pytest.mark.parametrize("use_distance", [True, False])
def test_rtao_packed_stream_bit_exact(ctx, oracle, use_distance, packed, tq_bits):
    """k_rtao_rays_w (b200_ao_packed, the default stream on the 4-wide tree): packed fp32x2 box tests, the ray in shared memory, one-word
    stack entries whose entry distance sits in the 7 (small scenes) or 4 free index bits -- the oracle's AO image bit for bit, like the
    stream it replaces (packed = False).  The long AO radius makes the stacks deep and the hit-distance culling matter."""
    for name in ("random", "helix"):
        data, width = DATASETS[name]()
        sc, osc = _scene_pair(ctx, oracle, data, width)
        cam = lv.make_camera(120, 80)
        ctx.set_new_settings({"ambient_occlusion_samples_per_frame": 8, "ambient_occlusion_distance_based": use_distance,
                              "use_jittered_primary_rays": True, "ambient_occlusion_radius": 0.3, "b200_ao_packed": packed, "b200_ao_tq_bits": tq_bits})
        try:
            ao, st = ctx.render_rtao(sc, cam, 0)
        finally:
            ctx.set_new_settings({"b200_ao_packed": True, "b200_ao_tq_bits": 0, "ambient_occlusion_radius": 0.1})
        opts = lvo.default_options(ao_strength=1.0, ao_spp=8, ao_use_distance=int(use_distance), ao_jitter_primary=1, ao_radius=0.3)
        ref, ost = osc.render_rtao(cam, opts, 0)
        assert st["rays_ao"] == ost["rays_ao"] > 0 and np.array_equal(ao.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("use_distance", [True, False])
def test_rtao_quantised_nodes_bit_exact(ctx, oracle, use_distance):
    """b200_ao_qnodes (experimental, off by default): 32-byte nodes with 16-bit outward-rounded child boxes -- same AO image."""
    data, width = DATASETS["random"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(120, 80)
    ctx.set_new_settings({"ambient_occlusion_samples_per_frame": 8, "ambient_occlusion_distance_based": use_distance,
                          "use_jittered_primary_rays": True, "ambient_occlusion_radius": 0.3, "b200_ao_qnodes": True})
    try:
        ao, st = ctx.render_rtao(sc, cam, 0)
    finally:
        ctx.set_new_settings({"b200_ao_qnodes": False, "ambient_occlusion_radius": 0.1})
    opts = lvo.default_options(ao_strength=1.0, ao_spp=8, ao_use_distance=int(use_distance), ao_jitter_primary=1, ao_radius=0.3)
    ref, ost = osc.render_rtao(cam, opts, 0)
    assert st["rays_ao"] == ost["rays_ao"] > 0 and np.array_equal(ao.view(np.uint32), ref.view(np.uint32))


def test_rtao_queue_falls_back_for_multi_record_leaves(ctx, oracle):
    """The leaf-queue kernel relies on one-record leaves; a scene built with larger leaves takes the leaf-vote kernel."""
    data, width = DATASETS["random"]()
    ctx.set_option("b200_bvh_leaf_size", 4)
    try:
        sc, osc = _scene_pair(ctx, oracle, data, width)
    finally:
        ctx.set_option("b200_bvh_leaf_size", 1)
    cam = lv.make_camera(120, 80)
    ctx.set_new_settings({"ambient_occlusion_samples_per_frame": 4, "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True})
    ao, _ = ctx.render_rtao(sc, cam, 0)
    ref, _ = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=4), 0)
    assert np.array_equal(ao.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("queue", [True, False])
def test_rtao_degenerate_segments(ctx, oracle, queue):
    """Zero-length segments (spheres) have no tangent: the AO rays of a hit on one have NaN directions.  They hit nothing (AO 1, as the
    oracle's arithmetic gives), never enter the traversal, and still deliver a result -- in a scene made of nothing else too, where every
    ray a warp fetches is invalid.  Mixed with regular and duplicated segments the frame is the oracle's, bit for bit."""
    cam = lv.make_camera(96, 64)
    settings = {"ambient_occlusion_samples_per_frame": 4, "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": False,
                "ambient_occlusion_radius": 0.1, "b200_ao_queue": queue}
    pts = np.array([[0.0, 0.0, 0.0], [0.3, 0.2, -0.1], [-0.3, -0.2, 0.1]], np.float32)
    pos = np.repeat(pts, 2, axis=0)
    only = (pos, np.linspace(0, 1, 6).astype(np.float32), np.arange(6, dtype=np.uint32).reshape(3, 2))
    fresh = lv.Context(0, lib_path=ctx.lib_path)   # a fresh per-ray result buffer (no results of earlier frames in it)
    try:
        fresh.set_new_settings(settings)
        sc = fresh.create_scene(*only, 0.08)
        ao, st = fresh.render_rtao(sc, cam, 0)
        assert st["pixels_hit"] > 0 and st["rays_ao"] == 4 * st["pixels_hit"]
        assert np.array_equal(ao, np.ones_like(ao))
        sc.close()
    finally:
        fresh.close()
    rng = np.random.default_rng(5)
    n = 120
    p0 = (rng.random((n, 3)) - 0.5) * 0.8
    p1 = p0 + rng.standard_normal((n, 3)) * 0.1
    p0[n // 2:] = p0[:n - n // 2]; p1[n // 2:] = p1[:n - n // 2]       # exact duplicates
    p1[::5] = p0[::5]                                                  # zero-length
    pos = np.empty((2 * n, 3), np.float32); pos[0::2], pos[1::2] = p0, p1
    data = (pos, rng.random(2 * n).astype(np.float32), np.arange(2 * n, dtype=np.uint32).reshape(n, 2))
    sc, osc = _scene_pair(ctx, oracle, data, 0.04)
    ctx.set_new_settings(settings)
    try:
        ao, st = ctx.render_rtao(sc, cam, 0)
    finally:
        ctx.set_option("b200_ao_queue", True)
    ref, ost = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=4, ao_use_distance=1, ao_jitter_primary=0, ao_radius=0.1), 0)
    assert st["rays_ao"] == ost["rays_ao"] > 0 and np.array_equal(ao.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("name,ao", [("helix", False), ("helix", True), ("random", True), ("single", False)])
def test_tubes_parity(ctx, oracle, name, ao):
    data, width = DATASETS[name]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(160, 100)
    tf = scenes.standard_transfer_function(opacity=(0.4, 1.0))
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 1.0 if ao else 0.0, "ambient_occlusion_samples_per_frame": 4,
                          "use_jittered_primary_rays": True, "ambient_occlusion_distance_based": True,
                          "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    opts = lvo.default_options(ao_strength=1.0 if ao else 0.0, ao_spp=4)
    img, st = ctx.render_tubes(sc, cam, 0)
    ao_ref = osc.render_rtao(cam, opts, 0)[0] if ao else None
    ref, ost = osc.render_tubes(cam, opts, tf, ao_tex=ao_ref)
    assert np.isfinite(img).all()
    assert np.abs(img - ref).max() <= TOL
    assert st["rays_primary"] == ost["rays"] + (cam.width * cam.height if ao else 0)


def test_tubes_jittered_accumulation(ctx, oracle):
    data, width = DATASETS["helix"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(96, 64)
    tf = scenes.standard_transfer_function()
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 0.0, "num_samples_per_frame": 2, "num_accumulated_frames": 4})
    opts = lvo.default_options(num_samples_per_frame=2, use_jittered_rays=1)
    img = np.zeros((64, 96, 4), np.float32)
    ref = np.zeros((64, 96, 4), np.float32)
    for f in range(3):
        img, _ = ctx.render_tubes(sc, cam, f, out=img)
        ref, _ = osc.render_tubes(cam, opts, tf, frame_number=f, rgba=ref)
    assert np.abs(img - ref).max() <= TOL
    ctx.set_new_settings({"num_samples_per_frame": 1, "num_accumulated_frames": 1})


def _lists(heads, nodes, cam, oracle, tw=2, th=8):
    return lvo.per_pixel_lists(heads, nodes, cam, lvo.default_options(tile_w=tw, tile_h=th), oracle)


@pytest.mark.parametrize("name", ["helix", "random"])
def test_ppll_integer_path_bit_exact(ctx, oracle, name):
    data, width = DATASETS[name]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(100, 70)   # not a multiple of the 2x8 addressing tile -> padded start-offset buffer
    tf = scenes.standard_transfer_function(opacity=(0.1, 0.6))
    ctx.set_transfer_function(tf)
    ctx.set_option("ambient_occlusion_strength", 0.0)
    ctx.ppll_clear(cam, 0)
    st = ctx.ppll_gather(sc, cam)
    got = ctx.ppll_read()
    ref = osc.ppll_gather(cam, lvo.default_options(), tf)
    assert got["counter"] == ref["counter"] == st["frags_generated"] and st["frags_dropped"] == 0
    assert got["padded"] == ref["padded"]
    assert np.array_equal(got["heads"] == 0xFFFFFFFF, ref["heads"] == 0xFFFFFFFF)
    a = _lists(got["heads"], got["nodes"], cam, oracle)
    b = _lists(ref["heads"], ref["nodes"], cam, oracle)
    assert a == b, "per-pixel multisets of (depth bits, colour) differ"


@pytest.mark.parametrize("mode", ["priority_queue", "bitonic", "insertion", "quicksort_hybrid"])
def test_ppll_resolve_parity(ctx, oracle, mode):
    data, width = DATASETS["random"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(128, 96)
    tf = scenes.standard_transfer_function(opacity=(0.1, 0.6))
    ctx.set_transfer_function(tf)
    ctx.set_option("ambient_occlusion_strength", 0.0)
    img, st = ctx.render_ppll(sc, cam, max_frags=256, sort_mode=mode)
    ref = osc.ppll_gather(cam, lvo.default_options(), tf)
    m = lv.SORT_MODES[mode]
    ref_img, rst = lvo.ppll_resolve(oracle, cam, lvo.default_options(), ref["heads"], ref["nodes"], 256, m, canonical=True)
    assert st["frags_sorted"] == rst["frags_sorted"] and st["max_depth_complexity"] == rst["max_depth_complexity"]
    # a list whose fragments all quantise to alpha 0 (0.001 <= a < 0.5/255) resolves to 0/0 = NaN in the reference's
    # blendFTB too (LinkedListSort.glsl:57); both sides must agree on where that happens
    nan = np.isnan(ref_img)
    assert np.array_equal(np.isnan(img), nan)
    assert np.abs(img[~nan] - ref_img[~nan]).max() <= TOL
    assert np.array_equal(img[~nan].view(np.uint32), ref_img[~nan].view(np.uint32)), "resolve expected bit-exact vs canonical oracle order"


@pytest.mark.parametrize("variant", [{"b200_ppll_reg_sort": True}, {"b200_ppll_binned_resolve": True}, {"b200_ppll_resolve_tile": 256},
                                     {"b200_ppll_reg_sort": True, "b200_ppll_resolve_tile": 512}], ids=lambda v: "+".join(v))
def test_ppll_resolve_variants_bit_exact(ctx, oracle, variant):
    """The optional resolve kernels -- warp bitonic sort in registers (4 / 8 keys per lane), smaller shared key tiles, the count-binned one --
    give the default kernel's frame bit for bit, on lists deep enough for every length class (a dense scene seen through a small frame)."""
    data = scenes.random_segments(5000, 0.35, seed=13)
    sc, osc = _scene_pair(ctx, oracle, data, 0.03)
    cam = lv.make_camera(96, 64)
    tf = scenes.standard_transfer_function(opacity=(0.2, 0.7))
    ctx.set_transfer_function(tf)
    ctx.set_option("ambient_occlusion_strength", 0.0)
    size = 64 * 96 * 64   # no overflow: which fragments an overflowing buffer drops is a race, in the reference too
    want, wst = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic", linked_list_size=size)
    ctx.set_new_settings(variant)
    try:
        img, st = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic", linked_list_size=size)
    finally:
        ctx.set_new_settings({"b200_ppll_reg_sort": False, "b200_ppll_binned_resolve": False, "b200_ppll_resolve_tile": 1024})
    assert st["frags_dropped"] == wst["frags_dropped"] == 0
    assert st["frags_sorted"] == wst["frags_sorted"] and st["max_depth_complexity"] == wst["max_depth_complexity"] > 130
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(img), nan) and np.array_equal(img[~nan].view(np.uint32), want[~nan].view(np.uint32))
    g = osc.ppll_gather(cam, lvo.default_options(), tf, linked_list_size=size)
    ref, _ = lvo.ppll_resolve(oracle, cam, lvo.default_options(), g["heads"], g["nodes"], 256, lv.SORT_MODES["bitonic"], canonical=True)
    assert np.array_equal(np.isnan(ref), nan) and np.array_equal(img[~nan].view(np.uint32), ref[~nan].view(np.uint32))


@pytest.mark.parametrize("mode", ["raster", "raster_contiguous"])
@pytest.mark.parametrize("name,eye_z", [("random", 0.8), ("helix", 0.8), ("random", 0.05)])
def test_ppll_raster_gather_bit_exact(ctx, oracle, name, eye_z, mode):
    """b200_ppll_gather_mode = raster (object-order gather, one warp per segment) and raster_contiguous (+ count / scan / fill: every list
    one contiguous run, index-addressed resolve): per pixel the same multiset of fragments as the oracle's all-hits enumeration, the
    same counters, the same resolved frame -- also with the camera inside the data."""
    data, width = DATASETS[name]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(128, 96, eye=(0.02, -0.01, eye_z))
    tf = scenes.standard_transfer_function(opacity=(0.1, 0.6))
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 0.0, "b200_ppll_gather_mode": mode})
    size = 300 * 128 * 96
    try:
        img, st = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic", linked_list_size=size)
        mine = ctx.ppll_read()
    finally:
        ctx.set_option("b200_ppll_gather_mode", "raycast")
    opts = lvo.default_options()
    g = osc.ppll_gather(cam, opts, tf, linked_list_size=size)
    assert st["frags_generated"] == g["counter"] == mine["counter"] and st["frags_dropped"] == 0
    assert lvo.per_pixel_lists(mine["heads"], mine["nodes"], cam, opts, oracle) == lvo.per_pixel_lists(g["heads"], g["nodes"], cam, opts, oracle)
    if st["max_depth_complexity"] <= 256:
        ref, rst = lvo.ppll_resolve(oracle, cam, opts, g["heads"], g["nodes"], 256, lv.SORT_MODES["bitonic"], canonical=True)
        nan = np.isnan(ref)
        assert np.array_equal(np.isnan(img), nan) and np.array_equal(img[~nan].view(np.uint32), ref[~nan].view(np.uint32))


def test_ppll_overflow_is_counted_not_fatal(ctx, oracle):
    data, width = DATASETS["helix"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(96, 64)
    tf = scenes.standard_transfer_function(opacity=(0.2, 0.5))
    ctx.set_transfer_function(tf)
    ref = osc.ppll_gather(cam, lvo.default_options(), tf)
    budget = ref["counter"] // 3
    img, st = ctx.render_ppll(sc, cam, max_frags=64, sort_mode="priority_queue", linked_list_size=budget)
    assert st["frags_generated"] == ref["counter"] and st["frags_stored"] == budget and st["frags_dropped"] == ref["counter"] - budget
    assert st["frags_sorted"] + st["frags_truncated"] == budget and np.isfinite(img).all()


def test_ppll_truncation_at_max_frags(ctx, oracle):
    data, width = DATASETS["random"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(64, 48)
    tf = scenes.standard_transfer_function(opacity=(0.1, 0.3))
    ctx.set_transfer_function(tf)
    img, st = ctx.render_ppll(sc, cam, max_frags=4, sort_mode="bitonic")
    assert st["frags_truncated"] > 0 and st["frags_sorted"] + st["frags_truncated"] == st["frags_stored"]
    assert np.isfinite(img).all()


def test_owned_tiles_match_host_enumeration(ctx):
    """lv_get_owned_tiles (C side) == linevis_b200.sharding.owned_tiles (host side used for the NCCL gather)."""
    from linevis_b200 import sharding
    for (W, H, ts, world) in [(3840, 2160, 64, 8), (200, 136, 32, 4), (100, 70, 16, 3)]:
        for r in range(world):
            c = lv.Context(0)
            c.set_tile_shard(r, world, ts)
            assert np.array_equal(c.owned_tiles(W, H), sharding.owned_tiles(W, H, ts, r, world))
            c.close()


def test_pack_unpack_tiles_kernels(ctx):
    import torch
    from linevis_b200 import sharding
    W, H, ts, world = 200, 136, 32, 3
    img = torch.rand((H, W, 4), device="cuda")
    out = torch.zeros_like(img)
    for r in range(world):
        c = lv.Context(0)
        c.set_tile_shard(r, world, ts)
        n = len(c.owned_tiles(W, H))
        packed = torch.zeros((n, ts * ts, 4), device="cuda")
        c.pack_owned_tiles(img, W, H, packed)
        c.synchronize()
        ref = sharding.pack_tiles_torch(img, sharding.owned_tiles(W, H, ts, r, world), ts, n)
        assert torch.equal(packed, ref)
        c.unpack_tiles(packed, r, world, W, H, out)
        c.close()
    assert torch.equal(out, img)


def test_tile_sharding_reassembles_full_frame(ctx, oracle):
    """world_size 4 emulated on one GPU: every rank renders its Morton-interleaved tiles; union == unsharded frame."""
    data, width = DATASETS["helix"]()
    pos, attr, seg = data
    cam = lv.make_camera(200, 136)
    tf = scenes.standard_transfer_function(opacity=(0.4, 1.0))
    settings = {"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 4, "num_samples_per_frame": 1,
                "num_accumulated_frames": 1, "use_jittered_primary_rays": False}
    ctx.set_transfer_function(tf)
    ctx.set_new_settings(settings)
    sc = ctx.create_scene(pos, attr, seg, width)
    full, _ = ctx.render_tubes(sc, cam, 0)
    fullp, _ = ctx.render_ppll(sc, cam, 64, "priority_queue")
    acc = np.full_like(full, np.nan)
    accp = np.full_like(full, np.nan)
    total_rays = 0
    for r in range(4):
        c = lv.Context(0)
        c.set_transfer_function(tf)
        c.set_new_settings(settings)
        c.set_tile_shard(r, 4, 32)
        s = c.create_scene(pos, attr, seg, width)
        out = np.full_like(full, np.nan)
        out, st = c.render_tubes(s, cam, 0, out=out)
        outp = np.full_like(full, np.nan)
        outp, _ = c.render_ppll(s, cam, 64, "priority_queue", out=outp)
        m = ~np.isnan(out[..., 0])
        assert np.isnan(acc[m]).all(), "tiles of different ranks overlap"
        acc[m] = out[m]
        accp[m] = outp[m]
        total_rays += st["rays_primary"]
        s.close(); c.close()
    assert not np.isnan(acc).any()
    assert np.abs(acc - full).max() <= 2e-4 and np.array_equal(accp, fullp)   # see test_gpu_fullsize: AO texels of unowned neighbours


@pytest.mark.parametrize("ao", [False, True])
def test_depth_cues_parity(ctx, oracle, ao):
    """USE_DEPTH_CUES: min/max view depth of the line vertices (DepthCues/ComputeDepthValues.glsl) + the mix towards grey in
    blinnPhongShadingTube (Utils/Lighting.glsl:183-187), for the tube pass and for the PPLL fragments."""
    data, width = DATASETS["helix"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(150, 90, eye=(0.1, 0.15, 0.7))
    tf = scenes.standard_transfer_function(opacity=(0.4, 1.0))
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"depth_cue_strength": 0.8, "ambient_occlusion_strength": 1.0 if ao else 0.0, "ambient_occlusion_samples_per_frame": 4,
                          "use_jittered_primary_rays": True, "ambient_occlusion_distance_based": True, "ambient_occlusion_radius": 0.1,
                          "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    try:
        opts = lvo.default_options(depth_cue_strength=0.8, ao_strength=1.0 if ao else 0.0, ao_spp=4)
        img, _ = ctx.render_tubes(sc, cam, 0)
        ao_ref = osc.render_rtao(cam, opts, 0)[0] if ao else None
        ref, _ = osc.render_tubes(cam, opts, tf, ao_tex=ao_ref)
        assert np.abs(img - ref).max() <= TOL and np.array_equal(img.view(np.uint32), ref.view(np.uint32))
        plain, _ = osc.render_tubes(cam, lvo.default_options(ao_strength=1.0 if ao else 0.0, ao_spp=4), tf, ao_tex=ao_ref)
        assert np.abs(plain - ref).max() > 0.02, "depth cues must change the image"
        ctx.set_option("ambient_occlusion_strength", 0.0)
        ctx.ppll_clear(cam, 0)
        ctx.ppll_gather(sc, cam)
        got = ctx.ppll_read()
        refg = osc.ppll_gather(cam, lvo.default_options(depth_cue_strength=0.8), tf)
        assert got["counter"] == refg["counter"]
        assert _lists(got["heads"], got["nodes"], cam, oracle) == _lists(refg["heads"], refg["nodes"], cam, oracle)
    finally:
        ctx.set_new_settings({"depth_cue_strength": 0.0, "ambient_occlusion_strength": 0.0})


def test_peer_frame_single_process(ctx, oracle):
    """lv_frame_alloc / lv_ipc_export and rendering into a raw library-owned frame (the peer-memory assembly path of N > 1 runs;
    opening the handle needs a second process and is covered by tools/p2p_check.py under torchrun)."""
    import torch
    from linevis_b200.sharding import PeerFrame
    data, width = DATASETS["helix"]()
    sc = ctx.create_scene(*data, width)
    cam = lv.make_camera(160, 100)
    ctx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.4, 1.0)))
    ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 4, "num_samples_per_frame": 1,
                          "num_accumulated_frames": 1})
    ref, _ = ctx.render_tubes(sc, cam)
    pf = PeerFrame(ctx, 160, 100, 0, 1, torch.device("cuda", 0))
    try:
        assert len(ctx.ipc_export(pf.ptr)) == 64
        ctx.render_tubes(sc, cam, 0, out=pf.ptr, stats=False)
        pf.fence()
        ctx.synchronize()
        got = pf.tensor().cpu().numpy()
    finally:
        pf.close()
        ctx.set_option("ambient_occlusion_strength", 0.0)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_frame_to_rgba8(ctx, oracle):
    """lv_frame_to_rgba8: the frame in the reference's sceneTexture format (packUnorm4x8 per pixel), from a device frame into host memory."""
    import torch
    data, width = DATASETS["helix"]()
    sc = ctx.create_scene(*data, width)
    cam = lv.make_camera(160, 100)
    ctx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.4, 1.0)))
    ctx.set_option("ambient_occlusion_strength", 0.0)
    frame = torch.zeros((100, 160, 4), dtype=torch.float32, device="cuda")
    ctx.render_tubes(sc, cam, 0, out=frame, stats=False)
    got = ctx.frame_to_rgba8(frame, 160, 100)
    f = frame.cpu().numpy()
    q = np.floor(np.clip(f, 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint32)   # round(clamp(c, 0, 1) * 255), halves never occur exactly off 0.5 ties
    want = q[..., 0] | (q[..., 1] << 8) | (q[..., 2] << 16) | (q[..., 3] << 24)
    assert (np.abs(((got[..., None] >> np.array([0, 8, 16, 24], np.uint32)) & 0xFF).astype(np.int64) - q.astype(np.int64)) <= 1).all()
    assert (got == want).mean() > 0.999
    with pytest.raises(lv.LineVisError):
        ctx.frame_to_rgba8(f, 160, 100)      # a host float frame is refused


@pytest.mark.parametrize("world", [2, 4])
def test_ao_sample_batch_shards_one_gpu_plays_the_ranks(ctx, oracle, world):
    """lv_sao_primary / lv_sao_trace / lv_sao_finish (AO rays sharded by sample batch): one context plays the ranks in turn, the exchanges are
    tensor copies; the union of the ranks' frames is the oracle's frame bit for bit and every rank traces the same number of AO rays."""
    import torch
    from linevis_b200 import sharding
    data, width = DATASETS["helix"]()
    sc, osc = _scene_pair(ctx, oracle, data, width)
    W, H, spp, tile = 160, 96, 8, 16
    spl = spp // world
    cam = lv.make_camera(W, H)
    tf = scenes.standard_transfer_function(opacity=(1.0, 1.0))
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": spp, "ambient_occlusion_iterations": 1,
                          "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True, "ambient_occlusion_radius": 0.1,
                          "num_samples_per_frame": 2, "num_accumulated_frames": 1})
    try:
        full, _ = ctx.render_tubes(sc, cam)
        lists = []
        for r in range(world):
            ctx.set_tile_shard(r, world, tile)
            ptr, n = ctx.sao_primary(sc, cam, 0)
            lists.append(sharding._device_floats(ptr, (max(n, 1), 12), "cuda")[:n].clone())
        counts = [h.shape[0] for h in lists]
        hits = torch.cat(lists, dim=0).contiguous()
        occs, rays = [], []
        for r in range(world):
            ctx.set_tile_shard(r, world, tile)
            occ = torch.zeros(hits.shape[0] * spl, device="cuda")
            ctx.sao_trace(sc, cam, 0, hits, hits.shape[0], r * spl, spl, occ)
            ctx.synchronize()
            occs.append(occ.view(hits.shape[0], spl))
        acc = np.full_like(full, np.nan)
        off = np.concatenate([[0], np.cumsum(counts)])
        for r in range(world):
            ctx.set_tile_shard(r, world, tile)
            ptr, n = ctx.sao_primary(sc, cam, 0)        # the context's own hit list is rank r's again (same pixels, maybe another order)
            assert n == counts[r]
            now = sharding._device_floats(ptr, (max(n, 1), 12), "cuda")[:n]
            pix_now = now[:, 7].contiguous().view(torch.int32).cpu().numpy()
            pix_then = lists[r][:, 7].contiguous().view(torch.int32).cpu().numpy()
            order = np.argsort(pix_then, kind="stable")[np.searchsorted(np.sort(pix_then), pix_now)]
            assert np.array_equal(pix_then[order], pix_now)
            idx = torch.from_numpy(order.astype(np.int64)).cuda()
            parts = torch.stack([occs[j][off[r]:off[r + 1]][idx] for j in range(world)], dim=0).contiguous()
            part = torch.full((H, W, 4), float("nan"), device="cuda")
            _, st = ctx.sao_finish(sc, cam, 0, parts, world, part)
            p = part.cpu().numpy()
            m = ~np.isnan(p[..., 0])
            assert not (m & ~np.isnan(acc[..., 0])).any()
            acc[m] = p[m]
        assert np.array_equal(acc.view(np.uint32), full.view(np.uint32))
        opts = lvo.default_options(ao_strength=1.0, ao_spp=spp, ao_use_distance=1, ao_jitter_primary=1, num_samples_per_frame=2, use_jittered_rays=1)
        ref, _ = osc.render_tubes(cam, opts, tf, ao_tex=osc.render_rtao(cam, opts, 0)[0])
        assert np.abs(acc - ref).max() <= TOL
    finally:
        ctx.set_tile_shard(0, 1, 64)
        ctx.set_new_settings({"ambient_occlusion_strength": 0.0, "num_samples_per_frame": 1, "ambient_occlusion_samples_per_frame": 4, "ambient_occlusion_iterations": 64})


def test_host_sah_builder_same_results(ctx, oracle):
    """b200_bvh_builder = sah (binned SAH on the host threads): another tree, the same closest hits and the same AO image bit for bit."""
    data, width = DATASETS["random"]()
    ctx.set_option("b200_bvh_builder", "sah")
    try:
        sc = ctx.create_scene(*data, width)
    finally:
        ctx.set_option("b200_bvh_builder", "lbvh")
    sc0, osc = _scene_pair(ctx, oracle, data, width)
    cam = lv.make_camera(120, 80)
    h, _ = ctx.trace_primary(sc, cam)
    h0, _ = ctx.trace_primary(sc0, cam)
    assert np.array_equal(h, h0)
    ctx.set_new_settings({"ambient_occlusion_samples_per_frame": 8, "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True,
                          "ambient_occlusion_radius": 0.3})
    try:
        ao, st = ctx.render_rtao(sc, cam, 0)
        ao0, st0 = ctx.render_rtao(sc0, cam, 0)
    finally:
        ctx.set_new_settings({"ambient_occlusion_radius": 0.1})
    assert st["rays_ao"] == st0["rays_ao"] > 0 and np.array_equal(ao.view(np.uint32), ao0.view(np.uint32))
    assert st["ao_traversal_steps"] != st0["ao_traversal_steps"]
    ref, _ = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=8, ao_use_distance=1, ao_jitter_primary=1, ao_radius=0.3), 0)
    assert np.array_equal(ao.view(np.uint32), ref.view(np.uint32))
