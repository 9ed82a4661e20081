"""BASELINE.json-sized inputs, checked through size-independent properties (the oracle would take minutes at these sizes):
linked-list structural invariants, counter identities, determinism / idempotence, leaf-size independence, tile-shard union."""
import numpy as np
import pytest

import linevis_b200 as lv
from linevis_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def helix100k():
    return scenes.helix_lines()            # config 2: 100 k segments


@pytest.fixture(scope="module")
def random1m():
    return scenes.random_segments(1_000_000, seed=2002)   # configs 3/4


def _walk_lists(heads, nodes):
    """Vectorised walk of all lists at once: returns per-pixel lengths and the number of node visits."""
    nxt = nodes["next"].astype(np.int64)
    cur = heads.astype(np.int64)
    NONE = 0xFFFFFFFF
    lengths = np.zeros(cur.shape, np.int64)
    visited = np.zeros(len(nodes), np.uint8)
    while True:
        m = cur != NONE
        if not m.any():
            break
        idx = cur[m]
        assert visited[idx].max() == 0, "a node is reachable from two lists / twice"
        visited[idx] = 1
        lengths[m] += 1
        cur[m] = nxt[idx]
    return lengths, int(visited.sum())


def test_config2_ppll_structure_and_counters(helix100k):
    pos, attr, seg = helix100k
    ctx = lv.Context(0)
    ctx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.1, 0.6)))
    sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
    cam = lv.make_camera(1920, 1080)
    img, st = ctx.render_ppll(sc, cam, max_frags=100, sort_mode="priority_queue")
    got = ctx.ppll_read()
    assert got["counter"] == st["frags_generated"] == st["frags_stored"] and st["frags_dropped"] == 0
    lengths, visited = _walk_lists(got["heads"], got["nodes"])
    assert visited == got["counter"], "every stored node belongs to exactly one list"
    assert lengths.max() == st["max_depth_complexity"]
    assert st["frags_sorted"] + st["frags_truncated"] == got["counter"]
    assert st["frags_sorted"] == int(np.minimum(lengths, 100).sum())
    depth = got["nodes"]["depth"]
    assert np.isfinite(depth).all() and depth.min() > 0.3 and depth.max() < 1.2      # camera at 0.8, data in a 0.5 box
    assert ((got["nodes"]["color"] >> 24) > 0).mean() > 0.99                          # alpha >= 0.001 survives the gather test
    # empty pixels keep the clear colour, covered ones are blended towards the tube colours
    pw, ph = got["padded"]
    empty = (lengths == 0)
    assert np.isfinite(img).all() and (img[..., 3] <= 1.0 + 1e-6).all()
    # idempotence: a second frame gives the same counters and the same image (node order may differ, the result may not)
    img2, st2 = ctx.render_ppll(sc, cam, max_frags=100, sort_mode="priority_queue")
    assert st2["frags_sorted"] == st["frags_sorted"] and np.array_equal(img, img2)
    # all correct-sort modes agree; the priority queue only differs where alpha saturates
    img_ins, _ = ctx.render_ppll(sc, cam, max_frags=100, sort_mode="insertion")
    img_bit, _ = ctx.render_ppll(sc, cam, max_frags=100, sort_mode="bitonic")
    assert np.array_equal(img_ins, img_bit)
    assert np.abs(img_ins - img).max() <= 0.011
    # the binned resolve variant is bit-identical
    ctx.set_option("b200_ppll_binned_resolve", True)
    img_b, st_b = ctx.render_ppll(sc, cam, max_frags=100, sort_mode="priority_queue")
    assert np.array_equal(img_b, img) and st_b["frags_sorted"] == st["frags_sorted"]
    sc.close(); ctx.close()


def test_config3_tubes_rtao_properties(random1m):
    pos, attr, seg = random1m
    cam = lv.make_camera(1920, 1080)
    frames, stats = [], []
    for leaf in (1, 4):
        ctx = lv.Context(0)
        ctx.set_transfer_function(scenes.standard_transfer_function())
        ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 16, "ambient_occlusion_iterations": 1,
                              "num_samples_per_frame": 1, "num_accumulated_frames": 1, "b200_bvh_leaf_size": leaf})
        sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
        img, st = ctx.render_tubes(sc, cam, 0)
        ao, sta = ctx.render_rtao(sc, cam, 0)
        assert sta["rays_ao"] == 16 * sta["pixels_hit"] and sta["rays_primary"] == 1920 * 1080
        assert ao.min() >= 0.0 and ao.max() == 1.0 and np.isfinite(img).all()
        img_again, _ = ctx.render_tubes(sc, cam, 0)
        assert np.array_equal(img, img_again), "frame is deterministic"
        frames.append(img); stats.append(st)
        sc.close(); ctx.close()
    # the result does not depend on the BVH (leaf size 1 vs 4): bit-identical frames and hit counts
    assert np.array_equal(frames[0], frames[1]) and stats[0]["pixels_hit"] == stats[1]["pixels_hit"]
    assert stats[0]["rays_ao"] == stats[1]["rays_ao"]


def test_config4_like_fragment_count_is_bvh_independent(random1m):
    pos, attr, seg = random1m
    cam = lv.make_camera(1280, 720)
    counts = []
    for leaf in (1, 8):
        ctx = lv.Context(0)
        ctx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.1, 0.6)))
        ctx.set_option("b200_bvh_leaf_size", leaf)
        sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
        img, st = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="priority_queue")
        counts.append((st["frags_generated"], st["frags_sorted"], st["max_depth_complexity"], int(np.isnan(img).sum()), float(np.nansum(img))))
        sc.close(); ctx.close()
    assert counts[0] == counts[1]


def test_config3_ao_image_is_kernel_independent(random1m):
    """Config 3 (1 M segments, 1080p, 16 spp): every AO ray-stream kernel -- leaf queue (default), leaf vote, quantised nodes, other
    stack layouts -- produces the same AO image bit for bit, with the same number of rays."""
    pos, attr, seg = random1m
    cam = lv.make_camera(1920, 1080)
    ctx = lv.Context(0)
    ctx.set_new_settings({"ambient_occlusion_samples_per_frame": 16, "ambient_occlusion_iterations": 1, "ambient_occlusion_distance_based": True,
                          "use_jittered_primary_rays": True})
    sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
    want, wst = ctx.render_rtao(sc, cam, 0)
    assert wst["rays_ao"] == 16 * wst["pixels_hit"] > 0
    for variant in ({"b200_ao_queue": False}, {"b200_ao_qnodes": True}, {"b200_ao_stack": 1}, {"b200_ao_queue": False, "b200_ao_stack": 0},
                    {"b200_ao_stack": 16, "b200_ao_min_blocks": 9}, {"b200_ao_packed": False}, {"b200_ao_tq_bits": 4}, {"b200_ao_wide": False}, {"b200_ao_wide": False, "b200_ao_raybuf": False},
                    {"b200_ao_raybuf": False}, {"b200_ao_raybuf": False, "b200_ao_min_blocks": 9}, {"b200_ao_wide_top": 85}, {"b200_ao_refill_below": 32},
                    {"b200_ao_wide": False, "b200_ao_raybuf": False, "b200_ao_queue": False}):
        ctx.set_new_settings(variant)
        got, st = ctx.render_rtao(sc, cam, 0)
        ctx.set_new_settings({"b200_ao_queue": True, "b200_ao_qnodes": False, "b200_ao_stack": 12, "b200_ao_min_blocks": 0, "b200_ao_wide": True,
                              "b200_ao_raybuf": True, "b200_ao_wide_top": 0, "b200_ao_refill_below": 0, "b200_ao_packed": True, "b200_ao_tq_bits": 0})
        assert st["rays_ao"] == wst["rays_ao"], variant
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), variant
    sc.close(); ctx.close()


def test_config4_like_resolve_is_kernel_independent(random1m):
    """1 M segments, MAX_NUM_FRAGS 256: the resolve variants (in-register sort, smaller key tiles, count-binned) give the default kernel's
    frame bit for bit on a frame with long lists (no fragment is dropped, so the lists are the same sets in every run)."""
    pos, attr, seg = random1m
    cam = lv.make_camera(1280, 720)
    ctx = lv.Context(0)
    ctx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.1, 0.6)))
    ctx.set_option("b200_expected_avg_depth_complexity", 40)
    sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
    want, wst = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic")
    assert wst["frags_dropped"] == 0 and wst["max_depth_complexity"] > 64
    for variant in ({"b200_ppll_reg_sort": True}, {"b200_ppll_resolve_tile": 256}, {"b200_ppll_reg_sort": True, "b200_ppll_resolve_tile": 512},
                    {"b200_ppll_binned_resolve": True}):
        ctx.set_new_settings(variant)
        got, st = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic")
        ctx.set_new_settings({"b200_ppll_reg_sort": False, "b200_ppll_resolve_tile": 1024, "b200_ppll_binned_resolve": False})
        assert st["frags_sorted"] == wst["frags_sorted"] and st["frags_dropped"] == 0, variant
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan) and np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32)), variant
    sc.close(); ctx.close()


def test_config4_like_raster_gather_equals_raycast_gather(random1m):
    """1 M segments: the object-order gather generates exactly the ray-cast gather's fragments (same count, same per-pixel list lengths)
    and -- nothing dropped or truncated, so every list is the same set -- the same resolved frame, bit for bit."""
    pos, attr, seg = random1m
    cam = lv.make_camera(1280, 720)
    ctx = lv.Context(0)
    ctx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.1, 0.6)))
    ctx.set_option("b200_expected_avg_depth_complexity", 40)
    sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
    out = {}
    for mode in ("raycast", "raster", "raster_contiguous"):
        ctx.set_option("b200_ppll_gather_mode", mode)
        img, st = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic")
        r = ctx.ppll_read()
        lengths, visited = _walk_lists(r["heads"], r["nodes"])
        assert visited == min(r["counter"], len(r["nodes"]))
        out[mode] = (img, st, lengths)
    a, sa, la = out["raycast"]
    nan = np.isnan(a)
    for mode in ("raster", "raster_contiguous"):
        b, sb, lb = out[mode]
        assert sa["frags_dropped"] == sb["frags_dropped"] == 0 and sa["frags_truncated"] == sb["frags_truncated"] == 0, mode
        assert sa["frags_generated"] == sb["frags_generated"] and sa["max_depth_complexity"] == sb["max_depth_complexity"], mode
        assert np.array_equal(la, lb), mode
        assert np.array_equal(np.isnan(b), nan) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32)), mode
    sc.close(); ctx.close()


@pytest.mark.parametrize("tube_jitter", [False, True])
def test_sharded_union_equals_full_frame_at_1080p(helix100k, tube_jitter):
    """8 ranks emulated on one GPU.  Without tube jitter the AO lookup sits on the pixel centre and neighbouring texels only
    enter with weights ~1e-4 (float error of the re-projection), so unowned neighbours (left at 1.0) may move a border pixel
    by <= 2e-4; with jittered tube rays the library renders a one-pixel AO ring around its tiles and the union is exact."""
    pos, attr, seg = helix100k
    cam = lv.make_camera(1920, 1080)
    tf = scenes.standard_transfer_function(opacity=(0.3, 1.0))
    settings = {"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 4, "num_samples_per_frame": 1,
                "num_accumulated_frames": 2 if tube_jitter else 1, "use_jittered_primary_rays": True}
    ctx = lv.Context(0)
    ctx.set_transfer_function(tf); ctx.set_new_settings(settings)
    sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
    full, st_full = ctx.render_tubes(sc, cam, 0)
    sc.close(); ctx.close()
    acc = np.full_like(full, np.nan)
    rays = 0
    for r in range(8):
        c = lv.Context(0)
        c.set_transfer_function(tf); c.set_new_settings(settings); c.set_tile_shard(r, 8, 64)
        s = c.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
        out, st = c.render_tubes(s, cam, 0, out=np.full_like(full, np.nan))
        m = ~np.isnan(out[..., 0])
        assert np.isnan(acc[m]).all()
        acc[m] = out[m]
        rays += st["rays_primary"] + st["rays_ao"]
        s.close(); c.close()
    assert not np.isnan(acc).any()
    if tube_jitter:
        assert np.array_equal(acc, full)
        assert rays >= st_full["rays_primary"] + st_full["rays_ao"]           # the AO ring is extra work
    else:
        assert np.abs(acc - full).max() <= 2e-4
        assert rays == st_full["rays_primary"] + st_full["rays_ao"], "work is partitioned, not duplicated"


# ----------------------------------------------------------------------------------------------------------------------------------
# BASELINE.json's own configurations against the oracle: a centre-crop camera of the full frame (same ray density, the whole scene),
# rendered by the CUDA path and by the oracle on the REFERENCE's CPU BVH library (oracle/_ref; the portable oracle BVH where that
# is not built).  Bars: AO image, PPLL counter / per-pixel multisets bit-exact; frames within 1e-3 (in practice bit-exact too).
def _crop_camera(W, H, sw, sh):
    import math
    assert sw * H == sh * W, "the crop keeps the frame's aspect ratio, so that pixels keep their solid angle"
    return lv.make_camera(sw, sh, fov_y=2.0 * math.atan(0.5 * sh / H))


@pytest.fixture(scope="module")
def oracle_best():
    from oracle import lvo
    try:
        o = lvo.Oracle("ref")
    except (FileNotFoundError, OSError):
        o = lvo.Oracle("own")
    o.set_num_threads()
    return o


def _tubes_rtao_crop_vs_oracle(data, W, H, sw, sh, spp, oracle):
    from oracle import lvo
    pos, attr, seg = data
    cam = _crop_camera(W, H, sw, sh)
    tf = scenes.standard_transfer_function()
    ctx = lv.Context(0)
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"depth_cue_strength": 0.0, "ambient_occlusion_strength": 1.0, "ambient_occlusion_gamma": 1.0,
                          "ambient_occlusion_samples_per_frame": spp, "ambient_occlusion_iterations": 1, "ambient_occlusion_radius": 0.1,
                          "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True,
                          "num_samples_per_frame": 1, "num_accumulated_frames": 1})     # bench.py's settings
    sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
    ao, sta = ctx.render_rtao(sc, cam, 0)
    img, st = ctx.render_tubes(sc, cam, 0)
    sc.close(); ctx.close()
    osc = oracle.scene(pos, attr, seg, scenes.LINE_WIDTH)
    opts = lvo.default_options(ao_strength=1.0, ao_spp=spp, ao_jitter_primary=1, ao_use_distance=1)
    rao, s1 = osc.render_rtao(cam, opts, 0)
    ref, s2 = osc.render_tubes(cam, opts, tf, ao_tex=rao)
    assert sta["pixels_hit"] == s1["pixels_hit"] > 0.2 * sw * sh and sta["rays_ao"] == s1["rays_ao"] == spp * s1["pixels_hit"]
    assert np.array_equal(ao.view(np.uint32), rao.view(np.uint32)), "AO image bit-exact"
    assert np.isfinite(img).all() and np.abs(img - ref).max() <= 1e-3
    assert np.array_equal(img.view(np.uint32), ref.view(np.uint32)), "frame expected bit-exact"


def test_config3_crop_equals_oracle(random1m, oracle_best):
    """Config 3 (1 M random segments, 1920x1080, tubes + 16-spp RTAO): the 480x270 centre crop, CUDA vs oracle."""
    _tubes_rtao_crop_vs_oracle(random1m, 1920, 1080, 480, 270, 16, oracle_best)


@pytest.fixture(scope="module")
def curl10m():
    import torch
    return scenes.curl_noise_streamlines(n_lines=20000, n_points=501, seed=3003, device=torch.device("cuda", 0))    # config 5


def test_config5_crop_equals_oracle(curl10m, oracle_best):
    """Config 5, the headline (10 M curl-noise segments, 3840x2160, tubes + 64-spp RTAO): the 384x216 centre crop, CUDA vs oracle
    (VulkanRayTracedAmbientOcclusion.glsl:178-319, TubeRayTracing.glsl:198-274)."""
    _tubes_rtao_crop_vs_oracle(curl10m, 3840, 2160, 384, 216, 64, oracle_best)


@pytest.mark.parametrize("mode", ["raster", "raycast", "raster_contiguous"])
def test_config4_crop_equals_oracle(random1m, oracle_best, mode):
    """Config 4 (1 M segments, 3840x2160, PPLL, MAX_NUM_FRAGS 256): the 480x270 centre crop -- fragment counter, per-pixel multisets of
    (depth bits, colour) and the resolved frame, CUDA vs oracle (LinkedListGather.glsl:33-72, LinkedListResolve.glsl:57-105)."""
    from oracle import lvo
    pos, attr, seg = random1m
    cam = _crop_camera(3840, 2160, 480, 270)
    tf = scenes.standard_transfer_function(opacity=(0.1, 0.6))
    ctx = lv.Context(0)
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_strength": 0.0, "b200_ppll_gather_mode": mode})
    sc = ctx.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
    size = 128 * 480 * 270        # the crop looks at the dense centre of the data: 76 fragments per pixel on average, nothing may be dropped
    img, st = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="priority_queue", linked_list_size=size)
    got = ctx.ppll_read()
    imgb, stb = ctx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic", linked_list_size=size)
    sc.close(); ctx.close()
    osc = oracle_best.scene(pos, attr, seg, scenes.LINE_WIDTH)
    opts = lvo.default_options()
    ref = osc.ppll_gather(cam, opts, tf, linked_list_size=size)
    assert got["counter"] == ref["counter"] == st["frags_generated"] > 10 * 480 * 270 and st["frags_dropped"] == 0
    assert got["padded"] == ref["padded"]
    assert np.array_equal(got["heads"] == 0xFFFFFFFF, ref["heads"] == 0xFFFFFFFF)
    a = lvo.per_pixel_multisets(got["heads"], got["nodes"])
    b = lvo.per_pixel_multisets(ref["heads"], ref["nodes"])
    assert a.shape == b.shape == (ref["counter"], 2) and np.array_equal(a, b), "per-pixel multisets of (depth bits, colour) differ"
    for mine, mst, m in ((img, st, 0), (imgb, stb, 5)):
        rimg, rst = lvo.ppll_resolve(oracle_best, cam, opts, ref["heads"], ref["nodes"], 256, m, canonical=True)
        assert mst["frags_sorted"] == rst["frags_sorted"] and mst["max_depth_complexity"] == rst["max_depth_complexity"]
        nan = np.isnan(rimg)
        assert np.array_equal(np.isnan(mine), nan)
        assert np.abs(mine[~nan] - rimg[~nan]).max() <= 1e-3
        assert np.array_equal(mine[~nan].view(np.uint32), rimg[~nan].view(np.uint32)), "resolved frame expected bit-exact"
