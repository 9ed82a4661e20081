"""The whole C ABI on the HOST: linevis_b200/csrc (lv_api.cu and every kernel, unchanged source) compiled against the SIMT emulator
tests/emu/emu_cuda.hpp and compared with the oracle bit for bit -- the same statements the -m gpu parity tests make, at sizes a
fiber-per-thread emulation finishes in seconds.  Covers what the per-function emulation (test_emu.py) cannot: the GPU BVH build,
the warp-packet traversals, the persistent AO ray stream (leaf-queue and leaf-vote kernels, every stack layout), the AO
prebaker, PPLL gather / resolve (plain and binned), depth cues, tile sharding.  A parity failure here is a logic bug in
the kernels; what only a GPU can show (memory model, performance) stays with the -m gpu tests."""
import os

import numpy as np
import pytest

import linevis_b200 as lv
from linevis_b200 import scenes
from oracle import lvo

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ectx():
    if not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("CUDA headers not found")
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(HERE, "emu", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    c = lv.Context(0, lib_path=mod.build())
    yield c
    c.close()


def _helix(n_lines=10, n_pts=25):
    return scenes.helix_lines(n_lines, n_pts), 0.012


def _random(n=400):
    return scenes.random_segments(n, 0.05, seed=11), 0.01


def _pair(ectx, oracle, data, width):
    return ectx.create_scene(*data, width), oracle.scene(*data, width)


@pytest.mark.parametrize("leaf", [1, 4])
@pytest.mark.parametrize("maker", [_helix, _random])
def test_bvh_build_and_primary_hits(ectx, oracle, maker, leaf):
    data, width = maker()
    ectx.set_option("b200_bvh_leaf_size", leaf)
    try:
        sc, osc = _pair(ectx, oracle, data, width)
    finally:
        ectx.set_option("b200_bvh_leaf_size", 1)
    cam = lv.make_camera(72, 48)
    hits, st = ectx.trace_primary(sc, cam)
    ref, _ = osc.trace_primary(cam)
    assert np.array_equal(hits["t"].view(np.uint32), ref["t"].view(np.uint32))
    assert np.array_equal(hits["prim"], ref["prim"]) and np.array_equal(hits["kind"], ref["kind"])
    assert st["pixels_hit"] == int((ref["prim"] != 0xFFFFFFFF).sum()) > 20
    # the emitted tree, walked from the root: every record is referenced by exactly one reachable leaf
    nodes = sc.bvh_nodes()
    seen = np.zeros(sc.info()["n_seg"], int)
    todo = [0]
    while todo:
        nd = nodes[todo.pop()]
        for ref_w, cnt, mn in ((nd["lref"], nd["lcount"], nd["lmin"]), (nd["rref"], nd["rcount"], nd["rmin"])):
            if not np.isfinite(mn).all():
                assert int(ref_w) == (0x80000000 | sc.info()["n_seg"]) and cnt == 1     # absent child -> the dummy record behind the last one
            elif cnt:
                first = int(ref_w) & 0x07FFFFFF
                assert ((int(ref_w) >> 27) & 15) + 1 == int(cnt) <= leaf
                seen[first:first + int(cnt)] += 1
            else:
                todo.append(int(ref_w))
    assert (seen == 1).all()


@pytest.mark.parametrize("queue,stack,minb", [(True, 12, 0), (True, 1, 9), (True, 8, 8), (False, 0, 10), (False, 1, 9), (False, 12, 9), (False, 16, 8)])
@pytest.mark.parametrize("use_distance", [True, False])
def test_rtao_every_kernel_variant(ectx, oracle, queue, stack, minb, use_distance):
    data, width = _random()
    sc, osc = _pair(ectx, oracle, data, width)
    cam = lv.make_camera(56, 36)
    ectx.set_new_settings({"ambient_occlusion_samples_per_frame": 6, "ambient_occlusion_distance_based": use_distance,
                           "use_jittered_primary_rays": True, "ambient_occlusion_radius": 0.4,
                           "b200_ao_stack": stack, "b200_ao_queue": queue, "b200_ao_min_blocks": minb})
    try:
        ao, st = ectx.render_rtao(sc, cam, 0)
        ao2, _ = ectx.render_rtao(sc, cam, 1, out=ao.copy())
    finally:
        ectx.set_new_settings({"b200_ao_stack": 12, "b200_ao_queue": True, "b200_ao_min_blocks": 0, "ambient_occlusion_radius": 0.1})
    opts = lvo.default_options(ao_strength=1.0, ao_spp=6, ao_use_distance=int(use_distance), ao_jitter_primary=1, ao_radius=0.4)
    ref, ost = osc.render_rtao(cam, opts, 0)
    ref2, _ = osc.render_rtao(cam, opts, 1, ao=ref.copy())
    assert st["rays_ao"] == ost["rays_ao"] > 500
    assert np.array_equal(ao.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(ao2.view(np.uint32), ref2.view(np.uint32))      # running mean over frames


def test_rtao_deep_stack_spills_out_of_shared_memory(ectx, oracle):
    # a long chain of collinear segments gives a deep, one-sided LBVH: the 12 shared stack entries overflow into the local part
    n = 300
    pos = np.zeros((n + 1, 3), np.float32)
    pos[:, 0] = np.linspace(-0.25, 0.25, n + 1) ** 3 * 16
    seg = np.stack([np.arange(n), np.arange(1, n + 1)], axis=1).astype(np.uint32)
    data, width = (pos, np.linspace(0, 1, n + 1).astype(np.float32), seg), 0.05
    sc, osc = _pair(ectx, oracle, data, width)
    cam = lv.make_camera(64, 24)
    ectx.set_new_settings({"ambient_occlusion_samples_per_frame": 4, "ambient_occlusion_distance_based": True, "ambient_occlusion_radius": 1.0,
                           "use_jittered_primary_rays": False})
    try:
        ao, st = ectx.render_rtao(sc, cam, 0)
    finally:
        ectx.set_new_settings({"ambient_occlusion_radius": 0.1, "use_jittered_primary_rays": True})
    ref, _ = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=4, ao_radius=1.0, ao_jitter_primary=0), 0)
    assert st["pixels_hit"] > 30 and np.array_equal(ao.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("mode", ["plain", "ao", "transparent_jitter", "depth_cues"])
def test_tube_frames(ectx, oracle, mode):
    data, width = _helix()
    sc, osc = _pair(ectx, oracle, data, width)
    cam = lv.make_camera(72, 48)
    opaque = mode != "transparent_jitter"
    tf = scenes.standard_transfer_function(opacity=(1.0, 1.0) if opaque else (0.3, 0.8))
    ectx.set_transfer_function(tf)
    settings = {"ambient_occlusion_strength": 1.0 if mode == "ao" else 0.0, "ambient_occlusion_samples_per_frame": 4,
                "use_jittered_primary_rays": True, "ambient_occlusion_distance_based": True, "depth_cue_strength": 0.8 if mode == "depth_cues" else 0.0,
                "num_samples_per_frame": 2 if mode == "transparent_jitter" else 1, "num_accumulated_frames": 4 if mode == "transparent_jitter" else 1}
    ectx.set_new_settings(settings)
    opts = lvo.default_options(ao_strength=settings["ambient_occlusion_strength"], ao_spp=4, depth_cue_strength=settings["depth_cue_strength"],
                               num_samples_per_frame=settings["num_samples_per_frame"], use_jittered_rays=int(mode == "transparent_jitter"))
    try:
        img, st = ectx.render_tubes(sc, cam, 0)
        ao = osc.render_rtao(cam, opts, 0)[0] if mode == "ao" else None
        ref, _ = osc.render_tubes(cam, opts, tf, ao_tex=ao)
        assert np.array_equal(img.view(np.uint32), ref.view(np.uint32))
        if mode == "transparent_jitter":      # second accumulated frame
            img2, _ = ectx.render_tubes(sc, cam, 1, out=img.copy())
            ref2, _ = osc.render_tubes(cam, opts, tf, frame_number=1, rgba=ref.copy())
            assert np.array_equal(img2.view(np.uint32), ref2.view(np.uint32))
    finally:
        ectx.set_new_settings({"ambient_occlusion_strength": 0.0, "depth_cue_strength": 0.0, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    assert st["pixels_hit"] > 50


@pytest.mark.parametrize("variant", ["plain", "binned", "reg_sort", "tile256", "reg_sort+tile512", "raster", "raster_contiguous", "raster_contiguous+reg_sort",
                                     "raster_contiguous+binned"])
@pytest.mark.parametrize("sort_mode", ["priority_queue", "bitonic"])
def test_ppll(ectx, oracle, variant, sort_mode):
    # dense enough for every list-length class of the resolve kernels (insertion <= 64, warp bitonic above -- in shared memory, or in
    # registers with 4 / 8 keys per lane; binned 32 / 64 / 128 / 256)
    binned = "binned" in variant
    data = scenes.random_segments(5000, 0.35, seed=13)
    sc, osc = _pair(ectx, oracle, data, 0.03)
    cam = lv.make_camera(48, 32)
    tf = scenes.standard_transfer_function(opacity=(0.2, 0.7))
    ectx.set_transfer_function(tf)
    ectx.set_new_settings({"ambient_occlusion_strength": 0.0, "b200_ppll_binned_resolve": binned, "b200_ppll_reg_sort": "reg_sort" in variant,
                           "b200_ppll_resolve_tile": 256 if "tile256" in variant else (512 if "tile512" in variant else 1024),
                           "b200_ppll_gather_mode": variant.split("+")[0] if variant.startswith("raster") else "raycast"})
    try:
        img, st = ectx.render_ppll(sc, cam, max_frags=200, sort_mode=sort_mode, linked_list_size=64 * 48 * 32)
    finally:
        ectx.set_new_settings({"b200_ppll_binned_resolve": False, "b200_ppll_reg_sort": False, "b200_ppll_resolve_tile": 1024, "b200_ppll_gather_mode": "raycast"})
    opts = lvo.default_options()
    g = osc.ppll_gather(cam, opts, tf)
    mine = ectx.ppll_read()
    assert st["frags_generated"] == g["counter"] == mine["counter"] > 1000
    assert lvo.per_pixel_lists(mine["heads"], mine["nodes"], cam, opts, oracle) == lvo.per_pixel_lists(g["heads"], g["nodes"], cam, opts, oracle)
    ref, rst = lvo.ppll_resolve(oracle, cam, opts, g["heads"], g["nodes"], 200, lv.SORT_MODES[sort_mode], canonical=True)
    assert st["frags_sorted"] == rst["frags_sorted"] and st["max_depth_complexity"] == rst["max_depth_complexity"] > 130
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(img), nan) and np.array_equal(img[~nan].view(np.uint32), ref[~nan].view(np.uint32))


@pytest.mark.parametrize("mode", ["raster", "raster_contiguous"])
@pytest.mark.parametrize("eye_z", [0.8, 0.1, 0.0])
def test_ppll_raster_gather(ectx, oracle, eye_z, mode):
    """b200_ppll_gather_mode = raster: the object-order gather produces the ray-cast gather's fragments -- per pixel the same multiset of
    (colour, depth bits) -- also with the camera inside the data (segments at and behind the eye plane get no screen bound and fall back
    to the whole frame), and its overflow behaviour is the reference's (dropped, counted)."""
    data = scenes.random_segments(1500, 0.35, seed=21)
    sc, osc = _pair(ectx, oracle, data, 0.02)
    cam = lv.make_camera(56, 40, eye=(0.03, -0.02, eye_z))
    tf = scenes.standard_transfer_function(opacity=(0.2, 0.7))
    ectx.set_transfer_function(tf)
    opts = lvo.default_options(use_capped_tubes=int(eye_z != 0.1), use_halos=int(eye_z != 0.0))
    ectx.set_new_settings({"ambient_occlusion_strength": 0.0, "b200_ppll_gather_mode": mode, "use_capped_tubes": eye_z != 0.1, "use_halos": eye_z != 0.0})
    try:
        size = 400 * 56 * 40
        img, st = ectx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic", linked_list_size=size)
        mine = ectx.ppll_read()
        small, sst = ectx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic", linked_list_size=1000)
    finally:
        ectx.set_new_settings({"b200_ppll_gather_mode": "raycast", "use_capped_tubes": True, "use_halos": True})
    g = osc.ppll_gather(cam, opts, tf, linked_list_size=size)
    assert st["frags_generated"] == g["counter"] == mine["counter"] > 1000 and st["frags_dropped"] == 0
    assert lvo.per_pixel_lists(mine["heads"], mine["nodes"], cam, opts, oracle) == lvo.per_pixel_lists(g["heads"], g["nodes"], cam, opts, oracle)
    if mode == "raster_contiguous":      # every list is one run of the node buffer, ending at its head, linked slot by slot
        heads, nxt = mine["heads"].astype(np.int64), mine["nodes"]["next"].astype(np.int64)
        NONE = 0xFFFFFFFF
        runs = sorted((h, ) for h in heads[heads != NONE])
        covered = 0
        for (h,) in runs:
            n = 0
            while nxt[h - n] != NONE:
                assert nxt[h - n] == h - n - 1
                n += 1
            covered += n + 1
        assert covered == mine["counter"]
    if st["max_depth_complexity"] <= 256:
        ref, _ = lvo.ppll_resolve(oracle, cam, opts, g["heads"], g["nodes"], 256, lv.SORT_MODES["bitonic"], canonical=True)
        nan = np.isnan(ref)
        assert np.array_equal(np.isnan(img), nan) and np.array_equal(img[~nan].view(np.uint32), ref[~nan].view(np.uint32))
    assert sst["frags_generated"] == st["frags_generated"] and sst["frags_stored"] == 1000 and sst["frags_dropped"] == st["frags_generated"] - 1000


def test_ppll_overflow_and_truncation(ectx, oracle):
    data, width = _random(600)
    sc = ectx.create_scene(*data, 0.02)
    cam = lv.make_camera(48, 32)
    ectx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.2, 0.7)))
    full, st_full = ectx.render_ppll(sc, cam, max_frags=64)
    img, st = ectx.render_ppll(sc, cam, max_frags=64, linked_list_size=500)     # fragment buffer too small: dropped, counted, never fatal
    assert st["frags_generated"] == st_full["frags_generated"] and st["frags_stored"] == 500 and st["frags_dropped"] == st["frags_generated"] - 500
    img, st = ectx.render_ppll(sc, cam, max_frags=4)                            # lists longer than MAX_NUM_FRAGS are truncated
    assert st["frags_truncated"] > 0 and st["frags_sorted"] + st["frags_truncated"] == st_full["frags_sorted"]


def test_prebaker_and_static_lookup(ectx, oracle):
    d = scenes.helix_polylines(8, 21)
    width = 0.012
    sc = ectx.create_scene(d["pos"], d["attr"], d["seg"], width)
    sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"])
    osc = oracle.scene(d["pos"], d["attr"], d["seg"], width)
    osc.set_lines(d["tangent"], d["normal"])
    ectx.set_new_settings({"ambient_occlusion_mode": "RTAO (Prebaker)", "b200_prebaker_iterations": 2, "b200_prebaker_samples_per_frame": 3,
                           "b200_prebaker_subdivisions": 6, "b200_prebaker_param_segment_length": 0.03, "b200_prebaker_radius": 0.2,
                           "ambient_occlusion_strength": 0.9, "ambient_occlusion_gamma": 1.4, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    tf = scenes.standard_transfer_function(opacity=(0.5, 1.0))
    ectx.set_transfer_function(tf)
    cam = lv.make_camera(64, 40)
    try:
        bw, sl = oracle.ao_parametrize(d["pos"], d["line_offsets"], 0.03)
        ref_f = None
        for it in range(2):
            st = sc.ao_bake(1)
            ref_f, ost = osc.ao_bake_iteration(sl, it, factors=ref_f, radius=0.2, n_subdiv=6, spp=3)
            got = sc.ao_read()
            assert st["rays_ao"] == ost["rays"] == len(sl) * 6 * 3
            assert np.array_equal(got["factors"].reshape(-1).view(np.uint32), ref_f.view(np.uint32)), it
        assert np.array_equal(got["blending_weights"], bw) and np.array_equal(got["sampling_locations"], sl)
        img, st = ectx.render_tubes(sc, cam)
        assert st["rays_ao"] == 0
        osc.set_static_ao(ref_f, 6, bw)
        ref, _ = osc.render_tubes(cam, lvo.default_options(ao_strength=0.9, ao_gamma=1.4, use_static_ao=1), tf)
        assert np.array_equal(img.view(np.uint32), ref.view(np.uint32))
    finally:
        ectx.set_new_settings({"ambient_occlusion_mode": "RTAO (Screen Space)", "ambient_occlusion_strength": 0.0, "ambient_occlusion_gamma": 1.0})


def test_tile_shards_union_is_the_frame(ectx, oracle):
    data, width = _helix()
    sc = ectx.create_scene(*data, width)
    cam = lv.make_camera(96, 64)
    tf = scenes.standard_transfer_function(opacity=(1.0, 1.0))
    ectx.set_transfer_function(tf)
    ectx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 2, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    try:
        full, _ = ectx.render_tubes(sc, cam)
        acc = np.full_like(full, np.nan)
        rays = 0
        for r in range(3):
            ectx.set_tile_shard(r, 3, 16)
            part = np.full_like(full, np.nan)
            _, st = ectx.render_tubes(sc, cam, out=part)
            m = ~np.isnan(part[..., 0])
            assert not (m & ~np.isnan(acc[..., 0])).any()        # shards are disjoint
            acc[m] = part[m]
            rays += st["rays_primary"] + st["rays_ao"]
        assert not np.isnan(acc).any()
    finally:
        ectx.set_tile_shard(0, 1, 64)
        ectx.set_option("ambient_occlusion_strength", 0.0)
    # without jitter the AO lookup of a tile's border pixels may touch a neighbour shard's texel with weight ~1e-4
    assert np.abs(acc - full).max() < 2e-3 and (acc == full).mean() > 0.98


def test_balanced_tile_owners_union_is_the_frame(ectx, oracle):
    """lv_get_tile_costs / lv_set_tile_owners: the cost map of a frame (hit pixels per tile, Morton order) adds up to the frame's hit
    pixels; with the longest-processing-time ownership built from it (sharding.balance_tiles) the ranks' shards are still disjoint, cover
    the frame bit for bit (jittered AO rays: the apron makes sharded frames exact) and their AO-ray counts are closer than round-robin's."""
    from linevis_b200 import sharding
    data, width = _helix()
    sc = ectx.create_scene(*data, width)
    cam = lv.make_camera(96, 64)
    ectx.set_transfer_function(scenes.standard_transfer_function(opacity=(1.0, 1.0)))
    ectx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 2, "num_samples_per_frame": 2, "num_accumulated_frames": 1})
    world, tile = 3, 16
    try:
        ectx.set_tile_shard(0, 1, tile)
        full, fst = ectx.render_tubes(sc, cam)
        costs = ectx.tile_costs(96, 64)
        assert costs.size == 24 and 2 * int(costs.sum()) == fst["rays_ao"] > 0
        owners = sharding.balance_tiles(costs, world, tile * tile, 2)
        assert owners.shape == (24,) and set(owners.tolist()) == {0, 1, 2}

        def shards(own):
            acc = np.full_like(full, np.nan)
            rays = []
            for r in range(world):
                ectx.set_tile_shard(r, world, tile)
                if own is not None:
                    ectx.set_tile_owners(96, 64, own)
                    want = {(int(x), int(y)) for (x, y), o in zip(sharding.all_tiles(96, 64, tile), own) if o == r}
                    assert {tuple(t) for t in ectx.owned_tiles(96, 64).tolist()} == want
                part = np.full_like(full, np.nan)
                _, st = ectx.render_tubes(sc, cam, out=part)
                m = ~np.isnan(part[..., 0])
                assert not (m & ~np.isnan(acc[..., 0])).any()
                acc[m] = part[m]
                rays.append(st["rays_ao"])
            return acc, rays
        acc_rr, rays_rr = shards(None)
        acc_b, rays_b = shards(owners)
        assert np.array_equal(acc_b.view(np.uint32), full.view(np.uint32)) and np.array_equal(acc_rr.view(np.uint32), full.view(np.uint32))
        assert max(rays_b) - min(rays_b) <= max(rays_rr) - min(rays_rr)
        with pytest.raises(lv.LineVisError):
            ectx.set_tile_owners(96, 64, owners[:-1])
    finally:
        ectx.set_tile_shard(0, 1, 64)
        ectx.set_new_settings({"ambient_occlusion_strength": 0.0, "num_samples_per_frame": 1})


def _host_floats(addr, n):
    """numpy view of n floats at a 'device' address of the emulated library (host memory)"""
    import ctypes
    return np.ctypeslib.as_array((ctypes.c_float * max(n, 1)).from_address(addr))[:n]


@pytest.mark.parametrize("world,jitter", [(1, False), (2, True), (4, False)])
def test_ao_sample_batch_shards(ectx, oracle, world, jitter):
    """lv_sao_primary / lv_sao_trace / lv_sao_finish: the ranks' tiles + every rank tracing spp / world samples of ALL hit pixels, with the
    exchanges done by hand (the emulated 'device' memory is host memory) -- the union of the ranks' frames is the one-context frame bit
    for bit (the per-sample values are summed in sample order by the pixel's owner), and every rank traces the same number of rays."""
    data, width = _helix()
    sc = ectx.create_scene(*data, width)
    cam = lv.make_camera(96, 64)
    spp, tile = 8, 16
    spl = spp // world
    ectx.set_transfer_function(scenes.standard_transfer_function(opacity=(1.0, 1.0)))
    ectx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": spp, "ambient_occlusion_iterations": 1,
                           "num_samples_per_frame": 2 if jitter else 1, "num_accumulated_frames": 1})
    try:
        full, fst = ectx.render_tubes(sc, cam)
        # stage 1 on every rank (one context plays the ranks in turn, so the hit lists are copied out)
        lists = []
        for r in range(world):
            ectx.set_tile_shard(r, world, tile)
            ptr, n = ectx.sao_primary(sc, cam, 0)
            lists.append(_host_floats(ptr, n * 12).reshape(n, 12).copy())
        counts = [h.shape[0] for h in lists]
        hits = np.ascontiguousarray(np.concatenate(lists, axis=0))
        # stage 2: rank r traces samples [r spl, (r + 1) spl) of all records
        occs = []
        for r in range(world):
            ectx.set_tile_shard(r, world, tile)
            occ = np.zeros(hits.shape[0] * spl, np.float32)
            ectx.sao_trace(sc, cam, 0, hits, hits.shape[0], r * spl, spl, occ)
            ectx.synchronize()
            occs.append(occ.reshape(hits.shape[0], spl))
        # stage 3: the owner gets its records' values from every rank (the all-to-all) and finishes its tiles.  The context's own hit list
        # must be the owner's again, so stage 1 is repeated for it (deterministic up to the order of the list: match rows by pixel)
        acc = np.full_like(full, np.nan)
        off = np.concatenate([[0], np.cumsum(counts)])
        for r in range(world):
            ectx.set_tile_shard(r, world, tile)
            ptr, n = ectx.sao_primary(sc, cam, 0)
            assert n == counts[r]
            now = _host_floats(ptr, n * 12).reshape(n, 12)
            pix_now, pix_then = now[:, 7].view(np.uint32), lists[r][:, 7].view(np.uint32)
            order = np.argsort(pix_then, kind="stable")[np.searchsorted(np.sort(pix_then), pix_now)]   # row of `lists[r]` holding now[i]'s pixel
            assert np.array_equal(pix_then[order], pix_now)
            parts = np.ascontiguousarray(np.stack([occs[j][off[r]:off[r + 1]][order] for j in range(world)], axis=0))
            part = np.full_like(full, np.nan)
            ectx.sao_finish(sc, cam, 0, parts, world, part)
            m = ~np.isnan(part[..., 0])
            assert not (m & ~np.isnan(acc[..., 0])).any()
            acc[m] = part[m]
        assert not np.isnan(acc).any()
        if jitter or world == 1:   # with jittered tube rays the apron makes sharded frames exact; without, border texels differ by ~1e-4
            assert np.array_equal(acc.view(np.uint32), full.view(np.uint32))
        else:
            assert np.abs(acc - full).max() < 2e-3 and (acc == full).mean() > 0.98
    finally:
        ectx.set_tile_shard(0, 1, 64)
        ectx.set_new_settings({"ambient_occlusion_strength": 0.0, "num_samples_per_frame": 1, "num_accumulated_frames": 1, "ambient_occlusion_samples_per_frame": 4,
                               "ambient_occlusion_iterations": 64})


def test_ao_sample_batch_stage_errors(ectx):
    """The staged entry points refuse what they cannot do: sample ranges that are not one of spp / count equal batches, a part count that
    does not divide spp, lv_sao_finish before lv_sao_primary of that frame size, the stages with screen-space RTAO switched off."""
    data, width = _helix()
    sc = ectx.create_scene(*data, width)
    cam = lv.make_camera(48, 32)
    ectx.set_transfer_function(scenes.standard_transfer_function())
    ectx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 6})
    try:
        ptr, n = ectx.sao_primary(sc, cam, 0)
        assert n > 0
        occ = np.zeros(n * 6, np.float32)
        hits = _host_floats(ptr, n * 12).reshape(n, 12).copy()
        for first, count in ((0, 4), (1, 2), (4, 3), (0, 0), (6, 6)):
            with pytest.raises(lv.LineVisError):
                ectx.sao_trace(sc, cam, 0, hits, n, first, count, occ)
        ectx.sao_trace(sc, cam, 0, hits, n, 0, 6, occ); ectx.synchronize()
        out = np.zeros((32, 48, 4), np.float32)
        with pytest.raises(lv.LineVisError):
            ectx.sao_finish(sc, cam, 0, occ, 4, out)                       # 6 samples in 4 parts
        with pytest.raises(lv.LineVisError):
            ectx.sao_finish(sc, lv.make_camera(40, 24), 0, occ, 1, np.zeros((24, 40, 4), np.float32))   # no primary pass of that frame size
        img, _ = ectx.sao_finish(sc, cam, 0, occ, 1, out)
        full, _ = ectx.render_tubes(sc, cam)
        assert np.array_equal(img.view(np.uint32), full.view(np.uint32))
        ectx.set_option("ambient_occlusion_strength", 0.0)
        with pytest.raises(lv.LineVisError):
            ectx.sao_primary(sc, cam, 0)
    finally:
        ectx.set_new_settings({"ambient_occlusion_strength": 0.0, "ambient_occlusion_samples_per_frame": 4})


def _sample_shard_worker(rank, world, port, lib_path, q):
    import torch
    import torch.distributed as dist
    from linevis_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = lv.Context(0, lib_path=lib_path)
        data, width = _helix()
        sc = c.create_scene(*data, width)
        cam = lv.make_camera(96, 64)
        c.set_transfer_function(scenes.standard_transfer_function(opacity=(1.0, 1.0)))
        c.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 8, "ambient_occlusion_iterations": 1,
                            "num_samples_per_frame": 2, "num_accumulated_frames": 1})
        full = None
        if rank == 0:
            full, _ = c.render_tubes(sc, cam)
        c.set_tile_shard(rank, world, 16)
        ss = sharding.SampleShards(c, rank, world, 8, "cpu")
        part = np.full((64, 96, 4), np.nan, np.float32)
        _, st = ss.render(sc, cam, 0, part, stats=True)
        parts = [None] * world
        dist.all_gather_object(parts, (part, st["rays_ao"]))
        if rank == 0:
            acc = np.full_like(full, np.nan)
            for p, _ in parts:
                m = ~np.isnan(p[..., 0])
                assert not (m & ~np.isnan(acc[..., 0])).any()
                acc[m] = p[m]
            q.put((bool(np.array_equal(acc.view(np.uint32), full.view(np.uint32))), [r for _, r in parts]))
        sc.close(); c.close()
    finally:
        dist.destroy_process_group()


def test_ao_sample_batch_shards_over_gloo(ectx):
    """sharding.SampleShards with two processes over gloo on the emulated library: counts all-gather, hit-list all-gather, per-sample
    all-to-all; the union of the two ranks' frames is the one-context frame bit for bit and both ranks trace the same number of AO rays."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sample_shard_worker, args=(r, 2, port, ectx.lib_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, rays = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and rays[0] == rays[1] > 0


@pytest.mark.parametrize("gather", ["raycast", "raster", "raster_contiguous"])
def test_ppll_tile_shards_union_is_the_frame(ectx, oracle, gather):
    """PPLL with tile sharding, both gather modes: every rank gathers and resolves only its tiles; the union is the unsharded frame bit
    for bit, and the ranks' fragment counts add up."""
    data = scenes.random_segments(1200, 0.35, seed=5)
    sc = ectx.create_scene(*data, 0.02)
    cam = lv.make_camera(80, 56)
    ectx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.2, 0.7)))
    ectx.set_new_settings({"ambient_occlusion_strength": 0.0, "b200_ppll_gather_mode": gather})
    size = 300 * 80 * 56
    try:
        full, fst = ectx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic", linked_list_size=size)
        acc = np.full_like(full, np.nan)
        frags = pixels = 0
        for r in range(3):
            ectx.set_tile_shard(r, 3, 16)
            part = np.full_like(full, np.nan)
            _, st = ectx.render_ppll(sc, cam, max_frags=256, sort_mode="bitonic", linked_list_size=size, out=part)
            owned = np.zeros(full.shape[:2], bool)
            for tx, ty in ectx.owned_tiles(80, 56):
                owned[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16] = True
            assert np.isnan(part[~owned]).all()                    # nothing outside the rank's tiles is written
            acc[owned] = part[owned]
            frags += st["frags_generated"]; pixels += st["rays_primary"]
    finally:
        ectx.set_tile_shard(0, 1, 64)
        ectx.set_option("b200_ppll_gather_mode", "raycast")
    assert frags == fst["frags_generated"] and pixels == fst["rays_primary"] == 80 * 56
    nan = np.isnan(full)
    assert np.array_equal(np.isnan(acc), nan) and np.array_equal(acc[~nan].view(np.uint32), full[~nan].view(np.uint32))


def _tube_scene(ectx, oracle, n_lines=10, n_pts=25, width=0.03):
    d = scenes.helix_polylines(n_lines, n_pts)
    sc = ectx.create_scene(d["pos"], d["attr"], d["seg"], width)
    sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"])
    osc = oracle.scene(d["pos"], d["attr"], d["seg"], width)
    osc.set_lines(d["tangent"], d["normal"])
    return d, sc, osc, width


@pytest.mark.parametrize("use_distance,jitter,n_sub", [(True, True, 6), (False, False, 6), (True, False, 8)])
def test_triangle_tube_rtao(ectx, oracle, use_distance, jitter, n_sub):
    """b200_rtao_geometry = triangles: the AO pass against the reference's own tube mesh (mesh generator on the host, LBVH over the
    triangles, packet closest hit + barycentric fetch, AO ray stream with the triangle test) equals the oracle's restatement."""
    d, sc, osc, width = _tube_scene(ectx, oracle)
    tm = lvo.TubeMesh(oracle, d["pos"], d["line_offsets"], width, n_sub)
    cam = lv.make_camera(72, 48)
    ectx.set_new_settings({"b200_rtao_geometry": "triangles", "tube_num_subdivisions": n_sub, "ambient_occlusion_samples_per_frame": 5,
                           "ambient_occlusion_distance_based": use_distance, "use_jittered_primary_rays": jitter, "ambient_occlusion_radius": 0.3})
    try:
        ao, st = ectx.render_rtao(sc, cam, 0)
        ao2, _ = ectx.render_rtao(sc, cam, 1, out=ao.copy())
    finally:
        ectx.set_new_settings({"b200_rtao_geometry": "capsules", "tube_num_subdivisions": 6, "ambient_occlusion_radius": 0.1, "use_jittered_primary_rays": True})
    opts = lvo.default_options(ao_strength=1.0, ao_spp=5, ao_use_distance=int(use_distance), ao_jitter_primary=int(jitter), ao_radius=0.3,
                               tube_num_subdivisions=n_sub)
    ref, ost = tm.render_rtao(cam, opts, 0)
    ref2, _ = tm.render_rtao(cam, opts, 1, ao=ref.copy())
    assert st["pixels_hit"] == ost["pixels_hit"] > 80 and st["rays_ao"] == ost["rays_ao"]
    assert np.array_equal(ao.view(np.uint32), ref.view(np.uint32)) and np.array_equal(ao2.view(np.uint32), ref2.view(np.uint32))
    # ... and it is a different image than the analytic stand-in's
    cap, _ = osc.render_rtao(cam, opts, 0)
    assert np.abs(cap - ref).max() > 0.05


def test_triangle_tube_frame_and_prebaker(ectx, oracle):
    d, sc, osc, width = _tube_scene(ectx, oracle, 8, 21)
    tm = lvo.TubeMesh(oracle, d["pos"], d["line_offsets"], width, 6)
    cam = lv.make_camera(64, 40)
    tf = scenes.standard_transfer_function(opacity=(1.0, 1.0))
    ectx.set_transfer_function(tf)
    ectx.set_new_settings({"b200_rtao_geometry": "triangles", "ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 4,
                           "num_samples_per_frame": 1, "num_accumulated_frames": 1, "ambient_occlusion_radius": 0.2})
    try:
        # the tube pass itself stays analytic (as in the reference's analytic-intersection mode); only the AO texture comes from the mesh
        img, st = ectx.render_tubes(sc, cam, 0)
        opts = lvo.default_options(ao_strength=1.0, ao_spp=4, ao_radius=0.2)
        ao, _ = tm.render_rtao(cam, opts, 0)
        ref, _ = osc.render_tubes(cam, opts, tf, ao_tex=ao)
        assert np.array_equal(img.view(np.uint32), ref.view(np.uint32))
        # prebaker: same start frames, rays against the mesh
        ectx.set_new_settings({"ambient_occlusion_mode": "RTAO (Prebaker)", "b200_prebaker_iterations": 2, "b200_prebaker_samples_per_frame": 3,
                               "b200_prebaker_subdivisions": 6, "b200_prebaker_param_segment_length": 0.03, "b200_prebaker_radius": 0.2})
        bw, sl = oracle.ao_parametrize(d["pos"], d["line_offsets"], 0.03)
        ref_f = None
        for it in range(2):
            bst = sc.ao_bake(1)
            ref_f, ost = osc.ao_bake_iteration(sl, it, factors=ref_f, radius=0.2, n_subdiv=6, spp=3, tube_mesh=tm)
            assert bst["rays_ao"] == ost["rays"]
            assert np.array_equal(sc.ao_read()["factors"].reshape(-1).view(np.uint32), ref_f.view(np.uint32)), it
        capsule_f, _ = osc.ao_bake_iteration(sl, 0, radius=0.2, n_subdiv=6, spp=3)
        first_f, _ = osc.ao_bake_iteration(sl, 0, radius=0.2, n_subdiv=6, spp=3, tube_mesh=tm)
        assert not np.array_equal(capsule_f, first_f)
    finally:
        ectx.set_new_settings({"b200_rtao_geometry": "capsules", "ambient_occlusion_mode": "RTAO (Screen Space)", "ambient_occlusion_strength": 0.0,
                               "ambient_occlusion_radius": 0.1})



@pytest.mark.parametrize("n_sub,ao,jitter", [(6, False, False), (8, True, False), (6, True, True)])
def test_triangle_geometry_mode_of_the_tube_pass(ectx, oracle, n_sub, ao, jitter):
    """geometry_mode = "Triangle Mesh" (RayTracingGeometryMode::TRIANGLE_MESH): the tube pass traces the reference's triangulated tubes and
    shades with ClosestHitTubeTriangles (barycentric normal / tangent / attribute, cap flag) -- frame bit-exact against the oracle's
    restatement, with the RTAO texture traced against the same mesh, with jittered multi-sample frames and accumulation."""
    d, sc, osc, width = _tube_scene(ectx, oracle, 8, 21)
    tm = lvo.TubeMesh(oracle, d["pos"], d["line_offsets"], width, n_sub)
    cam = lv.make_camera(64, 40)
    tf = scenes.standard_transfer_function(opacity=(0.4, 1.0))
    ctx = ectx
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

def test_triangle_mode_needs_polylines(ectx):
    data, width = _helix()
    sc = ectx.create_scene(*data, width)
    ectx.set_option("b200_rtao_geometry", "triangles")
    try:
        with pytest.raises(lv.LineVisError) as e:
            ectx.render_rtao(sc, lv.make_camera(32, 24), 0)
        assert e.value.code == -6 and "lv_scene_set_lines" in str(e.value)
        with pytest.raises(lv.LineVisError):
            ectx.set_option("b200_rtao_geometry", "nurbs")
    finally:
        ectx.set_option("b200_rtao_geometry", "capsules")


@pytest.mark.parametrize("use_distance", [True, False])
def test_rtao_quantised_nodes(ectx, oracle, use_distance):
    """b200_ao_qnodes (experimental): the AO ray stream over 32-byte nodes with 16-bit quantised, outward-rounded child boxes gives the
    same AO image -- the accepted set is decided by the records' exact AABBs, an enclosing box is never missed (DESIGN.md rule 2)."""
    for data, width in (_random(), _helix(), ((np.array([[-0.2, 0, 0], [0.2, 0.05, 0]], np.float32), np.array([0.1, 0.9], np.float32),
                                                np.array([[0, 1]], np.uint32)), 0.05)):
        sc, osc = _pair(ectx, oracle, data, width)
        cam = lv.make_camera(56, 36)
        ectx.set_new_settings({"ambient_occlusion_samples_per_frame": 6, "ambient_occlusion_distance_based": use_distance,
                               "ambient_occlusion_radius": 0.4, "b200_ao_qnodes": True})
        try:
            ao, st = ectx.render_rtao(sc, cam, 0)
        finally:
            ectx.set_new_settings({"b200_ao_qnodes": False, "ambient_occlusion_radius": 0.1})
        ref, ost = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=6, ao_use_distance=int(use_distance), ao_radius=0.4), 0)
        assert st["rays_ao"] == ost["rays_ao"] and np.array_equal(ao.view(np.uint32), ref.view(np.uint32))


def test_host_sah_builder(ectx, oracle):
    """b200_bvh_builder = sah: the binned-SAH tree built on the host (lv_sah_host.hpp) is a valid tree for every traversal -- closest hits,
    AO image (child-pair nodes and the 4-wide collapse of them) and PPLL fragments equal the oracle's bit for bit -- and it really is
    another tree than the Morton radix tree (other step counts)."""
    for data, width in (_random(), _helix(), _random(3)):
        ectx.set_option("b200_bvh_builder", "sah")
        try:
            sc = ectx.create_scene(*data, width)
        finally:
            ectx.set_option("b200_bvh_builder", "lbvh")
        sc0, osc = _pair(ectx, oracle, data, width)
        cam = lv.make_camera(56, 36)
        hits, st = ectx.trace_primary(sc, cam)
        hits0, st0 = ectx.trace_primary(sc0, cam)
        assert np.array_equal(hits, hits0)
        ectx.set_new_settings({"ambient_occlusion_samples_per_frame": 5, "ambient_occlusion_radius": 0.4, "ambient_occlusion_distance_based": True})
        try:
            for wide in (True, False):
                ectx.set_option("b200_ao_wide", wide)
                ao, ast = ectx.render_rtao(sc, cam, 0)
                ao0, ast0 = ectx.render_rtao(sc0, cam, 0)
                assert ast["rays_ao"] == ast0["rays_ao"] and np.array_equal(ao.view(np.uint32), ao0.view(np.uint32))
                if data[2].shape[0] > 100:
                    assert ast["ao_traversal_steps"] != ast0["ao_traversal_steps"]
        finally:
            ectx.set_new_settings({"ambient_occlusion_radius": 0.1, "b200_ao_wide": True, "ambient_occlusion_samples_per_frame": 4})
        ref, ost = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=5, ao_radius=0.4, ao_use_distance=1), 0)
        assert np.array_equal(ao0.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("radius", [1, 16])
def test_ploc_builder(ectx, oracle, radius):
    """b200_bvh_builder = ploc: the tree built by parallel locally-ordered clustering is a valid BVH over the same records (every
    record referenced once, boxes enclose), and primary hits, AO image (wide tree collapsed from it) and PPLL fragments are the
    LBVH's / the oracle's bit for bit -- results do not depend on the tree (rule 2)."""
    for data, width in (_random(), _helix(), _random(2), _random(3)):
        ectx.set_new_settings({"b200_bvh_builder": "ploc", "b200_bvh_ploc_radius": radius})
        try:
            sc, osc = _pair(ectx, oracle, data, width)
        finally:
            ectx.set_new_settings({"b200_bvh_builder": "lbvh", "b200_bvh_ploc_radius": 16})
        n = data[2].shape[0]
        nodes = sc.bvh_nodes()
        assert len(nodes) == max(1, n - 1)
        seen, stack = np.zeros(n, np.int32), [0]
        while stack:
            nd = nodes[stack.pop()]
            for side in ("l", "r"):
                ref, cnt = int(nd[side + "ref"]), int(nd[side + "count"])
                if cnt:
                    seen[ref & 0x07FFFFFF] += 1
                else:
                    ch = nodes[ref]
                    assert (np.minimum(ch["lmin"], ch["rmin"]) >= nd[side + "min"]).all() and (np.maximum(ch["lmax"], ch["rmax"]) <= nd[side + "max"]).all()
                    stack.append(ref)
        assert (seen == 1).all()
        cam = lv.make_camera(56, 36)
        hits, _ = ectx.trace_primary(sc, cam); ref, _ = osc.trace_primary(cam)
        assert np.array_equal(hits["t"].view(np.uint32), ref["t"].view(np.uint32)) and np.array_equal(hits["prim"], ref["prim"])
        ectx.set_new_settings({"ambient_occlusion_samples_per_frame": 5, "ambient_occlusion_radius": 0.4, "ambient_occlusion_distance_based": True,
                               "use_jittered_primary_rays": True, "tube_num_subdivisions": 6, "b200_rtao_geometry": "capsules"})
        try:
            ao, st = ectx.render_rtao(sc, cam, 0)
        finally:
            ectx.set_new_settings({"ambient_occlusion_radius": 0.1})
        rao, ost = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=5, ao_radius=0.4), 0)
        assert st["rays_ao"] == ost["rays_ao"] and np.array_equal(ao.view(np.uint32), rao.view(np.uint32))


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("use_distance", [True, False])
def test_rtao_ray_batches(ectx, oracle, use_distance, wide):
    """b200_ao_raybuf: the AO stream generates its rays 32 at a time by the whole warp into a shared-memory batch and idle lanes take
    entries from it -- same rays, same AO image bit for bit (also when the stream is shorter than a batch, or ends inside one)."""
    for data, width, frame in ((_random(), 0.01, (56, 36)), (_helix(), 0.012, (56, 36)), (_random(3), 0.01, (24, 16))):
        data = data[0] if isinstance(data, tuple) and len(data) == 2 else data
        sc, osc = _pair(ectx, oracle, data, width)
        cam = lv.make_camera(*frame)
        for spp in (5, 1):
            ectx.set_new_settings({"ambient_occlusion_samples_per_frame": spp, "ambient_occlusion_distance_based": use_distance,
                                   "ambient_occlusion_radius": 0.4, "b200_ao_raybuf": True, "b200_ao_wide": wide, "b200_ao_refill_below": 32})
            try:
                ao, st = ectx.render_rtao(sc, cam, 0)
            finally:
                ectx.set_new_settings({"b200_ao_raybuf": True, "b200_ao_wide": True, "ambient_occlusion_radius": 0.1, "b200_ao_refill_below": 0})
            ref, ost = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=spp, ao_use_distance=int(use_distance), ao_radius=0.4), 0)
            assert st["rays_ao"] == ost["rays_ao"] and np.array_equal(ao.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("top,packed", [(0, True), (0, 4), (0, False), (85, True), (341, True)])
@pytest.mark.parametrize("use_distance", [True, False])
def test_rtao_wide_quantised_tree(ectx, oracle, use_distance, top, packed):
    """b200_ao_wide: the AO ray stream over the 4-wide quantised tree (NodeW4: collapse of the child-pair nodes, 16-bit outward-rounded
    boxes, magic-number dequantisation) gives the same AO image bit for bit, with the same number of rays, on a random soup, a helix
    and a single segment (a root with one real child)."""
    for data, width in (_random(), _helix(), _random(3), ((np.array([[-0.2, 0, 0], [0.2, 0.05, 0]], np.float32), np.array([0.1, 0.9], np.float32),
                                                           np.array([[0, 1]], np.uint32)), 0.05)):
        sc, osc = _pair(ectx, oracle, data, width)
        cam = lv.make_camera(56, 36)
        ectx.set_new_settings({"ambient_occlusion_samples_per_frame": 6, "ambient_occlusion_distance_based": use_distance,
                               "ambient_occlusion_radius": 0.4, "b200_ao_wide": True, "b200_ao_wide_top": top, "b200_ao_packed": bool(packed),
                               "b200_ao_tq_bits": 4 if packed == 4 else 0})
        try:
            ao, st = ectx.render_rtao(sc, cam, 0)
            ectx.set_option("b200_ao_wide", False)
            ao2, st2 = ectx.render_rtao(sc, cam, 0)
        finally:
            ectx.set_new_settings({"b200_ao_wide": True, "b200_ao_wide_top": 0, "ambient_occlusion_radius": 0.1, "b200_ao_packed": True, "b200_ao_tq_bits": 0})
        ref, ost = osc.render_rtao(cam, lvo.default_options(ao_strength=1.0, ao_spp=6, ao_use_distance=int(use_distance), ao_radius=0.4), 0)
        assert st["rays_ao"] == ost["rays_ao"] and np.array_equal(ao.view(np.uint32), ref.view(np.uint32))
        if data[2].shape[0] > 100:   # the wide tree really is another tree: fewer steps per ray than the child-pair nodes
            assert st["ao_traversal_steps"] < 0.75 * st2["ao_traversal_steps"], (st["ao_traversal_steps"], st2["ao_traversal_steps"])


@pytest.mark.parametrize("async_delivery", [False, True])
def test_rgba8_frame_format(ectx, oracle, async_delivery):
    """b200_frame_format = rgba8: the tube pass and the PPLL resolve deliver RGBA8 UNORM frames packed in their epilogues -- equal to
    packUnorm4x8 of the float frame, for a host frame (synchronous and b200_async_delivery + lv_synchronize) and over frame accumulation."""
    data, width = _helix()
    sc, osc = _pair(ectx, oracle, data, width)
    cam = lv.make_camera(64, 40)
    tf = scenes.standard_transfer_function(opacity=(0.3, 0.9))
    ectx.set_transfer_function(tf)
    ectx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": 4, "num_samples_per_frame": 1, "num_accumulated_frames": 2})
    try:
        f0, _ = ectx.render_tubes(sc, cam, 0)
        f1, _ = ectx.render_tubes(sc, cam, 1, out=f0.copy())
        pp, _ = ectx.render_ppll(sc, cam, max_frags=64, sort_mode="priority_queue")
        ectx.set_new_settings({"b200_frame_format": "rgba8", "b200_async_delivery": async_delivery})
        a0 = np.zeros((40, 64), np.uint32); a1 = np.zeros((40, 64), np.uint32); ap = np.zeros((40, 64), np.uint32)
        ectx.render_tubes(sc, cam, 0, out=a0)
        ectx.render_tubes(sc, cam, 1, out=a1)          # the running mean lives in the library's own float image
        ectx.render_ppll(sc, cam, max_frags=64, sort_mode="priority_queue", out=ap)
        ectx.synchronize()
    finally:
        ectx.set_new_settings({"b200_frame_format": "rgba32f", "b200_async_delivery": False, "ambient_occlusion_strength": 0.0, "num_accumulated_frames": 1})

    def pack(img):
        q = np.floor(np.clip(img, 0.0, 1.0).astype(np.float32) * np.float32(255.0) + np.float32(0.5)).astype(np.uint32)
        return q[..., 0] | (q[..., 1] << 8) | (q[..., 2] << 16) | (q[..., 3] << 24)
    assert np.array_equal(a0, pack(f0)) and np.array_equal(a1, pack(f1)) and not np.array_equal(a0, a1)
    assert np.array_equal(ap, pack(np.nan_to_num(pp)))


def test_frame_to_rgba8_and_library_owned_frames(ectx, oracle):
    """lv_frame_alloc / lv_frame_to_rgba8: rendering into a library-owned device frame and reading it back in the reference's
    RGBA8 UNORM output format (packUnorm4x8 per pixel); with a tile shard only the owned tiles are converted."""
    data, width = _helix()
    sc = ectx.create_scene(*data, width)
    cam = lv.make_camera(80, 48)
    ectx.set_transfer_function(scenes.standard_transfer_function(opacity=(0.4, 1.0)))
    ectx.set_new_settings({"ambient_occlusion_strength": 0.0, "num_samples_per_frame": 1, "num_accumulated_frames": 1})
    img, _ = ectx.render_tubes(sc, cam)
    frame = ectx.frame_alloc(80, 48)
    try:
        ectx.render_tubes(sc, cam, 0, out=frame, stats=False)
        got = ectx.frame_to_rgba8(frame, 80, 48)
        q = np.floor(np.clip(img, np.float32(0), np.float32(1)) * np.float32(255) + np.float32(0.5)).astype(np.uint32)
        want = q[..., 0] | (q[..., 1] << 8) | (q[..., 2] << 16) | (q[..., 3] << 24)
        assert np.array_equal(got, want) and len(np.unique(got)) > 20
        ectx.set_tile_shard(1, 2, 16)
        part = ectx.frame_to_rgba8(frame, 80, 48, out=np.zeros((48, 80), np.uint32))
        owned = np.zeros((48, 80), bool)
        for tx, ty in ectx.owned_tiles(80, 48):
            owned[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16] = True
        assert np.array_equal(part[owned], want[owned]) and (part[~owned] == 0).all()
        with pytest.raises(lv.LineVisError):
            ectx.frame_to_rgba8(img, 80, 48)          # a host float frame is refused
    finally:
        ectx.set_tile_shard(0, 1, 64)
        ectx.frame_free(frame)


def test_prebaker_vertex_slices_union_is_the_full_bake(ectx, oracle):
    """Multi-GPU baking: three contexts' worth of vertex slices (lv_ao_set_vertex_range), each baked for two iterations, put together
    equal the one-GPU bake bit for bit; slices leave the other vertices' factors untouched."""
    from linevis_b200.sharding import bake_vertex_range
    d = scenes.helix_polylines(6, 17)
    width = 0.012
    ectx.set_new_settings({"b200_prebaker_iterations": 2, "b200_prebaker_samples_per_frame": 2, "b200_prebaker_subdivisions": 6,
                           "b200_prebaker_param_segment_length": 0.03, "b200_prebaker_radius": 0.2})
    def make():
        sc = ectx.create_scene(d["pos"], d["attr"], d["seg"], width)
        sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"])
        return sc
    full = make()
    full.ao_bake(0)
    want = full.ao_read()
    n_param = want["n_param"]
    assert want["iterations_done"] == 2 and n_param > 20
    merged = np.zeros_like(want["factors"])
    rays = 0
    for r in range(3):
        first, count = bake_vertex_range(n_param, r, 3)
        sc = make()
        sc.ao_set_vertex_range(first, count)
        rays += sc.ao_bake(0)["rays_ao"]
        got = sc.ao_read()["factors"]
        assert np.array_equal(got[:first], np.zeros_like(got[:first])) and np.array_equal(got[first + count:], np.zeros_like(got[first + count:]))
        merged[first:first + count] = got[first:first + count]
    assert rays == 2 * n_param * 6 * 2
    assert np.array_equal(merged.view(np.uint32), want["factors"].view(np.uint32))
    assert [bake_vertex_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert bake_vertex_range(2, 3, 4) == (2, 0)


_HANG_SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
import linevis_b200 as lv
ctx = lv.Context(0, lib_path=%(lib)r)
# a zero-length segment next to a regular one that starts in the same point; with leaf size 2 the whole scene is ONE leaf and the root's
# right child is absent.  A hit on the zero-length segment has no tangent (0/0): its AO rays have NaN directions.
pos = np.array([[-0.05, -0.13, -0.09], [-0.05, -0.13, -0.09], [-0.05, -0.13, -0.09], [-0.02, -0.17, -0.10]], np.float32)
data = (pos, np.array([0.1, 0.2, 0.3, 0.4], np.float32), np.array([[0, 1], [2, 3]], np.uint32))
cam = lv.make_camera(43, 32, eye=(0.037, 0.027, 0.8))
for leaf, queue in ((2, True), (1, True), (1, False)):
    ctx.set_option("b200_bvh_leaf_size", leaf)
    sc = ctx.create_scene(*data, 0.04)
    ctx.set_new_settings({"ambient_occlusion_samples_per_frame": 3, "ambient_occlusion_radius": 0.05, "b200_ao_queue": queue})
    ao, st = ctx.render_rtao(sc, cam, 0)
    assert st["pixels_hit"] > 0 and st["rays_ao"] == 3 * st["pixels_hit"] and not np.isnan(ao).any()
one = (pos[:2], np.array([0.1, 0.2], np.float32), np.array([[0, 1]], np.uint32))     # a scene that is nothing but one zero-length segment
for queue in (False, True):   # EVERY AO ray is invalid here: each one still has to deliver its result (nothing hit), and the warps have to go on fetching
    c1 = lv.Context(0, lib_path=%(lib)r)   # a fresh context: its per-ray result buffer holds no results of earlier frames
    c1.set_new_settings({"ambient_occlusion_samples_per_frame": 3, "ambient_occlusion_radius": 0.05, "b200_ao_queue": queue})
    s1 = c1.create_scene(*one, 0.04)
    ao, st = c1.render_rtao(s1, cam, 0)
    assert st["pixels_hit"] > 0 and st["rays_ao"] == 3 * st["pixels_hit"] and np.array_equal(ao, np.ones_like(ao)), (queue, st, ao.min())
    s1.close(); c1.close()
sc = ctx.create_scene(*one, 0.04)
bad = lv.make_camera(16, 16)
bad.inv_view[5] = float("nan")
try:
    ctx.render_rtao(sc, bad, 0)
    raise SystemExit("NaN camera accepted")
except lv.LineVisError as e:
    assert e.code == -1
# a finite but degenerate camera (all-zero inverse projection): every primary ray has the direction normalize(0) = NaN.  The packet
# traversals then "hit" every box, the absent child's too -- which is a one-record leaf on an all-NaN dummy record, not a way back
# to the root: every pass terminates and sees nothing.
zero = lv.make_camera(24, 16)
for k in range(16):
    zero.inv_proj[k] = 0.0
ctx.set_transfer_function(np.array([[1, 1, 1, 1], [0, 0, 0, 1]], np.float32))
for leaf in (1, 2):
    ctx.set_option("b200_bvh_leaf_size", leaf)
    for d in (data, one):
        sc = ctx.create_scene(*d, 0.04)
        hits, st = ctx.trace_primary(sc, zero)
        assert st["pixels_hit"] == 0
        ao, st = ctx.render_rtao(sc, zero, 0)
        img, st = ctx.render_tubes(sc, zero, 0)
        img, st = ctx.render_ppll(sc, zero, max_frags=8)
        assert st["frags_generated"] == 0
print("ok")
'''


def test_nan_rays_cannot_hang_the_traversal(ectx):
    """Found by tools/fuzz_emu.py: a hit on a zero-length segment has a 0/0 tangent, its AO rays have NaN directions, and NaN passes every
    slab test -- also the one of an ABSENT child, whose word 0 points back at the root: an endless loop (on a GPU: a hung kernel) in
    single-leaf scenes.  Such rays now never enter the traversal (they hit nothing, which is also what the oracle's arithmetic
    gives), and NaN camera matrices are refused.  Run in a subprocess: a regression would hang, not fail."""
    import subprocess, sys
    root = os.path.dirname(HERE)
    script = _HANG_SCRIPT % {"root": root, "lib": os.path.join(HERE, "emu", "liblinevis_b200_emu.so")}
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_fuzz_smoke(ectx):
    """A short fixed-seed run of tools/fuzz_emu.py (random degenerate / duplicated / axis-aligned scenes, random cameras and settings,
    every pass and kernel variant incl. accumulation and tile shards) -- in a subprocess, because what it guards against includes hangs."""
    import subprocess, sys
    root = os.path.dirname(HERE)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_emu.py"), "--seed", "3", "--cases", "12", "--seconds", "140"],
                       capture_output=True, text=True, timeout=170)
    assert r.returncode == 0 and "12 cases, 0 mismatching" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
