"""Differential fuzzing of the product's kernels against the oracle WITHOUT a GPU: random small scenes (degenerate, duplicated and
axis-aligned segments included), random cameras / frame sizes / settings, every pass of the C ABI, run through the host emulation
of the library (tests/emu) and compared with the oracle bit for bit.  Prints the seed of any mismatch.

    python tools/fuzz_emu.py --seconds 300 [--seed 1]"""
import argparse, importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import linevis_b200 as lv
from linevis_b200 import scenes
from oracle import lvo

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=120)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--start", type=int, default=0, help="first case index (to replay one case: --start I --cases 1)")
ap.add_argument("--cases", type=int, default=0, help="stop after this many cases (0 = until --seconds)")
ap.add_argument("--basic", action="store_true", help="skip the accumulation / tile-shard phase")
ap.add_argument("--case-timeout", type=int, default=120, help="seconds before a case is reported as hanging (SIGALRM, exits)")
ap.add_argument("--verbose", action="store_true", help="print every case before it runs (a hang then shows its seed)")
args = ap.parse_args()
spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "emu", "build_emu.py"))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
LIB = mod.build()   # once: the sources may change under a long run
ctx = lv.Context(0, lib_path=LIB)
o = lvo.Oracle("own")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def same(a, b):
    nan = np.isnan(b)
    return np.array_equal(np.isnan(a), nan) and np.array_equal(bits(a[~nan]), bits(b[~nan]))


def random_scene(rng):
    kind = rng.integers(0, 4)
    if kind == 0:      # polylines (needed for prebaker / triangles)
        d = scenes.helix_polylines(int(rng.integers(1, 8)), int(rng.integers(3, 20)), seed=int(rng.integers(1, 1 << 30)))
        return d, (d["pos"], d["attr"], d["seg"])
    n = int(rng.integers(1, 300))
    p0 = (rng.random((n, 3)) - 0.5) * rng.choice([0.2, 0.5, 1.0])
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    if kind == 2:      # axis-aligned directions: rays / boxes with zero extents
        d = np.eye(3)[rng.integers(0, 3, n)] * rng.choice([-1, 1], (n, 1))
    p1 = p0 + d * rng.choice([0.01, 0.05, 0.3])
    if kind == 3:      # exact duplicates and zero-length segments: ties and degenerate capsules
        p0[n // 2:] = p0[:n - n // 2]; p1[n // 2:] = p1[:n - n // 2]
        p1[::7] = p0[::7]
    pos = np.empty((2 * n, 3), np.float32); pos[0::2], pos[1::2] = p0, p1
    return None, (pos, rng.random(2 * n).astype(np.float32), np.arange(2 * n, dtype=np.uint32).reshape(n, 2))


import signal


def _hang(signum, frame):
    print("HANG: case %d (seed %d) exceeded %d s" % (it - 1, args.seed * 1000003 + it - 1, args.case_timeout), flush=True)
    os._exit(3)


signal.signal(signal.SIGALRM, _hang)
t_end = time.time() + args.seconds
it, bad = args.start, 0
while time.time() < t_end and (args.cases == 0 or it < args.start + args.cases):
    seed = args.seed * 1000003 + it
    rng = np.random.default_rng(seed)
    it += 1
    signal.alarm(args.case_timeout)
    d, data = random_scene(rng)
    width = float(rng.choice([0.002, 0.01, 0.04]))
    W, H = int(rng.integers(8, 70)), int(rng.integers(8, 50))
    eye = (float(rng.normal(0, 0.1)), float(rng.normal(0, 0.1)), float(rng.choice([0.3, 0.8, 1.5])))
    cam = lv.make_camera(W, H, eye=eye)
    leaf = int(rng.choice([1, 1, 2, 4]))
    if args.verbose:
        print("case %d seed %d: n_seg %d, %dx%d, width %g, eye %s" % (it - 1, seed, data[2].shape[0], W, H, width, eye), flush=True)
    builder = str(rng.choice(["lbvh", "lbvh", "sah", "ploc"])) if leaf == 1 else "lbvh"
    ctx.set_option("b200_bvh_leaf_size", leaf); ctx.set_option("b200_bvh_builder", builder)
    sc = ctx.create_scene(*data, width); osc = o.scene(*data, width)
    ctx.set_option("b200_bvh_leaf_size", 1); ctx.set_option("b200_bvh_builder", "lbvh")
    spp = int(rng.integers(1, 7)); dist = bool(rng.integers(0, 2)); jit = bool(rng.integers(0, 2)); radius = float(rng.choice([0.05, 0.1, 0.5]))
    queue = bool(rng.integers(0, 2)); stack = int(rng.choice([0, 1, 8, 12, 16])); qn = bool(rng.integers(0, 2))
    wide = bool(rng.integers(0, 2)); raybuf = bool(rng.integers(0, 2)); wtop = int(rng.choice([0, 0, 85, 341])); wreps = int(rng.choice([1, 1, 2]))
    refill = int(rng.choice([24, 24, 28, 32, 8]))
    packed = bool(rng.integers(0, 3) > 0); tqb = int(rng.choice([0, 0, 4]))
    capped = bool(rng.integers(0, 4) > 0); halos = bool(rng.integers(0, 2))
    tf = scenes.standard_transfer_function(opacity=(float(rng.random()), float(rng.random())))
    ctx.set_transfer_function(tf)
    ctx.set_new_settings({"ambient_occlusion_samples_per_frame": spp, "ambient_occlusion_distance_based": dist, "use_jittered_primary_rays": jit,
                          "ambient_occlusion_radius": radius, "b200_ao_queue": queue, "b200_ao_stack": stack, "b200_ao_qnodes": qn,
                          "b200_ao_wide": wide, "b200_ao_raybuf": raybuf, "b200_ao_packed": packed, "b200_ao_tq_bits": tqb, "b200_tube_prepass": bool(rng.integers(0, 2)), "b200_ao_wide_top": wtop, "b200_ao_wide_reps": wreps, "b200_ao_refill_below": refill,
                          "use_capped_tubes": capped, "use_halos": halos, "ambient_occlusion_strength": 1.0, "num_samples_per_frame": 1,
                          "num_accumulated_frames": 1, "depth_cue_strength": 0.0, "b200_rtao_geometry": "capsules", "ambient_occlusion_mode": "RTAO (Screen Space)"})
    opts = lvo.default_options(ao_strength=1.0, ao_spp=spp, ao_use_distance=int(dist), ao_jitter_primary=int(jit), ao_radius=radius,
                               use_capped_tubes=int(capped), use_halos=int(halos))
    fails = []
    if args.verbose:
        print("   phase primary", flush=True)
    hits, _ = ctx.trace_primary(sc, cam); ref, _ = osc.trace_primary(cam, opts)
    if not (np.array_equal(bits(hits["t"]), bits(ref["t"])) and np.array_equal(hits["prim"], ref["prim"]) and np.array_equal(hits["kind"], ref["kind"])):
        fails.append("primary")
    if args.verbose:
        print("   phase rtao", flush=True)
    ao, _ = ctx.render_rtao(sc, cam, 0); rao, _ = osc.render_rtao(cam, opts, 0)
    if not same(ao, rao):
        fails.append("rtao")
    if args.verbose:
        print("   phase tubes", flush=True)
    img, _ = ctx.render_tubes(sc, cam, 0); rimg, _ = osc.render_tubes(cam, opts, tf, ao_tex=rao)
    if not same(img, rimg):
        fails.append("tubes")
    if spp % 2 == 0 and ctx.get_option("b200_ao_packed") == "true" and wide and raybuf and queue and stack == 12 and not qn and wtop == 0 and leaf == 1:
        # AO-sample-batch stages: one context plays two sample batches of the whole frame (lv_sao_*), exchanges by hand
        import ctypes
        ptr, n = ctx.sao_primary(sc, cam, 0)
        hl = np.ctypeslib.as_array((ctypes.c_float * max(n * 12, 1)).from_address(ptr))[:n * 12].reshape(n, 12).copy() if n else np.zeros((0, 12), np.float32)
        parts = np.zeros((2, n, spp // 2), np.float32)
        for part in range(2):
            occ = np.zeros(max(n * (spp // 2), 1), np.float32)
            ctx.sao_trace(sc, cam, 0, hl, n, part * (spp // 2), spp // 2, occ); ctx.synchronize()
            parts[part] = occ[:n * (spp // 2)].reshape(n, spp // 2)
        ptr2, n2 = ctx.sao_primary(sc, cam, 0)
        now = np.ctypeslib.as_array((ctypes.c_float * max(n2 * 12, 1)).from_address(ptr2))[:n2 * 12].reshape(n2, 12)
        if n2 != n:
            fails.append("sample shards: hit count")
        else:
            pt, pn = hl[:, 7].view(np.uint32), now[:, 7].view(np.uint32)
            order = np.argsort(pt, kind="stable")[np.searchsorted(np.sort(pt), pn)] if n else np.zeros(0, np.int64)
            simg, _ = ctx.sao_finish(sc, cam, 0, np.ascontiguousarray(parts[:, order]), 2, np.zeros((H, W, 4), np.float32))
            if not same(simg, rimg):
                fails.append("sample shards")
    ctx.set_option("ambient_occlusion_strength", 0.0)
    if args.verbose:
        print("   phase ppll", flush=True)
    mf = int(rng.choice([4, 32, 128]))
    mode = int(rng.choice([0, 5]))
    ctx.set_new_settings({"b200_ppll_reg_sort": bool(rng.integers(0, 2)), "b200_ppll_binned_resolve": bool(rng.integers(0, 3) == 0),
                          "b200_ppll_resolve_tile": int(rng.choice([256, 512, 1024])),
                          "b200_ppll_gather_mode": str(rng.choice(["raycast", "raster", "raster_contiguous"]))})
    pimg, pst = ctx.render_ppll(sc, cam, max_frags=mf, sort_mode=mode, linked_list_size=200 * W * H)
    po = lvo.default_options(use_capped_tubes=int(capped), use_halos=int(halos))
    g = osc.ppll_gather(cam, po, tf, linked_list_size=200 * W * H)
    rp, _ = lvo.ppll_resolve(o, cam, po, g["heads"], g["nodes"], mf, mode, canonical=True)
    if pst["frags_generated"] != g["counter"]:
        fails.append("ppll count %d vs %d" % (pst["frags_generated"], g["counter"]))
    elif mf >= pst["max_depth_complexity"] and not same(pimg, rp):
        fails.append("ppll frame")
    if d is not None:
        sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"]); osc.set_lines(d["tangent"], d["normal"])
        nsub = int(rng.choice([4, 6, 8]))
        ctx.set_new_settings({"b200_rtao_geometry": "triangles", "tube_num_subdivisions": nsub})
        tm = lvo.TubeMesh(o, d["pos"], d["line_offsets"], width, nsub)
        tao, _ = ctx.render_rtao(sc, cam, 0)
        opts.tube_num_subdivisions = nsub
        rtao, _ = tm.render_rtao(cam, opts, 0)
        if not same(tao, rtao):
            fails.append("triangles")
        ctx.set_new_settings({"b200_rtao_geometry": "capsules", "tube_num_subdivisions": 6, "ambient_occlusion_mode": "RTAO (Prebaker)",
                              "b200_prebaker_iterations": 1, "b200_prebaker_samples_per_frame": 2, "b200_prebaker_subdivisions": 6,
                              "b200_prebaker_param_segment_length": 0.05, "b200_prebaker_radius": radius, "b200_prebaker_distance_based": dist})
        sc.ao_bake(0)
        bw, sl = o.ao_parametrize(d["pos"], d["line_offsets"], 0.05)
        rf, _ = osc.ao_bake_iteration(sl, 0, radius=radius, n_subdiv=6, spp=2, use_distance=dist, capped=capped)
        if not np.array_equal(bits(sc.ao_read()["factors"].reshape(-1)), bits(rf)):
            fails.append("prebaker")
        ctx.set_option("ambient_occlusion_mode", "RTAO (Screen Space)")
    # accumulation over frames with jittered multi-sample rays + depth cues; then the same frame from tile shards (AO apron ring)
    if args.basic:
        if fails:
            bad += 1
            print("MISMATCH seed %d: %s" % (seed, fails), flush=True)
        sc.close()
        continue
    nspf = int(rng.integers(1, 4)); dcs = float(rng.choice([0.0, 0.8]))
    ctx.set_new_settings({"ambient_occlusion_mode": "RTAO (Screen Space)", "ambient_occlusion_strength": 1.0, "num_samples_per_frame": nspf,
                          "num_accumulated_frames": 3, "depth_cue_strength": dcs, "b200_ao_qnodes": False})
    o2 = lvo.default_options(ao_strength=1.0, ao_spp=spp, ao_use_distance=int(dist), ao_jitter_primary=int(jit), ao_radius=radius,
                             use_capped_tubes=int(capped), use_halos=int(halos), num_samples_per_frame=nspf, use_jittered_rays=1, depth_cue_strength=dcs)
    acc = np.zeros((H, W, 4), np.float32); racc = np.zeros((H, W, 4), np.float32); rao2 = np.zeros((H, W), np.float32)
    for f in range(2):
        acc, _ = ctx.render_tubes(sc, cam, f, out=acc)
        rao2, _ = osc.render_rtao(cam, o2, f, ao=rao2)
        racc, _ = osc.render_tubes(cam, o2, tf, ao_tex=rao2, frame_number=f, rgba=racc)
    if not same(acc, racc):
        fails.append("accumulated jittered frame")
    world = int(rng.integers(2, 4)); ts = int(rng.choice([16, 32]))
    merged = np.full((H, W, 4), np.nan, np.float32)
    for r in range(world):
        c2 = lv.Context(0, lib_path=LIB)
        c2.set_transfer_function(tf)
        c2.set_new_settings({"ambient_occlusion_samples_per_frame": spp, "ambient_occlusion_distance_based": dist, "use_jittered_primary_rays": jit,
                             "ambient_occlusion_radius": radius, "use_capped_tubes": capped, "use_halos": halos, "ambient_occlusion_strength": 1.0,
                             "num_samples_per_frame": nspf, "num_accumulated_frames": 3, "depth_cue_strength": dcs})
        c2.set_tile_shard(r, world, ts)
        s2 = c2.create_scene(*data, width)
        part = np.full((H, W, 4), np.nan, np.float32)
        part, _ = c2.render_tubes(s2, cam, 0, out=part)
        m = ~np.isnan(part[..., 0])
        merged[m] = part[m]
        s2.close(); c2.close()
    first, _ = ctx.render_tubes(sc, cam, 0)
    if np.isnan(merged).any() or not same(merged, first):
        fails.append("tile shards (world %d, tile %d)" % (world, ts))
    ctx.set_new_settings({"num_samples_per_frame": 1, "num_accumulated_frames": 1, "depth_cue_strength": 0.0})
    if fails:
        bad += 1
        print("MISMATCH seed %d: %s  (n_seg %d, %dx%d, width %g, leaf %d, spp %d dist %s jit %s radius %g queue %s stack %d qn %s capped %s)" %
              (seed, fails, data[2].shape[0], W, H, width, leaf, spp, dist, jit, radius, queue, stack, qn, capped), flush=True)
    sc.close()
signal.alarm(0)
print("fuzz: %d cases, %d mismatching" % (it - args.start, bad))
