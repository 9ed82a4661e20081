"""CPU tests of the oracle: known answers, literal-vs-canonical sort behaviour, and the cross-check of the oracle's own
BVH against the reference's madmann91/bvh library (oracle/_ref).  The reference holds no golden vectors for this path
("parity unpinned"); these are the pins this repo can provide."""
import math
import os
import subprocess

import numpy as np
import pytest

import linevis_b200 as lv
from linevis_b200 import scenes
from oracle import lvo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- RNG (integer exact) -------------------------------------------------------------------------------------------
def _tea_py(v0, v1):
    """Independent pure-Python restatement of RayTracingUtilities.glsl:134-149."""
    M = 0xFFFFFFFF
    s0 = 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & M
        v0 = (v0 + ((((v1 << 4) & M) + 0xA341316C) & M ^ ((v1 + s0) & M) ^ (((v1 >> 5) + 0xC8013EA4) & M))) & M
        v1 = (v1 + ((((v0 << 4) & M) + 0xAD90777D) & M ^ ((v0 + s0) & M) ^ (((v0 >> 5) + 0x7E95761E) & M))) & M
    return v0


def test_tea_matches_independent_python(oracle):
    rng = np.random.default_rng(0)
    for a, b in [(0, 0), (1, 2), (0xFFFFFFFF, 0xFFFFFFFF), (1920 * 1080 - 1, 63)] + [tuple(int(x) for x in rng.integers(0, 2**32, 2)) for _ in range(50)]:
        assert oracle.tea(a, b) == _tea_py(a, b)


def test_lcg_rnd_stream(oracle):
    l, f = oracle.rnd_stream(12345, 64)
    s = 12345
    for i in range(64):
        s = (1664525 * s + 1013904223) & 0xFFFFFFFF
        assert int(l[i]) == (s & 0x00FFFFFF)
        assert f[i] == np.float32(s & 0x00FFFFFF) / np.float32(16777216.0)
        assert 0.0 <= f[i] < 1.0


def test_golden_rng_vectors(oracle):
    g = np.load(os.path.join(GOLDEN, "rng_vectors.npz"))
    for (a, b), v in zip(g["tea_in"], g["tea_out"]):
        assert oracle.tea(int(a), int(b)) == int(v)
    l, f = oracle.rnd_stream(int(g["lcg_seed"]), len(g["lcg_out"]))
    assert np.array_equal(l, g["lcg_out"]) and np.array_equal(f, g["rnd_out"])


# ---- deterministic transcendental functions ------------------------------------------------------------------------
def test_det_pow_accuracy(oracle):
    rng = np.random.default_rng(1)
    for y in (1.7, 30.0, 1.0, 0.5, 2.2):
        for x in np.concatenate([rng.random(200), [1.0, 0.5, 1e-3, 1e-20, 0.999999]]):
            got = oracle.det_pow(float(np.float32(x)), y)
            want = float(np.float32(x)) ** y
            assert abs(got - want) <= 2e-5 * want + 1e-12, (x, y, got, want)   # float32 exp2(y*log2 x): abs err ~ 1e-7*|y log2 x|
    assert oracle.det_pow(0.0, 1.7) == 0.0 and oracle.det_pow(1.0, 30.0) == 1.0 and oracle.det_pow(-1.0, 2.0) == 0.0


def test_det_sincos_accuracy(oracle):
    for xi in np.linspace(0.0, 0.999999, 1001):
        c, s = oracle.det_sincos2pi(float(np.float32(xi)))
        a = 2.0 * math.pi * float(np.float32(xi))
        assert abs(c - math.cos(a)) < 5e-7 and abs(s - math.sin(a)) < 5e-7


def test_sample_hemisphere_unit_and_upper(oracle):
    rng = np.random.default_rng(2)
    for a, b in rng.random((200, 2)):
        v = oracle.sample_hemisphere(float(a), float(b))
        assert abs(np.linalg.norm(v) - 1.0) < 1e-6 and v[2] >= 0.0 and v[2] == np.float32(a)


# ---- intersection known answers ------------------------------------------------------------------------------------
def test_ray_cylinder_body_known_answer(oracle):
    # ray along -z through the axis of an x-aligned tube of radius 0.1: enters at z = 0.1 -> t = 0.9
    hit, t, kind = oracle.intersect_tube([0, 0, 1], [0, 0, -1], [-0.5, 0, 0], [0.5, 0, 0], 0.1)
    assert hit and kind == 0 and abs(t - 0.9) < 1e-6
    # offset by d: t = 1 - sqrt(r^2 - d^2)
    hit, t, kind = oracle.intersect_tube([0, 0.06, 1], [0, 0, -1], [-0.5, 0, 0], [0.5, 0, 0], 0.1)
    assert hit and kind == 0 and abs(t - (1 - math.sqrt(0.1**2 - 0.06**2))) < 1e-6
    # miss outside the radius
    assert not oracle.intersect_tube([0, 0.11, 1], [0, 0, -1], [-0.5, 0, 0], [0.5, 0, 0], 0.1)[0]


def test_ray_capsule_end_spheres(oracle):
    # beyond the segment end only the sphere at p1 is hit (hitKind 2), at p0 hitKind 1 (TubeRayTracing.glsl:479-488)
    hit, t, kind = oracle.intersect_tube([0.55, 0, 1], [0, 0, -1], [-0.5, 0, 0], [0.5, 0, 0], 0.1)
    assert hit and kind == 2 and abs(t - (1 - math.sqrt(0.1**2 - 0.05**2))) < 1e-6
    hit, t, kind = oracle.intersect_tube([-0.55, 0, 1], [0, 0, -1], [-0.5, 0, 0], [0.5, 0, 0], 0.1)
    assert hit and kind == 1
    # uncapped: nothing there
    assert not oracle.intersect_tube([0.55, 0, 1], [0, 0, -1], [-0.5, 0, 0], [0.5, 0, 0], 0.1, capped=False)[0]
    # looking down the axis: the open cylinder has no body hit, the cap sphere is hit at distance 1 - 0.5 - r
    hit, t, kind = oracle.intersect_tube([1.0, 0, 0], [-1, 0, 0], [-0.5, 0, 0], [0.5, 0, 0], 0.1)
    assert hit and kind == 2 and abs(t - 0.4) < 1e-6


def test_ray_origin_inside_and_behind(oracle):
    # origin inside the tube: far root is reported (RayIntersectionTestsVulkan.glsl:107-116)
    hit, t, kind = oracle.intersect_tube([0, 0, 0], [0, 0, -1], [-0.5, 0, 0], [0.5, 0, 0], 0.1)
    assert hit and kind == 0 and abs(t - 0.1) < 1e-6
    # tube entirely behind the ray
    assert not oracle.intersect_tube([0, 0, 1], [0, 0, 1], [-0.5, 0, 0], [0.5, 0, 0], 0.1)[0]


# ---- BVH plumbing pins from the reference's bvh library ------------------------------------------------------------
def test_reference_bvh_smoke_vectors():
    """submodules/bvh/test/{custom_primitive,simple_example}.cpp compiled unmodified from /root/reference
    (oracle/Makefile `ref`): their expected outputs are the only known answers the reference tree offers."""
    ref = os.path.join(os.path.dirname(lvo.REF_LIB))
    for exe, want in (("bvh_custom_primitive", "distance: 25"), ("bvh_simple_example", "distance: 1.5")):
        path = os.path.join(ref, exe)
        if not os.path.exists(path):
            pytest.skip("oracle/_ref not built")
        r = subprocess.run([path], capture_output=True, text=True)
        assert r.returncode == 0 and want in r.stdout


@pytest.mark.parametrize("gen", ["helix", "random"])
def test_own_bvh_equals_reference_bvh_and_bruteforce(oracle, oracle_ref, gen):
    pos, attr, seg = scenes.helix_lines(30, 61) if gen == "helix" else scenes.random_segments(3000, 0.03, seed=5)
    cam = lv.make_camera(96, 64)
    a, sa = oracle.scene(pos, attr, seg, 0.006).trace_primary(cam)
    b, sb = oracle_ref.scene(pos, attr, seg, 0.006).trace_primary(cam)
    c, _ = oracle.scene(pos, attr, seg, 0.006).trace_primary(cam, bruteforce=True)
    assert np.array_equal(a, b), "oracle BVH vs reference madmann91/bvh traversal"
    assert np.array_equal(a["t"], c["t"]) and (a["prim"] != 0xFFFFFFFF).sum() > 100
    assert sb["T"] > 0 and sb["I"] > 0


def test_rtao_own_vs_reference_bvh(oracle, oracle_ref):
    pos, attr, seg = scenes.helix_lines(30, 61)
    cam = lv.make_camera(64, 48)
    opts = lvo.default_options(ao_strength=1.0, ao_spp=4)
    a, sa = oracle.scene(pos, attr, seg, 0.006).render_rtao(cam, opts)
    b, sb = oracle_ref.scene(pos, attr, seg, 0.006).render_rtao(cam, opts)
    assert np.array_equal(a, b) and sa["rays_ao"] == sb["rays_ao"] == 4 * sa["pixels_hit"]
    assert 0.0 <= a.min() < 1.0 and a.max() == 1.0


# ---- shading / frames ----------------------------------------------------------------------------------------------
def test_single_tube_frame_known_properties(oracle):
    """Config 1 restated: 128x128 frame of a 1-segment scene, finite output + analytic properties."""
    pos = np.array([[-0.2, 0.0, 0.0], [0.2, 0.0, 0.0]], np.float32)
    sc = oracle.scene(pos, np.array([0.0, 1.0], np.float32), np.array([[0, 1]], np.uint32), 0.05)
    cam = lv.make_camera(128, 128)
    tf = scenes.standard_transfer_function()
    img, st = sc.render_tubes(cam, lvo.default_options(), tf)
    hits, _ = sc.trace_primary(cam)
    assert np.isfinite(img).all()
    miss = hits["prim"] == 0xFFFFFFFF
    assert np.allclose(img[miss], 1.0)                      # background = clear colour through the miss shader
    # centre pixel row: camera at z = 0.8 looks at the tube axis: t = 0.8 - r (up to the half-pixel offset)
    cy = hits[64, 64]
    assert cy["kind"] == 0 and abs(cy["t"] - (0.8 - 0.025)) < 2e-4
    # hit pixels form a horizontal band: projected half-height = r / (0.8 - ...) * 64 / tan(fov/2) ~ 4 px
    rows = np.nonzero((~miss).any(axis=1))[0]
    assert 6 <= len(rows) <= 10 and abs(rows.mean() - 63.5) <= 0.51
    # silhouette pixels are darkened towards the foreground colour (halo), centre pixels carry the TF colour
    assert img[64, 64, :3].max() > 0.2 and st["rays"] >= 128 * 128


def test_tubes_running_mean_over_frames(oracle):
    pos, attr, seg = scenes.helix_lines(10, 31)
    sc = oracle.scene(pos, attr, seg, 0.01)
    cam = lv.make_camera(48, 32)
    tf = scenes.standard_transfer_function()
    opts = lvo.default_options(num_samples_per_frame=2, use_jittered_rays=1)
    f0, _ = sc.render_tubes(cam, opts, tf, frame_number=0)
    f1_alone, _ = sc.render_tubes(cam, opts, tf, frame_number=1, rgba=np.zeros_like(f0))
    f1, _ = sc.render_tubes(cam, opts, tf, frame_number=1, rgba=f0.copy())
    # frame 1 = mix(prev, new, 1/2); with prev = 0 the stored value is new/2 (TubeRayTracing.glsl:269-272)
    assert np.allclose(f1, 0.5 * f0 + f1_alone, atol=1e-6)


# ---- PPLL ----------------------------------------------------------------------------------------------------------
def test_addr_gen_tiled_2x8(oracle):
    W = 16
    seen = set()
    for y in range(16):
        for x in range(W):
            a = oracle.addr_gen(x, y, W, 2, 8)
            want = ((x // 2) + (W // 2) * (y // 8)) * 16 + (x % 2) + (y % 8) * 2
            assert a == want
            seen.add(a)
    assert seen == set(range(16 * W))
    assert oracle.addr_gen(5, 7, 100, 1, 1) == 705


def test_pack_unorm4x8(oracle):
    assert oracle.pack_unorm4x8([0, 0, 0, 0]) == 0
    assert oracle.pack_unorm4x8([1, 1, 1, 1]) == 0xFFFFFFFF
    assert oracle.pack_unorm4x8([1.5, -1, 0.5, 0.25]) == (255 | (0 << 8) | (128 << 16) | (64 << 24))


def _rand_frags(rng, n):
    cols = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    depths = rng.permutation(n).astype(np.float32) * 0.01 + 0.5   # distinct depths
    return cols, depths


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 33, 100])
def test_all_correct_sorts_agree_with_canonical(oracle, n):
    """With distinct depths every correct sort of the reference (modes 1-4, 6, 7) gives the canonical result bit for bit,
    and the priority queue (mode 0) the canonical early-out result."""
    rng = np.random.default_rng(n)
    cols, depths = _rand_frags(rng, n)
    for mode in (0, 1, 2, 3, 4, 6, 7):
        a = oracle.sort_blend(cols, depths, 128, mode, canonical=False)
        b = oracle.sort_blend(cols, depths, 128, mode, canonical=True)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), mode


def test_reference_bitonic_only_sorts_powers_of_two(oracle):
    """LinkedListSort.glsl:241-263 as written skips out-of-range partners and stops at k <= fragsCount, so it is a
    sorting network only for power-of-two counts.  The product implements the canonical order instead."""
    rng = np.random.default_rng(3)
    for n in (2, 4, 8, 64):
        cols, depths = _rand_frags(rng, n)
        assert np.array_equal(oracle.sort_blend(cols, depths, 128, 5, False), oracle.sort_blend(cols, depths, 128, 5, True))
    cols = np.array([0xFF0000FF, 0xFF00FF00, 0xFFFF0000], np.uint32)   # opaque r, g, b
    depths = np.array([0.3, 0.2, 0.1], np.float32)                     # element 2 is nearest but never compared
    lit = oracle.sort_blend(cols, depths, 128, 5, False)
    can = oracle.sort_blend(cols, depths, 128, 5, True)
    assert np.allclose(can[:3], [0, 0, 1]) and not np.allclose(lit[:3], can[:3])


def test_priority_queue_early_out(oracle):
    # three opaque fragments: blending stops after the nearest (alpha 1 >= 0.99)
    cols = np.array([0xFF0000FF, 0xFF00FF00, 0xFFFF0000], np.uint32)
    depths = np.array([0.3, 0.1, 0.2], np.float32)
    out = oracle.sort_blend(cols, depths, 16, 0, False)
    assert np.allclose(out, [0, 1, 0, 1])


def test_ppll_gather_resolve_consistency(oracle):
    pos, attr, seg = scenes.helix_lines(20, 41)
    sc = oracle.scene(pos, attr, seg, 0.008)
    cam = lv.make_camera(70, 50)
    tf = scenes.standard_transfer_function(opacity=(0.1, 0.6))
    opts = lvo.default_options()
    g = sc.ppll_gather(cam, opts, tf)
    assert g["padded"] == (70, 56) and g["counter"] == len(g["nodes"]) > 0
    lists = lvo.per_pixel_lists(g["heads"], g["nodes"], cam, opts, oracle)
    assert sum(len(v) for v in lists.values()) == g["counter"]
    hits, _ = sc.trace_primary(cam)
    # the nearest fragment of each pixel is the closest hit of the primary pass
    for (x, y), lst in list(lists.items())[::7]:
        assert lst[0][0] == int(hits[y, x]["t"].view(np.uint32)) or hits[y, x]["prim"] != 0xFFFFFFFF
    img, st = lvo.ppll_resolve(oracle, cam, opts, g["heads"], g["nodes"], 100, 0, canonical=True)
    assert st["frags_sorted"] == g["counter"] and np.isfinite(img).all()
    empty = np.array([[(x, y) not in lists for x in range(70)] for y in range(50)])
    assert np.allclose(img[empty], 1.0)
    # overflow: budget smaller than the fragment count -> counter keeps counting, stored nodes capped
    g2 = sc.ppll_gather(cam, opts, tf, linked_list_size=g["counter"] // 2)
    assert g2["counter"] == g["counter"] and len(g2["nodes"]) == g["counter"] // 2


# ---- segment builder (a2) ------------------------------------------------------------------------------------------
def test_segments_from_polylines_matches_python_host_mirror(oracle):
    rng = np.random.default_rng(4)
    lines, attrs, offs = [], [], [0]
    for k in range(12):
        n = int(rng.integers(1, 30))
        p = np.cumsum(rng.standard_normal((n, 3)) * 0.01, axis=0)
        if k % 3 == 0 and n > 4:
            p[2] = p[1]; p[3] = p[1]       # repeated points -> |tangent| < 1e-4 -> skipped
        lines.append(p); attrs.append(rng.random(n)); offs.append(offs[-1] + n)
    lines.append(np.zeros((3, 3))); attrs.append(np.zeros(3)); offs.append(offs[-1] + 3)   # fully degenerate trajectory
    pos = np.concatenate(lines).astype(np.float32); attr = np.concatenate(attrs).astype(np.float32)
    po, ao, so, to, no = oracle.segments_from_polylines(pos, attr, offs)
    pp, ap, sp = scenes.segments_from_polylines(pos, attr, offs)
    assert np.array_equal(po, pp) and np.array_equal(ao, ap) and np.array_equal(so, sp)
    assert np.allclose(np.linalg.norm(to, axis=1), 1.0, atol=1e-5)
    assert np.allclose(np.einsum("ij,ij->i", to, no), 0.0, atol=1e-4)       # Gram-Schmidt normals
    assert so.max() < len(po) and (so[:, 1] == so[:, 0] + 1).all()


def test_depth_cue_range_known_answer(oracle):
    """DepthCues/ComputeDepthValues.glsl:60-76: view depth of the vertices inside the frustum, +- 1e-2, clamped to [near, far]."""
    pos = np.array([[-0.1, 0.0, 0.0], [0.1, 0.0, 0.2], [5.0, 0.0, 0.0], [5.1, 0.0, 0.0]], np.float32)   # last segment is off screen
    sc = oracle.scene(pos, np.zeros(4, np.float32), np.array([[0, 1], [2, 3]], np.uint32), 0.01)
    cam = lv.make_camera(64, 64)                      # camera at z = 0.8 looking down -z
    dmin, dmax = sc.depth_range(cam)
    assert abs(dmin - (0.6 - 0.01)) < 1e-6 and abs(dmax - (0.8 + 0.01)) < 1e-6
    # nothing in the frustum: the neutral element (far, near) of the reduction survives
    sc2 = oracle.scene(pos[2:], np.zeros(2, np.float32), np.array([[0, 1]], np.uint32), 0.01)
    assert sc2.depth_range(cam) == (100.0, np.float32(0.01))


def test_oracle_matches_committed_regression_vectors():
    """tests/golden/oracle_vectors.npz (written by tests/golden/make_oracle_vectors.py): the oracle recomputes every frozen output bit
    for bit -- intersections, polyline frames / parametrization, the tube mesh, closest hits, both AO images, shaded frames (screen-space
    and prebaked AO), baked factors, PPLL counts and the resolved frame."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_oracle_vectors", os.path.join(here, "golden", "make_oracle_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.compute()
    frozen = np.load(os.path.join(here, "golden", "oracle_vectors.npz"))
    assert set(frozen.files) == set(now)
    for k in frozen.files:
        a, b = np.ascontiguousarray(now[k]), np.ascontiguousarray(frozen[k])
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if a.dtype.kind == "f":     # bit-exact, NaN (0/0 of an all-transparent PPLL list) matched by position
            nan = np.isnan(a)
            assert np.array_equal(nan, np.isnan(b)) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32)), k
        else:
            assert np.array_equal(a, b), k
