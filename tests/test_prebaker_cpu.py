"""CPU tests of the object-space AO prebaker: the oracle's restatement of the parametrization / baker / lookup, pinned by
analytic properties, and the product's host-only lv_ao_parametrize against it (bit-exact)."""
import math

import numpy as np
import pytest

import linevis_b200 as lv
from linevis_b200 import scenes
from oracle import lvo


def _random_polylines(seed, n_lines=9, max_pts=40, step=0.01):
    rng = np.random.default_rng(seed)
    lines = []
    for _ in range(n_lines):
        n = int(rng.integers(2, max_pts))
        lines.append((np.cumsum(rng.standard_normal((n, 3)) * step, axis=0) + rng.random(3) * 0.3).astype(np.float32))
    pos = np.concatenate(lines)
    off = np.concatenate([[0], np.cumsum([len(l) for l in lines])]).astype(np.uint64)
    return pos, off


def test_parametrization_straight_line_known_answer(oracle):
    # 11 points, spacing 0.1 -> length 1.0; expected piece length 0.25 -> 4 pieces, 5 parametrization vertices
    pos = np.zeros((11, 3), np.float32)
    pos[:, 0] = np.arange(11, dtype=np.float32) * np.float32(0.1)
    bw, sl = oracle.ao_parametrize(pos, [0, 11], 0.25)
    assert len(sl) == 5
    np.testing.assert_allclose(sl, [0, 2.5, 5.0, 7.5, 10.0 - 1e-5], atol=2e-5)
    np.testing.assert_allclose(bw, np.minimum(np.arange(11) * 0.4, 4 - 1e-5), atol=2e-5)
    # a very coarse expected length still gives one piece
    bw1, sl1 = oracle.ao_parametrize(pos, [0, 11], 10.0)
    assert len(sl1) == 2 and bw1[-1] < 1.0


def test_parametrization_maps_are_inverse(oracle):
    pos, off = _random_polylines(3)
    bw, sl = oracle.ao_parametrize(pos, off, 0.004)
    assert (np.diff(sl) >= 0).all()
    base = 0
    for l in range(len(off) - 1):
        b, e = int(off[l]), int(off[l + 1])
        w = bw[b:e]
        assert (np.diff(w) >= 0).all() and w[0] == base
        pieces = int(math.ceil(float(w[-1] - base) - 1e-3))
        # interpolating the blending weight at sampling location k must give back k
        for k in range(pieces + 1):
            loc = float(sl[base + k])
            i = min(int(loc), e - 2)
            f = loc - i
            assert abs((w[i - b] * (1 - f) + w[i - b + 1] * f) - (base + k)) < 2e-2
        base += pieces + 1
    assert base == len(sl)


@pytest.mark.parametrize("expected", [0.001, 0.005, 0.05, 1.0])
def test_product_parametrization_equals_oracle(oracle, expected):
    pos, off = _random_polylines(5)
    bw, sl = oracle.ao_parametrize(pos, off, expected)
    w2, s2 = lv.Context.ao_parametrize(pos, off, expected)
    assert np.array_equal(bw, w2) and np.array_equal(sl, s2)


def test_polyline_frames_match_oracle(oracle):
    d = scenes.helix_polylines(24, 41)
    _, _, so, to, no = oracle.segments_from_polylines(d["pos"], d["attr"], d["line_offsets"])
    assert np.array_equal(so, d["seg"])
    assert np.abs(to - d["tangent"]).max() < 1e-6 and np.abs(no - d["normal"]).max() < 1e-6
    assert np.abs((d["tangent"] * d["normal"]).sum(1)).max() < 1e-5       # Gram-Schmidt: normal is orthogonal to the tangent


def test_baker_isolated_line_is_unoccluded_and_wall_occludes(oracle):
    # one straight tube alone: the capsule is convex, no AO ray leaving its surface can hit it -> every factor is 1
    pos = np.zeros((6, 3), np.float32); pos[:, 0] = np.linspace(-0.2, 0.2, 6)
    attr = np.zeros(6, np.float32)
    seg = np.stack([np.arange(5), np.arange(1, 6)], axis=1).astype(np.uint32)
    tangent, normal = scenes.polyline_frames(pos, [0, 6])
    osc = oracle.scene(pos, attr, seg, 0.01)
    osc.set_lines(tangent, normal)
    bw, sl = oracle.ao_parametrize(pos, [0, 6], 0.05)
    f, st = osc.ao_bake_iteration(sl, 0, radius=0.1, n_subdiv=8, spp=16)
    assert st["rays"] == len(sl) * 8 * 16 and np.all(f == 1.0)
    # a second parallel tube right above (+y) occludes the side facing it, not the far side
    pos2 = np.concatenate([pos, pos + np.array([0, 0.02, 0], np.float32)])
    seg2 = np.concatenate([seg, seg + 6]).astype(np.uint32)
    t2, n2 = scenes.polyline_frames(pos2, [0, 6, 12])
    osc2 = oracle.scene(pos2, np.zeros(12, np.float32), seg2, 0.01)
    osc2.set_lines(t2, n2)
    bw2, sl2 = oracle.ao_parametrize(pos2, [0, 6, 12], 0.05)
    f2, _ = osc2.ao_bake_iteration(sl2, 0, radius=0.1, n_subdiv=8, spp=64, use_distance=False)
    f2 = f2.reshape(-1, 8)
    n_first = len(sl2) // 2
    mid = f2[n_first // 2]                                   # a vertex in the middle of the lower tube
    # direction of subdivision s: cos(a) normal + sin(a) binormal, binormal = tangent x normal
    nrm, tng = n2[2], t2[2]
    bnm = np.cross(tng, nrm)
    ups = [math.cos(2 * math.pi * s / 8) * nrm[1] + math.sin(2 * math.pi * s / 8) * bnm[1] for s in range(8)]
    assert mid[int(np.argmax(ups))] < 0.9 and mid[int(np.argmin(ups))] == 1.0


def test_running_mean_over_iterations(oracle):
    d = scenes.helix_polylines(12, 31)
    osc = oracle.scene(d["pos"], d["attr"], d["seg"], 0.01)
    osc.set_lines(d["tangent"], d["normal"])
    bw, sl = oracle.ao_parametrize(d["pos"], d["line_offsets"], 0.02)
    f0, _ = osc.ao_bake_iteration(sl, 0, spp=2)
    single1, _ = osc.ao_bake_iteration(sl, 1, factors=np.zeros_like(f0), spp=2)   # frame 1 mixes with "old" = 0 -> 0.5 * new
    f1, _ = osc.ao_bake_iteration(sl, 1, factors=f0.copy(), spp=2)
    np.testing.assert_allclose(f1, 0.5 * f0 + single1, atol=1e-6)


def test_static_lookup_interpolates(oracle):
    # 2 line points, 2 parametrization vertices, 4 subdivisions: factors chosen so that every interpolation axis is visible
    pos = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    osc = oracle.scene(pos, np.zeros(2, np.float32), np.array([[0, 1]], np.uint32), 0.01)
    factors = np.array([0.0, 0.2, 0.4, 0.6, 1.0, 1.0, 1.0, 1.0], np.float32)   # vertex 0: 0, .2, .4, .6; vertex 1: 1
    osc.set_static_ao(factors, 4, np.array([0.0, 1.0 - 1e-5], np.float32))
    two_pi = 2 * math.pi
    assert osc.static_ao_factor(0.0, 0.0) == 0.0
    assert abs(osc.static_ao_factor(0.0, two_pi / 4) - 0.2) < 1e-6
    assert abs(osc.static_ao_factor(0.0, two_pi / 8) - 0.1) < 1e-6             # halfway between subdivisions 0 and 1
    assert abs(osc.static_ao_factor(0.0, two_pi * 7 / 8) - 0.3) < 1e-6         # wraps around: between subdivision 3 (.6) and 0 (0)
    assert abs(osc.static_ao_factor(0.5, 0.0) - 0.5) < 1e-4                     # halfway along the line
    # strength / gamma: max(0, 1 - s + s * ao^gamma)
    assert abs(osc.static_ao_factor(0.0, two_pi / 4, strength=0.5, gamma=2.0) - (1 - 0.5 + 0.5 * 0.04)) < 1e-5
