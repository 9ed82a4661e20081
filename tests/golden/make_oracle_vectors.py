"""Regression fixtures of the CPU oracle (oracle/): small inputs -> outputs, frozen so that neither the oracle nor -- through the
parity tests -- the CUDA path can drift unnoticed.  NOT reference-generated (the reference's GLSL cannot run here):
these pin the oracle to ITSELF as of the commit that wrote them; what pins it to the reference are the cited restatements, the
analytic known answers and the structural tests.  Regenerate only on purpose:  python tests/golden/make_oracle_vectors.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import linevis_b200 as lv  # noqa: E402
from linevis_b200 import scenes  # noqa: E402
from oracle import lvo  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")


def compute():
    o = lvo.Oracle("own")
    out = {}
    rng = np.random.default_rng(424242)
    # IntersectionTube on random ray / segment pairs
    n = 96
    p0 = (rng.random((n, 3)) - 0.5).astype(np.float32)
    p1 = (p0 + (rng.random((n, 3)) - 0.5) * 0.3).astype(np.float32)
    ro = np.tile(np.array([0, 0, 2], np.float32), (n, 1))
    tgt = (p0 + p1) * 0.5 + (rng.random((n, 3)).astype(np.float32) - 0.5) * 0.08
    rd = tgt - ro
    rd = (rd / np.linalg.norm(rd, axis=1, keepdims=True)).astype(np.float32)
    res = [o.intersect_tube(ro[i], rd[i], p0[i], p1[i], 0.03, True) for i in range(n)]
    out.update(it_p0=p0, it_p1=p1, it_ro=ro, it_rd=rd, it_hit=np.array([r[0] for r in res]), it_t=np.array([r[1] for r in res], np.float32),
               it_kind=np.array([r[2] for r in res], np.int32))
    # polylines: frames, parametrization, tube mesh
    d = scenes.helix_polylines(5, 14)
    width = 0.02
    bw, sl = o.ao_parametrize(d["pos"], d["line_offsets"], 0.02)
    tm = lvo.TubeMesh(o, d["pos"], d["line_offsets"], width, 6)
    v, t = tm.arrays()
    out.update(pl_pos=d["pos"], pl_off=d["line_offsets"], pl_tangent=d["tangent"], pl_normal=d["normal"], par_bw=bw, par_sl=sl,
               mesh_v=v.view(np.uint32).reshape(-1, 8), mesh_t=t)
    # frames of a tiny scene
    osc = o.scene(d["pos"], d["attr"], d["seg"], width)
    osc.set_lines(d["tangent"], d["normal"])
    cam = lv.make_camera(32, 20)
    tf = scenes.standard_transfer_function(opacity=(0.3, 0.9))
    opts = lvo.default_options(ao_strength=1.0, ao_spp=4)
    hits, _ = osc.trace_primary(cam)
    ao, _ = osc.render_rtao(cam, opts, 0)
    ao_tri, _ = tm.render_rtao(cam, opts, 0)
    img, _ = osc.render_tubes(cam, opts, tf, ao_tex=ao)
    f, _ = osc.ao_bake_iteration(sl, 0, radius=0.2, n_subdiv=6, spp=2)
    osc.set_static_ao(f, 6, bw)
    img_static, _ = osc.render_tubes(cam, lvo.default_options(ao_strength=0.8, use_static_ao=1), tf)
    g = osc.ppll_gather(cam, lvo.default_options(), tf)
    lists = lvo.per_pixel_lists(g["heads"], g["nodes"], cam, lvo.default_options(), o)
    counts = np.zeros((20, 32), np.int32)
    for (x, y), l in lists.items():
        counts[y, x] = len(l)
    ppll, _ = lvo.ppll_resolve(o, cam, lvo.default_options(), g["heads"], g["nodes"], 32, 0, canonical=True)
    out.update(fr_hit_t=hits["t"], fr_hit_prim=hits["prim"], fr_hit_kind=hits["kind"], fr_ao=ao, fr_ao_tri=ao_tri, fr_img=img, fr_bake=f,
               fr_img_static=img_static, fr_ppll_counts=counts, fr_ppll=ppll)
    return out


if __name__ == "__main__":
    vec = compute()
    np.savez_compressed(OUT, **vec)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(vec), "arrays")
