"""ctypes wrapper of the CPU oracle (oracle/liblvo_oracle.so, oracle/_ref/liblvo_oracle_ref.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by linevis_b200/ (the product).  PARITY UNPINNED by the reference's own tests.
"""
import ctypes
import os
import subprocess

import numpy as np

from linevis_b200.camera import LvCamera

_HERE = os.path.dirname(os.path.abspath(__file__))
OWN_LIB = os.path.join(_HERE, "liblvo_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "liblvo_oracle_ref.so")


class LvoOptions(ctypes.Structure):
    _fields_ = [
        ("use_capped_tubes", ctypes.c_int32), ("use_halos", ctypes.c_int32),
        ("ao_strength", ctypes.c_float), ("ao_gamma", ctypes.c_float), ("ao_radius", ctypes.c_float),
        ("ao_spp", ctypes.c_uint32), ("ao_use_distance", ctypes.c_int32), ("ao_jitter_primary", ctypes.c_int32),
        ("tube_num_subdivisions", ctypes.c_uint32), ("num_samples_per_frame", ctypes.c_uint32),
        ("use_jittered_rays", ctypes.c_int32), ("use_deterministic_sampling", ctypes.c_int32),
        ("max_depth_complexity", ctypes.c_uint32), ("tile_w", ctypes.c_uint32), ("tile_h", ctypes.c_uint32),
        ("depth_cue_strength", ctypes.c_float), ("use_static_ao", ctypes.c_int32),
    ]


class LvoBakeOptions(ctypes.Structure):
    """AmbientOcclusionComputeRenderPass settings (VulkanAmbientOcclusionBaker.hpp:163-168)."""
    _fields_ = [("ao_radius", ctypes.c_float), ("num_tube_subdivisions", ctypes.c_uint32),
                ("samples_per_frame", ctypes.c_uint32), ("use_distance", ctypes.c_int32)]


def default_options(**kw):
    """Reference defaults (LineData.hpp:377-378, VulkanRayTracedAmbientOcclusion.hpp:150-153, LineData.cpp:52,
    VulkanRayTracer.hpp:137-142, LineRenderer.cpp:739-740); AO off until ao_strength > 0."""
    o = LvoOptions(1, 1, 0.0, 1.0, 0.1, 4, 1, 1, 6, 1, 0, 0, 1024, 2, 8, 0.0, 0)
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


HIT_DTYPE = np.dtype([("t", np.float32), ("prim", np.uint32), ("kind", np.uint32), ("pad", np.uint32)])
NODE_DTYPE = np.dtype([("color", np.uint32), ("depth", np.float32), ("next", np.uint32)])


def build(ref=False):
    """Compile the oracle with its committed recipe (oracle/Makefile)."""
    subprocess.run(["make", "-C", _HERE] + (["ref"] if ref else []), check=True, capture_output=True)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class Oracle:
    """One loaded oracle library.  backend = 'own' (portable BVH) or 'ref' (reference's madmann91/bvh)."""

    def __init__(self, backend="own"):
        path = OWN_LIB if backend == "own" else REF_LIB
        if not os.path.exists(path):
            if backend == "own" or os.path.isdir("/root/reference/submodules/bvh/include"):
                build(ref=(backend != "own"))
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.backend = backend
        self.lib = L = ctypes.CDLL(path)
        L.lvo_tea.restype = ctypes.c_uint32
        L.lvo_tea.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.lvo_det_pow.restype = ctypes.c_float
        L.lvo_det_pow.argtypes = [ctypes.c_float, ctypes.c_float]
        L.lvo_addr_gen.restype = ctypes.c_uint32
        L.lvo_addr_gen.argtypes = [ctypes.c_uint32] * 5
        L.lvo_pack_unorm4x8.restype = ctypes.c_uint32
        L.lvo_scene_create.restype = ctypes.c_void_p
        L.lvo_scene_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_float]
        L.lvo_scene_destroy.argtypes = [ctypes.c_void_p]
        L.lvo_scene_num_nodes.restype = ctypes.c_uint64
        L.lvo_scene_num_nodes.argtypes = [ctypes.c_void_p]
        L.lvo_backend_name.restype = ctypes.c_char_p
        L.lvo_ppll_gather.restype = ctypes.c_uint64
        L.lvo_num_threads.restype = ctypes.c_int
        L.lvo_det_acos.restype = ctypes.c_float
        L.lvo_det_acos.argtypes = [ctypes.c_float]
        L.lvo_ao_parametrize.restype = ctypes.c_uint64
        L.lvo_static_ao_factor.restype = ctypes.c_float
        L.lvo_static_ao_factor.argtypes = [ctypes.c_void_p] + [ctypes.c_float] * 4

    # ---- unit helpers
    def tea(self, a, b):
        return int(self.lib.lvo_tea(a & 0xFFFFFFFF, b & 0xFFFFFFFF))

    def rnd_stream(self, seed, n):
        l = np.zeros(n, np.uint32)
        f = np.zeros(n, np.float32)
        self.lib.lvo_rnd_stream(ctypes.c_uint32(seed), ctypes.c_uint32(n), _p(l, ctypes.c_uint32), _p(f, ctypes.c_float))
        return l, f

    def det_pow(self, x, y):
        return float(self.lib.lvo_det_pow(x, y))

    def det_sincos2pi(self, xi):
        c, s = ctypes.c_float(), ctypes.c_float()
        self.lib.lvo_det_sincos2pi(ctypes.c_float(xi), ctypes.byref(c), ctypes.byref(s))
        return c.value, s.value

    def addr_gen(self, x, y, vw, tw=2, th=8):
        return int(self.lib.lvo_addr_gen(x, y, vw, tw, th))

    def pack_unorm4x8(self, c):
        a = _f32(c)
        return int(self.lib.lvo_pack_unorm4x8(_p(a, ctypes.c_float)))

    def sample_hemisphere(self, a, b):
        out = np.zeros(3, np.float32)
        self.lib.lvo_sample_hemisphere(ctypes.c_float(a), ctypes.c_float(b), _p(out, ctypes.c_float))
        return out

    def intersect_tube(self, ro, rd, p0, p1, radius, capped=True):
        ro, rd, p0, p1 = _f32(ro), _f32(rd), _f32(p0), _f32(p1)
        t, k = ctypes.c_float(), ctypes.c_int()
        h = self.lib.lvo_intersect_tube(_p(ro, ctypes.c_float), _p(rd, ctypes.c_float), _p(p0, ctypes.c_float), _p(p1, ctypes.c_float),
                                        ctypes.c_float(radius), ctypes.c_int(int(capped)), ctypes.byref(t), ctypes.byref(k))
        return bool(h), t.value, k.value

    def sort_blend(self, colors, depths, max_frags, mode, canonical=False):
        colors = np.ascontiguousarray(colors, np.uint32)
        depths = _f32(depths)
        out = np.zeros(4, np.float32)
        self.lib.lvo_sort_blend(_p(colors, ctypes.c_uint32), _p(depths, ctypes.c_float), ctypes.c_uint32(len(colors)),
                                ctypes.c_uint32(max_frags), ctypes.c_int(mode), ctypes.c_int(int(canonical)), _p(out, ctypes.c_float))
        return out

    def segments_from_polylines(self, pos, attr, line_offsets):
        pos, attr = _f32(pos), _f32(attr)
        off = np.ascontiguousarray(line_offsets, np.uint64)
        n = pos.shape[0]
        po, ao = np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        to, no = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        so = np.zeros((max(n, 1), 2), np.uint32)
        npt, nseg = ctypes.c_uint64(), ctypes.c_uint64()
        self.lib.lvo_segments_from_polylines(_p(pos, ctypes.c_float), _p(attr, ctypes.c_float), _p(off, ctypes.c_uint64),
                                             ctypes.c_uint64(len(off) - 1), _p(po, ctypes.c_float), _p(ao, ctypes.c_float),
                                             _p(to, ctypes.c_float), _p(no, ctypes.c_float), _p(so, ctypes.c_uint32),
                                             ctypes.byref(npt), ctypes.byref(nseg))
        return po[:npt.value], ao[:npt.value], so[:nseg.value], to[:npt.value], no[:npt.value]

    def num_threads(self):
        return int(self.lib.lvo_num_threads())

    def set_num_threads(self, n=None):
        """Size the OpenMP pool explicitly (default: every host core this process may run on); torchrun exports OMP_NUM_THREADS=1."""
        if n is None:
            n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.lib.lvo_set_num_threads(ctypes.c_int(int(n)))
        return self.num_threads()

    def det_acos(self, x):
        return float(self.lib.lvo_det_acos(x))

    def ao_parametrize(self, pos, line_offsets, expected_param_segment_length):
        """recomputeStaticParametrization: (blending weights [n_pt], sampling locations [n_param])."""
        pos = _f32(pos)
        off = np.ascontiguousarray(line_offsets, np.uint64)
        bw = np.zeros(pos.shape[0], np.float32)
        args = (_p(pos, ctypes.c_float), _p(off, ctypes.c_uint64), ctypes.c_uint64(len(off) - 1), ctypes.c_float(expected_param_segment_length),
                _p(bw, ctypes.c_float))
        n = int(self.lib.lvo_ao_parametrize(*args, None, ctypes.c_uint64(0)))
        sl = np.zeros(max(n, 1), np.float32)
        self.lib.lvo_ao_parametrize(*args, _p(sl, ctypes.c_float), ctypes.c_uint64(n))
        return bw, sl[:n]

    def scene(self, pos, attr, seg_idx, line_width):
        return OracleScene(self, pos, attr, seg_idx, line_width)


class OracleScene:
    def __init__(self, oracle, pos, attr, seg_idx, line_width):
        self.o = oracle
        self.lib = oracle.lib
        pos, attr = _f32(pos), _f32(attr)
        seg = np.ascontiguousarray(seg_idx, np.uint32)
        self.n_seg = seg.shape[0]
        self.line_width = float(line_width)
        self.h = self.lib.lvo_scene_create(pos.ctypes.data, attr.ctypes.data, seg.ctypes.data,
                                           pos.shape[0], seg.shape[0], ctypes.c_float(line_width))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.lvo_scene_destroy(ctypes.c_void_p(self.h))
            self.h = None

    def num_nodes(self):
        return int(self.lib.lvo_scene_num_nodes(ctypes.c_void_p(self.h)))

    # ---- object-space AO prebaker (S6)
    def set_lines(self, tangent, normal):
        tangent, normal = _f32(tangent), _f32(normal)
        self.lib.lvo_scene_set_lines(ctypes.c_void_p(self.h), _p(tangent, ctypes.c_float), _p(normal, ctypes.c_float), ctypes.c_uint64(tangent.shape[0]))

    def ao_bake_iteration(self, sampling_locations, frame_number, factors=None, radius=0.1, n_subdiv=8, spp=4, use_distance=True, capped=True,
                          return_rays=False, tube_mesh=None):
        sl = _f32(sampling_locations)
        if factors is None:
            factors = np.zeros(sl.shape[0] * n_subdiv, np.float32)
        factors = _f32(factors)
        bo = LvoBakeOptions(radius, n_subdiv, spp, int(use_distance))
        stats = np.zeros(3, np.uint64)
        rays = np.zeros((sl.shape[0] * n_subdiv * spp, 6), np.float32) if return_rays else None
        self.lib.lvo_ao_bake_iteration(ctypes.c_void_p(self.h), ctypes.byref(bo), ctypes.c_int(int(capped)), _p(sl, ctypes.c_float),
                                       ctypes.c_uint64(sl.shape[0]), ctypes.c_uint32(frame_number), _p(factors, ctypes.c_float), _p(stats, ctypes.c_uint64),
                                       _p(rays, ctypes.c_float) if return_rays else None,
                                       ctypes.c_void_p(tube_mesh.h) if tube_mesh is not None else None)
        st = dict(T=int(stats[0]), I=int(stats[1]), rays=int(stats[2]))
        return (factors, st, rays) if return_rays else (factors, st)

    def shade_hits(self, cam, opts, tf, ro, rd, t, kind, prim, amin=0.0, amax=1.0, ao_tex=None):
        """closestHitTubeAnalytic for given hits -> [n, 5] (rgba, hitT)."""
        tf, ro, rd, t = _f32(tf), _f32(ro), _f32(rd), _f32(t)
        kind, prim = np.ascontiguousarray(kind, np.uint32), np.ascontiguousarray(prim, np.uint32)
        out = np.zeros((t.shape[0], 5), np.float32)
        aop = _p(_f32(ao_tex), ctypes.c_float) if ao_tex is not None else None
        self.lib.lvo_shade_hits(ctypes.c_void_p(self.h), ctypes.byref(cam), ctypes.byref(opts), _p(tf, ctypes.c_float), ctypes.c_uint32(tf.shape[0]),
                                ctypes.c_float(amin), ctypes.c_float(amax), aop, ctypes.c_uint64(t.shape[0]), _p(ro, ctypes.c_float),
                                _p(rd, ctypes.c_float), _p(t, ctypes.c_float), _p(kind, ctypes.c_uint32), _p(prim, ctypes.c_uint32), _p(out, ctypes.c_float))
        return out

    def set_static_ao(self, factors, n_subdiv, blending_weights):
        f, bw = _f32(factors), _f32(blending_weights)
        self.lib.lvo_scene_set_static_ao(ctypes.c_void_p(self.h), _p(f, ctypes.c_float), ctypes.c_uint64(f.size // n_subdiv), ctypes.c_uint32(n_subdiv),
                                         _p(bw, ctypes.c_float), ctypes.c_uint64(bw.shape[0]))

    def static_ao_factor(self, vertex_id, phi, strength=1.0, gamma=1.0):
        return float(self.lib.lvo_static_ao_factor(ctypes.c_void_p(self.h), strength, gamma, vertex_id, phi))

    def depth_range(self, cam):
        out = np.zeros(2, np.float32)
        self.lib.lvo_depth_range(ctypes.c_void_p(self.h), ctypes.byref(cam), _p(out, ctypes.c_float))
        return float(out[0]), float(out[1])

    def trace_primary(self, cam, opts=None, bruteforce=False):
        opts = opts or default_options()
        hits = np.zeros(cam.width * cam.height, HIT_DTYPE)
        stats = np.zeros(3, np.uint64)
        if bruteforce:
            self.lib.lvo_trace_primary_bruteforce(ctypes.c_void_p(self.h), ctypes.byref(cam), ctypes.byref(opts), hits.ctypes.data_as(ctypes.c_void_p))
        else:
            self.lib.lvo_trace_primary(ctypes.c_void_p(self.h), ctypes.byref(cam), ctypes.byref(opts),
                                       hits.ctypes.data_as(ctypes.c_void_p), _p(stats, ctypes.c_uint64))
        return hits.reshape(cam.height, cam.width), dict(T=int(stats[0]), I=int(stats[1]), rays=int(stats[2]))

    def render_rtao(self, cam, opts, frame_number=0, ao=None):
        if ao is None:
            ao = np.zeros((cam.height, cam.width), np.float32)
        ao = _f32(ao)
        stats = np.zeros(7, np.uint64)
        self.lib.lvo_render_rtao(ctypes.c_void_p(self.h), ctypes.byref(cam), ctypes.byref(opts), ctypes.c_uint32(frame_number),
                                 _p(ao, ctypes.c_float), _p(stats, ctypes.c_uint64))
        return ao, dict(T=int(stats[0]), I=int(stats[1]), rays_primary=int(stats[2]), rays_ao=int(stats[3]), pixels_hit=int(stats[4]),
                        T_ao=int(stats[5]), I_ao=int(stats[6]))

    def render_tubes(self, cam, opts, tf, amin=0.0, amax=1.0, ao_tex=None, frame_number=0, rgba=None):
        tf = _f32(tf)
        if rgba is None:
            rgba = np.zeros((cam.height, cam.width, 4), np.float32)
        rgba = _f32(rgba)
        stats = np.zeros(3, np.uint64)
        aop = _p(_f32(ao_tex), ctypes.c_float) if ao_tex is not None else None
        self._keep = ao_tex
        self.lib.lvo_render_tubes(ctypes.c_void_p(self.h), ctypes.byref(cam), ctypes.byref(opts), _p(tf, ctypes.c_float),
                                  ctypes.c_uint32(tf.shape[0]), ctypes.c_float(amin), ctypes.c_float(amax), aop,
                                  ctypes.c_uint32(frame_number), _p(rgba, ctypes.c_float), _p(stats, ctypes.c_uint64))
        return rgba, dict(T=int(stats[0]), I=int(stats[1]), rays=int(stats[2]))

    def ppll_gather(self, cam, opts, tf, amin=0.0, amax=1.0, linked_list_size=None, ao_tex=None):
        tf = _f32(tf)
        pw, ph = ctypes.c_uint32(), ctypes.c_uint32()
        self.lib.lvo_ppll_padded_size(cam.width, cam.height, opts.tile_w, opts.tile_h, ctypes.byref(pw), ctypes.byref(ph))
        pw, ph = pw.value, ph.value
        if linked_list_size is None:
            linked_list_size = 64 * pw * ph
        heads = np.zeros(pw * ph, np.uint32)
        nodes = np.zeros(linked_list_size, NODE_DTYPE)
        stats = np.zeros(4, np.uint64)
        aop = _p(_f32(ao_tex), ctypes.c_float) if ao_tex is not None else None
        counter = self.lib.lvo_ppll_gather(ctypes.c_void_p(self.h), ctypes.byref(cam), ctypes.byref(opts), _p(tf, ctypes.c_float),
                                           ctypes.c_uint32(tf.shape[0]), ctypes.c_float(amin), ctypes.c_float(amax), aop,
                                           ctypes.c_uint64(linked_list_size), _p(heads, ctypes.c_uint32),
                                           nodes.ctypes.data_as(ctypes.c_void_p), _p(stats, ctypes.c_uint64))
        counter = int(counter)
        return dict(counter=counter, heads=heads, nodes=nodes[:min(counter, linked_list_size)], padded=(pw, ph),
                    T=int(stats[0]), I=int(stats[1]), rays=int(stats[2]))


TUBE_VERTEX_DTYPE = np.dtype([("position", np.float32, 3), ("line_point", np.uint32), ("normal", np.float32, 3), ("phi", np.float32)])


class TubeMesh:
    """The reference's triangulated capped tubes (createCappedTriangleTubesRenderDataCPU) + a triangle BVH: the geometry the
    reference's RTAO passes are traced against (oracle/lvo_tritubes.hpp)."""

    def __init__(self, oracle, pos, line_offsets, line_width, num_subdivisions=6):
        self.lib = oracle.lib
        self.lib.lvo_tubemesh_create.restype = ctypes.c_void_p
        pos = _f32(pos)
        off = np.ascontiguousarray(line_offsets, np.uint64)
        self.h = self.lib.lvo_tubemesh_create(_p(pos, ctypes.c_float), _p(off, ctypes.c_uint64), ctypes.c_uint64(len(off) - 1),
                                              ctypes.c_float(0.5 * line_width), ctypes.c_int(num_subdivisions))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.lvo_tubemesh_destroy(ctypes.c_void_p(self.h))
            self.h = None

    def info(self):
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        self.lib.lvo_tubemesh_info(ctypes.c_void_p(self.h), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return dict(n_vertices=a.value, n_triangles=b.value, n_line_points=c.value)

    def arrays(self):
        i = self.info()
        v = np.zeros(i["n_vertices"], TUBE_VERTEX_DTYPE)
        t = np.zeros((i["n_triangles"], 3), np.uint32)
        self.lib.lvo_tubemesh_copy(ctypes.c_void_p(self.h), v.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p))
        return v, t

    def render_tubes(self, attr, cam, opts, tf, amin=0.0, amax=1.0, ao_tex=None, frame_number=0, rgba=None):
        """The tube pass in the reference's triangle-mesh geometry mode (ClosestHitTubeTriangles); attr = lineAttribute per input point."""
        tf, attr = _f32(tf), _f32(attr)
        if rgba is None:
            rgba = np.zeros((cam.height, cam.width, 4), np.float32)
        rgba = _f32(rgba)
        stats = np.zeros(3, np.uint64)
        ao = _f32(ao_tex) if ao_tex is not None else None
        self.lib.lvo_tubemesh_render_tubes(ctypes.c_void_p(self.h), _p(attr, ctypes.c_float), ctypes.byref(cam), ctypes.byref(opts), _p(tf, ctypes.c_float),
                                           ctypes.c_uint32(tf.shape[0]), ctypes.c_float(amin), ctypes.c_float(amax),
                                           _p(ao, ctypes.c_float) if ao is not None else None, ctypes.c_uint32(frame_number),
                                           _p(rgba, ctypes.c_float), _p(stats, ctypes.c_uint64))
        return rgba, dict(T=int(stats[0]), I=int(stats[1]), rays=int(stats[2]))

    def render_rtao(self, cam, opts, frame_number=0, ao=None):
        if ao is None:
            ao = np.zeros((cam.height, cam.width), np.float32)
        ao = _f32(ao)
        stats = np.zeros(5, np.uint64)
        self.lib.lvo_render_rtao_triangles(ctypes.c_void_p(self.h), ctypes.byref(cam), ctypes.byref(opts), ctypes.c_uint32(frame_number),
                                           _p(ao, ctypes.c_float), _p(stats, ctypes.c_uint64))
        return ao, dict(T=int(stats[0]), I=int(stats[1]), rays_primary=int(stats[2]), rays_ao=int(stats[3]), pixels_hit=int(stats[4]))


def ppll_resolve(oracle, cam, opts, heads, nodes, max_frags, sort_mode, canonical=True):
    heads = np.ascontiguousarray(heads, np.uint32)
    nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
    rgba = np.zeros((cam.height, cam.width, 4), np.float32)
    stats = np.zeros(3, np.uint64)
    oracle.lib.lvo_ppll_resolve(ctypes.byref(cam), ctypes.byref(opts), _p(heads, ctypes.c_uint32), nodes.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_uint32(max_frags), ctypes.c_int(sort_mode), ctypes.c_int(int(canonical)),
                                _p(rgba, ctypes.c_float), _p(stats, ctypes.c_uint64))
    return rgba, dict(frags_sorted=int(stats[0]), frags_truncated=int(stats[1]), max_depth_complexity=int(stats[2]))


def per_pixel_lists(heads, nodes, cam, opts, oracle):
    """Walk every pixel's list; returns {(x, y): sorted list of (depth_bits, color)} -- the order-insensitive
    multiset the integer-path parity is stated on."""
    pw = cam.width if cam.width % opts.tile_w == 0 else (cam.width // opts.tile_w + 1) * opts.tile_w
    out = {}
    depth_bits = nodes["depth"].view(np.uint32)
    for y in range(cam.height):
        for x in range(cam.width):
            off = int(heads[oracle.addr_gen(x, y, pw, opts.tile_w, opts.tile_h)])
            lst = []
            while off != 0xFFFFFFFF:
                lst.append((int(depth_bits[off]), int(nodes["color"][off])))
                off = int(nodes["next"][off])
            if lst:
                out[(x, y)] = sorted(lst)
    return out


def per_pixel_multisets(heads, nodes):
    """Vectorised form of per_pixel_lists for large frames: all lists are walked at once; returns a uint64 [n, 2] array of
    (start-offset slot, (depth bits << 32) | colour) rows sorted lexicographically -- equal arrays <=> equal per-pixel multisets."""
    nxt = nodes["next"].astype(np.int64)
    key = (nodes["depth"].view(np.uint32).astype(np.uint64) << np.uint64(32)) | nodes["color"].astype(np.uint64)
    cur = heads.astype(np.int64).ravel()
    slot = np.arange(cur.size, dtype=np.int64)
    rows = []
    live = cur != 0xFFFFFFFF
    cur, slot = cur[live], slot[live]
    while cur.size:
        rows.append(np.stack([slot.astype(np.uint64), key[cur]], axis=1))
        cur = nxt[cur]
        live = cur != 0xFFFFFFFF
        cur, slot = cur[live], slot[live]
    if not rows:
        return np.zeros((0, 2), np.uint64)
    r = np.concatenate(rows)
    return r[np.lexsort((r[:, 1], r[:, 0]))]
