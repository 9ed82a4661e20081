"""Thin object wrappers over the C ABI: Context (lv_ctx) and Scene (lv_scene).

Buffers handed in may be numpy arrays (host) or torch CUDA tensors (device); the library detects which.
Every method goes through the C ABI of include/linevis_b200.h -- the same entry points the C++ LineRenderer
adapter (linevis_b200/host/) binds -- and raises LineVisError on a non-zero status.
"""
import ctypes

import numpy as np

from . import capi
from .capi import LvStats, LineVisError, HIT_DTYPE, NODE_DTYPE, BVH_NODE_DTYPE, SORT_MODES, _ptr


class Context:
    """One renderer context per GPU (lv_ctx_create; replaces renderer construction in MainApp::setRenderer)."""

    def __init__(self, device=0, stream=None, lib_path=None):
        self.lib = capi.load_library(lib_path) if lib_path else capi.load_library()
        self.lib_path = lib_path
        h = ctypes.c_void_p()
        rc = self.lib.lv_ctx_create(ctypes.byref(h), int(device), ctypes.c_void_p(stream) if stream else None)
        if rc != capi.LV_OK:
            raise LineVisError(rc, self.lib.lv_last_global_error().decode())
        self.h = h
        self.device = device
        self._tile_size = 64

    def close(self):
        if getattr(self, "h", None):
            self.lib.lv_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != capi.LV_OK:
            raise LineVisError(rc, self.lib.lv_last_error(self.h).decode())

    # -- settings (LineRenderer::setNewSettings keys)
    def set_option(self, key, value):
        if isinstance(value, bool):
            value = "true" if value else "false"
        self._check(self.lib.lv_set_option(self.h, key.encode(), str(value).encode()))

    def set_new_settings(self, settings):
        """SettingsMap-style bulk update; same keys as the reference's replay scripts use."""
        for k, v in settings.items():
            self.set_option(k, v)

    def get_option(self, key):
        buf = ctypes.create_string_buffer(256)
        self._check(self.lib.lv_get_option(self.h, key.encode(), buf, 256))
        return buf.value.decode()

    def set_transfer_function(self, lut, attr_min=0.0, attr_max=1.0):
        lut = np.ascontiguousarray(lut, np.float32)
        assert lut.ndim == 2 and lut.shape[1] == 4
        self._check(self.lib.lv_set_transfer_function(self.h, _ptr(lut), lut.shape[0], attr_min, attr_max))

    # -- sharding
    def set_tile_shard(self, rank, world, tile_size=64):
        self._check(self.lib.lv_set_tile_shard(self.h, rank, world, tile_size))
        self._tile_size = tile_size

    def tile_costs(self, width, height):
        """Hit pixels of this context's last RTAO pass per tile of the frame (Morton order of the tile enumeration; 0 for other ranks' tiles)."""
        ts = self._tile_size
        n = ((width + ts - 1) // ts) * ((height + ts - 1) // ts)
        out = np.zeros(n, np.uint32)
        self._check(self.lib.lv_get_tile_costs(self.h, width, height, _ptr(out), n))
        return out

    def set_tile_owners(self, width, height, owners):
        """Explicit owner rank per tile (Morton order) for frames of this size; see sharding.balance_tiles."""
        owners = np.ascontiguousarray(owners, np.uint8)
        self._check(self.lib.lv_set_tile_owners(self.h, width, height, _ptr(owners), owners.size))

    def owned_tiles(self, width, height):
        n = ctypes.c_uint32()
        self._check(self.lib.lv_get_owned_tiles(self.h, width, height, None, ctypes.byref(n)))
        t = np.zeros((max(n.value, 1), 2), np.uint32)
        self._check(self.lib.lv_get_owned_tiles(self.h, width, height, _ptr(t), ctypes.byref(n)))
        return t[:n.value]

    def pack_owned_tiles(self, image, width, height, packed):
        self._check(self.lib.lv_pack_owned_tiles(self.h, _ptr(image), width, height, _ptr(packed)))

    def unpack_tiles(self, packed, src_rank, world, width, height, image):
        self._check(self.lib.lv_unpack_tiles(self.h, _ptr(packed), src_rank, world, width, height, _ptr(image)))

    def synchronize(self):
        self._check(self.lib.lv_synchronize(self.h))

    def frame_to_rgba8(self, frame, width, height, out=None):
        """lv_frame_to_rgba8: device RGBA32F frame -> RGBA8 UNORM (uint32 per pixel, the reference's sceneTexture format)."""
        if out is None:
            out = np.zeros((height, width), np.uint32)
        self._check(self.lib.lv_frame_to_rgba8(self.h, _ptr(frame), width, height, _ptr(out)))
        return out

    # -- peer-memory frame assembly (lv_frame_alloc / lv_ipc_*): returns raw device addresses (ints)
    def frame_alloc(self, width, height):
        p = ctypes.c_void_p()
        self._check(self.lib.lv_frame_alloc(self.h, width, height, ctypes.byref(p)))
        return p.value

    def frame_free(self, ptr):
        self._check(self.lib.lv_frame_free(self.h, ctypes.c_void_p(ptr)))

    def ipc_export(self, ptr):
        buf = ctypes.create_string_buffer(64)
        self._check(self.lib.lv_ipc_export(self.h, ctypes.c_void_p(ptr), buf))
        return buf.raw

    def ipc_open(self, handle):
        p = ctypes.c_void_p()
        self._check(self.lib.lv_ipc_open(self.h, ctypes.create_string_buffer(handle, 64), ctypes.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._check(self.lib.lv_ipc_close(self.h, ctypes.c_void_p(ptr)))

    # -- scene
    def create_scene(self, pos, attr, seg_idx, line_width=0.002):
        return Scene(self, pos, attr, seg_idx, line_width)

    @staticmethod
    def ao_parametrize(pos, line_offsets, expected_param_segment_length=0.001):
        """lv_ao_parametrize (host only): (blending weights [n_pt], sampling locations [n_param])."""
        lib = capi.load_library()
        pos = np.ascontiguousarray(pos, np.float32)
        off = np.ascontiguousarray(line_offsets, np.uint64)
        w = np.zeros(pos.shape[0], np.float32)
        n = ctypes.c_uint64()
        rc = lib.lv_ao_parametrize(_ptr(pos), _ptr(off), len(off) - 1, expected_param_segment_length, _ptr(w), None, 0, ctypes.byref(n))
        if rc != capi.LV_OK:
            raise LineVisError(rc, "lv_ao_parametrize")
        sl = np.zeros(max(n.value, 1), np.float32)
        lib.lv_ao_parametrize(_ptr(pos), _ptr(off), len(off) - 1, expected_param_segment_length, _ptr(w), _ptr(sl), sl.size, ctypes.byref(n))
        return w, sl[:n.value]

    @staticmethod
    def tube_mesh(pos, line_offsets, line_width, num_subdivisions=6, lib_path=None):
        """lv_tube_mesh (host only): the reference's triangulated capped tubes -> (vertices [n, 8] float32 rows, triangles [m, 3] uint32,
        number of line points)."""
        lib = capi.load_library(lib_path) if lib_path else capi.load_library()
        pos = np.ascontiguousarray(pos, np.float32)
        off = np.ascontiguousarray(line_offsets, np.uint64)
        nv, nt, nl = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        args = (_ptr(pos), _ptr(off), len(off) - 1, 0.5 * line_width, num_subdivisions)
        rc = lib.lv_tube_mesh(*args, None, 0, None, 0, ctypes.byref(nv), ctypes.byref(nt), ctypes.byref(nl))
        if rc != capi.LV_OK:
            raise LineVisError(rc, "lv_tube_mesh")
        v = np.zeros((max(nv.value, 1), 8), np.float32)
        t = np.zeros((max(nt.value, 1), 3), np.uint32)
        lib.lv_tube_mesh(*args, _ptr(v), nv.value, _ptr(t), nt.value, ctypes.byref(nv), ctypes.byref(nt), ctypes.byref(nl))
        return v[:nv.value], t[:nt.value], nl.value

    # -- frames
    def trace_primary(self, scene, cam, out=None):
        if out is None:
            out = np.zeros(cam.width * cam.height, HIT_DTYPE)
        st = LvStats()
        self._check(self.lib.lv_trace_primary(self.h, scene.h, ctypes.byref(cam), _ptr(out), ctypes.byref(st)))
        if isinstance(out, np.ndarray):
            out = out.reshape(cam.height, cam.width)
        return out, st.as_dict()

    def render_rtao(self, scene, cam, frame_number=0, out=None, stats=True):
        if out is None:
            out = np.zeros((cam.height, cam.width), np.float32)
        st = LvStats()
        self._check(self.lib.lv_render_rtao(self.h, scene.h, ctypes.byref(cam), frame_number, _ptr(out), ctypes.byref(st) if stats else None))
        return out, st.as_dict()

    def render_tubes(self, scene, cam, frame_number=0, out=None, stats=True):
        if out is None:
            out = np.zeros((cam.height, cam.width, 4), np.float32)
        st = LvStats()
        self._check(self.lib.lv_render_tubes(self.h, scene.h, ctypes.byref(cam), frame_number, _ptr(out), ctypes.byref(st) if stats else None))
        return out, st.as_dict()

    # AO-sample-batch shards: lv_render_tubes in three stages (include/linevis_b200.h); sharding.SampleShards drives them
    def sao_primary(self, scene, cam, frame_number=0):
        """-> (device address of this rank's hit list, number of 48-byte records)"""
        ptr, n = ctypes.c_void_p(), ctypes.c_uint32()
        self._check(self.lib.lv_sao_primary(self.h, scene.h, ctypes.byref(cam), frame_number, ctypes.byref(ptr), ctypes.byref(n)))
        return int(ptr.value or 0), int(n.value)

    def sao_trace(self, scene, cam, frame_number, hits, n_hits, sample_first, sample_count, occ):
        self._check(self.lib.lv_sao_trace(self.h, scene.h, ctypes.byref(cam), frame_number, _ptr(hits), n_hits, sample_first, sample_count, _ptr(occ)))

    def sao_finish(self, scene, cam, frame_number, occ_parts, n_parts, out, stats=True):
        st = LvStats()
        self._check(self.lib.lv_sao_finish(self.h, scene.h, ctypes.byref(cam), frame_number, _ptr(occ_parts), n_parts, _ptr(out), ctypes.byref(st) if stats else None))
        return out, st.as_dict()

    def render_ppll(self, scene, cam, max_frags=100, sort_mode="priority_queue", linked_list_size=0, out=None, stats=True):
        if out is None:
            out = np.zeros((cam.height, cam.width, 4), np.float32)
        mode = SORT_MODES[sort_mode] if isinstance(sort_mode, str) else int(sort_mode)
        st = LvStats()
        self._check(self.lib.lv_render_ppll(self.h, scene.h, ctypes.byref(cam), max_frags, mode, linked_list_size, _ptr(out),
                                            ctypes.byref(st) if stats else None))
        return out, st.as_dict()

    def ppll_clear(self, cam, linked_list_size=0):
        self._check(self.lib.lv_ppll_clear(self.h, ctypes.byref(cam), linked_list_size))

    def ppll_gather(self, scene, cam, stats=True):
        st = LvStats()
        self._check(self.lib.lv_ppll_gather(self.h, scene.h, ctypes.byref(cam), ctypes.byref(st) if stats else None))
        return st.as_dict()

    def ppll_resolve(self, cam, max_frags=100, sort_mode="priority_queue", out=None, stats=True):
        if out is None:
            out = np.zeros((cam.height, cam.width, 4), np.float32)
        mode = SORT_MODES[sort_mode] if isinstance(sort_mode, str) else int(sort_mode)
        st = LvStats()
        self._check(self.lib.lv_ppll_resolve(self.h, ctypes.byref(cam), max_frags, mode, _ptr(out), ctypes.byref(st) if stats else None))
        return out, st.as_dict()

    def ppll_read(self):
        cnt, pw, ph = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self.lib.lv_ppll_read(self.h, ctypes.byref(cnt), None, 0, None, 0, ctypes.byref(pw), ctypes.byref(ph)))
        heads = np.zeros(pw.value * ph.value, np.uint32)
        nodes = np.zeros(max(cnt.value, 1), NODE_DTYPE)
        self._check(self.lib.lv_ppll_read(self.h, ctypes.byref(cnt), _ptr(heads), heads.size, _ptr(nodes), nodes.size, None, None))
        return dict(counter=cnt.value, heads=heads, nodes=nodes[:cnt.value], padded=(pw.value, ph.value))


class Scene:
    """Device-resident segment soup + BVH (lv_scene_create; replaces getLinePassTubeAabbRenderData + TLAS build)."""

    def __init__(self, ctx, pos, attr, seg_idx, line_width):
        self.ctx = ctx
        h = ctypes.c_void_p()
        if isinstance(pos, np.ndarray):
            pos = np.ascontiguousarray(pos, np.float32)
            attr = np.ascontiguousarray(attr, np.float32)
            seg_idx = np.ascontiguousarray(seg_idx, np.uint32)
            n_pt, n_seg = pos.shape[0], seg_idx.shape[0]
            rc = ctx.lib.lv_scene_create(ctx.h, ctypes.byref(h), _ptr(pos), _ptr(attr), _ptr(seg_idx), n_pt, n_seg, line_width)
        else:  # torch CUDA tensors
            n_pt, n_seg = pos.shape[0], seg_idx.shape[0]
            rc = ctx.lib.lv_scene_create_device(ctx.h, ctypes.byref(h), _ptr(pos), _ptr(attr), _ptr(seg_idx), n_pt, n_seg, line_width)
        ctx._check(rc)
        self.h = h
        self.n_seg = n_seg
        self._n_pt = n_pt

    def info(self):
        ns, nn, ms = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_float()
        bb = np.zeros(6, np.float32)
        self.ctx._check(self.ctx.lib.lv_scene_info(self.h, ctypes.byref(ns), ctypes.byref(nn), ctypes.byref(ms), _ptr(bb)))
        return dict(n_seg=ns.value, n_nodes=nn.value, build_ms=ms.value, aabb=bb)

    def bvh_nodes(self):
        n = self.info()["n_nodes"]
        nodes = np.zeros(max(n, 1), BVH_NODE_DTYPE)
        self.ctx._check(self.ctx.lib.lv_scene_copy_bvh(self.h, _ptr(nodes), nodes.nbytes))
        return nodes[:n]

    # -- line frames + object-space AO prebaker ("RTAO (Prebaker)")
    def set_lines(self, pos, tangent, normal, line_offsets):
        """lv_scene_set_lines: per-point tangents / normals and the polyline structure (host arrays)."""
        pos, tangent, normal = (np.ascontiguousarray(a, np.float32) for a in (pos, tangent, normal))
        off = np.ascontiguousarray(line_offsets, np.uint64)
        self.ctx._check(self.ctx.lib.lv_scene_set_lines(self.h, _ptr(pos), _ptr(tangent), _ptr(normal), pos.shape[0], _ptr(off), len(off) - 1))

    def ao_bake(self, n_iterations=0, stats=True):
        st = LvStats()
        self.ctx._check(self.ctx.lib.lv_ao_bake(self.ctx.h, self.h, n_iterations, ctypes.byref(st) if stats else None))
        return st.as_dict()

    def ao_set_vertex_range(self, first_vertex, n_vertices=0):
        """Multi-GPU baking: this context bakes only the parametrization vertices [first_vertex, first_vertex + n_vertices)."""
        self.ctx._check(self.ctx.lib.lv_ao_set_vertex_range(self.h, int(first_vertex), int(n_vertices)))

    def ao_factors_ptr(self):
        """(device address, number of floats) of the baked factor buffer, for the slice exchange between ranks."""
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        self.ctx._check(self.ctx.lib.lv_ao_factors(self.h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def ao_bake_reset(self):
        self.ctx._check(self.ctx.lib.lv_ao_bake_reset(self.h))

    def ao_read(self):
        npar, nsub, done = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        lib = self.ctx.lib
        self.ctx._check(lib.lv_ao_read(self.h, None, 0, None, 0, None, 0, ctypes.byref(npar), ctypes.byref(nsub), ctypes.byref(done)))
        n_pt = self._n_pt
        f = np.zeros(max(npar.value * nsub.value, 1), np.float32)
        w = np.zeros(max(n_pt, 1), np.float32)
        sl = np.zeros(max(npar.value, 1), np.float32)
        self.ctx._check(lib.lv_ao_read(self.h, _ptr(f), f.size, _ptr(w), w.size, _ptr(sl), sl.size, None, None, None))
        return dict(factors=f[:npar.value * nsub.value].reshape(npar.value, nsub.value), blending_weights=w[:n_pt],
                    sampling_locations=sl[:npar.value], n_param=npar.value, n_subdiv=nsub.value, iterations_done=done.value)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.lv_scene_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
