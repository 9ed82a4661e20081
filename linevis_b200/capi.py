"""ctypes binding of the C ABI (include/linevis_b200.h) -- the same calls LineVis's C++ adapter makes.

The shared library is built in-tree by linevis_b200.build (nvcc, sm_100a).  There is NO CPU fallback: if the
library is missing, or no CUDA device is visible, the calls fail loudly.
"""
import ctypes
import os

import numpy as np

from .camera import LvCamera

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblinevis_b200.so")

LV_OK = 0
ERRORS = {-1: "LV_ERR_INVALID_ARGUMENT", -2: "LV_ERR_CUDA", -3: "LV_ERR_OUT_OF_MEMORY", -4: "LV_ERR_NO_DEVICE",
          -5: "LV_ERR_UNKNOWN_OPTION", -6: "LV_ERR_STATE"}

SORT_MODES = {"priority_queue": 0, "bubble": 1, "insertion": 2, "shell": 3, "max_heap": 4, "bitonic": 5,
              "quicksort": 6, "quicksort_hybrid": 7}

# every symbol include/linevis_b200.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = [
    "lv_ctx_create", "lv_ctx_destroy", "lv_last_error", "lv_last_global_error", "lv_abi_version", "lv_set_option",
    "lv_get_option", "lv_set_transfer_function", "lv_set_tile_shard", "lv_get_tile_costs", "lv_set_tile_owners", "lv_get_owned_tiles", "lv_pack_owned_tiles",
    "lv_unpack_tiles", "lv_scene_create", "lv_scene_create_device", "lv_scene_destroy", "lv_scene_info",
    "lv_scene_copy_bvh", "lv_render_tubes", "lv_sao_primary", "lv_sao_trace", "lv_sao_finish", "lv_render_ppll", "lv_trace_primary", "lv_render_rtao", "lv_ppll_clear",
    "lv_ppll_gather", "lv_ppll_resolve", "lv_ppll_read", "lv_synchronize",
    "lv_scene_set_lines", "lv_ao_parametrize", "lv_ao_bake", "lv_ao_bake_reset", "lv_ao_read",
    "lv_frame_alloc", "lv_frame_free", "lv_ipc_export", "lv_ipc_open", "lv_ipc_close", "lv_tube_mesh", "lv_frame_to_rgba8", "lv_ao_set_vertex_range", "lv_ao_factors",
]


class LvStats(ctypes.Structure):
    _fields_ = [
        ("rays_primary", ctypes.c_uint64), ("rays_ao", ctypes.c_uint64), ("traversal_steps", ctypes.c_uint64),
        ("intersections", ctypes.c_uint64), ("ao_traversal_steps", ctypes.c_uint64), ("ao_intersections", ctypes.c_uint64),
        ("pixels_hit", ctypes.c_uint64), ("frags_generated", ctypes.c_uint64),
        ("frags_stored", ctypes.c_uint64), ("frags_dropped", ctypes.c_uint64), ("frags_sorted", ctypes.c_uint64),
        ("frags_truncated", ctypes.c_uint64), ("max_depth_complexity", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
        ("ms_trace", ctypes.c_float), ("ms_rtao", ctypes.c_float), ("ms_rtao_rays", ctypes.c_float), ("ms_clear", ctypes.c_float),
        ("ms_gather", ctypes.c_float), ("ms_resolve", ctypes.c_float), ("ms_total", ctypes.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


HIT_DTYPE = np.dtype([("t", np.float32), ("prim", np.uint32), ("kind", np.uint32), ("pad", np.uint32)])
NODE_DTYPE = np.dtype([("color", np.uint32), ("depth", np.float32), ("next", np.uint32)])
BVH_NODE_DTYPE = np.dtype([("lmin", np.float32, 3), ("lref", np.uint32), ("lmax", np.float32, 3), ("lcount", np.uint32),
                           ("rmin", np.float32, 3), ("rref", np.uint32), ("rmax", np.float32, 3), ("rcount", np.uint32)])


class LineVisError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "LV_ERR"), code, message))
        self.code = code


_libs = {}


def load_library(path=None):
    """Load liblinevis_b200.so (or another build of the same C ABI, e.g. the host emulation the CPU tests use) and declare the
    prototypes.  Raises if it has not been built.  LINEVIS_B200_LIB names another CUDA build of the same sources for A/B
    measurements (e.g. the -fmad=true build of tools/build_variant.py); it is never a CPU path."""
    path = os.path.abspath(path or os.environ.get("LINEVIS_B200_LIB") or LIB_PATH)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError("%s not found: build it with `python -c 'import linevis_b200.build as b; b.build()'` "
                          "(nvcc, sm_100a).  linevis_b200 has no CPU fallback." % path)
    L = ctypes.CDLL(path)
    vp, u32, u64, f32, cp = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float, ctypes.c_char_p
    P = ctypes.POINTER
    L.lv_abi_version.restype = ctypes.c_int
    L.lv_last_error.restype = cp
    L.lv_last_error.argtypes = [vp]
    L.lv_last_global_error.restype = cp
    L.lv_ctx_create.argtypes = [P(vp), ctypes.c_int, vp]
    L.lv_ctx_destroy.argtypes = [vp]
    L.lv_set_option.argtypes = [vp, cp, cp]
    L.lv_get_option.argtypes = [vp, cp, cp, ctypes.c_size_t]
    L.lv_set_transfer_function.argtypes = [vp, vp, u32, f32, f32]
    L.lv_set_tile_shard.argtypes = [vp, u32, u32, u32]
    L.lv_get_tile_costs.argtypes = [vp, u32, u32, vp, u32]
    L.lv_set_tile_owners.argtypes = [vp, u32, u32, vp, u32]
    L.lv_get_owned_tiles.argtypes = [vp, u32, u32, vp, P(u32)]
    L.lv_pack_owned_tiles.argtypes = [vp, vp, u32, u32, vp]
    L.lv_unpack_tiles.argtypes = [vp, vp, u32, u32, u32, u32, vp]
    L.lv_scene_create.argtypes = [vp, P(vp), vp, vp, vp, u64, u64, f32]
    L.lv_scene_create_device.argtypes = [vp, P(vp), vp, vp, vp, u64, u64, f32]
    L.lv_scene_destroy.argtypes = [vp]
    L.lv_scene_info.argtypes = [vp, P(u64), P(u64), P(f32), vp]
    L.lv_scene_copy_bvh.argtypes = [vp, vp, ctypes.c_size_t]
    L.lv_render_tubes.argtypes = [vp, vp, P(LvCamera), u32, vp, P(LvStats)]
    L.lv_sao_primary.argtypes = [vp, vp, P(LvCamera), u32, P(vp), P(u32)]
    L.lv_sao_trace.argtypes = [vp, vp, P(LvCamera), u32, vp, u32, u32, u32, vp]
    L.lv_sao_finish.argtypes = [vp, vp, P(LvCamera), u32, vp, u32, vp, P(LvStats)]
    L.lv_render_ppll.argtypes = [vp, vp, P(LvCamera), u32, u32, u64, vp, P(LvStats)]
    L.lv_trace_primary.argtypes = [vp, vp, P(LvCamera), vp, P(LvStats)]
    L.lv_render_rtao.argtypes = [vp, vp, P(LvCamera), u32, vp, P(LvStats)]
    L.lv_ppll_clear.argtypes = [vp, P(LvCamera), u64]
    L.lv_ppll_gather.argtypes = [vp, vp, P(LvCamera), P(LvStats)]
    L.lv_ppll_resolve.argtypes = [vp, P(LvCamera), u32, u32, vp, P(LvStats)]
    L.lv_ppll_read.argtypes = [vp, P(u32), vp, ctypes.c_size_t, vp, ctypes.c_size_t, P(u32), P(u32)]
    L.lv_synchronize.argtypes = [vp]
    L.lv_ao_set_vertex_range.argtypes = [vp, u64, u64]
    L.lv_ao_factors.argtypes = [vp, P(vp), P(u64)]
    L.lv_frame_to_rgba8.argtypes = [vp, vp, u32, u32, vp]
    L.lv_tube_mesh.argtypes = [vp, vp, u64, f32, u32, vp, u64, vp, u64, P(u64), P(u64), P(u64)]
    L.lv_frame_alloc.argtypes = [vp, u32, u32, P(vp)]
    L.lv_frame_free.argtypes = [vp, vp]
    L.lv_ipc_export.argtypes = [vp, vp, vp]
    L.lv_ipc_open.argtypes = [vp, vp, P(vp)]
    L.lv_ipc_close.argtypes = [vp, vp]
    L.lv_scene_set_lines.argtypes = [vp, vp, vp, vp, u64, vp, u64]
    L.lv_ao_parametrize.argtypes = [vp, vp, u64, f32, vp, vp, u64, P(u64)]
    L.lv_ao_bake.argtypes = [vp, vp, u32, P(LvStats)]
    L.lv_ao_bake_reset.argtypes = [vp]
    L.lv_ao_read.argtypes = [vp, vp, ctypes.c_size_t, vp, ctypes.c_size_t, vp, ctypes.c_size_t, P(u32), P(u32), P(u32)]
    for name in ABI_SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is ctypes.c_int and name not in ("lv_abi_version",):
            fn.restype = ctypes.c_int
    _libs[path] = L
    return L


def _ptr(x):
    """Device or host pointer of a torch tensor / numpy array / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, np.ndarray):
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return ctypes.c_void_p(x.data_ptr())
    raise TypeError(type(x))
