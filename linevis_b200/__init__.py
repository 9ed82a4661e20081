"""linevis_b200 -- B200-native (sm_100a) implementation of LineVis's two hot paths behind a C ABI:
ray-traced tubes + RTAO, and per-pixel-linked-list OIT.  See DESIGN.md / INTEGRATION.md.

The product path is CUDA only.  Importing this package never imports anything from oracle/."""
from .camera import LvCamera, make_camera, look_at, perspective  # noqa: F401
from .api import Context, Scene  # noqa: F401
from .capi import LineVisError, SORT_MODES, HIT_DTYPE, NODE_DTYPE  # noqa: F401
from . import scenes  # noqa: F401

__all__ = ["Context", "Scene", "LvCamera", "make_camera", "look_at", "perspective", "LineVisError", "SORT_MODES", "scenes"]
