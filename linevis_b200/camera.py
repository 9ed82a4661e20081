"""Camera / frame description shared by the C ABI (include/linevis_b200.h: lv_camera) and the oracle.

Mirrors the camera part of LineData::LineUniformData (reference src/LineData/LineData.hpp:428-464,
filled at src/LineData/LineData.cpp:1275-1319).  sgl::Camera is not part of the reference tree, so the
matrices are explicit inputs; `make_camera` builds them the way glm::lookAt / glm::perspective do.
"""
import ctypes
import math

import numpy as np


class LvCamera(ctypes.Structure):
    _fields_ = [
        ("view", ctypes.c_float * 16),
        ("proj", ctypes.c_float * 16),
        ("inv_view", ctypes.c_float * 16),
        ("inv_proj", ctypes.c_float * 16),
        ("position", ctypes.c_float * 3),
        ("fov_y", ctypes.c_float),
        ("background", ctypes.c_float * 4),
        ("width", ctypes.c_uint32),
        ("height", ctypes.c_uint32),
        ("near_dist", ctypes.c_float),
        ("far_dist", ctypes.c_float),
    ]


def look_at(eye, center, up):
    """glm::lookAt (right-handed); returns a 4x4 row-major numpy matrix (float64)."""
    eye = np.asarray(eye, np.float64)
    f = np.asarray(center, np.float64) - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, np.asarray(up, np.float64))
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, f @ eye
    return m


def perspective(fov_y, aspect, near, far):
    """glm::perspective (RH, depth -1..1); 4x4 row-major numpy matrix (float64)."""
    t = math.tan(fov_y / 2.0)
    m = np.zeros((4, 4))
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = -(far + near) / (far - near)
    m[2, 3] = -(2.0 * far * near) / (far - near)
    m[3, 2] = -1.0
    return m


def _col_major(m):
    return np.ascontiguousarray(np.asarray(m, np.float64).T.reshape(16)).astype(np.float32)


def make_camera(width, height, eye=(0.0, 0.0, 0.8), center=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0),
                fov_y=2.0 * math.atan(0.5), near=0.01, far=100.0, background=(1.0, 1.0, 1.0, 1.0)):
    """Default = SURVEY.md 8d common camera: (0,0,0.8) looking -z, fovY = 2 atan(1/2), near .01, far 100."""
    v = look_at(eye, center, up)
    p = perspective(fov_y, width / float(height), near, far)
    cam = LvCamera()
    cam.view[:] = _col_major(v)
    cam.proj[:] = _col_major(p)
    cam.inv_view[:] = _col_major(np.linalg.inv(v))
    cam.inv_proj[:] = _col_major(np.linalg.inv(p))
    cam.position[:] = np.asarray(eye, np.float32)
    cam.fov_y = fov_y
    cam.background[:] = np.asarray(background, np.float32)
    cam.width, cam.height = int(width), int(height)
    cam.near_dist, cam.far_dist = near, far
    return cam
