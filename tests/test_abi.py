"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol include/linevis_b200.h
declares; the product has no CPU fallback and never touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

import linevis_b200 as lv
from linevis_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import linevis_b200.build as b
    b.build()
    return capi.load_library()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "linevis_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lv_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.ABI_SYMBOLS), declared ^ set(capi.ABI_SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lv_[a-z0-9_]+)$", nm, flags=re.M))
    assert declared <= exported, declared - exported
    assert lib.lv_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes the C side compiles to (checked against a tiny C program using the real header)
    src = r'''
#include <stdio.h>
#include "linevis_b200.h"
int main(void) { printf("%zu %zu %zu %zu\n", sizeof(lv_camera), sizeof(lv_stats), sizeof(lv_hit), sizeof(lv_ppll_node)); return 0; }
'''
    exe = "/tmp/lv_sizes"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    out = subprocess.run([exe], capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(lv.LvCamera), ctypes.sizeof(capi.LvStats), capi.HIT_DTYPE.itemsize, capi.NODE_DTYPE.itemsize]
    assert capi.NODE_DTYPE.itemsize == 12   # LinkedListFragmentNode, LinkedListHeader.glsl:36-43


def test_built_for_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lv.LineVisError) as e:
        lv.Context(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "linevis_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), (dp, f)
                assert not re.search(r"#include\s+[\"<][^\">]*lvo_", text) and "liblvo" not in text and "lvo_" not in text, (dp, f)
    ldd = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "lvo" not in ldd


def test_camera_matrices_are_consistent():
    import numpy as np
    cam = lv.make_camera(640, 360)
    v = np.array(cam.view).reshape(4, 4).T
    iv = np.array(cam.inv_view).reshape(4, 4).T
    p = np.array(cam.proj).reshape(4, 4).T
    ip = np.array(cam.inv_proj).reshape(4, 4).T
    assert np.allclose(v @ iv, np.eye(4), atol=1e-5) and np.allclose(p @ ip, np.eye(4), atol=1e-4)
    assert np.allclose(iv[:3, 3], [0, 0, 0.8]) and abs(cam.fov_y - 2 * np.arctan(0.5)) < 1e-7


def test_synthetic_scenes_shapes_and_normalisation():
    import numpy as np
    from linevis_b200 import scenes
    pos, attr, seg = scenes.helix_lines()
    assert seg.shape == (100000, 2) and pos.shape[0] == 400 * 251          # config 2: 100 k segments
    ext = pos.max(0) - pos.min(0)
    assert abs(ext.max() - 0.5) < 1e-5 and np.allclose(pos.max(0) + pos.min(0), 0, atol=1e-5)
    assert attr.min() >= 0 and attr.max() <= 1
    pos, attr, seg = scenes.random_segments(5000)
    d = np.linalg.norm(pos[seg[:, 1]] - pos[seg[:, 0]], axis=1)
    assert np.allclose(d, 0.01, atol=1e-6)
    pos, attr, seg = scenes.curl_noise_streamlines(50, 41)
    assert seg.shape[0] <= 50 * 40 and seg.shape[0] > 50 * 30 and np.isfinite(pos).all()
    lut = scenes.standard_transfer_function(64, (0.1, 0.6))
    assert lut.shape == (64, 4) and abs(lut[0, 3] - 0.1) < 1e-6 and abs(lut[-1, 3] - 0.6) < 1e-6
