"""In-tree build of the CUDA library (nvcc, sm_100a only) and of the C++ host adapter self-test."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "liblinevis_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # strict float32: no FMA contraction, IEEE div/sqrt, no flush-to-zero (parity with the CPU oracle)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    # host code too (the AO prebaker's arc-length parametrization runs on the host and must match the oracle bit for bit)
    "-Xcompiler", "-ffp-contract=off",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources():
    inc = os.path.join(_HERE, "..", "include", "linevis_b200.h")
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp", ".h"))] + [inc]


def build(force=False, verbose=False):
    """Compile linevis_b200/csrc/*.cu into linevis_b200/liblinevis_b200.so (cross-compiles without a GPU)."""
    if not force and not _stale(LIB, sources()):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "lv_api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB
