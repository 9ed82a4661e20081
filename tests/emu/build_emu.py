"""Builds tests/emu/liblinevis_b200_emu.so: the product's C ABI (linevis_b200/csrc/lv_api.cu + kernels, unchanged source) compiled
for the HOST against the SIMT emulator emu_cuda.hpp.  The only source transformation is syntactic: `kernel<<<grid, block, smem,
stream>>>(args)` becomes `EMU_LAUNCH(kernel, grid, block, smem, stream)(args)` and `extern __shared__` arrays become pointers to
the launch's dynamic shared memory.  TEST INFRASTRUCTURE ONLY."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(HERE, "..", "..", "linevis_b200", "csrc"))
CUDA_INC = "/usr/local/cuda/include"
OUT = os.path.join(HERE, "liblinevis_b200_emu.so")
GEN = os.path.join(HERE, "_gen")


def _split_top_level(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def transform(text):
    out, pos = "", 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            break
        j = text.index(">>>", i)
        # kernel expression: identifier with optional template argument list, scanning backwards from `<<<`
        k = i
        if text[k - 1] == ">":
            depth = 0
            while True:
                k -= 1
                if text[k] == ">":
                    depth += 1
                elif text[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while k > 0 and (text[k - 1].isalnum() or text[k - 1] in "_:"):
            k -= 1
        kernel = text[k:i]
        cfg = _split_top_level(text[i + 3:j])
        while len(cfg) < 4:
            cfg.append("0")
        out += text[pos:k] + "EMU_LAUNCH((%s), (%s), (%s), (%s), (%s))" % (kernel, cfg[0], cfg[1], cfg[2], cfg[3])
        pos = j + 3
    out += text[pos:]
    out = re.sub(r"extern\s+__shared__\s+unsigned long long\s+(\w+)\[\];", r"unsigned long long* \1 = static_cast<unsigned long long*>(emu::dyn_smem());", out)
    out = out.replace("cudaGetDeviceProperties(", "emu_get_device_properties(")
    return out


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, f) for f in ("emu_cuda.hpp", "emu_cudart.inc", "build_emu.py")]


def build(force=False):
    if os.environ.get("LV_EMU_LIB"):   # a prebuilt variant, e.g. the AddressSanitizer build of tools/emu_asan.sh
        return os.environ["LV_EMU_LIB"]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in sources()):
        return OUT
    os.makedirs(GEN, exist_ok=True)
    for f in os.listdir(CSRC):
        with open(os.path.join(CSRC, f)) as fh:
            t = transform(fh.read())
        with open(os.path.join(GEN, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w") as fh:
            fh.write(t)
    main = os.path.join(GEN, "emu_main.cpp")
    with open(main, "w") as fh:
        fh.write('#include "../emu_cuda.hpp"\n#include "../emu_cudart.inc"\n#include "lv_api.cpp"\n')
    # lv_api.cu includes "../../include/linevis_b200.h" relative to csrc: keep that path valid from _gen.
    # -Bsymbolic: the library's own cuda* / lv_* definitions win over same-named symbols of the real library / libcudart that may
    # already live in the process (the CPU suite also loads liblinevis_b200.so for the ABI checks)
    cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-DLV_HOST_EMU", "-I" + CUDA_INC, "-I" + os.path.join(HERE, "..", "..", "linevis_b200", "csrc"),
           "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v3", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-Wno-attributes", "-Wno-unknown-pragmas", "-Wno-subobject-linkage",
           "-o", OUT + ".tmp", main]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        errors = [l for l in r.stderr.splitlines() if "error" in l or "Error" in l]
        raise RuntimeError("emulation build failed:\n" + "\n".join(errors[:40]) + "\n...\n" + r.stderr[-3000:])
    os.replace(OUT + ".tmp", OUT)   # a process that has the previous build mapped keeps its own inode
    return OUT


if __name__ == "__main__":
    print(build(force=True))
