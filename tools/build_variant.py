"""Build another flavour of the CUDA library for A/B measurements (selected at run time with LINEVIS_B200_LIB=<path>):

    python tools/build_variant.py fmad      -> build/liblinevis_b200_fmad.so   (-fmad=true: FMA contraction on; SURVEY 7 hard part 5)

The shipped library stays the strict-IEEE build (DESIGN.md rule 1)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linevis_b200.build as b

VARIANTS = {"fmad": {"-fmad=false": "-fmad=true"}, "flush8": {}, "flush16": {}, "flush24": {}, "flush32": {}}
EXTRA = {"flush8": ["-DLV_PACKET_FLUSH=8"], "flush16": ["-DLV_PACKET_FLUSH=16"], "flush24": ["-DLV_PACKET_FLUSH=24"], "flush32": ["-DLV_PACKET_FLUSH=32"]}


def build(name, extra=()):
    out = os.path.join(os.path.dirname(b._HERE), "build", "liblinevis_b200_%s.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    flags = [VARIANTS[name].get(f, f) for f in b.NVCC_FLAGS] + list(extra) + EXTRA.get(name, [])
    cmd = [b._nvcc()] + flags + ["-o", out, os.path.join(b.CSRC, "lv_api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    for n in sys.argv[1:] or list(VARIANTS):
        print(build(n))
