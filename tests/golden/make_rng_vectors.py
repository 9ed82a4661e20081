"""Generates tests/golden/rng_vectors.npz from a pure-Python integer restatement of
Data/Shaders/Renderers/RayTracing/RayTracingUtilities.glsl:134-181 (tea, lcg, rnd), independent of oracle/ and of the CUDA code."""
import numpy as np

M = 0xFFFFFFFF


def tea(v0, v1):
    s0 = 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & M
        v0 = (v0 + (((((v1 << 4) & M) + 0xA341316C) & M) ^ ((v1 + s0) & M) ^ (((v1 >> 5) + 0xC8013EA4) & M))) & M
        v1 = (v1 + (((((v0 << 4) & M) + 0xAD90777D) & M) ^ ((v0 + s0) & M) ^ (((v0 >> 5) + 0x7E95761E) & M))) & M
    return v0


if __name__ == "__main__":
    rng = np.random.default_rng(20240607)
    tin = np.concatenate([np.array([[0, 0], [1, 0], [0, 1], [M, M], [2073599, 15], [8294399, 63]], np.uint64),
                          rng.integers(0, 2**32, (58, 2), dtype=np.uint64)]).astype(np.uint32)
    tout = np.array([tea(int(a), int(b)) for a, b in tin], np.uint32)
    seed = 0x9E3779B9
    s, lo, ro = seed, [], []
    for _ in range(256):
        s = (1664525 * s + 1013904223) & M
        lo.append(s & 0x00FFFFFF)
        ro.append(np.float32(s & 0x00FFFFFF) / np.float32(16777216.0))
    np.savez("tests/golden/rng_vectors.npz", tea_in=tin, tea_out=tout, lcg_seed=np.uint32(seed),
             lcg_out=np.array(lo, np.uint32), rnd_out=np.array(ro, np.float32))
