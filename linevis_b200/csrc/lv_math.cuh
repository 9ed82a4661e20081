// lv_math.cuh -- float3 helpers and deterministic transcendental functions for the sm_100a kernels.
//
// The whole library is compiled with -fmad=false: every product/sum below is a separately rounded IEEE
// float32 operation in the written order, so that hit decisions and shaded colours are reproducible
// bit-for-bit against a strict-IEEE CPU evaluation of the same GLSL (DESIGN.md "Float conventions").
// FMA is used only where it is spelled out (__fmaf_rn in the BVH slab test, which is conservative).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Per-thread device functions are declared LV_DEV.  In the product build that is __device__ __forceinline__.  With
// -DLV_HOST_EMU (tests/emu only, plain g++) the same source is compiled for the host, so that per-thread device code can be
// checked against the oracle in the CPU test suite; the handful of device intrinsics it uses get host shims below.
#ifdef LV_HOST_EMU
#include <cmath>
#include <cstring>
#define LV_DEV inline
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {   // PRMT, default mode: result byte i = byte (nibble i of s) of {y, x}
    const uint64_t v = (uint64_t(y) << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= uint32_t((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
// integer min / max overloads nvcc provides in the global namespace
inline int max(int a, int b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline unsigned max(unsigned a, unsigned b) { return a < b ? b : a; }
inline unsigned min(unsigned a, unsigned b) { return b < a ? b : a; }
#else
#define LV_DEV __device__ __forceinline__
#endif

namespace lv {

struct Vec3 { float x, y, z; };
struct Vec4 { float x, y, z, w; };

LV_DEV Vec3 v3(float x, float y, float z) { Vec3 r; r.x = x; r.y = y; r.z = z; return r; }
LV_DEV Vec4 v4(float x, float y, float z, float w) { Vec4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
LV_DEV Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
LV_DEV Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
LV_DEV Vec3 operator*(Vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
LV_DEV Vec3 operator*(float s, Vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
LV_DEV float dot3(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
LV_DEV Vec3 cross3(Vec3 a, Vec3 b) {
    return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
LV_DEV float length3(Vec3 a) { return sqrtf(dot3(a, a)); }
LV_DEV Vec3 normalize3(Vec3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return a * inv; }
// comparison-based min/max/clamp (GLSL leaves NaN handling open; fixed here)
LV_DEV float minf_(float a, float b) { return (b < a) ? b : a; }
LV_DEV float maxf_(float a, float b) { return (a < b) ? b : a; }
LV_DEV float clampf_(float x, float lo, float hi) { return minf_(maxf_(x, lo), hi); }
LV_DEV float mixf_(float a, float b, float t) { return a * (1.0f - t) + b * t; }
LV_DEV float smoothstepf_(float e0, float e1, float x) {
    float t = clampf_((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// column-major mat4 * vec4, summed left to right over the columns
LV_DEV Vec4 mat_mul(const float* m, Vec4 v) {
    Vec4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}

// ---- deterministic log2 / exp2 / pow / sincos (basic ops only; spec in DESIGN.md) ----
LV_DEV float det_log2(float x) {
    int eadj = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; eadj = -23; }
    uint32_t b = __float_as_uint(x);
    int e = int((b >> 23) & 0xffu) - 127 + eadj;
    float m = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float p = 0.111111111f;
    p = p * z + 0.142857143f;
    p = p * z + 0.2f;
    p = p * z + 0.333333333f;
    p = p * z + 1.0f;
    float lnm = (2.0f * s) * p;
    return float(e) + lnm * 1.44269504f;
}
LV_DEV float det_exp2(float z) {
    if (!(z >= -126.0f)) return 0.0f;
    if (z > 127.0f) z = 127.0f;
    float n = floorf(z + 0.5f);
    float r = z - n;
    float t = r * 0.693147182f;
    float p = 1.98412698e-4f;
    p = p * t + 1.38888889e-3f;
    p = p * t + 8.33333333e-3f;
    p = p * t + 4.16666667e-2f;
    p = p * t + 0.166666667f;
    p = p * t + 0.5f;
    p = p * t + 1.0f;
    p = p * t + 1.0f;
    float scale = __uint_as_float(uint32_t(int(n) + 127) << 23);
    return p * scale;
}
LV_DEV float det_pow(float x, float y) {
    if (!(x > 0.0f)) return 0.0f;
    return det_exp2(y * det_log2(x));
}
LV_DEV void det_sincos2pi(float xi, float& c, float& s) {
    float a = 4.0f * xi;
    float q = floorf(a + 0.5f);
    float r = (a - q) * 1.57079633f;
    float r2 = r * r;
    float ps = 2.75573192e-6f;
    ps = ps * r2 + -1.98412698e-4f;
    ps = ps * r2 + 8.33333333e-3f;
    ps = ps * r2 + -0.166666667f;
    ps = ps * r2 + 1.0f;
    float sr = r * ps;
    float pc = -2.75573192e-7f;
    pc = pc * r2 + 2.48015873e-5f;
    pc = pc * r2 + -1.38888889e-3f;
    pc = pc * r2 + 4.16666667e-2f;
    pc = pc * r2 + -0.5f;
    pc = pc * r2 + 1.0f;
    int qi = int(q) & 3;
    if (qi == 0) { c = pc; s = sr; }
    else if (qi == 1) { c = -sr; s = pc; }
    else if (qi == 2) { c = -pc; s = -sr; }
    else { c = sr; s = -pc; }
}

// acos for the tube angle phi (reference TubeRayTracing.glsl:554).  GLSL leaves acos undefined for |x| > 1 (the dot product of
// two normalised vectors can exceed 1 by an ulp): the argument is clamped.  |x| <= 0.5: pi/2 - asin(x); otherwise
// 2 asin(sqrt((1 - |x|) / 2)), mirrored for x < 0; asin(s) = s + s z P(z), z = s^2 (abs error 3e-7 vs libm).
LV_DEV float det_acos(float x) {
    x = clampf_(x, -1.0f, 1.0f);
    const float ax = fabsf(x);
    const bool small = ax <= 0.5f;
    float z, s;
    if (small) { z = x * x; s = x; }
    else { z = (1.0f - ax) * 0.5f; s = sqrtf(z); }
    float p = 3.380591050e-02f;
    p = p * z + 1.707774773e-02f;
    p = p * z + 3.111618385e-02f;
    p = p * z + 4.459802806e-02f;
    p = p * z + 7.500098646e-02f;
    p = p * z + 1.666666567e-01f;
    const float r = s + (s * z) * p;
    if (small) return 1.57079633f - r;
    if (x > 0.0f) return 2.0f * r;
    return 3.14159265f - 2.0f * r;
}

// ---- RNG: tea / lcg / rnd (reference Data/Shaders/Renderers/RayTracing/RayTracingUtilities.glsl:134-181) ----
LV_DEV uint32_t tea(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
LV_DEV uint32_t lcg(uint32_t& prev) {
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00FFFFFFu;
}
LV_DEV float rnd(uint32_t& seed) { return float(lcg(seed)) / float(0x01000000); }
// state after n calls of lcg(): the affine map x -> a x + c composed n times by square-and-multiply (mod 2^32).  The AO
// prebaker draws all random numbers of a vertex from ONE stream (VulkanAmbientOcclusionBaker.glsl:196,267); a ray in
// the middle of that stream jumps to its position instead of replaying it.
LV_DEV uint32_t lcg_skip(uint32_t state, uint32_t n) {
    uint32_t a = 1664525u, c = 1013904223u, acc_a = 1u, acc_c = 0u;
    while (n) {
        if (n & 1u) { acc_a = acc_a * a; acc_c = acc_c * a + c; }
        c = (a + 1u) * c;
        a = a * a;
        n >>= 1;
    }
    return acc_a * state + acc_c;
}

}  // namespace lv
