// lv_math.cuh -- float3 helpers and deterministic transcendental functions for the sm_100a kernels.
//
// The whole library is compiled with -fmad=false: every product/sum below is a separately rounded IEEE
// float32 operation in the written order, so that hit decisions and shaded colours are reproducible
// bit-for-bit against a strict-IEEE CPU evaluation of the same GLSL (DESIGN.md "Float conventions").
// FMA is used only where it is spelled out (__fmaf_rn in the BVH slab test, which is conservative).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lv {

struct Vec3 { float x, y, z; };
struct Vec4 { float x, y, z, w; };

__device__ __forceinline__ Vec3 v3(float x, float y, float z) { Vec3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ Vec4 v4(float x, float y, float z, float w) { Vec4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
__device__ __forceinline__ Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ Vec3 operator*(Vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ Vec3 operator*(float s, Vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot3(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ Vec3 cross3(Vec3 a, Vec3 b) {
    return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float length3(Vec3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ Vec3 normalize3(Vec3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return a * inv; }
// comparison-based min/max/clamp (GLSL leaves NaN handling open; fixed here)
__device__ __forceinline__ float minf_(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float maxf_(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float clampf_(float x, float lo, float hi) { return minf_(maxf_(x, lo), hi); }
__device__ __forceinline__ float mixf_(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float smoothstepf_(float e0, float e1, float x) {
    float t = clampf_((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// column-major mat4 * vec4, summed left to right over the columns
__device__ __forceinline__ Vec4 mat_mul(const float* m, Vec4 v) {
    Vec4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}

// ---- deterministic log2 / exp2 / pow / sincos (basic ops only; spec in DESIGN.md) ----
__device__ __forceinline__ float det_log2(float x) {
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
__device__ __forceinline__ float det_exp2(float z) {
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
__device__ __forceinline__ float det_pow(float x, float y) {
    if (!(x > 0.0f)) return 0.0f;
    return det_exp2(y * det_log2(x));
}
__device__ __forceinline__ void det_sincos2pi(float xi, float& c, float& s) {
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

// ---- RNG: tea / lcg / rnd (reference Data/Shaders/Renderers/RayTracing/RayTracingUtilities.glsl:134-181) ----
__device__ __forceinline__ uint32_t tea(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
__device__ __forceinline__ uint32_t lcg(uint32_t& prev) {
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00FFFFFFu;
}
__device__ __forceinline__ float rnd(uint32_t& seed) { return float(lcg(seed)) / float(0x01000000); }

}  // namespace lv
