/*
 * lvo_shaders.hpp -- CPU restatement of the LineVis GLSL shaders on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (linevis_b200/) may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the reference holds no golden vectors, known-answer tests or fixtures for this
 * path; its GLSL cannot be executed here (no Vulkan).  This file restates the cited
 * GLSL line by line in strict IEEE float32 (compile with -ffp-contract=off, no fast-math).  What can
 * be pinned is pinned in tests/: analytic known answers, the madmann91/bvh smoke vectors
 * (oracle/_ref), and cross-checks between this oracle's own BVH and the reference's bvh library.
 *
 * Every function cites the reference file:line (relative to the LineVis tree) it follows.
 *
 * Float conventions fixed "by specification" where GLSL leaves them open:
 *   normalize(v) = v * (1 / sqrt(dot(v,v)));   mat*vec = ((c0*x + c1*y) + c2*z) + c3*w;
 *   pow / sin / cos use the deterministic basic-op routines below (det_pow, det_sincos2pi);
 *   min/max/clamp are comparison based;  texture() = float linear filter, clamp-to-edge.
 */
#ifndef LVO_SHADERS_HPP
#define LVO_SHADERS_HPP

#include <cmath>
#include <cstdint>
#include <cstring>

namespace lvo {

// ----------------------------------------------------------------------------------------------
// vector helpers (GLSL semantics, fixed evaluation order)
// ----------------------------------------------------------------------------------------------
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

static inline vec3 V3(float x, float y, float z) { return vec3{x, y, z}; }
static inline vec3 operator+(vec3 a, vec3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 operator-(vec3 a, vec3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 operator*(vec3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline vec3 operator*(float s, vec3 a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline vec3 cross(vec3 a, vec3 b) {
    return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline float length(vec3 a) { return sqrtf(dot(a, a)); }
static inline vec3 normalize(vec3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
static inline float fmin_(float a, float b) { return (b < a) ? b : a; }
static inline float fmax_(float a, float b) { return (a < b) ? b : a; }
static inline float clamp(float x, float lo, float hi) { return fmin_(fmax_(x, lo), hi); }
static inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// column-major mat4 (glm): m[c*4+r]
static inline vec4 mul(const float* m, vec4 v) {
    vec4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}

// ----------------------------------------------------------------------------------------------
// deterministic transcendental functions (basic IEEE ops only; same spec as the CUDA side)
// ----------------------------------------------------------------------------------------------
static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

static inline float det_log2(float x) {  // x > 0
    int eadj = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; eadj = -23; }
    uint32_t b = f2u(x);
    int e = int((b >> 23) & 0xffu) - 127 + eadj;
    float m = u2f((b & 0x007fffffu) | 0x3f800000u);
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
static inline float det_exp2(float z) {
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
    float scale = u2f(uint32_t(int(n) + 127) << 23);
    return p * scale;
}
// GLSL pow(x, y) for x in [0, inf), y > 0
static inline float det_pow(float x, float y) {
    if (!(x > 0.0f)) return 0.0f;
    return det_exp2(y * det_log2(x));
}
// (cos, sin) of 2*pi*xi for xi in [0, 1)
static inline void det_sincos2pi(float xi, float& c, float& s) {
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
    float cr = pc;
    int qi = int(q) & 3;
    if (qi == 0) { c = cr; s = sr; }
    else if (qi == 1) { c = -sr; s = cr; }
    else if (qi == 2) { c = -cr; s = -sr; }
    else { c = sr; s = -cr; }
}

// acos(x) for the tube angle phi (TubeRayTracing.glsl:554).  GLSL leaves acos undefined for |x| > 1 (a dot product of two
// normalised vectors can exceed 1 by an ulp), so the argument is clamped.  |x| <= 0.5: pi/2 - asin(x);
// else 2 asin(sqrt((1-|x|)/2)) mirrored; asin(s) = s + s z P(z), z = s^2, P fitted on [0, 0.25] (abs error 3e-7 vs libm).
static inline float det_acos(float x) {
    x = clamp(x, -1.0f, 1.0f);
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

// ----------------------------------------------------------------------------------------------
// RNG -- Data/Shaders/Renderers/RayTracing/RayTracingUtilities.glsl:134-181 (integer exact)
// ----------------------------------------------------------------------------------------------
static inline uint32_t tea(uint32_t val0, uint32_t val1) {  // :134-149
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
static inline uint32_t lcg(uint32_t& prev) {  // :168-174
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00FFFFFFu;
}
static inline float rnd(uint32_t& seed) {  // :177-180
    return float(lcg(seed)) / float(0x01000000);
}

// ----------------------------------------------------------------------------------------------
// intersection -- Data/Shaders/Renderers/RayTracing/RayIntersectionTestsVulkan.glsl
// ----------------------------------------------------------------------------------------------
#define LVO_SQR(x) ((x) * (x))
static inline float squareVec(vec3 v) {  // :31-33
    return LVO_SQR(v.x) + LVO_SQR(v.y) + LVO_SQR(v.z);
}

// :39-72
static inline bool raySphereIntersection(vec3 rayOrigin, vec3 rayDirection, vec3 sphereCenter,
                                         float sphereRadius, float& hitT) {
    float A = LVO_SQR(rayDirection.x) + LVO_SQR(rayDirection.y) + LVO_SQR(rayDirection.z);
    float B = 2.0f * (rayDirection.x * (rayOrigin.x - sphereCenter.x)
                      + rayDirection.y * (rayOrigin.y - sphereCenter.y)
                      + rayDirection.z * (rayOrigin.z - sphereCenter.z));
    float C = LVO_SQR(rayOrigin.x - sphereCenter.x) + LVO_SQR(rayOrigin.y - sphereCenter.y)
              + LVO_SQR(rayOrigin.z - sphereCenter.z) - LVO_SQR(sphereRadius);
    float discriminant = LVO_SQR(B) - 4.0f * A * C;
    if (discriminant < 0.0f) return false;
    float discriminantSqrt = sqrtf(discriminant);
    float t0 = (-B - discriminantSqrt) / (2.0f * A);
    float t1 = (-B + discriminantSqrt) / (2.0f * A);
    hitT = t0;
    if (t0 >= 0.0f) hitT = t0;
    else if (t1 >= 0.0f) hitT = t1;
    else return false;
    return true;
}

// :78-119
static inline bool rayTubeIntersection(vec3 rayOrigin, vec3 rayDirection, vec3 tubeStart, vec3 tubeEnd,
                                       float tubeRadius, float& hitT) {
    vec3 tubeDirection = normalize(tubeEnd - tubeStart);
    vec3 deltaP = rayOrigin - tubeStart;
    vec3 dPerp = rayDirection - dot(rayDirection, tubeDirection) * tubeDirection;
    vec3 pPerp = deltaP - dot(deltaP, tubeDirection) * tubeDirection;
    float A = squareVec(dPerp);
    float B = 2.0f * dot(dPerp, pPerp);
    float C = squareVec(pPerp) - LVO_SQR(tubeRadius);
    float discriminant = LVO_SQR(B) - 4.0f * A * C;
    if (discriminant < 0.0f) return false;
    float discriminantSqrt = sqrtf(discriminant);
    float t0 = (-B - discriminantSqrt) / (2.0f * A);
    if (t0 >= 0.0f) {
        vec3 ip = rayOrigin + t0 * rayDirection;
        if (dot(tubeDirection, ip - tubeStart) > 0.0f && dot(tubeDirection, ip - tubeEnd) < 0.0f) {
            hitT = t0;
            return true;
        }
    }
    float t1 = (-B + discriminantSqrt) / (2.0f * A);
    if (t1 >= 0.0f) {
        vec3 ip = rayOrigin + t1 * rayDirection;
        if (dot(tubeDirection, ip - tubeStart) > 0.0f && dot(tubeDirection, ip - tubeEnd) < 0.0f) {
            hitT = t1;
            return true;
        }
    }
    return false;
}

// IntersectionTube main -- Data/Shaders/Renderers/RayTracing/TubeRayTracing.glsl:452-494.
// Returns true if reportIntersectionEXT would be called; the [tMin,tMax] acceptance is the caller's.
static inline bool intersectionTube(vec3 ro, vec3 rd, vec3 p0, vec3 p1, float lineRadius, bool cappedTubes,
                                    float& hitT, int& hitKind) {
    bool hasIntersection = false;
    hitT = 1e7f;
    hitKind = 0;
    float tubeT, sphere0T, sphere1T;
    if (rayTubeIntersection(ro, rd, p0, p1, lineRadius, tubeT)) {
        hitT = tubeT; hasIntersection = true; hitKind = 0;
    }
    if (cappedTubes) {
        bool h0 = raySphereIntersection(ro, rd, p0, lineRadius, sphere0T);
        bool h1 = raySphereIntersection(ro, rd, p1, lineRadius, sphere1T);
        if (h0 && sphere0T < hitT) { hasIntersection = true; hitT = sphere0T; hitKind = 1; }
        if (h1 && sphere1T < hitT) { hasIntersection = true; hitT = sphere1T; hitKind = 2; }
    }
    return hasIntersection;
}

// ----------------------------------------------------------------------------------------------
// Acceptance rule for a candidate segment.  The reference leaves this to the driver: an intersection shader runs for
// "rays that intersect the primitive's AABB" (conservatively, implementation-defined), and reportIntersectionEXT accepts
// hitT in [tMin, tMax].  Fixed here so that the accepted set is a pure function of (ray, segment, radius) and never of
// the acceleration structure: (1) the ray's [tmin, tmax] meets the segment's own AABB (src/LineData/LineDataFlow.cpp:
// 2230-2233) under the canonical slab test below, (2) IntersectionTube reports a hit, (3) hitT in [tmin, tmax].
// Canonical slab test: per plane t = fma(b, inv, c), one correctly rounded fused multiply-add, inv = 1/d with |d| clamped to
// 1e-30 and c = -(o * inv).  For a fixed ray t is monotone in b, so any box enclosing the segment's AABB passes whenever
// the segment's AABB does.
// ----------------------------------------------------------------------------------------------
static inline float safeInv(float d) {
    const float tiny = 1e-30f;
    if (fabsf(d) < tiny) d = std::signbit(d) ? -tiny : tiny;
    return 1.0f / d;
}
struct RayInv { float c[3], inv[3]; };
static inline RayInv makeRayInv(vec3 o, vec3 d) {
    RayInv r;
    r.inv[0] = safeInv(d.x); r.inv[1] = safeInv(d.y); r.inv[2] = safeInv(d.z);
    r.c[0] = -(o.x * r.inv[0]); r.c[1] = -(o.y * r.inv[1]); r.c[2] = -(o.z * r.inv[2]);
    return r;
}
static inline bool slabTest(const RayInv& r, const float* bmin, const float* bmax, float tmin, float tmax, float& tnear) {
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        float a = std::fmaf(bmin[k], r.inv[k], r.c[k]), b = std::fmaf(bmax[k], r.inv[k], r.c[k]);
        lo[k] = std::fmin(a, b); hi[k] = std::fmax(a, b);
    }
    float l = std::fmax(std::fmax(lo[0], lo[1]), std::fmax(lo[2], tmin));
    float h = std::fmin(std::fmin(hi[0], hi[1]), std::fmin(hi[2], tmax));
    tnear = l;
    return l <= h;
}
static inline bool acceptCandidate(vec3 ro, vec3 rd, const RayInv& ri, vec3 p0, vec3 p1, float radius, bool capped,
                                   float tmin, float tmax, float& t, int& kind) {
    const float bmin[3] = {std::fmin(p0.x, p1.x) - radius, std::fmin(p0.y, p1.y) - radius, std::fmin(p0.z, p1.z) - radius};
    const float bmax[3] = {std::fmax(p0.x, p1.x) + radius, std::fmax(p0.y, p1.y) + radius, std::fmax(p0.z, p1.z) + radius};
    float tn;
    if (!slabTest(ri, bmin, bmax, tmin, tmax, tn)) return false;
    return intersectionTube(ro, rd, p0, p1, radius, capped, t, kind) && t >= tmin && t <= tmax;
}

// ----------------------------------------------------------------------------------------------
// uniforms
// ----------------------------------------------------------------------------------------------
struct Uniforms {
    // LineUniformData (src/LineData/LineData.hpp:428-464)
    vec3 cameraPosition; float fieldOfViewY;
    float viewMatrix[16], projectionMatrix[16], inverseViewMatrix[16], inverseProjectionMatrix[16];
    vec4 backgroundColor, foregroundColor;
    float lineWidth;
    float ambientOcclusionStrength, ambientOcclusionGamma;
    float depthCueStrength, minDepth, maxDepth;   // USE_DEPTH_CUES; DepthMinMaxBuffer (Utils/Lighting.glsl:28-33)
    bool useDepthCues;
    uint32_t viewportW, viewportH;
    // transfer function (sgl TransferFunctionWindow; LUT by specification)
    const float* tfLut; uint32_t tfK; float minAttributeValue, maxAttributeValue;
    // shader defines
    bool useCappedTubes, useHalos, useAmbientOcclusion;
    // AO texture (result of the RTAO pass), W*H floats
    const float* aoTexture;
    // STATIC_AMBIENT_OCCLUSION_PREBAKING (Utils/AmbientOcclusion.glsl:30-37; uniforms LineData.hpp:452-458)
    bool staticAmbientOcclusionPrebaking;
    const float* ambientOcclusionFactors;            // [numParametrizationVertices * numAoTubeSubdivisions]
    const float* ambientOcclusionBlendingWeights;    // [numLineVertices]
    uint32_t numAoTubeSubdivisions, numLineVertices, numParametrizationVertices;
};

// Utils/TransferFunction.glsl:66-71 -- texture(sampler1D, posFloat), linear filter, clamp to edge
static inline vec4 transferFunction(const Uniforms& u, float attr) {
    float posFloat = clamp((attr - u.minAttributeValue) / (u.maxAttributeValue - u.minAttributeValue), 0.0f, 1.0f);
    float x = posFloat * float(u.tfK) - 0.5f;
    float fl = floorf(x);
    float w = x - fl;
    int i0 = int(fl), i1 = i0 + 1;
    int kmax = int(u.tfK) - 1;
    if (i0 < 0) i0 = 0; if (i0 > kmax) i0 = kmax;
    if (i1 < 0) i1 = 0; if (i1 > kmax) i1 = kmax;
    const float* a = u.tfLut + 4 * i0;
    const float* b = u.tfLut + 4 * i1;
    return vec4{mix(a[0], b[0], w), mix(a[1], b[1], w), mix(a[2], b[2], w), mix(a[3], b[3], w)};
}

// bilinear texture(sampler2D, uv).x with clamp-to-edge on a W*H float image
static inline float textureBilinear(const float* img, uint32_t W, uint32_t H, float uu, float vv) {
    float x = uu * float(W) - 0.5f, y = vv * float(H) - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float wx = x - fx, wy = y - fy;
    int x0 = int(fx), y0 = int(fy), x1 = x0 + 1, y1 = y0 + 1;
    int xm = int(W) - 1, ym = int(H) - 1;
    if (x0 < 0) x0 = 0; if (x0 > xm) x0 = xm; if (x1 < 0) x1 = 0; if (x1 > xm) x1 = xm;
    if (y0 < 0) y0 = 0; if (y0 > ym) y0 = ym; if (y1 < 0) y1 = 0; if (y1 > ym) y1 = ym;
    float a = mix(img[size_t(y0) * W + x0], img[size_t(y0) * W + x1], wx);
    float b = mix(img[size_t(y1) * W + x0], img[size_t(y1) * W + x1], wx);
    return mix(a, b, wy);
}

// Utils/AmbientOcclusion.glsl:84-99 (screen-space RTAO texture variant)
static inline float getAoFactor(const Uniforms& u, vec3 screenSpacePosition) {
    vec4 ndc = mul(u.projectionMatrix, vec4{screenSpacePosition.x, screenSpacePosition.y, screenSpacePosition.z, 1.0f});
    ndc.x = ndc.x / ndc.w; ndc.y = ndc.y / ndc.w; ndc.z = ndc.z / ndc.w;
    float aoFactor = textureBilinear(u.aoTexture, u.viewportW, u.viewportH, ndc.x * 0.5f + 0.5f, ndc.y * 0.5f + 0.5f);
    aoFactor = det_pow(aoFactor, u.ambientOcclusionGamma);
    return fmax_(0.0f, 1.0f - u.ambientOcclusionStrength + u.ambientOcclusionStrength * aoFactor);
}

// Utils/AmbientOcclusion.glsl:49-75 (STATIC_AMBIENT_OCCLUSION_PREBAKING variant)
static inline float getAoFactorStatic(const Uniforms& u, float interpolatedVertexId, float phi) {
    uint32_t lastLinePointIdx = uint32_t(interpolatedVertexId);
    uint32_t nextLinePointIdx = lastLinePointIdx + 1u < u.numLineVertices - 1u ? lastLinePointIdx + 1u : u.numLineVertices - 1u;
    float interpolationFactor = interpolatedVertexId - floorf(interpolatedVertexId);                    // fract
    float blendingWeightLast = u.ambientOcclusionBlendingWeights[lastLinePointIdx];
    float blendingWeightNext = u.ambientOcclusionBlendingWeights[nextLinePointIdx];
    float blendingWeight = mix(blendingWeightLast, blendingWeightNext, interpolationFactor);
    uint32_t lastVertexIdx = uint32_t(blendingWeight);
    uint32_t nextVertexIdx = lastVertexIdx + 1u < u.numParametrizationVertices - 1u ? lastVertexIdx + 1u : u.numParametrizationVertices - 1u;
    float interpolationFactorLine = blendingWeight - floorf(blendingWeight);
    const uint32_t N = u.numAoTubeSubdivisions;
    float circleIdxFlt = clamp(phi / 6.28318531f * float(N), 0.0f, float(N));                            // phi / (2.0 * M_PI) * N
    uint32_t circleIdxLast = (uint32_t(floorf(circleIdxFlt)) + N) % N;
    uint32_t circleIdxNext = (circleIdxLast + 1u) % N;
    float interpolationFactorCircle = circleIdxFlt - floorf(circleIdxFlt);
    float aoFactor00 = u.ambientOcclusionFactors[circleIdxLast + N * lastVertexIdx];
    float aoFactor01 = u.ambientOcclusionFactors[circleIdxLast + N * nextVertexIdx];
    float aoFactor10 = u.ambientOcclusionFactors[circleIdxNext + N * lastVertexIdx];
    float aoFactor11 = u.ambientOcclusionFactors[circleIdxNext + N * nextVertexIdx];
    float aoFactor0 = mix(aoFactor00, aoFactor01, interpolationFactorLine);
    float aoFactor1 = mix(aoFactor10, aoFactor11, interpolationFactorLine);
    float aoFactor = mix(aoFactor0, aoFactor1, interpolationFactorCircle);
    aoFactor = det_pow(aoFactor, u.ambientOcclusionGamma);
    return fmax_(0.0f, 1.0f - u.ambientOcclusionStrength + u.ambientOcclusionStrength * aoFactor);
}

// Utils/Antialiasing.glsl:1-3
static inline float getAntialiasingFactor(const Uniforms& u, float distance) {
    return distance / float(u.viewportH) * u.fieldOfViewY;
}

// Utils/Lighting.glsl:100-191 (no bands, no depth cues; AO branch when USE_AMBIENT_OCCLUSION && GEOMETRY_PASS_TUBE)
static inline vec4 blinnPhongShadingTube(const Uniforms& u, vec4 baseColor, vec3 fragmentPositionWorld,
                                         vec3 screenSpacePosition, float fragmentVertexId, float phi,
                                         vec3 fragmentNormal, vec3 fragmentTangent) {
    const vec3 ambientColor = V3(baseColor.x, baseColor.y, baseColor.z);
    const vec3 diffuseColor = ambientColor;
    float ambientOcclusionFactor = 1.0f;
    float kA, kD;
    const float kS = 0.3f, s = 30.0f;
    if (u.useAmbientOcclusion) {
        ambientOcclusionFactor = u.staticAmbientOcclusionPrebaking ? getAoFactorStatic(u, fragmentVertexId, phi)   // :118-125
                                                                   : getAoFactor(u, screenSpacePosition);
        kA = 0.2f + (1.0f - ambientOcclusionFactor) * 0.5f;
        kD = 0.9f * ambientOcclusionFactor;
    } else {
        kA = 0.1f;
        kD = 0.9f;
    }
    const vec3 Ia = kA * ambientColor;
    const vec3 n = normalize(fragmentNormal);
    const vec3 t = normalize(fragmentTangent);
    const vec3 v = normalize(u.cameraPosition - fragmentPositionWorld);
    const vec3 l = v;
    const vec3 h = normalize(v + l);
    vec3 helperVec = normalize(cross(t, l));
    vec3 newL = normalize(cross(helperVec, t));
    const float exponent = 1.7f;
    float cosNormal1 = det_pow(clamp(fabsf(dot(n, l)), 0.0f, 1.0f), exponent);
    float cosNormal2 = det_pow(clamp(fabsf(dot(n, newL)), 0.0f, 1.0f), exponent);
    float cosNormalCombined = 0.3f * cosNormal1 + 0.7f * cosNormal2;
    vec3 Id = (kD * cosNormalCombined) * diffuseColor;
    float isv = kS * det_pow(clamp(fabsf(dot(n, h)), 0.0f, 1.0f), s);
    vec3 phongColor = (Ia + Id) + V3(isv, isv, isv);
    if (u.useAmbientOcclusion) phongColor = phongColor * ambientOcclusionFactor;
    if (u.useDepthCues) {                                                              // Utils/Lighting.glsl:183-187
        float depthCueFactor = clamp((-screenSpacePosition.z - u.minDepth) / (u.maxDepth - u.minDepth), 0.0f, 1.0f);
        depthCueFactor = depthCueFactor * depthCueFactor * u.depthCueStrength;
        phongColor = V3(mix(phongColor.x, 0.5f, depthCueFactor), mix(phongColor.y, 0.5f, depthCueFactor), mix(phongColor.z, 0.5f, depthCueFactor));
    }
    return vec4{phongColor.x, phongColor.y, phongColor.z, baseColor.w};
}

struct HitColor { vec4 hitColor; float hitT; bool hasHit; };

// computeFragmentColor -- Data/Shaders/Renderers/RayTracing/RayHitCommon.glsl:74-543
// (variant: USE_CAPPED_TUBES, USE_HALOS, ANALYTIC_TUBE_INTERSECTIONS; no bands/multivar/stress/MLAT)
static inline HitColor computeFragmentColor(const Uniforms& u, vec3 fragmentPositionWorld, vec3 fragmentNormal,
                                            vec3 fragmentTangent, bool isCap, float phi, float fragmentVertexId,
                                            float fragmentAttribute) {
    vec4 fragmentColor = transferFunction(u, fragmentAttribute);                       // :127
    const vec3 n = normalize(fragmentNormal);                                          // :141
    const vec3 v = normalize(u.cameraPosition - fragmentPositionWorld);                // :142
    const vec3 t = normalize(fragmentTangent);                                         // :144
    vec3 helperVec = normalize(cross(t, v));                                           // :146
    vec3 newV = normalize(cross(helperVec, t));                                        // :147
    float ribbonPosition = 0.0f;
    if (u.useHalos) {
        if (u.useCappedTubes && isCap) {                                               // :193-229
            vec3 crossProdVn = cross(v, n);
            ribbonPosition = length(crossProdVn);
            vec3 crossProdVn2 = cross(newV, n);
            float ribbonPosition2 = length(crossProdVn2);
            if (dot(t, crossProdVn) < 0.0f) ribbonPosition2 = -ribbonPosition2;
            if (dot(t, crossProdVn) < 0.0f) ribbonPosition = -ribbonPosition;
            ribbonPosition2 = clamp(ribbonPosition2, -1.0f, 1.0f);
            if (fabsf(ribbonPosition2) < fabsf(ribbonPosition)) ribbonPosition = ribbonPosition2;
        } else {                                                                       // :353-372
            vec3 crossProdVn = cross(newV, n);
            ribbonPosition = length(crossProdVn);
            if (dot(t, crossProdVn) < 0.0f) ribbonPosition = -ribbonPosition;
            ribbonPosition = clamp(ribbonPosition, -1.0f, 1.0f);
        }
    }
    vec3 screenSpacePosition = V3(0, 0, 0);
    if ((u.useAmbientOcclusion && !u.staticAmbientOcclusionPrebaking) || u.useDepthCues) {   // :389-391
        vec4 sp = mul(u.viewMatrix, vec4{fragmentPositionWorld.x, fragmentPositionWorld.y, fragmentPositionWorld.z, 1.0f});
        screenSpacePosition = V3(sp.x, sp.y, sp.z);
    }
    fragmentColor = blinnPhongShadingTube(u, fragmentColor, fragmentPositionWorld, screenSpacePosition, fragmentVertexId, phi, n, t); // :415-426
    float absCoords = u.useHalos ? fabsf(ribbonPosition) : 0.0f;                       // :437-441
    float fragmentDepth = length(fragmentPositionWorld - u.cameraPosition);            // :443
    float EPSILON_OUTLINE = clamp(getAntialiasingFactor(u, fragmentDepth / u.lineWidth * 0.05f), 0.0f, 0.49f); // :451
    float EPSILON_WHITE = clamp(getAntialiasingFactor(u, fragmentDepth / u.lineWidth * 2.0f), 0.0f, 0.49f);    // :452
    const float WHITE_THRESHOLD = 0.7f;                                                // :488
    float coverage = u.useHalos ? 1.0f - smoothstep(1.0f - EPSILON_OUTLINE, 1.0f, absCoords) : 1.0f;          // :491-495
    float wmix = smoothstep(WHITE_THRESHOLD - EPSILON_WHITE, WHITE_THRESHOLD + EPSILON_WHITE, absCoords);
    HitColor out;
    out.hitColor = vec4{mix(fragmentColor.x, u.foregroundColor.x, wmix), mix(fragmentColor.y, u.foregroundColor.y, wmix),
                        mix(fragmentColor.z, u.foregroundColor.z, wmix), fragmentColor.w * coverage};          // :503-506
    out.hitT = length(fragmentPositionWorld - u.cameraPosition);                       // :540
    out.hasHit = true;
    return out;
}

// Per-segment line-point data the closest-hit shader reads besides position / attribute when USE_AMBIENT_OCCLUSION is on
// (linePointIndices :513, lineNormal :551): used by the static prebaked AO lookup only.
struct SegmentLineData { uint32_t idx0, idx1; vec3 n0, n1; };

// ClosestHitTubeAnalytic main -- Data/Shaders/Renderers/RayTracing/TubeRayTracing.glsl:512-613
static inline HitColor closestHitTubeAnalytic(const Uniforms& u, vec3 ro, vec3 rd, float hitT, int hitKind,
                                              vec3 p0, float a0, vec3 p1, float a1, const SegmentLineData* ld = nullptr) {
    vec3 fragmentPositionWorld = ro + rd * hitT;                                       // :517
    vec3 linePointInterpolated;
    float fragmentAttribute;
    float t;
    vec3 v = p1 - p0;                                                                  // :523
    if (hitKind == 0) {
        vec3 uu = fragmentPositionWorld - p0;
        t = dot(v, uu) / dot(v, v);
        linePointInterpolated = p0 + t * v;
        fragmentAttribute = (1.0f - t) * a0 + t * a1;
    } else if (hitKind == 1) {
        linePointInterpolated = p0; fragmentAttribute = a0; t = 0.0f;
    } else {
        linePointInterpolated = p1; fragmentAttribute = a1; t = 1.0f;
    }
    vec3 fragmentTangent = normalize(v);                                               // :544
    vec3 fragmentNormal = normalize(fragmentPositionWorld - linePointInterpolated);    // :545
    bool isCap = hitKind != 0;                                                         // :548
    float phi = 0.0f, fragmentVertexId = 0.0f;
    if (u.useAmbientOcclusion && u.staticAmbientOcclusionPrebaking && ld) {            // :550-562
        vec3 lineNormal = (1.0f - t) * ld->n0 + t * ld->n1;
        phi = det_acos(dot(fragmentNormal, lineNormal));
        float val = dot(lineNormal, cross(fragmentNormal, fragmentTangent));
        if (val < 0.0f) phi = 6.28318531f - phi;                                       // 2.0 * float(M_PI) - phi
        fragmentVertexId = (1.0f - t) * float(ld->idx0) + t * float(ld->idx1);
    }
    return computeFragmentColor(u, fragmentPositionWorld, fragmentNormal, fragmentTangent, isCap, phi, fragmentVertexId, fragmentAttribute);
}

// Depth range of one line vertex -- Data/Shaders/DepthCues/ComputeDepthValues.glsl:60-76 (min / max are folded by the caller,
// starting from (farDist, nearDist); the tree reduction of the shader is order independent)
static inline void depthRangeOfVertex(const float* viewMatrix, const float* projectionMatrix, float nearDist, float farDist,
                                      vec3 p, float& dmin, float& dmax) {
    vec4 sp = mul(viewMatrix, vec4{p.x, p.y, p.z, 1.0f});
    vec4 ndc = mul(projectionMatrix, sp);
    float nx = ndc.x / ndc.w, ny = ndc.y / ndc.w, nz = ndc.z / ndc.w;
    if (nx >= -1.0f && ny >= -1.0f && nz >= -1.0f && nx <= 1.0f && ny <= 1.0f && nz <= 1.0f) {
        float depth = clamp(-sp.z, nearDist, farDist);
        dmin = fmin_(dmin, depth - 1e-2f);
        dmax = fmax_(dmax, depth + 1e-2f);
    }
}

// Miss main -- TubeRayTracing.glsl:290-298
static inline HitColor missShader(const Uniforms& u) {
    return HitColor{u.backgroundColor, 0.0f, false};
}

// RayGen camera ray -- TubeRayTracing.glsl:202,219-226 (also VulkanRayTracedAmbientOcclusion.glsl:184-197)
static inline void cameraRay(const Uniforms& u, uint32_t px, uint32_t py, float xix, float xiy, vec3& ro, vec3& rd) {
    vec4 o = mul(u.inverseViewMatrix, vec4{0.0f, 0.0f, 0.0f, 1.0f});
    ro = V3(o.x, o.y, o.z);
    float ndcx = 2.0f * ((float(px) + xix) / float(u.viewportW)) - 1.0f;
    float ndcy = 2.0f * ((float(py) + xiy) / float(u.viewportH)) - 1.0f;
    vec4 tg = mul(u.inverseProjectionMatrix, vec4{ndcx, ndcy, 1.0f, 1.0f});
    vec3 nt = normalize(V3(tg.x, tg.y, tg.z));
    vec4 d = mul(u.inverseViewMatrix, vec4{nt.x, nt.y, nt.z, 0.0f});
    rd = V3(d.x, d.y, d.z);
}

// sampleHemisphere -- Data/Shaders/AO/RTAO/VulkanRayTracedAmbientOcclusion.glsl:151-156
static inline vec3 sampleHemisphere(float xix, float xiy) {
    float c, s;
    det_sincos2pi(xiy, c, s);
    float r = sqrtf(1.0f - xix * xix);
    return V3(c * r, s * r, xix);
}

// ----------------------------------------------------------------------------------------------
// PPLL -- Data/Shaders/Renderers/PPLL/*.glsl, Data/Shaders/Utils/TiledAddress.glsl
// ----------------------------------------------------------------------------------------------
// packUnorm4x8 (GLSL spec): round(clamp(c, 0, 1) * 255.0), x in the least significant byte
static inline uint32_t packUnorm4x8(vec4 c) {
    uint32_t r = uint32_t(floorf(clamp(c.x, 0.0f, 1.0f) * 255.0f + 0.5f));
    uint32_t g = uint32_t(floorf(clamp(c.y, 0.0f, 1.0f) * 255.0f + 0.5f));
    uint32_t b = uint32_t(floorf(clamp(c.z, 0.0f, 1.0f) * 255.0f + 0.5f));
    uint32_t a = uint32_t(floorf(clamp(c.w, 0.0f, 1.0f) * 255.0f + 0.5f));
    return r | (g << 8) | (b << 16) | (a << 24);
}
static inline vec4 unpackUnorm4x8(uint32_t p) {
    return vec4{float(p & 0xffu) / 255.0f, float((p >> 8) & 0xffu) / 255.0f,
                float((p >> 16) & 0xffu) / 255.0f, float((p >> 24) & 0xffu) / 255.0f};
}
// addrGen, ADDRESSING_TILED_NxM / linear -- Utils/TiledAddress.glsl:53-85 (default 2x8, LineRenderer.cpp:739-740)
static inline uint32_t addrGen(uint32_t x, uint32_t y, uint32_t viewportW, uint32_t tileN, uint32_t tileM) {
    if (tileN == 1 && tileM == 1) return x + viewportW * y;
    uint32_t surfaceWidth = viewportW / tileN;
    uint32_t tx = x / tileN, ty = y / tileM;
    uint32_t tileAddr1D = (tx + surfaceWidth * ty) * (tileN * tileM);
    uint32_t px = x & (tileN - 1), py = y & (tileM - 1);
    return tileAddr1D | (px + py * tileN);
}

}  // namespace lvo
#endif
