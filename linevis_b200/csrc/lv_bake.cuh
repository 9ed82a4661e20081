// lv_bake.cuh -- object-space RTAO prebaker (reference "RTAO (Prebaker)": src/Renderers/AmbientOcclusion/VulkanAmbientOcclusionBaker.cpp,
// compute shader Data/Shaders/AO/RTAO/VulkanAmbientOcclusionBaker.glsl:190-282).
//
// The reference's compute shader gives one thread a parametrization vertex and lets it loop over the tube subdivisions and
// the samples (numTubeSubdivisions x numAmbientOcclusionSamples dependent ray queries per thread).  Here the work is cut the
// other way round: k_bake_setup writes one 48-byte start frame per (vertex, subdivision) -- the same AoHit record the
// screen-space RTAO pass keeps per hit pixel -- and the persistent ray-stream kernel k_rtao_rays (lv_kernels.cuh) traces all
// n_param x n_subdiv x spp rays with its lane refill / leaf vote machinery; k_rtao_reduce folds the samples of a record in
// sample order and applies the running mean over iterations.  The only baker-specific piece inside k_rtao_rays is the
// random stream: all rays of a vertex draw from ONE LCG stream seeded with tea(vertex, frame) (:196,:267), so record
// (vertex, s) starts at stream position 2 s spp and its k-th ray at 2 (s spp + k) -- reached with lcg_skip, no replay.
#pragma once
#include "lv_math.cuh"
#include "lv_types.cuh"

namespace lv {

struct BakeParams {
    const float4* pt_pos;      // [n_line_pts] linePosition   (LinePointDataUnified, reference LineRenderData.hpp:99-106)
    const float4* pt_tan;      // [n_line_pts] lineTangent
    const float4* pt_nrm;      // [n_line_pts] lineNormal
    const float* sampling;     // [n_param] samplingLocations (VulkanAmbientOcclusionBaker.cpp:594-612)
    uint32_t n_line_pts, n_param, n_subdiv, spp, frame_number;
    float line_radius;
    uint32_t first_vertex, n_vertices;   // the slice of parametrization vertices this context bakes (multi-GPU: lv_ao_set_vertex_range)
};

LV_DEV Vec3 mix3_(Vec3 a, Vec3 b, float t) { return v3(mixf_(a.x, b.x, t), mixf_(a.y, b.y, t), mixf_(a.z, b.z, t)); }
LV_DEV Vec3 xyz_(float4 v) { return v3(v.x, v.y, v.z); }

// getInterpolatedLinePoint (:104-124) + the per-subdivision frame of main (:238-263) for one (vertex, subdivision) pair
LV_DEV AoHit bake_record(const BakeParams& B, uint32_t vertex, uint32_t sub) {
    const float loc = __ldg(B.sampling + vertex);
    const uint32_t lower = uint32_t(loc);
    const uint32_t upper = lower + 1u < B.n_line_pts - 1u ? lower + 1u : B.n_line_pts - 1u;
    const float f = loc - floorf(loc);
    const Vec3 tl = xyz_(__ldg(B.pt_tan + lower)), tu = xyz_(__ldg(B.pt_tan + upper));
    const Vec3 nl = xyz_(__ldg(B.pt_nrm + lower)), nu = xyz_(__ldg(B.pt_nrm + upper));
    const Vec3 position = mix3_(xyz_(__ldg(B.pt_pos + lower)), xyz_(__ldg(B.pt_pos + upper)), f);
    const Vec3 tangent = normalize3(mix3_(tl, tu, f));
    const Vec3 normal = normalize3(mix3_(nl, nu, f));
    const Vec3 binormal = normalize3(mix3_(cross3(tl, nl), cross3(tu, nu), f));
    float ca, sa;
    det_sincos2pi(float(sub) / float(B.n_subdiv), ca, sa);            // angle = sub / N * 2 pi  (:239-241)
    const Vec3 surface_n = ca * normal + sa * binormal;               // :258
    const Vec3 origin = position + (B.line_radius + 1e-6f) * surface_n;   // :259
    const uint32_t seed = lcg_skip(tea(vertex, B.frame_number), 2u * sub * B.spp);
    AoHit r;
    r.pos_off = make_float4(origin.x, origin.y, origin.z, 0.0f);
    r.nrm_px = make_float4(surface_n.x, surface_n.y, surface_n.z, __uint_as_float(sub + B.n_subdiv * vertex));
    r.tng = make_float4(tangent.x, tangent.y, tangent.z, __uint_as_float(seed));
    return r;
}

// AO ray number `sample` of a start frame (sampleHemisphere + frame transform; screen-space RTAO:
// VulkanRayTracedAmbientOcclusion.glsl:151-156,257,289-299; baker: VulkanAmbientOcclusionBaker.glsl:159-164,262-268).
template <bool BAKE>
LV_DEV void ao_ray_from_record(const AoHit* rec, uint32_t sample, uint32_t spp, uint32_t frame_number, Vec3& org, Vec3& dir) {
    const float4* hp = reinterpret_cast<const float4*>(rec);
    const float4 a = __ldg(hp), b = __ldg(hp + 1), c = __ldg(hp + 2);
    const Vec3 pos = v3(a.x, a.y, a.z), nrm = v3(b.x, b.y, b.z), tng = v3(c.x, c.y, c.z);
    const Vec3 btg = cross3(nrm, tng);                               // :257
    uint32_t seed;
    if (BAKE) seed = lcg_skip(__float_as_uint(c.w), 2u * sample);    // one stream per vertex (baker :196,267)
    else seed = tea(__float_as_uint(b.w), frame_number * spp + sample);   // :289-292
    const float xa = rnd(seed), xb = rnd(seed);
    float cs, sn;
    det_sincos2pi(xb, cs, sn);
    const float rr = sqrtf(1.0f - xa * xa);
    const Vec3 hs = v3(cs * rr, sn * rr, xa);                        // sampleHemisphere
    dir = normalize3((tng * hs.x + btg * hs.y) + nrm * hs.z);
    org = BAKE ? pos : pos + dir * a.w;                              // :299 (baker: the origin as set up by bake_record)
}

// SegAux of one record: the caller's point indices of its segment and the line normals there
LV_DEV SegAux make_seg_aux(uint2 idx, const float4* pt_nrm) {
    const float4 a = __ldg(pt_nrm + idx.x), b = __ldg(pt_nrm + idx.y);
    SegAux x;
    x.n0 = make_float4(a.x, a.y, a.z, __uint_as_float(idx.x));
    x.n1 = make_float4(b.x, b.y, b.z, __uint_as_float(idx.y));
    return x;
}

#if !defined(LV_HOST_EMU) || defined(LV_HOST_EMU_SIMT)
__global__ void k_bake_setup(const __grid_constant__ BakeParams B, AoHit* records) {
    const uint32_t total = B.n_vertices * B.n_subdiv;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        records[i] = bake_record(B, B.first_vertex + i / B.n_subdiv, i % B.n_subdiv);
}

__global__ void k_seg_aux(const uint32_t* prim_ids, const uint2* seg_idx, const float4* pt_nrm, uint32_t n_seg, SegAux* aux) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_seg; i += gridDim.x * blockDim.x)
        aux[i] = make_seg_aux(seg_idx[prim_ids[i]], pt_nrm);
}

// xyz triples -> float4 (w = 0), for the three per-point arrays
__global__ void k_expand_xyz(const float* xyz, uint32_t n, float4* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0f);
}
#endif

}  // namespace lv
