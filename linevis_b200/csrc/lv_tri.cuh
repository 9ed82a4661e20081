// lv_tri.cuh -- triangle-tube mode of the RTAO passes (b200_rtao_geometry = triangles): the reference traces its AO passes against
// the TRIANGULATED tubes (N-gon rings + hemisphere caps, lv_tubemesh.hpp), not against the analytic capsules; this mode does the
// same, so that an AO image can match the reference's geometry and not only its own analytic stand-in (DESIGN.md rule 5).
//
//   - TriRec: 48-byte triangle record (three positions, the vertex indices in the w lanes), Morton order like SegRec;
//   - LBVH build over the triangles' AABBs: k_tri_bounds / k_tri_morton / k_pack_tris / k_tri_fit + the topology kernels of
//     lv_bvh.cuh (k_radix_tree, k_emit_nodes, k_tree_depth), one triangle per leaf;
//   - acceptance rule, BVH independent like rule 2: the ray's [tMin, tMax] meets the triangle's own AABB under the canonical slab
//     test AND the double-sided Moeller-Trumbore test reports t in [tMin, tMax]; closest hit = smallest t, ties -> lowest
//     triangle index (the hardware's test is only specified as watertight; fixed by specification, mirrored by the oracle);
//   - k_rtao_primary_tri: camera ray, warp-packet closest hit, the shader's barycentric fetch
//     (reference Data/Shaders/AO/RTAO/VulkanRayTracedAmbientOcclusion.glsl:205-276) -> the same 48-byte AO start frame the
//     capsule path writes, so the AO ray stream (k_rtao_rays_q<.., PRIM = 1>) and k_rtao_reduce run unchanged on top.
#pragma once
#include "lv_trace.cuh"

namespace lv {

LV_DEV void tri_positions(const TriRec& r, Vec3& p0, Vec3& p1, Vec3& p2) {
    p0 = v3(r.a.x, r.a.y, r.a.z); p1 = v3(r.b.x, r.b.y, r.b.z); p2 = v3(r.c.x, r.c.y, r.c.z);
}

// double-sided Moeller-Trumbore; (u, v) = barycentric weights of vertex 1 and 2 (rayQueryGetIntersectionBarycentricsEXT)
LV_DEV bool tri_hit(Vec3 o, Vec3 d, const TriRec& r, float tmin, float tmax, float& t, float& u, float& v) {
    Vec3 p0, p1, p2;
    tri_positions(r, p0, p1, p2);
    const Vec3 e1 = p1 - p0, e2 = p2 - p0;
    const Vec3 pv = cross3(d, e2);
    const float det = dot3(e1, pv);
    if (det == 0.0f) return false;
    const float inv = 1.0f / det;
    const Vec3 tv = o - p0;
    u = dot3(tv, pv) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    const Vec3 qv = cross3(tv, e1);
    v = dot3(d, qv) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot3(e2, qv) * inv;
    return t >= tmin && t <= tmax;
}

// the triangle's own AABB (min / max of its vertices) against the ray's ORIGINAL interval: first half of the acceptance rule
LV_DEV bool tri_box_hit(const RayBox& rb, const TriRec& r, float tmin, float tmax) {
    float tn;
    return box_hit(rb, fminf(fminf(r.a.x, r.b.x), r.c.x), fminf(fminf(r.a.y, r.b.y), r.c.y), fminf(fminf(r.a.z, r.b.z), r.c.z),
                   fmaxf(fmaxf(r.a.x, r.b.x), r.c.x), fmaxf(fmaxf(r.a.y, r.b.y), r.c.y), fmaxf(fmaxf(r.a.z, r.b.z), r.c.z), tmin, tmax, tn);
}

LV_DEV TriRec load_tri(const TriRec* p) {
    TriRec r;
    const float4* q = reinterpret_cast<const float4*>(p);
    r.a = __ldg(q); r.b = __ldg(q + 1); r.c = __ldg(q + 2);
    return r;
}

#if !defined(LV_HOST_EMU) || defined(LV_HOST_EMU_SIMT)
// ---- LBVH over triangles: the record-specific kernels (the topology kernels are shared with the segment BVH)
__device__ __forceinline__ void tri_box(const float* vpos, const uint32_t* idx, uint32_t i, float* mn, float* mx) {
    const uint32_t a = idx[3 * size_t(i)], b = idx[3 * size_t(i) + 1], c = idx[3 * size_t(i) + 2];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float p = vpos[3 * size_t(a) + k], q = vpos[3 * size_t(b) + k], r = vpos[3 * size_t(c) + k];
        mn[k] = fminf(fminf(p, q), r); mx[k] = fmaxf(fmaxf(p, q), r);
    }
}

__global__ void k_tri_bounds(const float* vpos, const uint32_t* idx, uint32_t n, float* bounds) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float a[3], b[3];
        tri_box(vpos, idx, i, a, b);
#pragma unroll
        for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], a[k]); mx[k] = fmaxf(mx[k], b[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomic_min_f(bounds + k, mn[k]); atomic_max_f(bounds + 3 + k, mx[k]); }
    }
}

__global__ void k_tri_morton(const float* vpos, const uint32_t* idx, uint32_t n, const float* bounds, unsigned long long* keys, uint32_t* vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float mn[3], mx[3];
    tri_box(vpos, idx, i, mn, mx);
    unsigned long long code = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = bounds[k], ext = bounds[3 + k] - lo;
        const float c = 0.5f * (mn[k] + mx[k]);
        float u = ext > 0.0f ? (c - lo) / ext : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        const unsigned long long q = (unsigned long long)(fminf(u * 2097152.0f, 2097151.0f));
        code |= expand21(q) << (2 - k);
    }
    keys[i] = code;
    vals[i] = i;
}

__global__ void k_pack_tris(const float* vpos, const uint32_t* idx, const uint32_t* order, uint32_t n, TriRec* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = order[i];
    const uint32_t a = idx[3 * size_t(t)], b = idx[3 * size_t(t) + 1], c = idx[3 * size_t(t) + 2];
    TriRec r;
    r.a = make_float4(vpos[3 * size_t(a)], vpos[3 * size_t(a) + 1], vpos[3 * size_t(a) + 2], __uint_as_float(a));
    r.b = make_float4(vpos[3 * size_t(b)], vpos[3 * size_t(b) + 1], vpos[3 * size_t(b) + 2], __uint_as_float(b));
    r.c = make_float4(vpos[3 * size_t(c)], vpos[3 * size_t(c) + 1], vpos[3 * size_t(c) + 2], __uint_as_float(c));
    out[i] = r;
}

// bottom-up fit, one thread per triangle record (same protocol as k_fit: the second arrival at an inner node continues upward)
__global__ void k_tri_fit(const TriRec* tris, int n, const int2* children, const int* parent, float* boxes, unsigned int* flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TriRec s = tris[i];
    const float b[6] = {fminf(fminf(s.a.x, s.b.x), s.c.x), fminf(fminf(s.a.y, s.b.y), s.c.y), fminf(fminf(s.a.z, s.b.z), s.c.z),
                        fmaxf(fmaxf(s.a.x, s.b.x), s.c.x), fmaxf(fmaxf(s.a.y, s.b.y), s.c.y), fmaxf(fmaxf(s.a.z, s.b.z), s.c.z)};
    float* mine = boxes + 6 * size_t(n - 1 + i);
#pragma unroll
    for (int k = 0; k < 6; k++) mine[k] = b[k];
    if (n == 1) return;
    int node = parent[n - 1 + i];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(flags + node, 1u) == 0u) return;
        __threadfence();
        const int2 ch = children[node];
        const volatile float* lb = boxes + 6 * size_t((ch.x < 0) ? (n - 1 + (ch.x & 0x7fffffff)) : ch.x);
        const volatile float* rb = boxes + 6 * size_t((ch.y < 0) ? (n - 1 + (ch.y & 0x7fffffff)) : ch.y);
        float* o = boxes + 6 * size_t(node);
#pragma unroll
        for (int k = 0; k < 3; k++) { o[k] = fminf(lb[k], rb[k]); o[3 + k] = fmaxf(lb[3 + k], rb[3 + k]); }
        node = parent[node];
    }
}

// per-vertex attributes of the mesh for the barycentric fetch: (normal.xyz, as_float(vertexLinePointIndex: line point index, bit 31 = cap vertex))
__global__ void k_tri_vertex_attr(const float* vnrm, const uint32_t* vline, uint32_t n, float4* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = make_float4(vnrm[3 * size_t(i)], vnrm[3 * size_t(i) + 1], vnrm[3 * size_t(i) + 2], __uint_as_float(vline[i]));
}

// lineAttribute of the mesh's line points (the w lane of tri_line_tan): per-point attributes are recovered from the segment records
// (record r holds the attributes of the two points seg_idx[prim_ids[r]] names), then gathered through the mesh's source-point indices
__global__ void k_point_attr(const SegRec* segs, const uint32_t* prim_ids, const uint2* seg_idx, uint32_t n_seg, float* pt_attr) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_seg; r += gridDim.x * blockDim.x) {
        const uint2 ix = seg_idx[prim_ids[r]];
        pt_attr[ix.x] = segs[r].a.w; pt_attr[ix.y] = segs[r].b.w;
    }
}
__global__ void k_tri_line_attr(const float* pt_attr, const uint32_t* line_src, uint32_t n, float4* line_tan) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) line_tan[i].w = pt_attr[line_src[i]];
}

struct TriHitRec { float t, u, v; uint32_t idx, prim; };   // idx = record (BVH order), prim = triangle index of the mesh

// closest hit of a warp packet of coherent rays against the triangle BVH (same scheme as bvh_trace_packet; one-record leaves)
__device__ __forceinline__ bool bvh_trace_packet_tri(const SceneDev& S, bool active, Vec3 o, Vec3 d, float tmin, float tmax, TriHitRec& best,
                                                     uint32_t* stack, uint32_t& steps, uint32_t& isect) {
    best.t = tmax; best.u = best.v = 0.0f; best.idx = 0; best.prim = 0xFFFFFFFFu;
    bool found = false;
    if (S.n_tri == 0 || __ballot_sync(0xffffffffu, active) == 0u) return false;
    const uint32_t lane = threadIdx.x & 31;
    const RayBox rb = make_raybox(o, d);
    uint32_t node = 0;
    int sp = 0;
    while (true) {
        const Node64 nd = load_node(S.tri_nodes + node);
        steps += (lane == 0);
        const float tcull = best.t + S.line_width;   // tie-safe margin, as for the capsules (DESIGN.md rule 3)
        float tl, tr;
        const bool hl = active && box_hit(rb, nd.l0, nd.l1, tmin, tcull, tl);
        const bool hr = active && box_hit(rb, nd.r0, nd.r1, tmin, tcull, tr);
        const unsigned ml = __ballot_sync(0xffffffffu, hl), mr = __ballot_sync(0xffffffffu, hr);
        const uint32_t cw[2] = {__float_as_uint(nd.l0.w), __float_as_uint(nd.r0.w)};
        const unsigned mk[2] = {ml, mr};
        uint32_t inner[2]; int n_inner = 0;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            if (!mk[side]) continue;
            const uint32_t w = cw[side];
            if (w & kLeafBit) {
                const uint32_t ref = w & kRefMask, cnt = ((w >> 27) & 15u) + 1u;
                isect += (lane == 0) ? cnt : 0u;
                const bool mine = side ? hr : hl;
                for (uint32_t i = 0; i < cnt; i++) {
                    const TriRec r = load_tri(S.tris + ref + i);
                    float t, u, v;
                    if (mine && tri_box_hit(rb, r, tmin, tmax) && tri_hit(o, d, r, tmin, tmax, t, u, v)) {
                        if (!found || t <= best.t) {
                            const uint32_t prim = __ldg(S.tri_ids + ref + i);
                            if (!found || t < best.t || prim < best.prim) { best.t = t; best.u = u; best.v = v; best.idx = ref + i; best.prim = prim; found = true; }
                        }
                    }
                }
            } else inner[n_inner++] = w;
        }
        if (n_inner == 2) {
            const unsigned right_near = __ballot_sync(0xffffffffu, hr && (!hl || tr < tl));
            const unsigned left_near = __ballot_sync(0xffffffffu, hl && (!hr || tl <= tr));
            const bool rf = __popc(right_near) > __popc(left_near);
            if (lane == 0) stack[sp] = rf ? inner[0] : inner[1];
            sp++;
            node = rf ? inner[1] : inner[0];
        } else if (n_inner == 1) node = inner[0];
        else {
            if (sp == 0) break;
            --sp;
            __syncwarp();
            node = stack[sp];
        }
        __syncwarp();
    }
    return found;
}
#endif  // warp-collective code

// the RTAO shader's barycentric vertex fetch for a triangle hit (VulkanRayTracedAmbientOcclusion.glsl:213-276) -> AO start frame
LV_DEV Vec3 interp3(Vec3 a, Vec3 b, Vec3 c, Vec3 w) { return (a * w.x + b * w.y) + c * w.z; }   // BarycentricInterpolation.glsl:39-41
LV_DEV Vec3 xyz4(float4 v) { return v3(v.x, v.y, v.z); }

LV_DEV AoHit tri_ao_frame(const SceneDev& S, const TriRec& r, float u, float v, float subdiv_corr, uint32_t pixel) {
    const Vec3 w = v3(1.0f - u - v, u, v);
    Vec3 p0, p1, p2;
    tri_positions(r, p0, p1, p2);
    const float4 a0 = __ldg(S.tri_vattr + __float_as_uint(r.a.w)), a1 = __ldg(S.tri_vattr + __float_as_uint(r.b.w)), a2 = __ldg(S.tri_vattr + __float_as_uint(r.c.w));
    const uint32_t l0 = __float_as_uint(a0.w) & 0x7FFFFFFFu, l1 = __float_as_uint(a1.w) & 0x7FFFFFFFu, l2 = __float_as_uint(a2.w) & 0x7FFFFFFFu;
    const Vec3 pos = interp3(p0, p1, p2, w);
    const Vec3 nrm = normalize3(interp3(xyz4(a0), xyz4(a1), xyz4(a2), w));
    const Vec3 line_pos = interp3(xyz4(__ldg(S.tri_line_pos + l0)), xyz4(__ldg(S.tri_line_pos + l1)), xyz4(__ldg(S.tri_line_pos + l2)), w);
    const Vec3 tng = normalize3(interp3(xyz4(__ldg(S.tri_line_tan + l0)), xyz4(__ldg(S.tri_line_tan + l1)), xyz4(__ldg(S.tri_line_tan + l2)), w));
    const float offset = length3(line_pos - pos) / subdiv_corr;
    AoHit rec;
    rec.pos_off = make_float4(pos.x, pos.y, pos.z, offset);
    rec.nrm_px = make_float4(nrm.x, nrm.y, nrm.z, __uint_as_float(pixel));
    rec.tng = make_float4(tng.x, tng.y, tng.z, 0.0f);
    return rec;
}

// ClosestHitTubeTriangles + LineAttributesBarycentric.glsl:1-39 (TubeRayTracing.glsl:301-351): the tube pass's hit shader in the
// triangle-mesh geometry mode -- barycentric position / normal / tangent / attribute, cap flag from the vertices, then the common
// computeFragmentColor (shade_surface).  The prebaked-AO lookup (phi / vertex id interpolation) is not part of this mode.
LV_DEV Shaded shade_tri_hit(const FrameParams& P, const SceneDev& S, const TriRec& r, float u, float v) {
    const Vec3 w = v3(1.0f - u - v, u, v);
    Vec3 p0, p1, p2;
    tri_positions(r, p0, p1, p2);
    const float4 a0 = __ldg(S.tri_vattr + __float_as_uint(r.a.w)), a1 = __ldg(S.tri_vattr + __float_as_uint(r.b.w)), a2 = __ldg(S.tri_vattr + __float_as_uint(r.c.w));
    const uint32_t v0 = __float_as_uint(a0.w), v1 = __float_as_uint(a1.w), v2 = __float_as_uint(a2.w);
    const bool is_cap = ((v0 | v1 | v2) >> 31) != 0u;
    const float4 t0 = __ldg(S.tri_line_tan + (v0 & 0x7FFFFFFFu)), t1 = __ldg(S.tri_line_tan + (v1 & 0x7FFFFFFFu)), t2v = __ldg(S.tri_line_tan + (v2 & 0x7FFFFFFFu));
    const Vec3 pos = interp3(p0, p1, p2, w);
    const Vec3 nrm0 = normalize3(interp3(xyz4(a0), xyz4(a1), xyz4(a2), w));
    const Vec3 tan0 = normalize3(interp3(xyz4(t0), xyz4(t1), xyz4(t2v), w));
    const float attr = (t0.w * w.x + t1.w * w.y) + t2v.w * w.z;          // interpolateFloat, BarycentricInterpolation.glsl:35-37
    const Vec3 tg = normalize3(tan0);
    return shade_surface<false>(P, pos, nrm0, tan0, tg, normalize3(tg), is_cap, attr, 0.0f, nullptr);
}

}  // namespace lv
