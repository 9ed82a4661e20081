// lv_trace.cuh -- per-thread BVH traversal over 64-byte child-pair nodes (closest / any / all hits).
//
// Replaces traceRayEXT / rayQueryEXT against the driver's acceleration structure
// (reference TubeRayTracing.glsl:56,68; VulkanRayTracedAmbientOcclusion.glsl:164-165,201-204).
// Result-defining rules (mirrored by the CPU oracle): a candidate is accepted iff its reported hitT lies
// in [tmin, tmax]; the closest hit is the smallest hitT, ties go to the lowest caller-side segment index.
//
// A candidate is accepted iff (1) the ray's [tmin, tmax] interval meets the segment's own AABB under the canonical slab
// test below, (2) IntersectionTube reports a hit, (3) the reported hitT lies in [tmin, tmax].
//
// Cost model: one traversal step = one 64-byte node fetch (4 x LDG.128), one intersection =
// one 32-byte segment record fetch (2 x LDG.128).  `steps` / `isect` count exactly those.
#pragma once
#include "lv_shade.cuh"

namespace lv {

constexpr int kStackSize = 72;
constexpr uint32_t kLeafBit = 0x80000000u;   // child word: leaf = kLeafBit | (count-1) << 27 | first record; inner = node index
constexpr uint32_t kRefMask = 0x07FFFFFFu;

// Canonical slab test (DESIGN.md "Result-defining rules"): per plane t = fma(b, inv, c) -- ONE correctly rounded fused
// multiply-add -- with inv = 1/d (|d| clamped to 1e-30) and c = -(o * inv); hit iff
// max(lo_x, lo_y, lo_z, tmin) <= min(hi_x, hi_y, hi_z, tmax).  For a fixed ray t is a monotone function of b, so a box
// that encloses another can never be missed when the inner one is hit: the set of accepted candidates does not depend on
// the BVH topology.
struct RayBox {
    float ix, iy, iz;     // safe 1/d
    float cx, cy, cz;     // -(o * inv)
};

__device__ __forceinline__ float safe_inv(float d) {
    const float tiny = 1e-30f;
    if (fabsf(d) < tiny) d = (__float_as_uint(d) >> 31) ? -tiny : tiny;
    return 1.0f / d;
}
__device__ __forceinline__ RayBox make_raybox(Vec3 o, Vec3 d) {
    RayBox b;
    b.ix = safe_inv(d.x); b.iy = safe_inv(d.y); b.iz = safe_inv(d.z);
    b.cx = -(o.x * b.ix); b.cy = -(o.y * b.iy); b.cz = -(o.z * b.iz);
    return b;
}
__device__ __forceinline__ bool box_hit(const RayBox& rb, float mnx, float mny, float mnz, float mxx, float mxy, float mxz,
                                        float tmin, float tmax, float& tn) {
    float ax = __fmaf_rn(mnx, rb.ix, rb.cx), bx = __fmaf_rn(mxx, rb.ix, rb.cx);
    float ay = __fmaf_rn(mny, rb.iy, rb.cy), by = __fmaf_rn(mxy, rb.iy, rb.cy);
    float az = __fmaf_rn(mnz, rb.iz, rb.cz), bz = __fmaf_rn(mxz, rb.iz, rb.cz);
    float lo = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    float hi = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    tn = lo;
    return lo <= hi;
}
__device__ __forceinline__ bool box_hit(const RayBox& rb, float4 mn, float4 mx, float tmin, float tmax, float& tn) {
    return box_hit(rb, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, tmin, tmax, tn);
}
// the segment's own AABB (min/max(p0,p1) -+ r, reference src/LineData/LineDataFlow.cpp:2230-2233) against the ray's
// ORIGINAL interval: part of the acceptance rule, and a cheap reject in front of the three quadratics
__device__ __forceinline__ bool seg_box_hit(const RayBox& rb, const SegRec& s, float r, float tmin, float tmax) {
    float tn;
    return box_hit(rb, fminf(s.a.x, s.b.x) - r, fminf(s.a.y, s.b.y) - r, fminf(s.a.z, s.b.z) - r,
                   fmaxf(s.a.x, s.b.x) + r, fmaxf(s.a.y, s.b.y) + r, fmaxf(s.a.z, s.b.z) + r, tmin, tmax, tn);
}

__device__ __forceinline__ Node64 load_node(const Node64* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    Node64 n; n.l0 = __ldg(q); n.l1 = __ldg(q + 1); n.r0 = __ldg(q + 2); n.r1 = __ldg(q + 3);
    return n;
}
__device__ __forceinline__ SegRec load_seg(const SegRec* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    SegRec s; s.a = __ldg(q); s.b = __ldg(q + 1);
    return s;
}

struct HitRec {
    float t;
    uint32_t idx;    // record index in BVH order
    uint32_t prim;   // caller-side segment index
    uint32_t kind;
};

// MODE 0: closest hit, MODE 1: any hit (terminate on first accepted candidate).
template <int MODE>
__device__ __forceinline__ bool bvh_trace(const SceneDev& S, Vec3 o, Vec3 d, float tmin, float tmax, bool capped,
                                          HitRec& best, uint32_t& steps, uint32_t& isect) {
    best.t = tmax; best.idx = 0; best.prim = 0xFFFFFFFFu; best.kind = 0;
    if (S.n_seg == 0) return false;
    const RayQ rq = make_rayq(o, d);
    const RayBox rb = make_raybox(o, d);
    uint32_t stack[kStackSize];
    int sp = 0;
    uint32_t node = 0;
    bool found = false;
    while (true) {
        const Node64 nd = load_node(S.nodes + node);
        steps++;
        float tl, tr;
        // Boxes are culled against best.t + one tube diameter, not best.t: the reference's float32 quadratic reports hitT
        // with an error of up to a few % of r, so a candidate that TIES the current best (adjacent capsules share an end
        // sphere) may have a box entry slightly beyond it; the tie rule must still see it, whatever the BVH looks like.
        const float tcull = MODE == 0 ? best.t + S.line_width : best.t;
        bool hl = box_hit(rb, nd.l0, nd.l1, tmin, tcull, tl);
        bool hr = box_hit(rb, nd.r0, nd.r1, tmin, tcull, tr);
        const uint32_t lw = __float_as_uint(nd.l0.w), rw = __float_as_uint(nd.r0.w);   // absent children have a box that never hits
        const uint32_t lref = lw & kRefMask, rref = rw & kRefMask;
        const uint32_t lcnt = (lw & kLeafBit) ? ((lw >> 27) & 15u) + 1u : 0u, rcnt = (rw & kLeafBit) ? ((rw >> 27) & 15u) + 1u : 0u;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            bool h = side ? hr : hl;
            uint32_t cnt = side ? rcnt : lcnt, ref = side ? rref : lref;
            if (h && cnt) {
                isect += cnt;
                for (uint32_t i = 0; i < cnt; i++) {
                    SegRec s = load_seg(S.segs + ref + i);
                    float t; uint32_t kind;
                    if (seg_box_hit(rb, s, S.radius, tmin, tmax) && capsule_hit(rq, s, S.radius, capped, t, kind) && t >= tmin && t <= tmax) {
                        if (MODE == 1) { best.t = t; best.idx = ref + i; best.kind = kind; return true; }
                        if (!found || t <= best.t) {
                            uint32_t prim = __ldg(S.prim_ids + ref + i);
                            if (!found || t < best.t || prim < best.prim) {
                                best.t = t; best.idx = ref + i; best.prim = prim; best.kind = kind; found = true;
                            }
                        }
                    }
                }
                if (side) hr = false; else hl = false;
            }
        }
        if (hl && hr) {
            uint32_t nearn = lref, farn = rref;
            if (tr < tl) { nearn = rref; farn = lref; }
            if (sp < kStackSize) stack[sp++] = farn;
            node = nearn;
        } else if (hl) node = lref;
        else if (hr) node = rref;
        else {
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    return found;
}

// Closest hit for a WARP PACKET of coherent rays (the 8x4 pixel patch of camera rays a warp owns).  The warp walks the
// BVH together with one shared stack: a child is visited if any lane's box test (against that lane's own best hit plus the
// tie margin) passes, near child first by majority vote, and every node / record is fetched once per warp.  Each lane keeps
// its own closest hit under the same rules as bvh_trace<0> (acceptance rule, ties -> lowest segment index), so the result is
// identical; only the order in which candidates are met differs, and the result does not depend on that order.
// Must be called by all 32 lanes; `active` = this lane has a ray.  `stack` = kStackSize words of shared memory per warp.
__device__ __forceinline__ bool bvh_trace_packet(const SceneDev& S, bool active, Vec3 o, Vec3 d, float tmin, float tmax, bool capped,
                                                 HitRec& best, uint32_t* stack, uint32_t& steps, uint32_t& isect) {
    best.t = tmax; best.idx = 0; best.prim = 0xFFFFFFFFu; best.kind = 0;
    bool found = false;
    if (S.n_seg == 0 || __ballot_sync(0xffffffffu, active) == 0u) return false;
    const uint32_t lane = threadIdx.x & 31;
    const RayQ rq = make_rayq(o, d);
    const RayBox rb = make_raybox(o, d);
    uint32_t node = 0;
    int sp = 0;
    while (true) {
        const Node64 nd = load_node(S.nodes + node);
        steps += (lane == 0);
        const float tcull = best.t + S.line_width;
        float tl, tr;
        const bool hl = active && box_hit(rb, nd.l0, nd.l1, tmin, tcull, tl);
        const bool hr = active && box_hit(rb, nd.r0, nd.r1, tmin, tcull, tr);
        const unsigned ml = __ballot_sync(0xffffffffu, hl), mr = __ballot_sync(0xffffffffu, hr);
        const uint32_t cw[2] = {__float_as_uint(nd.l0.w), __float_as_uint(nd.r0.w)};
        const unsigned mk[2] = {ml, mr};
        uint32_t inner[2]; int n_inner = 0; int inner_side[2];
#pragma unroll
        for (int side = 0; side < 2; side++) {
            if (!mk[side]) continue;
            const uint32_t w = cw[side];
            if (w & kLeafBit) {
                const uint32_t ref = w & kRefMask, cnt = ((w >> 27) & 15u) + 1u;
                isect += (lane == 0) ? cnt : 0u;
                const bool mine = side ? hr : hl;
                for (uint32_t i = 0; i < cnt; i++) {
                    const SegRec s = load_seg(S.segs + ref + i);
                    float t; uint32_t kind;
                    if (mine && seg_box_hit(rb, s, S.radius, tmin, tmax) && capsule_hit(rq, s, S.radius, capped, t, kind) && t >= tmin && t <= tmax) {
                        if (!found || t <= best.t) {
                            const uint32_t prim = __ldg(S.prim_ids + ref + i);
                            if (!found || t < best.t || prim < best.prim) { best.t = t; best.idx = ref + i; best.prim = prim; best.kind = kind; found = true; }
                        }
                    }
                }
            } else { inner_side[n_inner] = side; inner[n_inner++] = w; }
        }
        if (n_inner == 2) {
            // near child first: majority of the lanes that hit both (or either)
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
        (void)inner_side;
        __syncwarp();
    }
    return found;
}

// All candidates with hitT in [tmin, tmax]; f(record_index, t, kind, SegRec) per accepted candidate.
template <class F>
__device__ __forceinline__ void bvh_trace_all(const SceneDev& S, Vec3 o, Vec3 d, float tmin, float tmax, bool capped,
                                              uint32_t& steps, uint32_t& isect, F&& f) {
    if (S.n_seg == 0) return;
    const RayQ rq = make_rayq(o, d);
    const RayBox rb = make_raybox(o, d);
    uint32_t stack[kStackSize];
    int sp = 0;
    uint32_t node = 0;
    while (true) {
        const Node64 nd = load_node(S.nodes + node);
        steps++;
        float tl, tr;
        bool hl = box_hit(rb, nd.l0, nd.l1, tmin, tmax, tl);
        bool hr = box_hit(rb, nd.r0, nd.r1, tmin, tmax, tr);
        const uint32_t lw = __float_as_uint(nd.l0.w), rw = __float_as_uint(nd.r0.w);   // absent children have a box that never hits
        const uint32_t lref = lw & kRefMask, rref = rw & kRefMask;
        const uint32_t lcnt = (lw & kLeafBit) ? ((lw >> 27) & 15u) + 1u : 0u, rcnt = (rw & kLeafBit) ? ((rw >> 27) & 15u) + 1u : 0u;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            bool h = side ? hr : hl;
            uint32_t cnt = side ? rcnt : lcnt, ref = side ? rref : lref;
            if (h && cnt) {
                isect += cnt;
                for (uint32_t i = 0; i < cnt; i++) {
                    SegRec s = load_seg(S.segs + ref + i);
                    float t; uint32_t kind;
                    if (seg_box_hit(rb, s, S.radius, tmin, tmax) && capsule_hit(rq, s, S.radius, capped, t, kind) && t >= tmin && t <= tmax) f(ref + i, t, kind, s);
                }
                if (side) hr = false; else hl = false;
            }
        }
        if (hl && hr) { if (sp < kStackSize) stack[sp++] = rref; node = lref; }
        else if (hl) node = lref;
        else if (hr) node = rref;
        else {
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
}

}  // namespace lv
