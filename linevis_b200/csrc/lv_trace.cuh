// lv_trace.cuh -- per-thread BVH traversal over 64-byte child-pair nodes (closest / any / all hits).
//
// Replaces traceRayEXT / rayQueryEXT against the driver's acceleration structure
// (reference TubeRayTracing.glsl:56,68; VulkanRayTracedAmbientOcclusion.glsl:164-165,201-204).
// Result-defining rules (mirrored by the CPU oracle): a candidate is accepted iff its reported hitT lies
// in [tmin, tmax]; the closest hit is the smallest hitT, ties go to the lowest caller-side segment index.
//
// Cost model: one traversal step = one 64-byte node fetch (4 x LDG.128), one intersection =
// one 32-byte segment record fetch (2 x LDG.128).  `steps` / `isect` count exactly those.
#pragma once
#include "lv_shade.cuh"

namespace lv {

constexpr int kStackSize = 72;

struct RayBox {
    float ix, iy, iz;     // safe 1/d
    float ox, oy, oz;     // -o/d
};

__device__ __forceinline__ float safe_inv(float d) {
    const float tiny = 1e-30f;
    if (fabsf(d) < tiny) d = (__float_as_uint(d) >> 31) ? -tiny : tiny;
    return 1.0f / d;
}
__device__ __forceinline__ RayBox make_raybox(Vec3 o, Vec3 d) {
    RayBox b;
    b.ix = safe_inv(d.x); b.iy = safe_inv(d.y); b.iz = safe_inv(d.z);
    b.ox = -(o.x * b.ix); b.oy = -(o.y * b.iy); b.oz = -(o.z * b.iz);
    return b;
}
// slab test; entry distance in tn.  Slightly widened so the box test never rejects what the capsule test accepts.
__device__ __forceinline__ bool box_hit(const RayBox& rb, float4 mn, float4 mx, float tmin, float tmax, float& tn) {
    float ax = __fmaf_rn(mn.x, rb.ix, rb.ox), bx = __fmaf_rn(mx.x, rb.ix, rb.ox);
    float ay = __fmaf_rn(mn.y, rb.iy, rb.oy), by = __fmaf_rn(mx.y, rb.iy, rb.oy);
    float az = __fmaf_rn(mn.z, rb.iz, rb.oz), bz = __fmaf_rn(mx.z, rb.iz, rb.oz);
    float lo = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    float hi = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    tn = lo;
    return lo * 0.9999995f <= hi * 1.0000005f;
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
        bool hl = box_hit(rb, nd.l0, nd.l1, tmin, best.t, tl);
        bool hr = box_hit(rb, nd.r0, nd.r1, tmin, best.t, tr);
        uint32_t lref = __float_as_uint(nd.l0.w), lcnt = __float_as_uint(nd.l1.w);
        uint32_t rref = __float_as_uint(nd.r0.w), rcnt = __float_as_uint(nd.r1.w);
        // an absent child is encoded as inner reference 0 (the root is nobody's child); its inverted box must not be trusted
        hl = hl && (lcnt | lref);
        hr = hr && (rcnt | rref);
#pragma unroll
        for (int side = 0; side < 2; side++) {
            bool h = side ? hr : hl;
            uint32_t cnt = side ? rcnt : lcnt, ref = side ? rref : lref;
            if (h && cnt) {
                isect += cnt;
                for (uint32_t i = 0; i < cnt; i++) {
                    SegRec s = load_seg(S.segs + ref + i);
                    float t; uint32_t kind;
                    if (capsule_hit(rq, s, S.radius, capped, t, kind) && t >= tmin && t <= tmax) {
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
        uint32_t lref = __float_as_uint(nd.l0.w), lcnt = __float_as_uint(nd.l1.w);
        uint32_t rref = __float_as_uint(nd.r0.w), rcnt = __float_as_uint(nd.r1.w);
        // an absent child is encoded as inner reference 0 (the root is nobody's child); its inverted box must not be trusted
        hl = hl && (lcnt | lref);
        hr = hr && (rcnt | rref);
#pragma unroll
        for (int side = 0; side < 2; side++) {
            bool h = side ? hr : hl;
            uint32_t cnt = side ? rcnt : lcnt, ref = side ? rref : lref;
            if (h && cnt) {
                isect += cnt;
                for (uint32_t i = 0; i < cnt; i++) {
                    SegRec s = load_seg(S.segs + ref + i);
                    float t; uint32_t kind;
                    if (capsule_hit(rq, s, S.radius, capped, t, kind) && t >= tmin && t <= tmax) f(ref + i, t, kind, s);
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
