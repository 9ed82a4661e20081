// lv_trace.cuh -- BVH traversal building blocks over 64-byte child-pair nodes: canonical slab test, loads, the warp-packet
// closest-hit traversal.  (The AO ray stream and the PPLL all-hits packet live in lv_kernels.cuh.)
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

LV_DEV float safe_inv(float d) {
    const float tiny = 1e-30f;
    if (fabsf(d) < tiny) d = (__float_as_uint(d) >> 31) ? -tiny : tiny;
    return 1.0f / d;
}
LV_DEV RayBox make_raybox(Vec3 o, Vec3 d) {
    RayBox b;
    b.ix = safe_inv(d.x); b.iy = safe_inv(d.y); b.iz = safe_inv(d.z);
    b.cx = -(o.x * b.ix); b.cy = -(o.y * b.iy); b.cz = -(o.z * b.iz);
    return b;
}
LV_DEV bool box_hit(const RayBox& rb, float mnx, float mny, float mnz, float mxx, float mxy, float mxz,
                                        float tmin, float tmax, float& tn) {
    float ax = __fmaf_rn(mnx, rb.ix, rb.cx), bx = __fmaf_rn(mxx, rb.ix, rb.cx);
    float ay = __fmaf_rn(mny, rb.iy, rb.cy), by = __fmaf_rn(mxy, rb.iy, rb.cy);
    float az = __fmaf_rn(mnz, rb.iz, rb.cz), bz = __fmaf_rn(mxz, rb.iz, rb.cz);
    float lo = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    float hi = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    tn = lo;
    return lo <= hi;
}
LV_DEV bool box_hit(const RayBox& rb, float4 mn, float4 mx, float tmin, float tmax, float& tn) {
    return box_hit(rb, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, tmin, tmax, tn);
}
// the segment's own AABB (min/max(p0,p1) -+ r, reference src/LineData/LineDataFlow.cpp:2230-2233) against the ray's
// ORIGINAL interval: part of the acceptance rule, and a cheap reject in front of the three quadratics
LV_DEV bool seg_box_hit(const RayBox& rb, const SegRec& s, float r, float tmin, float tmax) {
    float tn;
    return box_hit(rb, fminf(s.a.x, s.b.x) - r, fminf(s.a.y, s.b.y) - r, fminf(s.a.z, s.b.z) - r,
                   fmaxf(s.a.x, s.b.x) + r, fmaxf(s.a.y, s.b.y) + r, fmaxf(s.a.z, s.b.z) + r, tmin, tmax, tn);
}

// 256-bit read-only loads (LDG.E.256, new on sm_100): a 64-byte node is 2 load instructions instead of 4, a 32-byte
// segment record 1 instead of 2.  With every lane on a different node the L1 handles one wavefront per lane PER LOAD
// INSTRUCTION, and that wavefront rate -- not DRAM, not issue -- was what bound k_rtao_rays (profiles/r1e: L1 76 %).
#ifdef LV_HOST_EMU
LV_DEV void ldg256(const void* p, float4& a, float4& b) { a = static_cast<const float4*>(p)[0]; b = static_cast<const float4*>(p)[1]; }
#else
LV_DEV void ldg256(const void* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
#endif
LV_DEV Node64 load_node(const Node64* p) {
    Node64 n;
    ldg256(p, n.l0, n.l1);
    ldg256(reinterpret_cast<const char*>(p) + 32, n.r0, n.r1);
    return n;
}
LV_DEV SegRec load_seg(const SegRec* p) {
    SegRec s;
    ldg256(p, s.a, s.b);
    return s;
}

struct HitRec {
    float t;
    uint32_t idx;    // record index in BVH order
    uint32_t prim;   // caller-side segment index
    uint32_t kind;
};

#if !defined(LV_HOST_EMU) || defined(LV_HOST_EMU_SIMT)   // warp-collective code: needs a GPU or the SIMT emulator (tests/emu/emu_cuda.hpp)
// Closest hit for a WARP PACKET of coherent rays (the 8x4 pixel patch of camera rays a warp owns).  The warp walks the
// BVH together with one shared stack: a child is visited if any lane's box test (against that lane's own best hit plus the
// tie margin) passes, near child first by majority vote, and every node / record is fetched once per warp.  Each lane keeps
// its own closest hit (acceptance rule above, ties -> lowest segment index); the order in which candidates are met does not
// influence the result.
// Leaf tests are DEFERRED and BATCHED: at a leaf only the lanes whose own box test passed want the record (ncu on k_tubes: the
// capsule test ran with 5 - 9 of 32 lanes, 40 % of the kernel's warp instructions).  They append (lane, record) to a warp queue; as
// soon as kPacketFlush entries are queued the lanes run ONE test each -- entry i on lane i, the owner's ray fetched with shuffles -- and hands the
// result to the owner through a 64-bit shared word: key = (hit distance bits << 32) | caller-side segment index, merged with
// atomicMin, which IS the result rule (smallest distance, ties -> lowest segment index).  The record index / hit kind of the winning
// key are written by the lane that holds it after the batch.  Owners cull with the distance they know so far.
// Must be called by all 32 lanes; `active` = this lane has a ray.  `scratch` = one PacketScratch of shared memory per warp.
// Queued tests at which a batch runs.  32 = only full batches: best for dense data (config 5: 3.79 -> 3.14 ms for the two packet kernels),
// but a packet that meets fewer than 32 candidates on its whole way then never learns a hit distance to cull with (config 3: 1.55 ->
// 1.88 ms).  LV_PACKET_FLUSH is the compromise measured on both.
#ifndef LV_PACKET_FLUSH
#define LV_PACKET_FLUSH 12
#endif
constexpr uint32_t kPacketFlush = LV_PACKET_FLUSH;
struct PacketScratch {
    unsigned long long key[32];   // per lane: closest accepted hit so far
    uint32_t aux[32];             // ... its record index (BVH order) | hit kind << 28
    uint32_t queue[64];           // ring of deferred tests: record index | owner lane << 27
    uint32_t stack[kStackSize];
};

__device__ __forceinline__ void packet_flush(const SceneDev& S, PacketScratch& sc, uint32_t lane, uint32_t n, uint32_t& q_head, uint32_t& q_count,
                                             const RayQ& rq, const RayBox& rb, float tmin, float tmax, bool capped, float& best_t) {
    const uint32_t e = sc.queue[(q_head + (lane < n ? lane : 0u)) & 63u];
    const uint32_t owner = e >> 27, rec = e & kRefMask;
    RayQ r2; RayBox b2;
    r2.o.x = __shfl_sync(0xffffffffu, rq.o.x, owner); r2.o.y = __shfl_sync(0xffffffffu, rq.o.y, owner); r2.o.z = __shfl_sync(0xffffffffu, rq.o.z, owner);
    r2.d.x = __shfl_sync(0xffffffffu, rq.d.x, owner); r2.d.y = __shfl_sync(0xffffffffu, rq.d.y, owner); r2.d.z = __shfl_sync(0xffffffffu, rq.d.z, owner);
    r2.dd = __shfl_sync(0xffffffffu, rq.dd, owner);
    b2.ix = __shfl_sync(0xffffffffu, rb.ix, owner); b2.iy = __shfl_sync(0xffffffffu, rb.iy, owner); b2.iz = __shfl_sync(0xffffffffu, rb.iz, owner);
    b2.cx = __shfl_sync(0xffffffffu, rb.cx, owner); b2.cy = __shfl_sync(0xffffffffu, rb.cy, owner); b2.cz = __shfl_sync(0xffffffffu, rb.cz, owner);
    const float tmin2 = __shfl_sync(0xffffffffu, tmin, owner);
    unsigned long long mykey = ~0ull;
    uint32_t kind = 0;
    if (lane < n) {
        const SegRec s = load_seg(S.segs + rec);
        float t;
        const float4 ax4 = __ldg(S.seg_axes + rec);   // = seg_axis(s), stored at scene creation
        if (seg_box_hit(b2, s, S.radius, tmin2, tmax) && capsule_hit(r2, s, v3(ax4.x, ax4.y, ax4.z), S.radius, capped, t, kind) && t >= tmin2 && t <= tmax) {
            mykey = (static_cast<unsigned long long>(__float_as_uint(t)) << 32) | __ldg(S.prim_ids + rec);   // t > 0: the bit pattern orders like the value
            atomicMin(&sc.key[owner], mykey);
        }
    }
    __syncwarp();
    if (mykey != ~0ull && sc.key[owner] == mykey) sc.aux[owner] = rec | (kind << 28);   // the winner (segment indices are unique) names its record
    __syncwarp();
    best_t = __uint_as_float(uint32_t(sc.key[lane] >> 32));
    q_head = (q_head + n) & 63u;
    q_count -= n;
}

__device__ __forceinline__ bool bvh_trace_packet(const SceneDev& S, bool active, Vec3 o, Vec3 d, float tmin, float tmax, bool capped,
                                                 HitRec& best, PacketScratch& sc, uint32_t& steps, uint32_t& isect) {
    best.t = tmax; best.idx = 0; best.prim = 0xFFFFFFFFu; best.kind = 0;
    if (S.n_seg == 0 || __ballot_sync(0xffffffffu, active) == 0u) return false;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const RayQ rq = make_rayq(o, d);
    const RayBox rb = make_raybox(o, d);
    const unsigned long long no_hit = (static_cast<unsigned long long>(__float_as_uint(tmax)) << 32) | 0xFFFFFFFFull;
    sc.key[lane] = no_hit;
    sc.aux[lane] = 0u;
    __syncwarp();
    float best_t = tmax;
    uint32_t q_head = 0, q_count = 0;   // uniform over the warp
    uint32_t node = 0;
    int sp = 0;
    while (true) {
        const Node64 nd = load_node(S.nodes + node);
        steps += (lane == 0);
        const float tcull = best_t + S.line_width;
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
                const uint32_t nm = __popc(mk[side]);
                for (uint32_t i = 0; i < cnt; i++) {
                    if (mine) sc.queue[(q_head + q_count + __popc(mk[side] & lt_mask)) & 63u] = (ref + i) | (lane << 27);
                    q_count += nm;
                    __syncwarp();
                    if (q_count >= kPacketFlush) packet_flush(S, sc, lane, q_count < 32u ? q_count : 32u, q_head, q_count, rq, rb, tmin, tmax, capped, best_t);
                }
            } else inner[n_inner++] = w;
        }
        if (n_inner == 2) {
            // near child first: majority of the lanes that hit both (or either)
            const unsigned right_near = __ballot_sync(0xffffffffu, hr && (!hl || tr < tl));
            const unsigned left_near = __ballot_sync(0xffffffffu, hl && (!hr || tl <= tr));
            const bool rf = __popc(right_near) > __popc(left_near);
            if (lane == 0) sc.stack[sp] = rf ? inner[0] : inner[1];
            sp++;
            node = rf ? inner[1] : inner[0];
        } else if (n_inner == 1) node = inner[0];
        else {
            if (sp == 0) break;
            --sp;
            __syncwarp();
            node = sc.stack[sp];
        }
        __syncwarp();
    }
    if (q_count) packet_flush(S, sc, lane, q_count, q_head, q_count, rq, rb, tmin, tmax, capped, best_t);
    const unsigned long long k = sc.key[lane];
    const uint32_t a = sc.aux[lane];
    __syncwarp();   // the scratch is reused by the next packet of this warp
    if (k == no_hit) return false;
    best.t = __uint_as_float(uint32_t(k >> 32)); best.prim = uint32_t(k); best.idx = a & kRefMask; best.kind = a >> 28;
    return true;
}
#endif  // warp-collective code

}  // namespace lv
