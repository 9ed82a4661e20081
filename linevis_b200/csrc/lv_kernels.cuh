// lv_kernels.cuh -- frame kernels: primary closest hit, tube ray-gen (S1-S4), RTAO (S5), PPLL clear/gather/resolve (S8-S10).
//
// Thread mapping shared by all per-pixel kernels: the owned image tiles (tile_size x tile_size, Morton-ordered
// list in FrameParams) are cut into 16x8-pixel blocks of 128 threads; a warp covers an 8x4 pixel patch so that
// its primary rays are coherent and its framebuffer / list-head accesses fall into full 32-byte sectors.
#pragma once
#include "lv_trace.cuh"
#include "lv_bake.cuh"
#include "lv_tri.cuh"

namespace lv {

constexpr int kBlockThreads = 128;
constexpr uint32_t kNone = 0xFFFFFFFFu;

__device__ __forceinline__ bool thread_pixel(const FrameParams& P, uint32_t& x, uint32_t& y) {
    const uint32_t bx = P.tile_size >> 4, by = P.tile_size >> 3;
    const uint32_t bpt = bx * by;
    const uint32_t tile = blockIdx.x / bpt, sub = blockIdx.x - tile * bpt;
    const uint2 t = P.tiles[tile];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    x = t.x * P.tile_size + (sub % bx) * 16 + (warp & 1) * 8 + (lane & 7);
    y = t.y * P.tile_size + (sub / bx) * 8 + (warp >> 1) * 4 + (lane >> 3);
    return x < P.W && y < P.H;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void flush_counter(unsigned long long* dst, unsigned long long v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
}

// packUnorm4x8 (GLSL 4.60 spec 8.4): round(clamp(c, 0, 1) * 255) per channel
__device__ __forceinline__ uint32_t pack_unorm4x8(Vec4 c) {
    uint32_t r = uint32_t(floorf(clampf_(c.x, 0.0f, 1.0f) * 255.0f + 0.5f));
    uint32_t g = uint32_t(floorf(clampf_(c.y, 0.0f, 1.0f) * 255.0f + 0.5f));
    uint32_t b = uint32_t(floorf(clampf_(c.z, 0.0f, 1.0f) * 255.0f + 0.5f));
    uint32_t a = uint32_t(floorf(clampf_(c.w, 0.0f, 1.0f) * 255.0f + 0.5f));
    return r | (g << 8) | (b << 16) | (a << 24);
}

// RayGen camera ray (reference TubeRayTracing.glsl:202,219-226)
__device__ __forceinline__ void camera_ray(const FrameParams& P, uint32_t px, uint32_t py, float xix, float xiy, Vec3& ro, Vec3& rd) {
    Vec4 o = mat_mul(P.inv_view, v4(0.0f, 0.0f, 0.0f, 1.0f));
    ro = v3(o.x, o.y, o.z);
    float nx = 2.0f * ((float(px) + xix) / float(P.W)) - 1.0f;
    float ny = 2.0f * ((float(py) + xiy) / float(P.H)) - 1.0f;
    Vec4 tg = mat_mul(P.inv_proj, v4(nx, ny, 1.0f, 1.0f));
    Vec3 nt = normalize3(v3(tg.x, tg.y, tg.z));
    Vec4 d = mat_mul(P.inv_view, v4(nt.x, nt.y, nt.z, 0.0f));
    rd = v3(d.x, d.y, d.z);
}

// ------------------------------------------------------------------------------------------------
// Depth cues: min / max view depth over the line vertices inside the frustum (+- 1e-2), reference
// Data/Shaders/DepthCues/ComputeDepthValues.glsl:60-98 + MinMaxDepthReduction (LineRenderer.cpp:410-431).  min/max are exact
// and order independent, so one pass with warp shuffles + integer atomics on the (positive) float bit patterns replaces
// the reference's multi-pass tree reduction.  out = {min, max}, initialised to {farDist, nearDist} by the host.
// the cylinder axis of every record, once per scene: the capsule tests of the AO batches and of the packets' leaf batches load it
// (16 bytes) instead of re-deriving it -- an IEEE sqrt and an IEEE division, about 30 instructions -- for every (ray, record) pair
__global__ void k_seg_axes(const SegRec* segs, uint32_t n, float4* axes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Vec3 a = seg_axis(segs[i]);
    axes[i] = make_float4(a.x, a.y, a.z, 0.0f);
}

__global__ void k_depth_range(const __grid_constant__ FrameParams P, const SegRec* segs, uint32_t n_seg, float* out) {
    float dmin = P.far_dist, dmax = P.near_dist;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_seg; i += gridDim.x * blockDim.x) {
        const SegRec s = segs[i];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float4 p = e ? s.b : s.a;
            const Vec4 sp = mat_mul(P.view, v4(p.x, p.y, p.z, 1.0f));
            const Vec4 ndc = mat_mul(P.proj, sp);
            const float nx = ndc.x / ndc.w, ny = ndc.y / ndc.w, nz = ndc.z / ndc.w;
            if (nx >= -1.0f && ny >= -1.0f && nz >= -1.0f && nx <= 1.0f && ny <= 1.0f && nz <= 1.0f) {
                const float depth = clampf_(-sp.z, P.near_dist, P.far_dist);
                dmin = minf_(dmin, depth - 1e-2f);
                dmax = maxf_(dmax, depth + 1e-2f);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = minf_(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = maxf_(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // near > 1e-2 is not guaranteed, so the values may be negative: use the sign-aware float atomics of the BVH builder
        atomic_min_f(out, dmin);
        atomic_max_f(out + 1, dmax);
    }
}

// ------------------------------------------------------------------------------------------------
// closest-hit only (parity / debugging entry point lv_trace_primary)
__global__ void __launch_bounds__(kBlockThreads)
k_primary(const __grid_constant__ FrameParams P, const __grid_constant__ SceneDev S, lv_hit* hits, Counters* C) {
    __shared__ PacketScratch s_scratch[kBlockThreads / 32];
    uint32_t x, y;
    const bool valid = thread_pixel(P, x, y);
    uint32_t steps = 0, isect = 0, nhit = 0;
    Vec3 ro = v3(0, 0, 0), rd = v3(0, 0, 1);
    if (valid) camera_ray(P, x, y, 0.5f, 0.5f, ro, rd);
    HitRec h;
    const bool hit = bvh_trace_packet(S, valid, ro, rd, 0.0001f, 1000.0f, P.use_capped != 0, h, s_scratch[threadIdx.x >> 5], steps, isect);
    if (valid) {
        lv_hit out; out.t = 0.0f; out.prim = kNone; out.kind = 0; out.pad = 0;
        if (hit) { out.t = h.t; out.prim = h.prim; out.kind = h.kind; nhit = 1; }
        hits[size_t(y) * P.W + x] = out;
    }
    flush_counter(&C->rays_primary, valid ? 1 : 0);
    flush_counter(&C->steps, steps);
    flush_counter(&C->isect, isect);
    flush_counter(&C->pixels_hit, nhit);
}

// ------------------------------------------------------------------------------------------------
// S1 ray-gen with the traceRayTransparent loop, S3/S4 shading and the running mean over frames
// (reference TubeRayTracing.glsl:61-82,198-274).  `image` is the accumulation image (float RGBA).
// First hit of the tube pass's first sample, traced AHEAD of k_tubes on a second stream (b200_tube_prepass): the closest-hit
// traversal does not depend on the AO image, so it runs beside k_rtao_primary and in the shadow of the AO ray stream's tail (a
// persistent kernel drains unevenly: ~0.3 ms during which most SMs idle -- 1 % of a one-GPU frame, 7 % of an 8-GPU one).
// first[pixel] = (t bits, record index | hit kind << 28 | hit << 31).
__global__ void __launch_bounds__(kBlockThreads)
k_tube_first(const __grid_constant__ FrameParams P, const __grid_constant__ SceneDev S, uint2* first, Counters* C) {
    __shared__ PacketScratch s_scratch[kBlockThreads / 32];
    uint32_t x, y;
    const bool valid = thread_pixel(P, x, y);
    uint32_t steps = 0, isect = 0;
    float xix = 0.5f, xiy = 0.5f;
    if (P.use_jitter) {   // sample 0 of k_tubes
        uint32_t seed = P.det_sampling ? tea(19u, P.frame_number * P.spp) : tea(x + y * P.W, P.frame_number * P.spp);
        xix = rnd(seed); xiy = rnd(seed);
    }
    Vec3 ro = v3(0, 0, 0), rd = v3(0, 0, 1);
    if (valid) camera_ray(P, x, y, xix, xiy, ro, rd);
    HitRec h;
    const bool hit = bvh_trace_packet(S, valid, ro, rd, 0.0001f, 1000.0f, P.use_capped != 0, h, s_scratch[threadIdx.x >> 5], steps, isect);
    if (valid) first[size_t(y) * P.W + x] = hit ? make_uint2(__float_as_uint(h.t), h.idx | (h.kind << 28) | 0x80000000u) : make_uint2(0u, 0u);
    flush_counter(&C->steps, steps);
    flush_counter(&C->isect, isect);
}

// out8 != nullptr (b200_frame_format = rgba8): the frame is ALSO stored as RGBA8 UNORM -- the reference's own sceneTexture format
// (TubeRayTracing.glsl:42, src/Widgets/DataView.cpp:100-108) -- in the same epilogue; `image` then is the library's own float
// accumulation image and out8 the delivered frame (4 B / pixel: a quarter of the peer-store and read-back traffic).
// PRIM = 1: the reference's triangle-mesh geometry mode (geometry_mode = "Triangle Mesh"): closest hit against the triangulated tubes,
// ClosestHitTubeTriangles (lv_tri.cuh: shade_tri_hit).
template <bool SAO, int PRIM = 0>
__global__ void __launch_bounds__(kBlockThreads)
k_tubes(const __grid_constant__ FrameParams P, const __grid_constant__ SceneDev S, float4* image, Counters* C, uint32_t* out8, const uint2* first) {
    __shared__ PacketScratch s_scratch[kBlockThreads / 32];
    PacketScratch& stack = s_scratch[threadIdx.x >> 5];
    uint32_t x, y;
    const bool valid = thread_pixel(P, x, y);
    uint32_t steps = 0, isect = 0, rays = 0, nhit = 0;
    float fr = 0.0f, fg = 0.0f, fb = 0.0f, fa = 0.0f;
    const uint32_t nspp = P.use_jitter ? P.spp : 1u;
    for (uint32_t si = 0; si < nspp; si++) {
        float xix = 0.5f, xiy = 0.5f;
        if (P.use_jitter) {
            uint32_t seed = P.det_sampling ? tea(19u, P.frame_number * P.spp + si)
                                           : tea(x + y * P.W, P.frame_number * P.spp + si);
            xix = rnd(seed); xiy = rnd(seed);
        }
        Vec3 ro = v3(0, 0, 0), rd = v3(0, 0, 1);
        if (valid) camera_ray(P, x, y, xix, xiy, ro, rd);
        float cr = 0.0f, cg = 0.0f, cb = 0.0f, ca = 0.0f;
        float tmin = 0.0001f;
        bool live = valid;
        // traceRayTransparent (:61-82): the warp's rays are traced as a packet; lanes drop out at a miss or alpha > 0.99
        for (uint32_t hi = 0; hi < P.max_depth; hi++) {
            if (__ballot_sync(0xffffffffu, live) == 0u) break;
            HitRec h;
            TriHitRec th;
            bool hit;
            if (PRIM == 1) hit = bvh_trace_packet_tri(S, live, ro, rd, tmin, 1000.0f, th, stack.stack, steps, isect);
            else if (first && hi == 0 && si == 0) {   // traced ahead by k_tube_first (same ray, same acceptance rule)
                const uint2 f = valid ? first[size_t(y) * P.W + x] : make_uint2(0u, 0u);
                hit = (f.y >> 31) != 0u;
                h.t = __uint_as_float(f.x); h.idx = f.y & kRefMask; h.kind = (f.y >> 28) & 3u; h.prim = 0u;
            } else hit = bvh_trace_packet(S, live, ro, rd, tmin, 1000.0f, P.use_capped != 0, h, stack, steps, isect);
            if (live) {
                rays++;
                Vec4 hc; float hit_t;
                if (hit && PRIM == 1) {
                    const Shaded sh = shade_tri_hit(P, S, load_tri(S.tris + th.idx), th.u, th.v);
                    hc = sh.color; hit_t = sh.hit_t;
                    if (hi == 0 && si == 0) nhit = 1;
                } else if (hit) {
                    const SegRec s = load_seg(S.segs + h.idx);
                    const Shaded sh = shade_hit<SAO>(P, ro, rd, h.t, h.kind, s, SAO && S.seg_aux ? S.seg_aux + h.idx : nullptr);
                    hc = sh.color; hit_t = sh.hit_t;
                    if (hi == 0 && si == 0) nhit = 1;
                } else {  // Miss (TubeRayTracing.glsl:290-298)
                    hc = v4(P.bg[0], P.bg[1], P.bg[2], P.bg[3]); hit_t = 0.0f;
                }
                tmin = hit_t + maxf_(hit_t * 1e-5f, 1e-7f);
                cr = cr + (1.0f - ca) * hc.w * hc.x;
                cg = cg + (1.0f - ca) * hc.w * hc.y;
                cb = cb + (1.0f - ca) * hc.w * hc.z;
                ca = ca + (1.0f - ca) * hc.w;
                if (!hit || ca > 0.99f) live = false;
            }
        }
        fr += cr; fg += cg; fb += cb; fa += ca;
    }
    if (valid) {
        if (P.use_jitter) { float dn = float(P.spp); fr /= dn; fg /= dn; fb /= dn; fa /= dn; }
        float4* px = image + size_t(y) * P.W + x;
        if (P.frame_number != 0) {
            float4 prev = *px;
            float a = 1.0f / float(P.frame_number + 1);
            fr = mixf_(prev.x, fr, a); fg = mixf_(prev.y, fg, a); fb = mixf_(prev.z, fb, a); fa = mixf_(prev.w, fa, a);
        }
        *px = make_float4(fr, fg, fb, fa);
        if (out8) out8[size_t(y) * P.W + x] = pack_unorm4x8(v4(fr, fg, fb, fa));
    }
    flush_counter(&C->rays_primary, rays);
    flush_counter(&C->steps, steps);
    flush_counter(&C->isect, isect);
    flush_counter(&C->pixels_hit, nhit);
}

// ------------------------------------------------------------------------------------------------
// S5 (reference Data/Shaders/AO/RTAO/VulkanRayTracedAmbientOcclusion.glsl:178-319) in three kernels:
//   k_rtao_primary  one (optionally jittered) camera ray per pixel; a miss writes aoFactor = 1 through the running
//                   mean, a hit appends the shading frame of the hit point to a compact list;
//   k_rtao_rays     persistent ray-stream kernel over (hit pixel, sample) pairs, see below;
//   k_rtao_reduce   sums the per-sample occlusion values of each hit pixel IN SAMPLE ORDER (exactly the shader's
//                   loop, :283-306) and folds the mean into the accumulation image (:313-317).
__device__ __forceinline__ void apron_mark_owned(const FrameParams& P, uint32_t x, uint32_t y, unsigned int stamp) {
    P.apron_marks[size_t(y) * P.W + x] = stamp;
}

// PRIM = 0: analytic capsules (the default); PRIM = 1: the reference's triangulated tubes (lv_tri.cuh)
template <int PRIM>
__global__ void __launch_bounds__(kBlockThreads)
k_rtao_primary(const __grid_constant__ FrameParams P, const __grid_constant__ SceneDev S, float* ao, AoHit* hit_list,
               unsigned int* hit_count, Counters* C, unsigned int* apron_mark, unsigned int apron_stamp) {
    uint32_t x, y;
    bool valid;
    if (apron_mark) {
        // apron launch (tile-sharded + jittered tube rays only): the one-pixel ring around every owned tile, because the
        // tube pass looks the AO image up bilinearly at a jittered position.  A ring pixel shared by several owned tiles, or
        // owned itself, is claimed once through the stamp array.
        const uint32_t ts = P.tile_size, ring = 4 * ts + 4, per_tile = (ring + kBlockThreads - 1) / kBlockThreads;
        const uint32_t tile = blockIdx.x / per_tile, t = (blockIdx.x - tile * per_tile) * kBlockThreads + threadIdx.x;
        const uint2 tl = P.tiles[tile];
        int lx, ly;
        if (t < ts + 2) { lx = int(t) - 1; ly = -1; }
        else if (t < 2 * ts + 4) { lx = int(t - (ts + 2)) - 1; ly = int(ts); }
        else if (t < 3 * ts + 4) { lx = -1; ly = int(t - (2 * ts + 4)); }
        else { lx = int(ts); ly = int(t - (3 * ts + 4)); }
        const long long gx = (long long)tl.x * ts + lx, gy = (long long)tl.y * ts + ly;
        valid = t < ring && gx >= 0 && gy >= 0 && gx < (long long)P.W && gy < (long long)P.H;
        x = uint32_t(gx); y = uint32_t(gy);
        if (valid) valid = atomicExch(apron_mark + size_t(y) * P.W + x, apron_stamp) != apron_stamp;
    } else valid = thread_pixel(P, x, y);
    uint32_t steps = 0, isect = 0;
    bool hit = false;
    AoHit rec;
    if (valid && !apron_mark && apron_stamp) apron_mark_owned(P, x, y, apron_stamp);
    __shared__ PacketScratch s_scratch[kBlockThreads / 32];
    Vec3 ro = v3(0, 0, 0), rd = v3(0, 0, 1);
    if (valid) {
        uint32_t seed = tea(x + y * P.W, P.frame_number);
        float xix = 0.5f, xiy = 0.5f;
        if (P.ao_jitter) { xix = rnd(seed); xiy = rnd(seed); }
        camera_ray(P, x, y, xix, xiy, ro, rd);
    }
    HitRec h;
    TriHitRec th;
    if (PRIM == 1) hit = bvh_trace_packet_tri(S, valid, ro, rd, 0.0001f, 1000.0f, th, s_scratch[threadIdx.x >> 5].stack, steps, isect);
    else hit = bvh_trace_packet(S, valid, ro, rd, 0.0001f, 1000.0f, P.use_capped != 0, h, s_scratch[threadIdx.x >> 5], steps, isect);
    if (valid) {
        if (hit && PRIM == 1) {
            rec = tri_ao_frame(S, load_tri(S.tris + th.idx), th.u, th.v, P.subdiv_corr, y * P.W + x);   // the shader's barycentric fetch (:213-276)
        } else if (hit) {
            // analytic stand-in for the barycentric vertex fetch of the triangle-mesh path (:222-276)
            const SegRec s = load_seg(S.segs + h.idx);
            const Vec3 p0 = v3(s.a.x, s.a.y, s.a.z), p1 = v3(s.b.x, s.b.y, s.b.z);
            const Vec3 pos = ro + rd * h.t;
            const Vec3 seg = p1 - p0;
            Vec3 centre;
            if (h.kind == 0) { float u = dot3(seg, pos - p0) / dot3(seg, seg); centre = p0 + u * seg; }
            else if (h.kind == 1) centre = p0; else centre = p1;
            const Vec3 nrm = normalize3(pos - centre);
            const Vec3 tng = normalize3(seg);
            const float offset = length3(centre - pos) / P.subdiv_corr;
            rec.pos_off = make_float4(pos.x, pos.y, pos.z, offset);
            rec.nrm_px = make_float4(nrm.x, nrm.y, nrm.z, __uint_as_float(y * P.W + x));
            rec.tng = make_float4(tng.x, tng.y, tng.z, 0.0f);
        } else {
            float v = 1.0f;
            float* p = ao + size_t(y) * P.W + x;
            if (P.frame_number != 0) v = mixf_(*p, v, 1.0f / float(P.frame_number + 1));
            *p = v;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
        const uint32_t lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(hit_count, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit) hit_list[base + __popc(m & ((1u << lane) - 1u))] = rec;
    }
    flush_counter(&C->rays_primary, valid ? 1 : 0);
    flush_counter(&C->steps, steps);
    flush_counter(&C->isect, isect);
    flush_counter(&C->ao_pixels_hit, hit ? 1 : 0);
}

// Persistent ray-stream kernel for the AO rays (the dominant kernel of path (a)).
//
// The first version of this kernel gave each warp one pixel's samples and ran an if-if traversal; ncu showed it
// issue-bound at 3.7 of 32 lanes active (profiles/r1a_*): incoherent short rays finish at very different times and leaf
// work diverges from box work.  This version keeps the lanes busy instead (11.6 of 32, profiles/r1c_*):
//   - rays are numbered r = hit_slot * spp + sample; a lane whose ray has finished takes the next number from a global
//     counter (one atomic per warp refill, ranks by ballot/popc) as soon as fewer than ao_refill_below lanes of its warp
//     are live; consecutive numbers share a pixel, i.e. rays start out coherent;
//   - every iteration all lanes that hold an inner node take ONE step; a lane that reaches a leaf postpones it (leaves
//     travel as encoded child words, also on the stack) and waits until ao_leaf_vote lanes hold one, then those lanes
//     intersect their leaves together;
//   - popped entries carry their box entry distance and are skipped when the closest hit found meanwhile is nearer.
// Tried and dropped (measured, no gain): sorting the rays of 128 neighbouring hit pixels into 64 direction bins,
// warp-private chunks of consecutive ray numbers, a second compaction pass over box-test survivors inside the leaf phase,
// precomputing the ray directions in a separate full-utilisation kernel (the refill is not where the time goes).
// The per-ray result (4 B) goes to occ[r]; sample-ordered summation happens in k_rtao_reduce.
constexpr uint32_t kDone = 0x7FFFFFFFu;
// rq.dd = |d|^2 is NaN exactly when a component of the direction is NaN (it cannot be 0 or inf for a normalised direction)
__device__ __forceinline__ bool ao_ray_valid(const RayQ& rq) { return rq.dd == rq.dd; }
constexpr int kAoStack = 72;

// BAKE = object-space prebaker (lv_bake.cuh): records are (parametrization vertex, tube subdivision) frames, the ray origin is the
// record's position itself and the random numbers come from the vertex's LCG stream instead of a per-sample TEA seed.
// STACK = layout of the per-lane traversal stack (b200_ao_stack):
//   0  two local-memory arrays (node word, entry distance) -- the measured default;
//   1  one local-memory array of packed 64-bit entries (one LDL/STL.64 per pop / push instead of two 32-bit ones);
//   K >= 2  the first K packed entries in SHARED memory, laid out [entry][thread] so that a warp's accesses never
//      conflict whatever the lanes' stack depths are (local memory is interleaved per 32-bit word: lanes at different
//      depths touch different 128-byte lines, one L1 wavefront each), deeper entries spill to a local array.
//      K x 8 B x 128 threads = K KiB of shared memory per block.

constexpr int kAoStackWide = 96;   // 4-wide tree: a step pushes up to three entries (lv_scene checks the tree's exact need against this)
template <int K, int CAP = kAoStack> struct AoStack {
    static constexpr int kAoSmemStack = K;
    unsigned long long e[CAP - kAoSmemStack];
#ifdef LV_HOST_EMU   // host emulation: plain indexing instead of shared-window addresses + inline PTX
    unsigned long long* sm;
    __device__ __forceinline__ void init() {
        __shared__ unsigned long long s_stack[kAoSmemStack][kBlockThreads];
        sm = &s_stack[0][threadIdx.x];
    }
    __device__ __forceinline__ void put(int i, uint32_t node, float t) {
        const unsigned long long v = (static_cast<unsigned long long>(__float_as_uint(t)) << 32) | node;
        if (i < kAoSmemStack) sm[size_t(i) * kBlockThreads] = v; else e[i - kAoSmemStack] = v;
    }
    __device__ __forceinline__ unsigned long long get(int i) const { return i < kAoSmemStack ? sm[size_t(i) * kBlockThreads] : e[i - kAoSmemStack]; }
#else
    uint32_t sm;   // shared-window address of this thread's column: entry i at sm + i * kBlockThreads * 8
    __device__ __forceinline__ void init() {
        __shared__ unsigned long long s_stack[kAoSmemStack][kBlockThreads];
        sm = uint32_t(__cvta_generic_to_shared(&s_stack[0][threadIdx.x]));
    }
    __device__ __forceinline__ void put(int i, uint32_t node, float t) {
        const unsigned long long v = (static_cast<unsigned long long>(__float_as_uint(t)) << 32) | node;
        if (i < kAoSmemStack) asm volatile("st.shared.u64 [%0], %1;" ::"r"(sm + uint32_t(i) * (kBlockThreads * 8u)), "l"(v) : "memory");
        else e[i - kAoSmemStack] = v;
    }
    __device__ __forceinline__ unsigned long long get(int i) const {
        unsigned long long v;
        if (i < kAoSmemStack) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(sm + uint32_t(i) * (kBlockThreads * 8u)) : "memory");
        else v = e[i - kAoSmemStack];
        return v;
    }
#endif
};
template <int CAP> struct AoStack<1, CAP> {
    unsigned long long e[CAP];
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void put(int i, uint32_t node, float t) { e[i] = (static_cast<unsigned long long>(__float_as_uint(t)) << 32) | node; }
    __device__ __forceinline__ unsigned long long get(int i) const { return e[i]; }
};
// pop until an entry whose box entry distance is still within reach; returns kDone if the stack runs empty
template <int STACK, int CAP>
__device__ __forceinline__ uint32_t ao_stack_pop(const AoStack<STACK, CAP>& st, int& sp, float best) {
    while (sp > 0) {
        --sp;
        const unsigned long long v = st.get(sp);
        if (__uint_as_float(uint32_t(v >> 32)) <= best) return uint32_t(v);
    }
    return 0x7FFFFFFFu;   // kDone
}

template <int MIN_BLOCKS, bool BAKE, int STACK>
__global__ void __launch_bounds__(kBlockThreads, MIN_BLOCKS)
k_rtao_rays(const __grid_constant__ FrameParams P, const __grid_constant__ SceneDev S, float* occ, const AoHit* hit_list,
            const unsigned int* hit_count, unsigned long long* work_counter, Counters* C) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t spp = P.ao_spp;
    const unsigned long long total = (unsigned long long)(*hit_count) * spp;
    const bool capped = P.use_capped != 0;
    const bool any_mode = P.ao_use_distance == 0;
    const float radius = S.radius;
    uint32_t steps = 0, isect = 0, rays = 0;

    uint32_t stk_node[STACK == 0 ? kAoStack : 1];
    float stk_t[STACK == 0 ? kAoStack : 1];
    AoStack<STACK == 0 ? 1 : STACK> pst;   // unused (and eliminated) with STACK == 0
    if (STACK != 0) pst.init();
    int sp = 0;
    uint32_t cur = kDone;
    bool exhausted = false;          // no more rays to fetch
    unsigned long long ray_id = 0;
    RayQ rq; RayBox rb;
    float best = 0.0f;
    bool found = false;
    rq.o = v3(0, 0, 0); rq.d = v3(0, 0, 1); rq.dd = 1.0f; rb = make_raybox(rq.o, rq.d);

    while (true) {
        // ---- refill idle lanes
        const bool idle = (cur == kDone) && !exhausted;
        const unsigned need = __ballot_sync(0xffffffffu, idle);
        if (need) {
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if (int(lane) == leader) base = atomicAdd(work_counter, (unsigned long long)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (idle) {
                ray_id = base + __popc(need & ((1u << lane) - 1u));
                if (ray_id >= total) exhausted = true;
                else {
                    const uint32_t slot = uint32_t(ray_id / spp), sample = uint32_t(ray_id - (unsigned long long)slot * spp);
                    Vec3 org, dir;
                    ao_ray_from_record<BAKE>(hit_list + slot, sample, spp, P.frame_number, org, dir);
                    rq = make_rayq(org, dir);
                    rb = make_raybox(org, dir);
                    best = P.ao_radius; found = false;
                    // a ray without a direction (NaN: the start frame of a hit on a zero-length segment has no tangent) hits nothing; it must
                    // not enter the traversal either -- NaN passes every slab test, including the one of an absent child
                    sp = 0; cur = ao_ray_valid(rq) ? 0u : kDone;                     // root
                    rays++;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, cur != kDone) == 0u) {
            // nobody traverses: the stream is drained, or every ray just fetched was invalid -- those are finished here (nothing hit) and the
            // warp fetches again (leaving instead would drop their results, and the rest of the stream if this is the last warp running)
            if (!exhausted && ray_id < total) { occ[ray_id] = P.ao_radius; ray_id = total; }
            if (__ballot_sync(0xffffffffu, !exhausted) == 0u) break;
            continue;
        }

        // ---- trace until too many lanes of the warp are idle again
        while (true) {
            // A: ONE inner-node step for every lane that holds an inner node.  Lanes that already hold a leaf wait; waiting
            // for the slowest lane to reach a leaf (classic while-while) left 3/4 of the lanes idle in this phase.
            if (cur != kDone && !(cur & kLeafBit)) {
                const Node64 nd = load_node(S.nodes + cur);
                steps++;
                float tl, tr;
                bool hl = box_hit(rb, nd.l0, nd.l1, 0.0f, best, tl);
                bool hr = box_hit(rb, nd.r0, nd.r1, 0.0f, best, tr);
                const uint32_t cl = __float_as_uint(nd.l0.w), cr = __float_as_uint(nd.r0.w);
                if (hl && hr) {
                    const bool swap = tr < tl;
                    if (STACK == 0) { if (sp < kAoStack) { stk_node[sp] = swap ? cl : cr; stk_t[sp] = swap ? tl : tr; sp++; } }
                    else if (sp < kAoStack) { pst.put(sp, swap ? cl : cr, swap ? tl : tr); sp++; }
                    cur = swap ? cr : cl;
                } else if (hl) cur = cl;
                else if (hr) cur = cr;
                else {
                    cur = kDone;
                    if (STACK == 0) { while (sp > 0) { --sp; if (stk_t[sp] <= best) { cur = stk_node[sp]; break; } } }
                    else cur = ao_stack_pop(pst, sp, best);
                }
            }
            // B: postponed leaves are intersected once enough lanes hold one (vote), or when nobody can step any more
            const bool at_leaf = (cur != kDone) && (cur & kLeafBit);
            const unsigned leaf_mask = __ballot_sync(0xffffffffu, at_leaf);
            const unsigned inner_mask = __ballot_sync(0xffffffffu, cur != kDone && !(cur & kLeafBit));
            if (at_leaf && (__popc(leaf_mask) >= P.ao_leaf_vote || inner_mask == 0u)) {
                const uint32_t ref = cur & kRefMask, cnt = ((cur >> 27) & 15u) + 1u;
                isect += cnt;
                bool stop = false;
                for (uint32_t i = 0; i < cnt; i++) {
                    const SegRec s = load_seg(S.segs + ref + i);
                    float t; uint32_t kind;
                    // a one-record leaf's box in its parent IS the record's own AABB (same float expressions): already tested
                    if ((cnt == 1u || seg_box_hit(rb, s, radius, 0.0f, P.ao_radius)) && capsule_hit(rq, s, radius, capped, t, kind) && t >= 0.0f && t <= P.ao_radius) {
                        if (!found || t < best) { best = t; found = true; }
                        if (any_mode) { stop = true; break; }
                    }
                }
                cur = kDone;
                if (stop) sp = 0;
                else if (STACK == 0) { while (sp > 0) { --sp; if (stk_t[sp] <= best) { cur = stk_node[sp]; break; } } }
                else cur = ao_stack_pop(pst, sp, best);
            }
            if (cur == kDone && !exhausted && ray_id < total) {
                // ray finished: traceAoRay result (:158-175)
                occ[ray_id] = found ? (any_mode ? 0.0f : best) : P.ao_radius;   // the numerator of traceAoRay's t / radius; k_rtao_reduce divides
                ray_id = total;   // written
            }
            if (__popc(__ballot_sync(0xffffffffu, cur != kDone)) < P.ao_refill_below) break;
        }
    }
    flush_counter(&C->rays_ao, rays);
    flush_counter(&C->ao_steps, steps);
    flush_counter(&C->ao_isect, isect);
}

// ------------------------------------------------------------------------------------------------
// AO ray stream with a warp-level LEAF QUEUE (b200_ao_queue, scenes built with one-record leaves).
//
// In k_rtao_rays a lane that reaches a leaf waits until ao_leaf_vote lanes hold one; ncu (profiles/r1g) shows what that costs
// once the stack traffic is out of the way: the box-step block runs with 18 of 32 lanes, the capsule test with 9, and issue
// slots are the limiter (78 % issue-active).  Here a lane does not wait: it appends (lane, record) to a 64-entry ring in
// shared memory, pops its stack and keeps traversing with the hit distance it knows so far.  As soon as 32 entries are queued,
// ALL lanes run one capsule test each -- entry i on lane i, the owner's ray fetched with 7 shuffles -- and hand the result
// back through two shared-memory words per lane (atomicMin on the hit distance bits, a pending-entry counter).  The owner
// folds the result into its `best` after the batch; until then it may walk a few nodes a tighter `best` would have culled
// (the result does not depend on that: a closest hit is a minimum over accepted candidates, any-hit is an OR).  A ray is
// finished when its stack is empty and none of its entries is pending.  A partial batch is flushed when fewer than
// ao_leaf_vote lanes can still step, so that rays waiting for their last leaves do not starve.
constexpr int kLeafQueue = 64;
constexpr uint32_t kNoHitBits = 0x7F800000u;   // +inf

// NM = node format: 0 the 64-byte child-pair nodes (Node64); 1 the 32-byte quantised child-pair nodes (NodeQ, measured slower: the
// integer-to-float conversions of its dequantisation cost more issue slots than the second load); 2 the 64-byte 4-WIDE quantised
// nodes (NodeW4): four boxes per fetch, about half the dependent node fetches per ray, up to three pushes per step.  1 and 2: capsules only.
// TOP (NM = 2 only): the first min(TOP, S.w_top) wide nodes -- whole top levels of the breadth-first tree, the nodes every ray walks --
// are staged into shared memory once per persistent block with ONE bulk-async copy (cp.async.bulk.shared::cluster.global, completion on
// an mbarrier: the TMA engine moves the bytes, no thread touches them) and steps on them are served from there.
template <int MIN_BLOCKS, bool BAKE, int STACK, int PRIM = 0, int NM = 0, int TOP = 0, bool RBUF = false>
__global__ void __launch_bounds__(kBlockThreads, MIN_BLOCKS)
k_rtao_rays_q(const __grid_constant__ FrameParams P, const __grid_constant__ SceneDev S, float* occ, const AoHit* hit_list,
              const unsigned int* hit_count, unsigned long long* work_counter, Counters* C) {
    __shared__ uint32_t s_queue[kBlockThreads / 32][kLeafQueue];   // record index | owner lane << 27
    __shared__ uint32_t s_hit[kBlockThreads / 32][32];             // nearest accepted hit (float bits) delivered to a lane by the current batch
    __shared__ int s_pend[kBlockThreads / 32][32];                 // queue entries of a lane that are not processed yet
    __shared__ float s_rbuf[RBUF ? kBlockThreads / 32 : 1][9][32];  // RBUF: the warp's batch of generated rays (origin, direction, safe 1/direction)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* queue = s_queue[warp];
    uint32_t* whit = s_hit[warp];
    int* wpend = s_pend[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t spp = P.ao_spp;
    const unsigned long long total = (unsigned long long)(*hit_count) * spp;
    const bool capped = P.use_capped != 0;
    const bool any_mode = P.ao_use_distance == 0;
    const float radius = S.radius;
    uint32_t steps = 0, isect = 0, rays = 0;

    __shared__ __align__(128) NodeW4 s_top[TOP > 0 ? TOP : 1];
    __shared__ __align__(8) unsigned long long s_mbar;
    uint32_t n_top = 0;
    if (NM == 2 && TOP > 0) {
        n_top = S.w_top < uint32_t(TOP) ? S.w_top : uint32_t(TOP);
#ifdef LV_HOST_EMU
        for (uint32_t i = threadIdx.x; i < n_top * 16u; i += kBlockThreads) reinterpret_cast<uint32_t*>(s_top)[i] = reinterpret_cast<const uint32_t*>(S.wnodes)[i];
        __syncthreads();
#else
        const uint32_t bar = uint32_t(__cvta_generic_to_shared(&s_mbar)), dst = uint32_t(__cvta_generic_to_shared(s_top));
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (n_top) {
            if (threadIdx.x == 0) {
                const uint32_t bytes = n_top * 64u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(S.wnodes), "r"(bytes), "r"(bar) : "memory");
            }
            uint32_t done = 0;   // every thread waits for phase 0 of the barrier: the copy's bytes have landed
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
        }
#endif
    }
    constexpr bool QN = NM != 0;                       // quantised boxes: the record's exact AABB is tested in the leaf batch
    constexpr int kCap = NM == 2 ? kAoStackWide : kAoStack;
    AoStack<STACK, kCap> pst;
    pst.init();
    int sp = 0;
    uint32_t cur = kDone;
    bool exhausted = false, has_ray = false;
    unsigned long long ray_id = 0;
    RayQ rq; RayBox rb;
    float best = 0.0f;
    bool found = false;
    rq.o = v3(0, 0, 0); rq.d = v3(0, 0, 1); rq.dd = 1.0f; rb = make_raybox(rq.o, rq.d);
    uint32_t q_head = 0, q_count = 0;   // uniform over the warp
    uint32_t w4nx = kW4SelLo, w4ny = kW4SelLo, w4nz = kW4SelLo;
    wpend[lane] = 0;
    __syncwarp();

    // Ray generation is decoupled from the refill (RBUF): the whole warp generates the next 32 rays of the stream together -- TEA seed,
    // LCG, hemisphere sample, the three IEEE reciprocals of the slab test: ~600 instructions per ray, which the lanes that happened to
    // be idle used to execute alone (10 of 32 lanes, 8 % of the kernel's warp instructions) -- into a shared-memory batch; a lane
    // without a ray then just takes the next entry.  Refilling is cheap that way, so it can happen as soon as lanes are idle.
    uint32_t rb_next = 32u;                // next unread entry of the warp's batch (32 = empty); uniform over the warp
    unsigned long long rb_base = 0;        // ray number of entry 0
    bool stream_done = false;              // the last batch reached the end of the stream
    float (*rbuf)[32] = s_rbuf[RBUF ? warp : 0];

    while (true) {
        // ---- refill lanes without a ray
        unsigned need = __ballot_sync(0xffffffffu, !has_ray && !exhausted);
        while (need) {
            if (RBUF) {
                if (rb_next >= 32u) {
                    if (stream_done) { if (!has_ray) exhausted = true; break; }
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(work_counter, 32ull);
                    rb_base = __shfl_sync(0xffffffffu, base, 0);
                    rb_next = 0u;
                    const unsigned long long id = rb_base + lane;
                    Vec3 org = v3(0, 0, 0), dir = v3(0, 0, 1);
                    if (id < total) {
                        const uint32_t slot = uint32_t(id / spp), sample = uint32_t(id - (unsigned long long)slot * spp);
                        ao_ray_from_record<BAKE>(hit_list + slot, sample, spp, P.frame_number, org, dir);
                    }
                    const RayBox gb = make_raybox(org, dir);
                    rbuf[0][lane] = org.x; rbuf[1][lane] = org.y; rbuf[2][lane] = org.z;
                    rbuf[3][lane] = dir.x; rbuf[4][lane] = dir.y; rbuf[5][lane] = dir.z;
                    rbuf[6][lane] = gb.ix; rbuf[7][lane] = gb.iy; rbuf[8][lane] = gb.iz;
                    if (rb_base + 32ull >= total) stream_done = true;
                    __syncwarp();
                }
                const uint32_t avail = 32u - rb_next, rank = __popc(need & lt_mask);
                if (((need >> lane) & 1u) && rank < avail) {
                    const uint32_t e = rb_next + rank;
                    ray_id = rb_base + e;
                    if (ray_id >= total) exhausted = true;
                    else {
                        const Vec3 org = v3(rbuf[0][e], rbuf[1][e], rbuf[2][e]), dir = v3(rbuf[3][e], rbuf[4][e], rbuf[5][e]);
                        rq = make_rayq(org, dir);
                        rb.ix = rbuf[6][e]; rb.iy = rbuf[7][e]; rb.iz = rbuf[8][e];
                        rb.cx = -(org.x * rb.ix); rb.cy = -(org.y * rb.iy); rb.cz = -(org.z * rb.iz);   // as in make_raybox
                        if (NM == 2) {   // which bound of each axis the ray enters first (rb.i* is never 0 or NaN for a ray that is traversed)
                            w4nx = rb.ix >= 0.0f ? kW4SelLo : kW4SelHi; w4ny = rb.iy >= 0.0f ? kW4SelLo : kW4SelHi; w4nz = rb.iz >= 0.0f ? kW4SelLo : kW4SelHi;
                        }
                        best = P.ao_radius; found = false;
                        sp = 0; cur = ao_ray_valid(rq) ? 0u : kDone;                 // root (a NaN ray hits nothing and must not be traversed, see k_rtao_rays)
                        has_ray = true;
                        rays++;
                    }
                }
                rb_next += min(uint32_t(__popc(need)), avail);
                __syncwarp();   // the entries have been read before the next batch overwrites them
                need = __ballot_sync(0xffffffffu, !has_ray && !exhausted);
                continue;
            }
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if (int(lane) == leader) base = atomicAdd(work_counter, (unsigned long long)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if ((need >> lane) & 1u) {
                ray_id = base + __popc(need & lt_mask);
                if (ray_id >= total) exhausted = true;
                else {
                    const uint32_t slot = uint32_t(ray_id / spp), sample = uint32_t(ray_id - (unsigned long long)slot * spp);
                    Vec3 org, dir;
                    ao_ray_from_record<BAKE>(hit_list + slot, sample, spp, P.frame_number, org, dir);
                    rq = make_rayq(org, dir);
                    rb = make_raybox(org, dir);
                    if (NM == 2) {   // which bound of each axis the ray enters first (rb.i* is never 0 or NaN for a ray that is traversed)
                        w4nx = rb.ix >= 0.0f ? kW4SelLo : kW4SelHi; w4ny = rb.iy >= 0.0f ? kW4SelLo : kW4SelHi; w4nz = rb.iz >= 0.0f ? kW4SelLo : kW4SelHi;
                    }
                    best = P.ao_radius; found = false;
                    sp = 0; cur = ao_ray_valid(rq) ? 0u : kDone;                     // root (a NaN ray hits nothing and must not be traversed, see k_rtao_rays)
                    has_ray = true;
                    rays++;
                }
            }
            break;   // one pass: every idle lane has fetched a ray number
        }
        if (__ballot_sync(0xffffffffu, has_ray) == 0u) break;

        while (true) {
            // A: one inner-node step for every lane that holds an inner node
            if (NM == 2) {
              // ao_wide_reps steps per pass of the loop: the queue / refill bookkeeping below is paid once per pass
              for (int rep = 0; rep < P.ao_wide_reps; rep++)
              if (cur != kDone && !(cur & kLeafBit)) {
                steps++;
                float4 ha, hb2, hc, hd;
                if (TOP > 0 && cur < n_top) {                                          // a top-level node: from the staged copy
                    const float4* tp = reinterpret_cast<const float4*>(s_top + cur);
                    ha = tp[0]; hb2 = tp[1]; hc = tp[2]; hd = tp[3];
                } else {
                    ldg256(S.wnodes + cur, ha, hb2);                                   // children 0, 1
                    ldg256(reinterpret_cast<const char*>(S.wnodes + cur) + 32, hc, hd);   // children 2, 3
                }
                const float ox = S.w_origin[0], oy = S.w_origin[1], oz = S.w_origin[2], sx = S.w_scale[0], sy = S.w_scale[1], sz = S.w_scale[2];
                // Per axis the ray's direction sign says which bound is entered first: t = fma(b, inv, c) is monotone in b, so
                // min(t(lo), t(hi)) = t(near bound) bit for bit -- the canonical slab test without its six min / max, and the
                // selection costs nothing because the byte-permute that builds the float does it (w4n*: PRMT selectors of the ray).
                const uint32_t fx = w4nx ^ 0x0022u, fy = w4ny ^ 0x0022u, fz = w4nz ^ 0x0022u;
                const float kInf = __int_as_float(0x7f800000);
#define LV_W4_CHILD(T, WX, WY, WZ, CW) do { \
                    const uint32_t wx = __float_as_uint(WX), wy = __float_as_uint(WY), wz = __float_as_uint(WZ); \
                    const float nx = __fmaf_rn(w4_dequant_packed(wx, w4nx, sx, ox), rb.ix, rb.cx), gx = __fmaf_rn(w4_dequant_packed(wx, fx, sx, ox), rb.ix, rb.cx); \
                    const float ny = __fmaf_rn(w4_dequant_packed(wy, w4ny, sy, oy), rb.iy, rb.cy), gy = __fmaf_rn(w4_dequant_packed(wy, fy, sy, oy), rb.iy, rb.cy); \
                    const float nz = __fmaf_rn(w4_dequant_packed(wz, w4nz, sz, oz), rb.iz, rb.cz), gz = __fmaf_rn(w4_dequant_packed(wz, fz, sz, oz), rb.iz, rb.cz); \
                    const float lo = fmaxf(fmaxf(nx, ny), fmaxf(nz, 0.0f)), hi = fminf(fminf(gx, gy), fminf(gz, best)); \
                    T = (lo <= hi) & (__float_as_uint(CW) != kAbsentChild) ? lo : kInf; } while (0)
                float t0, t1, t2, t3;
                LV_W4_CHILD(t0, ha.x, ha.y, ha.z, hd.x);
                LV_W4_CHILD(t1, ha.w, hb2.x, hb2.y, hd.y);
                LV_W4_CHILD(t2, hb2.z, hb2.w, hc.x, hd.z);
                LV_W4_CHILD(t3, hc.y, hc.z, hc.w, hd.w);
#undef LV_W4_CHILD
                const uint32_t c0 = __float_as_uint(hd.x), c1 = __float_as_uint(hd.y), c2 = __float_as_uint(hd.z), c3 = __float_as_uint(hd.w);
                // descend into the nearest hit child, push the other hit children (entries carry their entry distance and are culled
                // against the closest hit when popped, so their order only matters for speed)
                int near = 0; float tnear = t0; uint32_t cnear = c0;
                if (t1 < tnear) { tnear = t1; cnear = c1; near = 1; }
                if (t2 < tnear) { tnear = t2; cnear = c2; near = 2; }
                if (t3 < tnear) { tnear = t3; cnear = c3; near = 3; }
                if (tnear == kInf) cur = ao_stack_pop(pst, sp, best);
                else {
                    if (near != 0 && t0 != kInf && sp < kCap) { pst.put(sp, c0, t0); sp++; }
                    if (near != 1 && t1 != kInf && sp < kCap) { pst.put(sp, c1, t1); sp++; }
                    if (near != 2 && t2 != kInf && sp < kCap) { pst.put(sp, c2, t2); sp++; }
                    if (near != 3 && t3 != kInf && sp < kCap) { pst.put(sp, c3, t3); sp++; }
                    cur = cnear;
                }
              }
            } else if (cur != kDone && !(cur & kLeafBit)) {
                steps++;
                float tl, tr;
                bool hl, hr;
                uint32_t cl, cr;
                if (NM == 1) {
                    float4 qa, qb;
                    ldg256(S.qnodes + cur, qa, qb);   // the whole node in one request
                    const uint32_t w0 = __float_as_uint(qa.x), w1 = __float_as_uint(qa.y), w2 = __float_as_uint(qa.z);
                    const uint32_t w4 = __float_as_uint(qb.x), w5 = __float_as_uint(qb.y), w6 = __float_as_uint(qb.z);
                    cl = __float_as_uint(qa.w); cr = __float_as_uint(qb.w);
                    const float ox = S.q_origin[0], oy = S.q_origin[1], oz = S.q_origin[2], sx = S.q_scale[0], sy = S.q_scale[1], sz = S.q_scale[2];
                    hl = cl != kAbsentChild &&
                         box_hit(rb, __fmaf_rn(float(w0 & 0xffffu), sx, ox), __fmaf_rn(float(w0 >> 16), sy, oy), __fmaf_rn(float(w1 & 0xffffu), sz, oz),
                                 __fmaf_rn(float(w1 >> 16), sx, ox), __fmaf_rn(float(w2 & 0xffffu), sy, oy), __fmaf_rn(float(w2 >> 16), sz, oz), 0.0f, best, tl);
                    hr = cr != kAbsentChild &&
                         box_hit(rb, __fmaf_rn(float(w4 & 0xffffu), sx, ox), __fmaf_rn(float(w4 >> 16), sy, oy), __fmaf_rn(float(w5 & 0xffffu), sz, oz),
                                 __fmaf_rn(float(w5 >> 16), sx, ox), __fmaf_rn(float(w6 & 0xffffu), sy, oy), __fmaf_rn(float(w6 >> 16), sz, oz), 0.0f, best, tr);
                } else {
                    const Node64 nd = load_node((PRIM == 1 ? S.tri_nodes : S.nodes) + cur);
                    hl = box_hit(rb, nd.l0, nd.l1, 0.0f, best, tl);
                    hr = box_hit(rb, nd.r0, nd.r1, 0.0f, best, tr);
                    cl = __float_as_uint(nd.l0.w); cr = __float_as_uint(nd.r0.w);
                }
                if (hl && hr) {
                    const bool swap = tr < tl;
                    if (sp < kAoStack) { pst.put(sp, swap ? cl : cr, swap ? tl : tr); sp++; }
                    cur = swap ? cr : cl;
                } else if (hl) cur = cl;
                else if (hr) cur = cr;
                else cur = ao_stack_pop(pst, sp, best);
            }
            // E: lanes that hold a leaf queue its record and go on with their stack
            const bool at_leaf = (cur & kLeafBit) != 0u;   // kDone has bit 31 clear
            const unsigned leaf_mask = __ballot_sync(0xffffffffu, at_leaf);
            if (leaf_mask) {
                if (at_leaf) {
                    queue[(q_head + q_count + __popc(leaf_mask & lt_mask)) & (kLeafQueue - 1)] = (cur & kRefMask) | (lane << 27);
                    wpend[lane] += 1;
                    cur = ao_stack_pop(pst, sp, best);
                }
                q_count += __popc(leaf_mask);
                __syncwarp();
            }
            // B: a full batch of 32 capsule tests, or a partial one when too few lanes can still step
            const unsigned movable = __ballot_sync(0xffffffffu, cur != kDone);
            if (q_count >= 32u || (q_count != 0u && __popc(movable) < P.ao_leaf_vote)) {
                const uint32_t n = q_count < 32u ? q_count : 32u;
                whit[lane] = kNoHitBits;
                __syncwarp();
                const uint32_t e = queue[(q_head + (lane < n ? lane : 0u)) & (kLeafQueue - 1)];
                const uint32_t owner = e >> 27;
                RayQ r2;
                r2.o.x = __shfl_sync(0xffffffffu, rq.o.x, owner); r2.o.y = __shfl_sync(0xffffffffu, rq.o.y, owner); r2.o.z = __shfl_sync(0xffffffffu, rq.o.z, owner);
                r2.d.x = __shfl_sync(0xffffffffu, rq.d.x, owner); r2.d.y = __shfl_sync(0xffffffffu, rq.d.y, owner); r2.d.z = __shfl_sync(0xffffffffu, rq.d.z, owner);
                r2.dd = __shfl_sync(0xffffffffu, rq.dd, owner);
                RayBox rb2;
                if (QN) {   // quantised node boxes are looser than the record's own AABB: the acceptance rule's slab test is done here, with the owner's ray
                    rb2.ix = __shfl_sync(0xffffffffu, rb.ix, owner); rb2.iy = __shfl_sync(0xffffffffu, rb.iy, owner); rb2.iz = __shfl_sync(0xffffffffu, rb.iz, owner);
                    rb2.cx = __shfl_sync(0xffffffffu, rb.cx, owner); rb2.cy = __shfl_sync(0xffffffffu, rb.cy, owner); rb2.cz = __shfl_sync(0xffffffffu, rb.cz, owner);
                }
                if (lane < n) {
                    isect++;
                    float t;
                    bool accepted;
                    // the record's own AABB was the child box tested in its parent (one-record leaves): the acceptance rule's slab test is done
                    if (PRIM == 1) {
                        float u, v;
                        accepted = tri_hit(r2.o, r2.d, load_tri(S.tris + (e & kRefMask)), 0.0f, P.ao_radius, t, u, v);
                    } else {
                        const SegRec s = load_seg(S.segs + (e & kRefMask));
                        uint32_t kind;
                        accepted = (!QN || seg_box_hit(rb2, s, radius, 0.0f, P.ao_radius)) && capsule_hit(r2, s, radius, capped, t, kind) && t >= 0.0f && t <= P.ao_radius;
                    }
                    if (accepted) atomicMin(&whit[owner], __float_as_uint(t));
                    atomicSub(&wpend[owner], 1);
                }
                __syncwarp();
                const uint32_t hb = whit[lane];
                if (hb != kNoHitBits) {
                    const float t = __uint_as_float(hb);
                    if (!found || t < best) { best = t; found = true; }
                    if (any_mode) { sp = 0; cur = kDone; }
                }
                q_head = (q_head + n) & (kLeafQueue - 1);
                q_count -= n;
            }
            // a ray is finished when its stack is empty and none of its queue entries is pending: traceAoRay result (:158-175)
            if (has_ray && cur == kDone && wpend[lane] == 0) {
                occ[ray_id] = found ? (any_mode ? 0.0f : best) : P.ao_radius;   // the numerator of traceAoRay's t / radius; k_rtao_reduce divides
                has_ray = false;
            }
            if (__popc(__ballot_sync(0xffffffffu, has_ray)) < P.ao_refill_below) break;
        }
    }
    flush_counter(&C->rays_ao, rays);
    flush_counter(&C->ao_steps, steps);
    flush_counter(&C->ao_isect, isect);
}

}  // namespace lv
#include "lv_aostream.cuh"
namespace lv {

__global__ void k_rtao_reduce(const __grid_constant__ FrameParams P, const float* occ, const AoHit* hit_list,
                              const unsigned int* hit_count, float* ao) {
    const uint32_t n_hit = *hit_count;
    const uint32_t spp = P.ao_spp, spl = P.ao_spp_local;
    // AO-sample-batch shards (spl < spp): `occ` holds spp / spl parts, part j = samples [j spl, (j + 1) spl) of every record, traced by rank j
    const size_t part_stride = size_t(n_hit) * spl;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_hit; slot += gridDim.x * blockDim.x) {
        const uint32_t pixel = __float_as_uint(hit_list[slot].nrm_px.w);
        const float* q = occ + size_t(slot) * spl;
        // the stream stores hit distances (radius for a miss, 0 for an any-hit): t / radius (:170) is taken here, coalesced and by all
        // lanes, instead of by the few lanes of a warp whose ray has just finished.  Summed in sample order, whoever traced the sample.
        float sum = 0.0f;
        for (uint32_t j = 0; j * spl < spp; j++)
            for (uint32_t i = 0; i < spl; i++) sum += q[j * part_stride + i] / P.ao_radius;
        float v = sum / float(spp);
        float* p = ao + pixel;
        if (P.frame_number != 0) v = mixf_(*p, v, 1.0f / float(P.frame_number + 1));
        *p = v;
    }
}

// ------------------------------------------------------------------------------------------------
// PPLL.  addrGen: reference Data/Shaders/Utils/TiledAddress.glsl:53-85.
__device__ __forceinline__ uint32_t addr_gen(const FrameParams& P, uint32_t x, uint32_t y) {
    if (P.addr_tw == 1 && P.addr_th == 1) return x + P.padded_w * y;
    // tile sizes are powers of two (checked by lv_set_option): the divisions of the reference are shifts
    const uint32_t sw = P.padded_w >> P.addr_tw_log2;
    const uint32_t tx = x >> P.addr_tw_log2, ty = y >> P.addr_th_log2;
    const uint32_t base = (tx + sw * ty) << (P.addr_tw_log2 + P.addr_th_log2);
    return base | ((x & (P.addr_tw - 1)) + ((y & (P.addr_th - 1)) << P.addr_tw_log2));
}

// S9 gather.  The fragment source is the all-hits enumeration of the pixel-centre ray: every accepted candidate
// in [1e-4, 1000] is shaded (S3) and appended.  One thread owns one pixel, so the list head lives in a register and is
// written once (same final startOffset/next structure as the reference's atomicExchange chain); the global fragment
// counter is bumped once per converged warp group.
//
// Traversal is a WARP PACKET: the 32 rays of an 8x4 pixel patch differ by a few tube radii, so the warp walks the BVH
// together with one shared stack -- a node is visited if any lane's box test hits, every node / record is fetched once per
// warp (broadcast load), and there is no divergence between box work and leaf work.  Accepted hits are queued per lane in
// shared memory and shaded in lockstep rounds (all lanes that have a queued hit shade one), because shading is by far
// the longest divergent section.  First version (one independent traversal per thread, shading inside the leaf loop):
// 4.5 of 32 lanes active, 250 ms on config 4 (profiles/r1a).
constexpr int kGatherQueue = 16;   // queued hits per lane; >= the largest leaf (b200_bvh_leaf_size <= 16)

struct GatherState { uint32_t head, stored, gen; };

// Shade the queued hits of all lanes in lockstep rounds, then allocate the surviving fragments of the whole warp with ONE
// atomicAdd (warp scan of the per-lane counts) so that each pixel's new nodes are CONTIGUOUS in the fragment buffer: the
// resolve pass then walks runs of up to kGatherQueue adjacent 12-byte nodes instead of one 32-byte sector per node.
template <bool SAO>
__device__ __forceinline__ void gather_flush(const FrameParams& P, const SceneDev& S, Vec3 ro, Vec3 rd, uint32_t lane,
                                             uint2 (*q)[kBlockThreads], uint32_t& qn, GatherState& g, lv_ppll_node* nodes,
                                             unsigned long long* frag_counter, unsigned long long list_size) {
    uint32_t kept = 0;
    for (uint32_t i = 0; __ballot_sync(0xffffffffu, i < qn); i++) {
        if (i < qn) {
            const uint2 e = q[i][threadIdx.x];
            const SegRec s = load_seg(S.segs + (e.x & kRefMask));
            const Shaded sh = shade_hit<SAO>(P, ro, rd, __uint_as_float(e.y), e.x >> 28, s, SAO && S.seg_aux ? S.seg_aux + (e.x & kRefMask) : nullptr);
            if (!(sh.color.w < 0.001f)) {                                       // LinkedListGather.glsl:38
                q[kept][threadIdx.x] = make_uint2(pack_unorm4x8(sh.color), __float_as_uint(sh.hit_t));   // kept <= i: slot is free
                kept++;
            }
        }
    }
    qn = 0;
    uint32_t incl = kept;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (int(lane) >= o) incl += v; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0u) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(frag_counter, (unsigned long long)total);   // fragCounter, LinkedListGather.glsl:55
    base = __shfl_sync(0xffffffffu, base, 0) + (incl - kept);
    g.gen += kept;
    for (uint32_t j = 0; j < kept; j++) {
        const unsigned long long idx = base + j;
        if (idx < list_size) {                                                   // :57
            const uint2 e = q[j][threadIdx.x];
            lv_ppll_node nd; nd.color = e.x; nd.depth = __uint_as_float(e.y); nd.next = g.head;
            nodes[idx] = nd;
            g.head = uint32_t(idx);
            g.stored++;
        }
    }
}

template <bool SAO>
__global__ void __launch_bounds__(kBlockThreads)
k_ppll_gather(const __grid_constant__ FrameParams P, const __grid_constant__ SceneDev S, uint32_t* heads, uint32_t* counts,
              lv_ppll_node* nodes, unsigned long long* frag_counter, unsigned long long list_size, Counters* C) {
    __shared__ uint32_t s_stack[kBlockThreads / 32][kStackSize];
    __shared__ uint2 s_queue[kGatherQueue][kBlockThreads];   // (record | kind << 28, t bits); [slot][thread] is conflict-free
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* stack = s_stack[warp];
    uint32_t x, y;
    const bool valid = thread_pixel(P, x, y);
    uint32_t steps = 0, isect = 0;
    // the pixel's list so far (lv_ppll_clear: empty): a second gather between two clears APPENDS, like the reference's atomicExchange chain
    GatherState g; g.head = kNone; g.stored = 0; g.gen = 0;
    uint32_t list_addr = 0;
    if (valid) { list_addr = addr_gen(P, x, y); g.head = heads[list_addr]; g.stored = counts[list_addr]; }
    Vec3 ro = v3(0, 0, 0), rd = v3(0, 0, 1);
    if (valid) camera_ray(P, x, y, 0.5f, 0.5f, ro, rd);
    const RayQ rq = make_rayq(ro, rd);
    const RayBox rb = make_raybox(ro, rd);
    const float tmin = 0.0001f, tmax = 1000.0f;
    const bool capped = P.use_capped != 0;
    uint32_t qn = 0;
    if (S.n_seg != 0 && __ballot_sync(0xffffffffu, valid)) {
        uint32_t node = 0;
        int sp = 0;
        while (true) {
            const Node64 nd = load_node(S.nodes + node);      // same address in every lane: one broadcast fetch per warp
            steps += (lane == 0);
            float tn;
            const bool hl = valid && box_hit(rb, nd.l0, nd.l1, tmin, tmax, tn);
            const bool hr = valid && box_hit(rb, nd.r0, nd.r1, tmin, tmax, tn);
            const uint32_t cw[2] = {__float_as_uint(nd.l0.w), __float_as_uint(nd.r0.w)};
            const bool any[2] = {__ballot_sync(0xffffffffu, hl) != 0u, __ballot_sync(0xffffffffu, hr) != 0u};
            uint32_t inner[2]; int n_inner = 0;
#pragma unroll
            for (int side = 0; side < 2; side++) {
                if (!any[side]) continue;
                const uint32_t w = cw[side];
                if (w & kLeafBit) {
                    const uint32_t ref = w & kRefMask, cnt = ((w >> 27) & 15u) + 1u;
                    isect += (lane == 0) ? cnt : 0u;
                    if (__ballot_sync(0xffffffffu, qn + cnt > uint32_t(kGatherQueue)))
                        gather_flush<SAO>(P, S, ro, rd, lane, s_queue, qn, g, nodes, frag_counter, list_size);
                    const bool mine = side ? hr : hl;
                    for (uint32_t i = 0; i < cnt; i++) {
                        const SegRec s = load_seg(S.segs + ref + i);
                        float t; uint32_t kind;
                        if (mine && seg_box_hit(rb, s, S.radius, tmin, tmax) && capsule_hit(rq, s, S.radius, capped, t, kind) && t >= tmin && t <= tmax) {
                            s_queue[qn][threadIdx.x] = make_uint2((ref + i) | (kind << 28), __float_as_uint(t));
                            qn++;
                        }
                    }
                } else inner[n_inner++] = w;
            }
            if (n_inner == 2) { if (lane == 0) stack[sp] = inner[1]; sp++; node = inner[0]; }
            else if (n_inner == 1) node = inner[0];
            else {
                if (sp == 0) break;
                --sp;
                __syncwarp();
                node = stack[sp];
            }
            __syncwarp();
        }
        gather_flush<SAO>(P, S, ro, rd, lane, s_queue, qn, g, nodes, frag_counter, list_size);
    }
    if (valid) {
        heads[list_addr] = g.head;
        counts[list_addr] = g.stored;
    }
    flush_counter(&C->rays_primary, valid ? 1 : 0);
    flush_counter(&C->steps, steps);
    flush_counter(&C->isect, isect);
    flush_counter(&C->frags_generated, g.gen);
}

// Per-pixel camera ray of the pixel centre, computed once per frame for the object-order gather: [2 * pixel] = (rd.xyz, |rd|^2),
// [2 * pixel + 1] = (safe 1/rd).  The same expressions as camera_ray / make_rayq / make_raybox, evaluated once per pixel instead of
// once per (pixel, segment) candidate.
__global__ void __launch_bounds__(kBlockThreads)
k_pixel_rays(const __grid_constant__ FrameParams P, float4* rays) {
    uint32_t x, y;
    if (!thread_pixel(P, x, y)) return;
    Vec3 ro, rd;
    camera_ray(P, x, y, 0.5f, 0.5f, ro, rd);
    const RayQ rq = make_rayq(ro, rd);
    const RayBox rb = make_raybox(ro, rd);
    float4* o = rays + 2 * (size_t(y) * P.W + x);
    o[0] = make_float4(rd.x, rd.y, rd.z, rq.dd);
    o[1] = make_float4(rb.ix, rb.iy, rb.iz, 0.0f);
}

// S9 gather, OBJECT ORDER (b200_ppll_gather_mode = raster, the default; like the reference, whose gather pass is a rasterisation of the tube
// geometry, LinkedListGather.glsl).  One warp takes one segment: a conservative screen rectangle of the capsule (the hull of the two
// end spheres' projected bounding boxes, padded by a pixel) thinned out by the 2-D distance to the projected axis; the surviving
// candidate pixels are tested 32 at a time with EXACTLY the predicates of the ray-cast gather -- the pixel-centre camera ray (read from
// the per-pixel ray table), the record's own-AABB slab test, capsule_hit, t in [1e-4, 1000] -- so the set of fragments per pixel is the
// same, bit for bit; only the (race-dependent, in the reference too) order inside a list differs.  Hits are queued again and SHADED 32
// at a time, so that both expensive stages run with full warps (measured on config 4 before the second queue: 30 of 32 lanes in the
// test, 21.5 in the shading).  Everything that depends on the segment alone -- its AABB, cylinder axis, the tangent normalisations of
// the shader -- is computed once per segment (warp-uniform), not once per fragment.  Fragments are appended the reference's way: one
// counter bump per warp round, atomicExch on the pixel's head.  No BVH is involved; the work is proportional to the screen area of the
// tubes.  Tile-sharded frames: candidates outside this rank's tiles are skipped (owned_tiles: one byte per tile of the frame).
// STAGE (b200_ppll_gather_mode = raster_contiguous): fragments go to a staging array as (colour, depth, pixel address) and only the
// per-pixel counts are bumped; an exclusive scan of the counts and k_ppll_fill then place every pixel's fragments in ONE contiguous
// run of the node buffer (next = previous slot, head = last slot: still the reference's linked structure), which the resolve pass
// reads by index instead of chasing pointers (k_ppll_resolve<.., CONTIG>).
constexpr int kRasterQueue = 64;
template <bool SAO, bool STAGE, int MIN_BLOCKS>
__global__ void __launch_bounds__(kBlockThreads, MIN_BLOCKS)
k_ppll_gather_raster(const __grid_constant__ FrameParams P, const __grid_constant__ SceneDev S, uint32_t* heads, uint32_t* counts,
                     lv_ppll_node* nodes, unsigned long long* frag_counter, unsigned long long list_size, Counters* C,
                     unsigned long long* work_counter, const unsigned char* owned_tiles, uint32_t tiles_x, unsigned long long n_pixels,
                     const float4* pixel_rays) {
    __shared__ uint32_t s_cand[kBlockThreads / 32][kRasterQueue];   // per warp: queued candidate pixels (x | y << 16)
    __shared__ uint32_t s_hpix[kBlockThreads / 32][kRasterQueue];   // per warp: queued hits -- pixel,
    __shared__ uint32_t s_ht[kBlockThreads / 32][kRasterQueue];     //   t bits,
    __shared__ uint32_t s_hkind[kBlockThreads / 32][kRasterQueue];  //   hit kind
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* queue = s_cand[warp];
    uint32_t* hpix = s_hpix[warp]; uint32_t* ht = s_ht[warp]; uint32_t* hkind = s_hkind[warp];
    const float tmin = 0.0001f, tmax = 1000.0f;
    const bool capped = P.use_capped != 0;
    uint32_t isect = 0, gen = 0;
    Vec3 ro;
    { Vec4 o = mat_mul(P.inv_view, v4(0.0f, 0.0f, 0.0f, 1.0f)); ro = v3(o.x, o.y, o.z); }   // camera_ray's origin
    // radius of the end spheres in view space: scaled by the largest column of the view matrix's 3x3 (1 for a rigid camera), a little generous
    float sc2 = 0.0f;
#pragma unroll
    for (int cI = 0; cI < 3; cI++) sc2 = fmaxf(sc2, P.view[4 * cI] * P.view[4 * cI] + P.view[4 * cI + 1] * P.view[4 * cI + 1] + P.view[4 * cI + 2] * P.view[4 * cI + 2]);
    const float rv = S.radius * sqrtf(sc2) * 1.001f;
    // how much the clip coordinates change over one radius (w: 1 x rv for a perspective matrix), and the clip w below which a point counts as near
    const float dw = rv * (fabsf(P.proj[3]) + fabsf(P.proj[7]) + fabsf(P.proj[11]));
    const float dcx = rv * (fabsf(P.proj[0]) + fabsf(P.proj[4]) + fabsf(P.proj[8])), dcy = rv * (fabsf(P.proj[1]) + fabsf(P.proj[5]) + fabsf(P.proj[9]));
    const float thr = 4.0f * dw + 1e-6f;
    const bool persp = P.proj[3] == 0.0f && P.proj[7] == 0.0f && P.proj[11] == -1.0f && P.proj[15] == 0.0f && P.proj[1] == 0.0f && P.proj[4] == 0.0f;
    const float focal_px = fmaxf(fabsf(P.proj[0]) * 0.5f * float(P.W), fabsf(P.proj[5]) * 0.5f * float(P.H));
    // persistent warps fetch chunks of kRasterChunk consecutive records (Morton order: neighbours in space) from a global counter --
    // the screen area of a segment, i.e. its cost, varies by orders of magnitude
    constexpr uint32_t kRasterChunk = 8;
    while (true) {
      unsigned long long first = 0;
      if (lane == 0) first = atomicAdd(work_counter, (unsigned long long)kRasterChunk);
      first = __shfl_sync(0xffffffffu, first, 0);
      if (first >= S.n_seg) break;
      const uint32_t last = uint32_t(first + kRasterChunk < S.n_seg ? first + kRasterChunk : S.n_seg);
      for (uint32_t seg = uint32_t(first); seg < last; seg++) {
        const SegRec s = load_seg(S.segs + seg);          // same address in every lane: one broadcast fetch
        // view-space end points; clip coordinates are linear along the axis
        Vec4 va = mat_mul(P.view, v4(s.a.x, s.a.y, s.a.z, 1.0f)), vb = mat_mul(P.view, v4(s.b.x, s.b.y, s.b.z, 1.0f));
        const Vec4 ca = mat_mul(P.proj, va), cb = mat_mul(P.proj, vb);
        if (fmaxf(ca.w, cb.w) + dw <= 0.0f) continue;       // the whole capsule is behind the eye plane: no forward ray reaches it
        bool full_frame = false;
        if (fminf(ca.w, cb.w) < thr) {
            // Near the eye plane the projection has no bound.  Split the axis at clip w = thr: the FRONT part (w >= thr) projects to a
            // bounded rectangle; the NEAR part (its reachable points have 0 < w <= thr + dw) can only be on screen -- |clip x| <= w and
            // |clip y| <= w -- if its axis comes within that distance (+ a radius) of the view axis.  If it does: whole frame.
            const bool a_front = ca.w >= cb.w;
            const Vec4 vF = a_front ? va : vb, vN = a_front ? vb : va, cF = a_front ? ca : cb, cN = a_front ? cb : ca;
            const float span = cF.w - cN.w;                                               // >= 0
            const float ts = (cF.w > thr && span > 0.0f) ? (cF.w - thr) / span : 0.0f;     // S: where the near part starts (w = thr, or F)
            const float te = (cN.w < -dw && span > 0.0f) ? (cF.w + dw) / span : 1.0f;      // E: where it stops being reachable (w = -dw, or N)
            const float sx = cF.x + ts * (cN.x - cF.x), sy = cF.y + ts * (cN.y - cF.y), e_x = cF.x + te * (cN.x - cF.x), e_y = cF.y + te * (cN.y - cF.y);
            const float lim = thr + dw;
            const float near_x = ((sx < 0.0f) != (e_x < 0.0f)) ? 0.0f : fminf(fabsf(sx), fabsf(e_x));
            const float near_y = ((sy < 0.0f) != (e_y < 0.0f)) ? 0.0f : fminf(fabsf(sy), fabsf(e_y));
            const bool near_off_screen = near_x > lim + dcx || near_y > lim + dcy;        // a NaN anywhere: not provably off screen
            if (!near_off_screen) full_frame = true;
            else if (!(cF.w > thr)) continue;                                             // no front part, near part off screen
            else {                                                                        // bound the front part F..S only
                const Vec4 vS = v4(vF.x + ts * (vN.x - vF.x), vF.y + ts * (vN.y - vF.y), vF.z + ts * (vN.z - vF.z), 1.0f);
                va = vF; vb = vS;
            }
        }
        // lanes 0..15: the 8 corners of the view-space bounding cube of each end sphere, projected
        const Vec4 vp = ((lane >> 3) & 1u) ? vb : va;
        const Vec4 cl = mat_mul(P.proj, v4(vp.x + ((lane & 1u) ? rv : -rv), vp.y + ((lane & 2u) ? rv : -rv), vp.z + ((lane & 4u) ? rv : -rv), 1.0f));
        float fx = (cl.x / cl.w * 0.5f + 0.5f) * float(P.W) - 0.5f;       // continuous pixel coordinate: pixel px has its centre at fx = px
        float fy = (cl.y / cl.w * 0.5f + 0.5f) * float(P.H) - 0.5f;
        const bool bad = !(cl.w > 1e-6f) || !(fx == fx) || !(fy == fy);     // at / behind the eye plane, or not a number: no bound
        // the projected sphere centre of this lane's end point and the distance of this lane's corner from it: the projected capsule lies
        // within max(distance) of the projected axis (a straight 2-D segment), which culls most of a diagonal segment's rectangle below
        const Vec4 cc = mat_mul(P.proj, vp);
        const float ccx = (cc.x / cc.w * 0.5f + 0.5f) * float(P.W) - 0.5f, ccy = (cc.y / cc.w * 0.5f + 0.5f) * float(P.H) - 0.5f;
        float rad = sqrtf((fx - ccx) * (fx - ccx) + (fy - ccy) * (fy - ccy));
        if (persp && cc.w > rv) {
            // a tighter bound for a perspective matrix (clip w = -z): a point c + d of the sphere, |d| <= rv, projects within
            // F rv sqrt(1 + |c.xy|^2 / w^2) / (w - rv) pixels of the projected centre (Cauchy-Schwarz on w d.xy - c.xy d.w)
            const float rs = focal_px * rv * sqrtf(1.0f + (vp.x * vp.x + vp.y * vp.y) / (cc.w * cc.w)) / (cc.w - rv) * 1.001f;
            if (rs < rad) rad = rs;                                                              // NaN: keeps the corner bound
        }
        fx = fminf(fmaxf(fx, -4.0f), float(P.W) + 4.0f); fy = fminf(fmaxf(fy, -4.0f), float(P.H) + 4.0f);
        float xmin = fx, xmax = fx, ymin = fy, ymax = fy;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
            rad = fmaxf(rad, __shfl_xor_sync(0xffffffffu, rad, o));
        }
        const float ax = __shfl_sync(0xffffffffu, ccx, 0), ay = __shfl_sync(0xffffffffu, ccy, 0);      // lanes 0-7: end point a, 8-15: b
        const float ex = __shfl_sync(0xffffffffu, ccx, 8) - ax, ey = __shfl_sync(0xffffffffu, ccy, 8) - ay;
        const float l2 = ex * ex + ey * ey, inv_l2 = l2 > 0.0f ? 1.0f / l2 : 0.0f;
        // Slack on top of the geometric bound.  The reference's float32 quadratics accept rays that pass OUTSIDE the capsule: the
        // discriminant (d.oc)^2 - |d|^2 (|oc|^2 - r^2) carries a rounding error of a few ulp of |oc|^2 = D^2 (D = distance from the eye),
        // which moves the accepted silhouette out by about k eps D^2 / (2 r) in world units.  Measured on config 4 (D 0.8, r 1e-3, 2.7
        // pixel radius): a slack of 0.25 pixel = k 5 loses 542 of 158.9 M fragments, 0.5 pixel = k 10 loses none; k = 16 is used, times
        // the pixels per world unit at the segment, plus b200_ppll_raster_slack pixels for the rounding of the projection itself.
        const float dmax = sqrtf(fmaxf((va.x * va.x + va.y * va.y) + va.z * va.z, (vb.x * vb.x + vb.y * vb.y) + vb.z * vb.z)) + rv;
        const float px_per_unit = persp ? focal_px / fmaxf(fminf(ca.w, cb.w), 1e-6f) : focal_px;
        const float slack = P.raster_slack + (16.0f * 5.96e-8f) * dmax * dmax / (2.0f * S.radius) * px_per_unit;   // r = 0: inf, no cull (and no hits)
        const float lim2 = (rad + slack) * (rad + slack);
        int x0 = int(floorf(xmin)) - 1, x1 = int(ceilf(xmax)) + 1, y0 = int(floorf(ymin)) - 1, y1 = int(ceilf(ymax)) + 1;
        const bool no_bound = full_frame || __ballot_sync(0xffffffffu, bad) != 0u;
        if (no_bound) { x0 = 0; y0 = 0; x1 = int(P.W) - 1; y1 = int(P.H) - 1; }
        x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, int(P.W) - 1); y1 = min(y1, int(P.H) - 1);
        if (x0 > x1 || y0 > y1) continue;                  // off screen
        if (owned_tiles) {
            // tile-sharded frame: a segment whose rectangle touches none of this rank's tiles is done here -- without this every rank
            // enumerated every segment's candidates and only skipped the pixels (8 GPUs: gather 5.1 -> 3.6 ms per rank for 14.7 / 8 =
            // 1.8 ms of work; a one-segment-per-lane prefilter in front of the warp-wide set-up made it 4.0 ms and was dropped: what is
            // left is the per-segment flush of the two queues, whose rounds are mostly empty when a rank owns an eighth of a segment's pixels)
            const uint32_t tx0 = uint32_t(x0) / P.tile_size, ty0 = uint32_t(y0) / P.tile_size;
            const uint32_t ntx = uint32_t(x1) / P.tile_size - tx0 + 1u, nt = ntx * (uint32_t(y1) / P.tile_size - ty0 + 1u);
            bool any = false;
            for (uint32_t t0 = 0; t0 < nt && !any; t0 += 32u) {
                const uint32_t t = t0 + lane;
                const bool mine = t < nt && owned_tiles[(ty0 + t / ntx) * tiles_x + tx0 + t % ntx] != 0;
                any = __ballot_sync(0xffffffffu, mine) != 0u;
            }
            if (!any) continue;
        }
        const uint32_t bw = uint32_t(x1 - x0 + 1), area = bw * uint32_t(y1 - y0 + 1);
        // row / column of candidate i: an exact float reciprocal division while (i + 0.5) / bw cannot come within rounding of an
        // integer (area < 2^22), the integer division otherwise (whole-frame rectangles)
        const bool fdiv = area < (1u << 22);
        const float inv_bw = 1.0f / float(bw);
        // segment-only terms of the acceptance test and of the shader (warp-uniform)
        const float r = S.radius;
        const float blx = fminf(s.a.x, s.b.x) - r, bly = fminf(s.a.y, s.b.y) - r, blz = fminf(s.a.z, s.b.z) - r;   // the record's own AABB,
        const float bhx = fmaxf(s.a.x, s.b.x) + r, bhy = fmaxf(s.a.y, s.b.y) + r, bhz = fmaxf(s.a.z, s.b.z) + r;   // as in seg_box_hit
        const SegShade pre = seg_shade(s);
        const Vec3 axis = pre.tan0;                         // normalize(p1 - p0): cylinder_hit's axis is the shader's tangent
        // The cheap tests (ownership, 2-D cull) thin the rectangle out; their survivors are queued per warp and taken 32 at a time, so
        // that the expensive parts -- slab test + capsule test, then shading -- run with full warps.
        uint32_t q_count = 0, h_count = 0;                  // uniform over the warp; <= 31 between rounds
        for (uint32_t base = 0; base < area; base += 32u) {
            const uint32_t i = base + lane;
            uint32_t cx = 0, cy = 0;
            if (i < area) {
                const uint32_t row = fdiv ? uint32_t((float(i) + 0.5f) * inv_bw) : i / bw;
                cx = uint32_t(x0) + (i - row * bw); cy = uint32_t(y0) + row;
            }
            // tile-sharded frames: only the pixels of this rank's tiles (1 byte per tile of the frame)
            bool cand = i < area && (!owned_tiles || owned_tiles[(cy / P.tile_size) * tiles_x + cx / P.tile_size]);
            if (cand && !no_bound) {   // 2-D distance from the projected axis; a NaN anywhere keeps the candidate
                const float qx = float(cx) - ax, qy = float(cy) - ay;
                const float tt = fminf(fmaxf((qx * ex + qy * ey) * inv_l2, 0.0f), 1.0f);
                const float dx = qx - tt * ex, dy = qy - tt * ey;
                if (dx * dx + dy * dy > lim2) cand = false;
            }
            const unsigned cm = __ballot_sync(0xffffffffu, cand);
            if (cand) queue[q_count + __popc(cm & lt_mask)] = cx | (cy << 16);
            q_count += __popc(cm);
            __syncwarp();
            const bool last_round = base + 32u >= area;
            while (q_count >= 32u || (last_round && (q_count | h_count) != 0u)) {
                // ---- test stage: the newest n candidates (n = 0 only when the rectangle is done and hits are still queued)
                const uint32_t n = q_count < 32u ? q_count : 32u;
                q_count -= n;
                const uint32_t e = queue[q_count + (lane < n ? lane : 0u)];
                bool hit = false;
                float t = 0.0f; uint32_t kind = 0;
                if (lane < n) {
                    const float4* pr = pixel_rays + 2 * (size_t(e >> 16) * P.W + (e & 0xffffu));
                    const float4 r0 = __ldg(pr), r1 = __ldg(pr + 1);
                    RayQ rq; rq.o = ro; rq.d = v3(r0.x, r0.y, r0.z); rq.dd = r0.w;
                    RayBox rb; rb.ix = r1.x; rb.iy = r1.y; rb.iz = r1.z;
                    rb.cx = -(ro.x * rb.ix); rb.cy = -(ro.y * rb.iy); rb.cz = -(ro.z * rb.iz);
                    isect++;
                    float tn;
                    hit = box_hit(rb, blx, bly, blz, bhx, bhy, bhz, tmin, tmax, tn) && capsule_hit(rq, s, axis, r, capped, t, kind) && t >= tmin && t <= tmax;
                }
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                if (hit) { const uint32_t k = h_count + __popc(hm & lt_mask); hpix[k] = e; ht[k] = __float_as_uint(t); hkind[k] = kind; }
                h_count += __popc(hm);
                __syncwarp();
                // ---- shade stage: a full warp of hits, or what is left once the last candidates have been tested
                const bool drained = last_round && q_count == 0u;
                while (h_count >= 32u || (drained && h_count != 0u)) {
                    const uint32_t m = h_count < 32u ? h_count : 32u;
                    h_count -= m;
                    const uint32_t hi = h_count + (lane < m ? lane : 0u);
                    const uint32_t hp = hpix[hi];
                    const uint32_t px = hp & 0xffffu, py = hp >> 16;
                    bool keep = false;
                    uint32_t col = 0;
                    float depth = 0.0f;
                    if (lane < m) {
                        const float4 r0 = __ldg(pixel_rays + 2 * (size_t(py) * P.W + px));
                        const Shaded sh = shade_hit<SAO>(P, ro, v3(r0.x, r0.y, r0.z), __uint_as_float(ht[hi]), hkind[hi], s, pre, SAO && S.seg_aux ? S.seg_aux + seg : nullptr);
                        if (!(sh.color.w < 0.001f)) { keep = true; col = pack_unorm4x8(sh.color); depth = sh.hit_t; }   // LinkedListGather.glsl:38
                    }
                    const unsigned km = __ballot_sync(0xffffffffu, keep);
                    if (km) {
                        unsigned long long idx = 0;
                        const int leader = __ffs(km) - 1;
                        if (int(lane) == leader) idx = atomicAdd(frag_counter, (unsigned long long)__popc(km));   // fragCounter, LinkedListGather.glsl:55
                        idx = __shfl_sync(0xffffffffu, idx, leader) + __popc(km & lt_mask);
                        if (keep) {
                            gen++;
                            if (idx < list_size) {                                                                 // :57
                                const uint32_t a = addr_gen(P, px, py);
                                lv_ppll_node nd; nd.color = col; nd.depth = depth;
                                nd.next = STAGE ? a : atomicExch(heads + a, uint32_t(idx));                        // :60
                                nodes[idx] = nd;                                                                   // STAGE: `nodes` is the staging array
                                atomicAdd(counts + a, 1u);
                            }
                        }
                    }
                    __syncwarp();   // the entries just read may be overwritten by the next push
                }
            }
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&C->rays_primary, n_pixels);   // the pixels the pass covers (what the ray-cast gather counts as rays)
    flush_counter(&C->isect, isect);
    flush_counter(&C->frags_generated, gen);
}

// staged fragments -> contiguous per-pixel runs (offs = exclusive scan of counts; cursor zeroed).  The order inside a run is the order
// of arrival (a race, like the reference's list order).
__global__ void k_ppll_fill(const lv_ppll_node* stage, const unsigned long long* frag_counter, unsigned long long list_size, const uint32_t* offs,
                            unsigned int* cursor, const uint32_t* counts, uint32_t* heads, lv_ppll_node* nodes) {
    const unsigned long long n = *frag_counter < list_size ? *frag_counter : list_size;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        lv_ppll_node f = stage[i];
        const uint32_t a = f.next;
        const uint32_t k = atomicAdd(cursor + a, 1u), slot = offs[a] + k;
        f.next = k ? slot - 1u : kNone;
        nodes[slot] = f;
        if (k + 1u == counts[a]) heads[a] = slot;
    }
}

// S10 resolve.  Per warp: 32 pixels.  Lanes walk their own lists (32 independent pointer chases in flight) into a
// shared-memory tile of kResolveCap 64-bit keys (depth bits << 32 | colour); lists are packed back to back, as many
// pixels per round as fit.  Each packed list is then sorted by the whole warp with an all-ascending bitonic network
// and blended front to back by its owning lane, in sorted order, with the reference's arithmetic
// (LinkedListSort.glsl:45-58; early-out at alpha >= 0.99 for the priority-queue mode, :217-218).
constexpr int kResolveCap = 1024;     // keys per warp (8 KiB)
constexpr int kResolveWarps = 4;
constexpr int kResolveInsertionMax = 64;  // lists up to this length are sorted by their own lane

__device__ __forceinline__ void cmpxchg(unsigned long long* s, uint32_t i, uint32_t l) {
    unsigned long long a = s[i], b = s[l];
    if (a > b) { s[i] = b; s[l] = a; }
}
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* s, uint32_t n, uint32_t lane) {
    uint32_t m = 2;
    while (m < n) m <<= 1;
    const uint32_t half = m >> 1;
    for (uint32_t k = 2; k <= m; k <<= 1) {
        const uint32_t hk = k >> 1;
        for (uint32_t idx = lane; idx < half; idx += 32) {   // flip: i <-> mirrored partner inside the k-block
            const uint32_t blk = idx / hk, t = idx - blk * hk;
            const uint32_t i = blk * k + t, l = blk * k + (k - 1 - t);
            if (l < n) cmpxchg(s, i, l);
        }
        __syncwarp();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t idx = lane; idx < half; idx += 32) {
                const uint32_t i = 2 * idx - (idx & (j - 1)), l = i + j;
                if (l < n) cmpxchg(s, i, l);
            }
            __syncwarp();
        }
    }
}

// The same network with the keys in REGISTERS (KPL per lane, n <= 32 * KPL; b200_ppll_reg_sort): a compare-exchange is a pair of
// selects (partner in the same lane) or a 64-bit shuffle plus a select (partner in another lane) instead of two shared-memory
// round trips, runtime index arithmetic and a __syncwarp per pass.  Network position p = lane * KPL + r, so the passes with distance
// < KPL stay inside a lane; the keys are fetched at [r * 32 + lane] (conflict-free; which key starts where is irrelevant to a sort)
// and come back to [p] in ascending order.  Missing keys are +inf padding.
template <int KPL>
__device__ __forceinline__ void warp_bitonic_sort_reg(unsigned long long* s, uint32_t n, uint32_t lane) {
    unsigned long long k[KPL];
#pragma unroll
    for (int r = 0; r < KPL; r++) { const uint32_t e = uint32_t(r) * 32u + lane; k[r] = e < n ? s[e] : ~0ull; }
#pragma unroll
    for (int kk = 2; kk <= 32 * KPL; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= KPL) {
                const uint32_t pl = uint32_t(j / KPL);                                          // partner = lane ^ pl, same register
                const bool up = kk >= 32 * KPL || (lane & uint32_t(kk / KPL)) == 0u;              // ascending block: (p & kk) == 0
                const bool keep_min = ((lane & pl) == 0u) == up;
#pragma unroll
                for (int r = 0; r < KPL; r++) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[r], int(pl));
                    k[r] = ((o < k[r]) == keep_min) ? o : k[r];   // equal keys: taking the partner's copy changes nothing
                }
            } else {
#pragma unroll
                for (int r = 0; r < KPL; r++) {
                    if ((r & j) == 0) {
                        const int q = r | j;
                        const bool up = kk < KPL ? (r & kk) == 0 : (kk >= 32 * KPL || (lane & uint32_t(kk / KPL)) == 0u);
                        const unsigned long long a = k[r], b = k[q];
                        const bool sw = (a > b) == up;                // equal keys may swap: no effect
                        k[r] = sw ? b : a; k[q] = sw ? a : b;
                    }
                }
            }
        }
    }
    __syncwarp();   // every lane has fetched its keys before any is overwritten
#pragma unroll
    for (int r = 0; r < KPL; r++) { const uint32_t p2 = lane * uint32_t(KPL) + uint32_t(r); if (p2 < n) s[p2] = k[r]; }
}

// CAP: keys per warp in the shared tile (b200_ppll_resolve_tile; must hold the longest list, max_frags <= CAP).  A smaller tile
// lets more warps live on an SM -- the walk is a dependent pointer chase, so short-list frames are bound by warps in flight.
// CONTIG: every list is a contiguous run ending at its head (k_ppll_fill): node i of the walk is nodes[head - i], no pointer chase.
template <bool REGSORT, int CAP = kResolveCap, bool CONTIG = false>
__global__ void __launch_bounds__(kBlockThreads)
k_ppll_resolve(const __grid_constant__ FrameParams P, const uint32_t* heads, const uint32_t* counts, const lv_ppll_node* nodes,
               uint32_t max_frags, int early_out, float4* image, Counters* C, const uint32_t* order, const unsigned int* n_sorted, uint32_t* out8) {
    __shared__ unsigned long long s_keys[kResolveWarps * CAP];
    __shared__ float s_unorm[256];   // unpackUnorm4x8: float(b) / 255.0f, tabulated once (4 IEEE divisions per fragment otherwise)
    for (uint32_t i = threadIdx.x; i < 256u; i += kBlockThreads) s_unorm[i] = float(i) / 255.0f;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* tile = s_keys + warp * CAP;
    uint32_t x, y;
    bool valid;
    if (order) {   // binned mode: this kernel only takes the first n_sorted[0] pixels of `order` (lists longer than 256 keys)
        const uint32_t slot = blockIdx.x * kBlockThreads + threadIdx.x;
        valid = slot < n_sorted[0];
        const uint32_t pixel = valid ? order[slot] : 0u;
        x = pixel % P.W; y = pixel / P.W;
    } else valid = thread_pixel(P, x, y);
    uint32_t head = kNone, total = 0;
    if (valid) { const uint32_t a = addr_gen(P, x, y); head = heads[a]; total = counts[a]; }
    const uint32_t cnt = min(total, max_frags);
    if (valid && cnt == 0) {   // discard -> clear colour
        if (out8) out8[size_t(y) * P.W + x] = pack_unorm4x8(v4(P.bg[0], P.bg[1], P.bg[2], P.bg[3]));
        else image[size_t(y) * P.W + x] = make_float4(P.bg[0], P.bg[1], P.bg[2], P.bg[3]);
    }
    unsigned remaining = __ballot_sync(0xffffffffu, cnt > 0);
    while (remaining) {
        const uint32_t c = ((remaining >> lane) & 1u) ? cnt : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (int(lane) >= o) incl += v; }
        const uint32_t excl = incl - c;
        const bool sel = c > 0 && incl <= uint32_t(CAP);
        if (sel) {
            unsigned long long* mine = tile + excl;
            if (CONTIG) {
                // the list is the run nodes[head - c + 1 .. head]: four independent node fetches in flight per lane, then their inserts
                const bool ins = c <= uint32_t(kResolveInsertionMax);
                for (uint32_t i = 0; i < c; i += 4u) {
                    lv_ppll_node n4[4];
#pragma unroll
                    for (uint32_t u = 0; u < 4u; u++) if (i + u < c) n4[u] = nodes[head - (i + u)];
#pragma unroll
                    for (uint32_t u = 0; u < 4u; u++) {
                        if (i + u < c) {
                            const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(n4[u].depth)) << 32) | n4[u].color;
                            uint32_t j = i + u;
                            if (ins) while (j > 0 && mine[j - 1] > key) { mine[j] = mine[j - 1]; j--; }
                            mine[j] = key;
                        }
                    }
                }
            } else {
            lv_ppll_node nd = nodes[head];
            if (c <= uint32_t(kResolveInsertionMax)) {
                // short list: this lane insertion-sorts its own slice while the next node is in flight (32 lists in parallel)
                for (uint32_t i = 0; i < c; i++) {
                    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(nd.depth)) << 32) | nd.color;
                    if (i + 1 < c) nd = nodes[nd.next];
                    uint32_t j = i;
                    while (j > 0 && mine[j - 1] > key) { mine[j] = mine[j - 1]; j--; }
                    mine[j] = key;
                }
            } else {
                for (uint32_t i = 0; i < c; i++) {
                    mine[i] = (static_cast<unsigned long long>(__float_as_uint(nd.depth)) << 32) | nd.color;
                    if (i + 1 < c) nd = nodes[nd.next];
                }
            }
            }
        }
        __syncwarp();
        const unsigned selmask = __ballot_sync(0xffffffffu, sel);
        // long lists: cooperative bitonic sort, one list at a time
        for (unsigned m = __ballot_sync(0xffffffffu, sel && c > uint32_t(kResolveInsertionMax)); m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const uint32_t n = __shfl_sync(0xffffffffu, c, src), base = __shfl_sync(0xffffffffu, excl, src);
            if (REGSORT && n <= 128u) warp_bitonic_sort_reg<4>(tile + base, n, lane);
            else if (REGSORT && n <= 256u) warp_bitonic_sort_reg<8>(tile + base, n, lane);
            else warp_bitonic_sort(tile + base, n, lane);
        }
        __syncwarp();
        if (sel) {
            float r = 0.0f, g = 0.0f, b = 0.0f, a = 0.0f;
            for (uint32_t i = 0; i < c; i++) {
                if (early_out && !(a < 0.99f)) break;
                const uint32_t col = uint32_t(tile[excl + i] & 0xffffffffull);
                const float sr = s_unorm[col & 0xffu], sg = s_unorm[(col >> 8) & 0xffu];
                const float sb = s_unorm[(col >> 16) & 0xffu], sa = s_unorm[col >> 24];
                r = r + (1.0f - a) * sa * sr;
                g = g + (1.0f - a) * sa * sg;
                b = b + (1.0f - a) * sa * sb;
                a = a + (1.0f - a) * sa;
            }
            r = r / a; g = g / a; b = b / a;
            // BACK_TO_FRONT_STRAIGHT_ALPHA over the clear colour (PerPixelLinkedListLineRenderer.cpp:70)
            const Vec4 px = v4(r * a + P.bg[0] * (1.0f - a), g * a + P.bg[1] * (1.0f - a), b * a + P.bg[2] * (1.0f - a), a + P.bg[3] * (1.0f - a));
            if (out8) out8[size_t(y) * P.W + x] = pack_unorm4x8(px);   // b200_frame_format = rgba8: the resolved pixel in the sceneTexture's format
            else image[size_t(y) * P.W + x] = make_float4(px.x, px.y, px.z, px.w);
        }
        remaining &= ~selmask;
        __syncwarp();
    }
    flush_counter(&C->frags_sorted, cnt);
    flush_counter(&C->frags_truncated, total - cnt);
    unsigned mx = total;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx) atomicMax(&C->max_depth_complexity, mx);
}

// ------------------------------------------------------------------------------------------------
// S10 resolve, count-binned variant (default).  Lists of very different lengths in one warp cost max(n)^2 for everybody, so
// the owned pixels are first counting-sorted by list length, longest first (k_ppll_bin_*: histogram, scan, scatter; empty
// pixels get the clear colour right there).  k_ppll_resolve_binned then gives every warp 32 pixels with (nearly) EQUAL
// list length: the 32 lists live interleaved in shared memory (element i of lane l at [i*32 + l]: all lanes touch the same
// row at the same time, no bank conflicts), every lane chases its own list, sorts it (insertion sort while the next node
// is in flight up to 32 keys, Shell sort with Ciura's gaps beyond) and blends it.  One instantiation per length class
// 1..32 / 33..64 / 65..128 / 129..256 so that short lists keep high occupancy; lists longer than 256 (possible only with
// max_frags > 256) take the cooperative kernel above, restricted to those pixels.
// n_sorted[k] = number of pixels whose list is longer than kBinBounds[k]; the order array is sorted by descending length,
// so class k owns the slots [n_sorted[k-1], n_sorted[k]).
constexpr int kBinClasses = 5;
__device__ __constant__ int kBinBounds[kBinClasses] = {256, 128, 64, 32, 0};

__global__ void k_ppll_bin_count(const __grid_constant__ FrameParams P, const uint32_t* counts, uint32_t max_frags, unsigned int* hist) {
    uint32_t x, y;
    const bool valid = thread_pixel(P, x, y);
    const uint32_t c = valid ? min(counts[addr_gen(P, x, y)], max_frags) : 0u;
    if (c) {   // empty pixels are not binned (they would all hammer one address); equal lengths in a warp share one atomic
        const unsigned m = __match_any_sync(__activemask(), c);
        if ((threadIdx.x & 31) == uint32_t(__ffs(m) - 1)) atomicAdd(hist + c, (unsigned)__popc(m));
    }
}
__global__ void k_ppll_bin_scan(uint32_t max_frags, const unsigned int* hist, unsigned int* offsets, unsigned int* n_sorted) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned acc = 0;
        int k = 0;
        for (int len = int(max_frags); len >= 1; len--) {
            while (k < kBinClasses && len <= kBinBounds[k]) n_sorted[k++] = acc;
            offsets[len] = acc;
            acc += hist[len];
        }
        while (k < kBinClasses) n_sorted[k++] = acc;
    }
}
__global__ void k_ppll_bin_scatter(const __grid_constant__ FrameParams P, const uint32_t* counts, uint32_t max_frags, unsigned int* offsets,
                                   uint32_t* order, float4* image) {
    uint32_t x, y;
    const bool valid = thread_pixel(P, x, y);
    if (!valid) return;
    const uint32_t c = min(counts[addr_gen(P, x, y)], max_frags);
    if (c == 0) { image[size_t(y) * P.W + x] = make_float4(P.bg[0], P.bg[1], P.bg[2], P.bg[3]); return; }   // discard -> clear colour
    const uint32_t lane = threadIdx.x & 31;
    const unsigned m = __match_any_sync(__activemask(), c);
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (int(lane) == leader) base = atomicAdd(offsets + c, (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    order[base + __popc(m & ((1u << lane) - 1u))] = y * P.W + x;
}

// CLASS k serves the slots [n_sorted[k-1], n_sorted[k]) (k >= 1), lists of at most MAXN = kBinBounds[k-1] keys.
template <int MAXN, int WARPS, int CLASS>
__global__ void __launch_bounds__(WARPS * 32)
k_ppll_resolve_binned(const __grid_constant__ FrameParams P, const uint32_t* heads, const uint32_t* counts, const lv_ppll_node* nodes,
                      const uint32_t* order, const unsigned int* n_sorted, uint32_t max_frags, int early_out, float4* image, Counters* C) {
    extern __shared__ unsigned long long s_dyn[];                 // WARPS * MAXN * 32 keys, then 256 floats
    float* s_unorm = reinterpret_cast<float*>(s_dyn + size_t(WARPS) * MAXN * 32);
    for (uint32_t i = threadIdx.x; i < 256u; i += WARPS * 32) s_unorm[i] = float(i) / 255.0f;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* mine = s_dyn + size_t(warp) * MAXN * 32 + lane;   // element i at mine[i * 32]
    const uint32_t first = n_sorted[CLASS - 1], last = n_sorted[CLASS];
    uint32_t sorted = 0, trunc = 0, mx = 0;
    for (uint32_t base = first + (blockIdx.x * WARPS + warp) * 32u; base < last; base += gridDim.x * WARPS * 32u) {
        const uint32_t slot = base + lane;
        if (slot < last) {
            const uint32_t pixel = order[slot];
            const uint32_t x = pixel % P.W, y = pixel / P.W;
            const uint32_t a = addr_gen(P, x, y);
            const uint32_t total = counts[a], c = min(total, max_frags);
            sorted += c; trunc += total - c; mx = max(mx, total);
            lv_ppll_node nd = nodes[heads[a]];
            if (MAXN <= 32) {
                for (uint32_t i = 0; i < c; i++) {
                    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(nd.depth)) << 32) | nd.color;
                    if (i + 1 < c) nd = nodes[nd.next];
                    uint32_t j = i;
                    while (j > 0 && mine[(j - 1) * 32] > key) { mine[j * 32] = mine[(j - 1) * 32]; j--; }
                    mine[j * 32] = key;
                }
            } else {
                for (uint32_t i = 0; i < c; i++) {
                    mine[i * 32] = (static_cast<unsigned long long>(__float_as_uint(nd.depth)) << 32) | nd.color;
                    if (i + 1 < c) nd = nodes[nd.next];
                }
                const uint32_t gaps[6] = {132u, 57u, 23u, 10u, 4u, 1u};   // Ciura
#pragma unroll
                for (int gi = 0; gi < 6; gi++) {
                    const uint32_t gap = gaps[gi];
                    if (gap >= uint32_t(MAXN)) continue;
                    for (uint32_t i = gap; i < c; i++) {
                        const unsigned long long key = mine[i * 32];
                        uint32_t j = i;
                        while (j >= gap && mine[(j - gap) * 32] > key) { mine[j * 32] = mine[(j - gap) * 32]; j -= gap; }
                        mine[j * 32] = key;
                    }
                }
            }
            float r = 0.0f, g = 0.0f, b = 0.0f, al = 0.0f;
            for (uint32_t i = 0; i < c; i++) {
                if (early_out && !(al < 0.99f)) break;
                const uint32_t col = uint32_t(mine[i * 32] & 0xffffffffull);
                const float sr = s_unorm[col & 0xffu], sg = s_unorm[(col >> 8) & 0xffu];
                const float sb = s_unorm[(col >> 16) & 0xffu], sa = s_unorm[col >> 24];
                r = r + (1.0f - al) * sa * sr;
                g = g + (1.0f - al) * sa * sg;
                b = b + (1.0f - al) * sa * sb;
                al = al + (1.0f - al) * sa;
            }
            r = r / al; g = g / al; b = b / al;
            image[size_t(y) * P.W + x] = make_float4(r * al + P.bg[0] * (1.0f - al), g * al + P.bg[1] * (1.0f - al),
                                                      b * al + P.bg[2] * (1.0f - al), al + P.bg[3] * (1.0f - al));
        }
        __syncwarp();
    }
    flush_counter(&C->frags_sorted, sorted);
    flush_counter(&C->frags_truncated, trunc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx) atomicMax(&C->max_depth_complexity, mx);
}

__global__ void k_fill_f32(float* p, size_t n, float v) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = v;
}

// RGBA32F frame -> RGBA8 UNORM (the reference's sceneTexture format), owned tiles only
__global__ void k_frame_to_rgba8(const __grid_constant__ FrameParams P, const float4* image, uint32_t* out) {
    uint32_t x, y;
    if (!thread_pixel(P, x, y)) return;
    const float4 c = image[size_t(y) * P.W + x];
    out[size_t(y) * P.W + x] = pack_unorm4x8(v4(c.x, c.y, c.z, c.w));
}

// ------------------------------------------------------------------------------------------------
// tile pack / unpack around the multi-GPU framebuffer gather
__global__ void k_pack_tiles(const float4* image, uint32_t W, uint32_t H, const uint2* tiles, uint32_t tile_size, float4* packed) {
    const uint32_t tile = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tile_size * tile_size) return;
    const uint32_t x = tiles[tile].x * tile_size + i % tile_size, y = tiles[tile].y * tile_size + i / tile_size;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x < W && y < H) v = image[size_t(y) * W + x];
    packed[size_t(tile) * tile_size * tile_size + i] = v;
}
__global__ void k_unpack_tiles(const float4* packed, uint32_t W, uint32_t H, const uint2* tiles, uint32_t tile_size, float4* image) {
    const uint32_t tile = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tile_size * tile_size) return;
    const uint32_t x = tiles[tile].x * tile_size + i % tile_size, y = tiles[tile].y * tile_size + i / tile_size;
    if (x < W && y < H) image[size_t(y) * W + x] = packed[size_t(tile) * tile_size * tile_size + i];
}

}  // namespace lv
