// lv_bvh.cuh -- GPU BVH build over line segments (replaces the driver-built VkAccelerationStructureKHR,
// reference src/LineData/LineData.cpp:879-907,1057-1075; AABBs as in src/LineData/LineDataFlow.cpp:2230-2233).
//
// LBVH: 63-bit Morton codes of the segment-box centres -> radix sort -> Karras' parallel binary radix tree
// -> bottom-up box fit -> 64-byte child-pair nodes.  Every inner node of a radix tree covers a contiguous
// range of the sorted records, so a subtree with <= leaf_max records is referenced directly as a leaf
// (first record, count) and never materialised.
#pragma once
#ifndef LV_HOST_EMU   // the host emulation (tests/emu) supplies cub::DeviceRadixSort::SortPairs itself
#include <cub/cub.cuh>
#endif
#include "lv_types.cuh"
#include "lv_math.cuh"

namespace lv {

__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}

__global__ void k_init_bounds(float* b) {
    if (threadIdx.x < 3) b[threadIdx.x] = __int_as_float(0x7f800000);
    else if (threadIdx.x < 6) b[threadIdx.x] = __int_as_float(0xff800000);
}

__device__ __forceinline__ void seg_box(const float* pos, const uint32_t* idx, uint32_t i, float r, float* mn, float* mx) {
    uint32_t a = idx[2 * i], b = idx[2 * i + 1];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float p = pos[3 * size_t(a) + k], q = pos[3 * size_t(b) + k];
        mn[k] = fminf(p, q) - r; mx[k] = fmaxf(p, q) + r;
    }
}

__global__ void k_scene_bounds(const float* pos, const uint32_t* idx, uint32_t n, float r, float* bounds) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float a[3], b[3];
        seg_box(pos, idx, i, r, a, b);
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

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const float* pos, const uint32_t* idx, uint32_t n, float r, const float* bounds,
                         unsigned long long* keys, uint32_t* vals, bool cubic) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float mn[3], mx[3];
    seg_box(pos, idx, i, r, mn, mx);
    unsigned long long code = 0;
    // ONE scale for the three axes (the largest extent): Morton cells are cubes whatever the aspect of the scene box, so the radix
    // tree's splits halve the longest side first instead of cutting a short axis of a flat box (config 5's 2:1:2 box) too early.
    // b200_bvh_morton = per_axis keeps the old normalisation (each axis to [0, 1]).
    const float ext_max = fmaxf(fmaxf(bounds[3] - bounds[0], bounds[4] - bounds[1]), bounds[5] - bounds[2]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float lo = bounds[k], ext = cubic ? ext_max : bounds[3 + k] - lo;
        float c = 0.5f * (mn[k] + mx[k]);
        float u = ext > 0.0f ? (c - lo) / ext : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        unsigned long long q = (unsigned long long)(fminf(u * 2097152.0f, 2097151.0f));
        code |= expand21(q) << (2 - k);
    }
    keys[i] = code;
    vals[i] = i;
}

__global__ void k_pack_segments(const float* pos, const float* attr, const uint32_t* idx, const uint32_t* order, uint32_t n, SegRec* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = order[i];
    uint32_t a = idx[2 * size_t(s)], b = idx[2 * size_t(s) + 1];
    SegRec rec;
    rec.a = make_float4(pos[3 * size_t(a)], pos[3 * size_t(a) + 1], pos[3 * size_t(a) + 2], attr[a]);
    rec.b = make_float4(pos[3 * size_t(b)], pos[3 * size_t(b) + 1], pos[3 * size_t(b) + 2], attr[b]);
    out[i] = rec;
}

__device__ __forceinline__ int delta_k(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll(a ^ b);
}

// Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees" (HPG 2012), one thread per inner node.
__global__ void k_radix_tree(const unsigned long long* keys, int n, int2* children, int2* ranges, int* parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta_k(keys, n, i, i + 1) - delta_k(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta_k(keys, n, i, i - d);
    int lmax = 2;
    while (delta_k(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta_k(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta_k(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta_k(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int lc = gamma, rc = gamma + 1;
    int lcode = (first == gamma) ? (lc | 0x80000000) : lc;
    int rcode = (last == gamma + 1) ? (rc | 0x80000000) : rc;
    children[i] = make_int2(lcode, rcode);
    ranges[i] = make_int2(first, last);
    // parent array: inner nodes at [0, n-1), leaves at [n-1, 2n-1)
    parent[(first == gamma) ? (n - 1 + lc) : lc] = i;
    parent[(last == gamma + 1) ? (n - 1 + rc) : rc] = i;
    if (i == 0) parent[0] = -1;
}

// bottom-up fit: one thread per leaf record; the second thread to arrive at an inner node continues upward.
__global__ void k_fit(const SegRec* segs, int n, float r, const int2* children, const int* parent, float* boxes, unsigned int* flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SegRec s = segs[i];
    float b[6] = {fminf(s.a.x, s.b.x) - r, fminf(s.a.y, s.b.y) - r, fminf(s.a.z, s.b.z) - r,
                  fmaxf(s.a.x, s.b.x) + r, fmaxf(s.a.y, s.b.y) + r, fmaxf(s.a.z, s.b.z) + r};
    float* mine = boxes + 6 * size_t(n - 1 + i);
#pragma unroll
    for (int k = 0; k < 6; k++) mine[k] = b[k];
    if (n == 1) return;
    int node = parent[n - 1 + i];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(flags + node, 1u) == 0u) return;  // first arrival: the sibling will finish this node
        __threadfence();
        int2 ch = children[node];
        const volatile float* lb = boxes + 6 * size_t((ch.x < 0) ? (n - 1 + (ch.x & 0x7fffffff)) : ch.x);
        const volatile float* rb = boxes + 6 * size_t((ch.y < 0) ? (n - 1 + (ch.y & 0x7fffffff)) : ch.y);
        float* o = boxes + 6 * size_t(node);
#pragma unroll
        for (int k = 0; k < 3; k++) { o[k] = fminf(lb[k], rb[k]); o[3 + k] = fmaxf(lb[3 + k], rb[3 + k]); }
        node = parent[node];
    }
}

// upper bound of the traversal stack depth: longest leaf-to-root parent chain
__global__ void k_tree_depth(int n, const int* parent, unsigned int* max_depth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int d = 0;
    if (i < n && n > 1) {
        int node = parent[n - 1 + i];
        while (node >= 0) { d++; node = parent[node]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d) atomicMax(max_depth, d);
}

__device__ __forceinline__ void emit_child(int code, int n, int leaf_max, const int2* ranges, const float* boxes, float4& mnref, float4& mxcnt) {
    uint32_t ref, cnt; const float* b;
    if (code < 0) { ref = uint32_t(code & 0x7fffffff); cnt = 1; b = boxes + 6 * size_t(n - 1 + ref); }
    else {
        int2 rg = ranges[code];
        int size = rg.y - rg.x + 1;
        b = boxes + 6 * size_t(code);
        if (size <= leaf_max) { ref = uint32_t(rg.x); cnt = uint32_t(size); }
        else { ref = uint32_t(code); cnt = 0; }
    }
    // child word: leaf = bit31 | (count-1) << 27 | first record, inner = node index; the count is repeated in the max.w lane
    mnref = make_float4(b[0], b[1], b[2], __uint_as_float(cnt ? (0x80000000u | ((cnt - 1u) << 27) | ref) : ref));
    mxcnt = make_float4(b[3], b[4], b[5], __uint_as_float(cnt));
}

__global__ void k_emit_nodes(int n, int leaf_max, const int2* children, const int2* ranges, const float* boxes, Node64* nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 1) {
        if (i == 0) {
            const float* b = boxes;
            Node64 nd;
            nd.l0 = make_float4(b[0], b[1], b[2], __uint_as_float(0x80000000u));
            nd.l1 = make_float4(b[3], b[4], b[5], __uint_as_float(1u));
            // absent child: a point box at +inf, which no finite ray's slab test passes.  A NaN ray passes EVERY min/max slab test, so the
            // word must not be an inner-node index (0 would lead back to the root: an endless loop); it references the dummy record the
            // scene keeps behind its last one (all NaN, never hit) as a one-record leaf.
            nd.r0 = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(0x80000000u | uint32_t(n)));
            nd.r1 = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(1u));
            nodes[0] = nd;
        }
        return;
    }
    if (i >= n - 1) return;
    Node64 nd;
    if (i == 0 && n <= leaf_max) {  // whole scene fits one leaf
        const float* b = boxes;
        nd.l0 = make_float4(b[0], b[1], b[2], __uint_as_float(0x80000000u | ((uint32_t(n) - 1u) << 27)));
        nd.l1 = make_float4(b[3], b[4], b[5], __uint_as_float(uint32_t(n)));
        nd.r0 = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(0x80000000u | uint32_t(n)));   // absent child -> the dummy record (see above)
        nd.r1 = make_float4(INFINITY, INFINITY, INFINITY, __uint_as_float(1u));
        nodes[0] = nd;
        return;
    }
    int2 ch = children[i];
    emit_child(ch.x, n, leaf_max, ranges, boxes, nd.l0, nd.l1);
    emit_child(ch.y, n, leaf_max, ranges, boxes, nd.r0, nd.r1);
    nodes[i] = nd;
}

// Node64 -> NodeQ (b200_ao_qnodes).  q_min rounds down, q_max rounds up, then both are corrected against the traversal's own
// dequantisation fma(q, scale, origin) until the quantised box encloses the exact one.
__device__ __forceinline__ uint32_t quantize_bound(float b, float o, float s, bool up) {
    float g = (b - o) / s;
    g = up ? ceilf(g) : floorf(g);
    int q = int(fminf(fmaxf(g, 0.0f), 65535.0f));
    if (up) { while (q < 65535 && __fmaf_rn(float(q), s, o) < b) q++; }
    else { while (q > 0 && __fmaf_rn(float(q), s, o) > b) q--; }
    return uint32_t(q);
}
__global__ void k_quantize_nodes(const Node64* nodes, uint32_t n, float ox, float oy, float oz, float sx, float sy, float sz, NodeQ* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Node64 nd = nodes[i];
    NodeQ q;
    const float4 mn[2] = {nd.l0, nd.r0}, mx[2] = {nd.l1, nd.r1};
#pragma unroll
    for (int c = 0; c < 2; c++) {
        uint32_t* w = q.w + 4 * c;
        if (mn[c].x == INFINITY) { w[0] = w[1] = w[2] = 0u; w[3] = kAbsentChild; continue; }   // absent child (a point box at +inf in Node64)
        const uint32_t ax = quantize_bound(mn[c].x, ox, sx, false), ay = quantize_bound(mn[c].y, oy, sy, false), az = quantize_bound(mn[c].z, oz, sz, false);
        const uint32_t bx = quantize_bound(mx[c].x, ox, sx, true), by = quantize_bound(mx[c].y, oy, sy, true), bz = quantize_bound(mx[c].z, oz, sz, true);
        w[0] = ax | (ay << 16); w[1] = az | (bx << 16); w[2] = by | (bz << 16);
        w[3] = __float_as_uint(mn[c].w);
    }
    out[i] = q;
}

// ---- PLOC builder (b200_bvh_builder = ploc) -----------------------------------------------------------------------------------
// Parallel locally-ordered clustering (Meister & Bittner 2018; the reference library's LocallyOrderedClusteringBuilder,
// submodules/bvh/include/bvh/locally_ordered_clustering_builder.hpp, is the CPU form): the records in Morton order are the initial
// clusters; every round each cluster looks `radius` neighbours to either side for the partner with the smallest merged surface area,
// mutual nearest neighbours merge into a new child-pair node, the survivors are compacted in order, until one cluster is left.
// A cluster = its box, the child word that refers to it (leaf | record, or node index) and the height of its subtree.
struct PlocCluster {
    float4 lo;   // min.xyz, as_float(child word)
    float4 hi;   // max.xyz, as_float(subtree height)
};

__global__ void k_ploc_init(const SegRec* segs, uint32_t n, float r, PlocCluster* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SegRec s = segs[i];
    PlocCluster c;
    c.lo = make_float4(fminf(s.a.x, s.b.x) - r, fminf(s.a.y, s.b.y) - r, fminf(s.a.z, s.b.z) - r, __uint_as_float(0x80000000u | i));   // the record's own AABB, as in k_fit
    c.hi = make_float4(fmaxf(s.a.x, s.b.x) + r, fmaxf(s.a.y, s.b.y) + r, fmaxf(s.a.z, s.b.z) + r, __uint_as_float(0u));
    out[i] = c;
}

__global__ void k_ploc_nearest(const PlocCluster* cl, uint32_t m, int radius, uint32_t* nn) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float4 lo = cl[i].lo, hi = cl[i].hi;
    const int j0 = int(i) - radius < 0 ? 0 : int(i) - radius, j1 = int(i) + radius >= int(m) ? int(m) - 1 : int(i) + radius;
    float best = INFINITY; uint32_t arg = i == 0 ? 1u : i - 1u;
    for (int j = j0; j <= j1; j++) {
        if (j == int(i)) continue;
        const float4 l2 = cl[j].lo, h2 = cl[j].hi;
        const float dx = fmaxf(hi.x, h2.x) - fminf(lo.x, l2.x), dy = fmaxf(hi.y, h2.y) - fminf(lo.y, l2.y), dz = fmaxf(hi.z, h2.z) - fminf(lo.z, l2.z);
        const float a = dx * dy + dy * dz + dz * dx;     // the same expression for (i, j) and (j, i): mutual choices are consistent
        if (a < best) { best = a; arg = uint32_t(j); }  // ties: the lowest index
    }
    nn[i] = arg;
}

// flags[i]: bit 0 = cluster i survives the round (it is not the higher-indexed half of a merging pair), bit 32 = it merges (lower half)
__global__ void k_ploc_flags(const uint32_t* nn, uint32_t m, unsigned long long* flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t j = nn[i];
    const bool mutual = nn[j] == i;
    flags[i] = (mutual && j < i ? 0ull : 1ull) | (mutual && i < j ? (1ull << 32) : 0ull);
}

// scan[i] = exclusive sums of flags (low word: output slot; high word: how many merges precede).  Nodes are numbered backwards from
// n_inner - 1, so that the last merge -- the root -- is node 0.
__global__ void k_ploc_apply(const PlocCluster* in, const uint32_t* nn, const unsigned long long* flags, const unsigned long long* scan, uint32_t m,
                             uint32_t created, uint32_t n_inner, Node64* nodes, PlocCluster* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long f = flags[i];
    if (!(f & 1ull)) return;                               // merged into its partner
    PlocCluster c = in[i];
    if (f >> 32) {
        const PlocCluster d = in[nn[i]];
        const uint32_t node = n_inner - 1u - (created + uint32_t(scan[i] >> 32));
        Node64 nd;
        nd.l0 = c.lo; nd.l1 = make_float4(c.hi.x, c.hi.y, c.hi.z, __uint_as_float((__float_as_uint(c.lo.w) & 0x80000000u) ? 1u : 0u));
        nd.r0 = d.lo; nd.r1 = make_float4(d.hi.x, d.hi.y, d.hi.z, __uint_as_float((__float_as_uint(d.lo.w) & 0x80000000u) ? 1u : 0u));
        nodes[node] = nd;
        const uint32_t h = max(__float_as_uint(c.hi.w), __float_as_uint(d.hi.w)) + 1u;
        c.lo = make_float4(fminf(c.lo.x, d.lo.x), fminf(c.lo.y, d.lo.y), fminf(c.lo.z, d.lo.z), __uint_as_float(node));
        c.hi = make_float4(fmaxf(c.hi.x, d.hi.x), fmaxf(c.hi.y, d.hi.y), fmaxf(c.hi.z, d.hi.z), __uint_as_float(h));
    }
    out[uint32_t(scan[i])] = c;
}

// ---- 4-wide quantised tree (NodeW4) -------------------------------------------------------------------------------------------
// Dequantisation of a 16-bit grid coordinate q: 0x4B000000 | q is the float 8388608 + q (exact), so ONE logic operation + ONE fma
// replace an integer-to-float conversion (a quarter-rate instruction; it is what made the 32-byte NodeQ kernel lose) + an fma:
//     bound = fma(as_float(0x4B000000 | q), scale, origin_m),   origin_m = origin - 8388608 * scale (rounded once, on the host).
// The builder rounds against exactly this expression, so its rounding errors are part of the grid, not of the bound.
LV_DEV float w4_dequant(uint32_t q, float scale, float origin_m) { return __fmaf_rn(__uint_as_float(0x4B000000u | q), scale, origin_m); }
// the same value from a packed word lo | hi << 16: `sel` = 0x7610 takes the low half, 0x7632 the high half (one PRMT)
LV_DEV float w4_dequant_packed(uint32_t w, uint32_t sel, float scale, float origin_m) {
    return __fmaf_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, sel)), scale, origin_m);
}
constexpr uint32_t kW4SelLo = 0x7610u, kW4SelHi = 0x7632u;

__device__ __forceinline__ uint32_t w4_quantize(float b, float o, float s, float om, bool up) {
    float g = (b - o) / s;
    g = up ? ceilf(g) : floorf(g);
    int q = int(fminf(fmaxf(g, 0.0f), 65535.0f));
    if (up) { while (q < 65535 && w4_dequant(uint32_t(q), s, om) < b) q++; while (q > 0 && w4_dequant(uint32_t(q - 1), s, om) >= b) q--; }
    else { while (q > 0 && w4_dequant(uint32_t(q), s, om) > b) q--; while (q < 65535 && w4_dequant(uint32_t(q + 1), s, om) <= b) q++; }
    return uint32_t(q);
}

struct W4Grid { float o[3], s[3], om[3]; };

// One breadth-first round of the collapse: every queue item (binary node, wide-node index, stack need of the path so far) becomes a
// wide node.  Its children start as the binary node's two; while there are fewer than four, the inner child with the largest surface
// area is replaced by its own two children (the usual greedy collapse of a binary BVH into a wide one).  Inner children get
// consecutive wide indices from `alloc` and go to the next round's queue.  `need` = the worst-case traversal stack: a step pushes
// every hit child but the one it descends into, so a path needs sum (children - 1) entries.
__global__ void k_w4_round(const Node64* nodes, const uint3* in_q, uint32_t n_in, uint3* out_q, unsigned int* out_count, unsigned int* alloc,
                           NodeW4* wnodes, unsigned int* need, const W4Grid G, uint32_t n_rec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const uint3 item = in_q[i];
    float4 lo[4], hi[4];
    int n = 0;
    {
        const Node64 nd = nodes[item.x];
        if (nd.l0.x != INFINITY) { lo[n] = nd.l0; hi[n] = nd.l1; n++; }
        if (nd.r0.x != INFINITY) { lo[n] = nd.r0; hi[n] = nd.r1; n++; }
    }
    while (n < 4) {
        int pick = -1; float best = -1.0f;
        for (int k = 0; k < n; k++) {
            if (__float_as_uint(lo[k].w) & 0x80000000u) continue;   // a leaf
            const float dx = hi[k].x - lo[k].x, dy = hi[k].y - lo[k].y, dz = hi[k].z - lo[k].z;
            const float a = dx * dy + dy * dz + dz * dx;
            if (a > best) { best = a; pick = k; }
        }
        if (pick < 0) break;
        const Node64 nd = nodes[__float_as_uint(lo[pick].w)];
        const bool hl = nd.l0.x != INFINITY, hr = nd.r0.x != INFINITY;
        if (hl && hr) { lo[pick] = nd.l0; hi[pick] = nd.l1; lo[n] = nd.r0; hi[n] = nd.r1; n++; }
        else if (hl) { lo[pick] = nd.l0; hi[pick] = nd.l1; }
        else if (hr) { lo[pick] = nd.r0; hi[pick] = nd.r1; }
        else break;
    }
    int n_inner = 0;
    for (int k = 0; k < n; k++) n_inner += (__float_as_uint(lo[k].w) & 0x80000000u) ? 0 : 1;
    uint32_t wbase = 0, qbase = 0;
    if (n_inner) { wbase = atomicAdd(alloc, (unsigned)n_inner); qbase = atomicAdd(out_count, (unsigned)n_inner); }
    const uint32_t path = item.z + uint32_t(n > 0 ? n - 1 : 0);
    atomicMax(need, path);
    NodeW4 w;
    int inner_i = 0;
    for (int k = 0; k < 4; k++) {
        uint32_t* bw = w.w + 3 * k;
        uint32_t& cw = w.w[12 + k];
        // an absent child is an EMPTY box (lo = 65535 > hi = 0 on every axis: its near bound lies behind its far bound whatever the ray's signs
        // are, so the slab test fails without a look at the child word); the word refers to the all-NaN dummy record, like k_emit_nodes'
        if (k >= n) { bw[0] = bw[1] = bw[2] = 0x0000FFFFu; cw = 0x80000000u | n_rec; continue; }
        const uint32_t ax = w4_quantize(lo[k].x, G.o[0], G.s[0], G.om[0], false), ay = w4_quantize(lo[k].y, G.o[1], G.s[1], G.om[1], false),
                       az = w4_quantize(lo[k].z, G.o[2], G.s[2], G.om[2], false);
        const uint32_t bx = w4_quantize(hi[k].x, G.o[0], G.s[0], G.om[0], true), by = w4_quantize(hi[k].y, G.o[1], G.s[1], G.om[1], true),
                       bz = w4_quantize(hi[k].z, G.o[2], G.s[2], G.om[2], true);
        bw[0] = ax | (bx << 16); bw[1] = ay | (by << 16); bw[2] = az | (bz << 16);
        const uint32_t word = __float_as_uint(lo[k].w);
        if (word & 0x80000000u) cw = word;
        else {
            cw = wbase + uint32_t(inner_i);
            out_q[qbase + uint32_t(inner_i)] = make_uint3(word, wbase + uint32_t(inner_i), path);
            inner_i++;
        }
    }
    wnodes[item.y] = w;
}

}  // namespace lv
