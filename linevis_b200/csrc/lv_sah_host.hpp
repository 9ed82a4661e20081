// lv_sah_host.hpp -- top-down binned-SAH builder on the HOST (b200_bvh_builder = sah; SURVEY 8f rank 1, builder quality).
//
// What it is for: a measurement.  The scene's default tree is the GPU Morton radix tree (10 ms for 10 M segments); the reference's CPU
// library offers a binned surface-area-heuristic builder (submodules/bvh/include/bvh/binned_sah_builder.hpp) whose trees need fewer
// traversal steps.  This builder makes such a tree for the device traversals -- same Node64 layout, one record per leaf, the records
// stay in their Morton order and are referenced by index -- so that the question "what is a SAH-quality tree worth to the AO ray stream"
// has a measured answer.  It runs on the host threads at scene creation (seconds for 10 M segments), never per frame.
//
// Algorithm (the classic one): a node covers a set of records; their AABB centroids are binned into 16 bins along each axis, the split
// plane with the smallest  area(left) * count(left) + area(right) * count(right)  over the 3 x 15 candidates wins, the index array is
// partitioned in place; sets whose centroids coincide are halved.  Nodes are numbered in preorder (a subtree of m leaves owns m - 1
// consecutive nodes), so the two halves of a node can be built by different threads without any shared counter.
// Leaf boxes are the records' own AABBs in exactly the float expressions of k_fit / seg_box_hit (rule 2 of DESIGN.md).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <future>
#include <vector>
#include "lv_types.cuh"

namespace lvsah {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], b.lo[k]); hi[k] = std::fmax(hi[k], b.hi[k]); } }
    float half_area() const { const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dy * dz + dz * dx; }
};

struct Builder {
    static constexpr int kBins = 16;
    const std::vector<Box>& boxes;
    std::vector<uint32_t>& idx;
    lv::Node64* nodes;
    Builder(const std::vector<Box>& b, std::vector<uint32_t>& i, lv::Node64* n) : boxes(b), idx(i), nodes(n) {}

    static float as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

    // builds the subtree over idx[b, e) (e - b >= 2) into nodes[base ..], returns its box and depth
    Box build(uint32_t b, uint32_t e, uint32_t base, uint32_t depth, uint32_t& depth_out, int par_levels) {
        Box cb; cb.reset();   // centroid bounds (twice the centroid: lo + hi)
        for (uint32_t i = b; i < e; i++) {
            const Box& x = boxes[idx[i]];
            for (int k = 0; k < 3; k++) { const float c = x.lo[k] + x.hi[k]; cb.lo[k] = std::fmin(cb.lo[k], c); cb.hi[k] = std::fmax(cb.hi[k], c); }
        }
        uint32_t mid = b + (e - b) / 2;
        bool found = false;
        if (e - b > 2) {
            float best = INFINITY; int best_axis = -1, best_bin = -1;
            for (int axis = 0; axis < 3; axis++) {
                const float ext = cb.hi[axis] - cb.lo[axis];
                if (!(ext > 0.0f)) continue;
                const float scale = float(kBins) / ext;
                Box bb[kBins]; uint32_t cnt[kBins];
                for (int k = 0; k < kBins; k++) { bb[k].reset(); cnt[k] = 0; }
                for (uint32_t i = b; i < e; i++) {
                    const Box& x = boxes[idx[i]];
                    const int k = std::max(0, std::min(kBins - 1, int((x.lo[axis] + x.hi[axis] - cb.lo[axis]) * scale)));
                    bb[k].grow(x); cnt[k]++;
                }
                float right_area[kBins]; uint32_t right_cnt[kBins];
                Box acc; acc.reset(); uint32_t n = 0;
                for (int k = kBins - 1; k > 0; k--) { acc.grow(bb[k]); n += cnt[k]; right_area[k] = acc.half_area(); right_cnt[k] = n; }
                acc.reset(); n = 0;
                for (int k = 0; k < kBins - 1; k++) {
                    acc.grow(bb[k]); n += cnt[k];
                    if (n == 0 || right_cnt[k + 1] == 0) continue;
                    const float cost = acc.half_area() * float(n) + right_area[k + 1] * float(right_cnt[k + 1]);
                    if (cost < best) { best = cost; best_axis = axis; best_bin = k; }
                }
            }
            if (best_axis >= 0) {
                const float lo = cb.lo[best_axis], scale = float(kBins) / (cb.hi[best_axis] - lo);
                uint32_t* first = idx.data() + b;
                uint32_t* m = std::partition(first, idx.data() + e, [&](uint32_t p) {
                    const Box& x = boxes[p];
                    return std::max(0, std::min(kBins - 1, int((x.lo[best_axis] + x.hi[best_axis] - lo) * scale))) <= best_bin;
                });
                const uint32_t split = uint32_t(m - idx.data());
                if (split > b && split < e) { mid = split; found = true; }
            }
        }
        (void)found;   // no usable plane (coinciding centroids): the set is halved as it lies
        const uint32_t nl = mid - b, nr = e - mid;
        Box lb, rb; uint32_t ld = depth, rd = depth;
        uint32_t lword, rword;
        auto left = [&]() {
            if (nl == 1) { lb = boxes[idx[b]]; lword = 0x80000000u | idx[b]; }
            else { lword = base + 1; lb = build(b, mid, base + 1, depth + 1, ld, par_levels - 1); }
        };
        auto right = [&]() {
            if (nr == 1) { rb = boxes[idx[mid]]; rword = 0x80000000u | idx[mid]; }
            else { rword = base + nl; rb = build(mid, e, base + nl, depth + 1, rd, par_levels - 1); }   // the left subtree owns nl - 1 nodes after `base`
        };
        if (par_levels > 0 && e - b > 65536) {
            auto fut = std::async(std::launch::async, left);
            right();
            fut.get();
        } else { left(); right(); }
        lv::Node64 nd;
        nd.l0 = make_float4(lb.lo[0], lb.lo[1], lb.lo[2], as_float(lword)); nd.l1 = make_float4(lb.hi[0], lb.hi[1], lb.hi[2], as_float(nl == 1 ? 1u : 0u));
        nd.r0 = make_float4(rb.lo[0], rb.lo[1], rb.lo[2], as_float(rword)); nd.r1 = make_float4(rb.hi[0], rb.hi[1], rb.hi[2], as_float(nr == 1 ? 1u : 0u));
        nodes[base] = nd;
        depth_out = depth;
        if (nl > 1) depth_out = std::max(depth_out, ld);
        if (nr > 1) depth_out = std::max(depth_out, rd);
        Box u = lb; u.grow(rb);
        return u;
    }
};

// segs: n records of 8 floats (SegRec: p0.xyz, attr0, p1.xyz, attr1) in the scene's record order; nodes: room for n - 1 (n >= 2).
// Returns the depth of the tree (edges from the root to the deepest inner node).
inline uint32_t build(const float* segs, uint32_t n, float r, lv::Node64* nodes, int threads_log2 = 5) {
    std::vector<Box> boxes(n);
    for (uint32_t i = 0; i < n; i++) {
        const float* s = segs + size_t(i) * 8;
        for (int k = 0; k < 3; k++) { boxes[i].lo[k] = std::fmin(s[k], s[4 + k]) - r; boxes[i].hi[k] = std::fmax(s[k], s[4 + k]) + r; }   // as k_fit
    }
    std::vector<uint32_t> idx(n);
    for (uint32_t i = 0; i < n; i++) idx[i] = i;
    Builder B(boxes, idx, nodes);
    uint32_t depth = 0;
    B.build(0, n, 0, 0, depth, threads_log2);
    return depth;
}

}  // namespace lvsah
