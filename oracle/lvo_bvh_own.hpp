/*
 * lvo_bvh_own.hpp -- the oracle's own BVH backend (binned-SAH build, stack traversal).
 * TEST INFRASTRUCTURE ONLY (see lvo_shaders.hpp).  Stands in for the Vulkan driver's opaque acceleration
 * structure (src/LineData/LineData.cpp:879-907, 1057-1075); portable, needs nothing outside this repo.
 * The alternative backend lvo_bvh_ref.hpp runs the same drivers on the reference's submodules/bvh library.
 */
#ifndef LVO_BVH_OWN_HPP
#define LVO_BVH_OWN_HPP
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "lvo_shaders.hpp"

namespace lvo {

struct Segment { vec3 p0; float a0; vec3 p1; float a1; };

struct Node {
    float bmin[3], bmax[3];
    uint32_t left;   // inner: index of left child (right = left + 1); leaf: first primitive
    uint32_t count;  // 0 for inner nodes
};

struct Scene {
    std::vector<Segment> segs;
    std::vector<Node> nodes;
    std::vector<uint32_t> primIdx;
    float lineWidth;
};

struct Box {
    float mn[3], mx[3];
    void reset() { for (int k = 0; k < 3; k++) { mn[k] = 3.4e38f; mx[k] = -3.4e38f; } }
    void grow(const Box& o) { for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], o.mn[k]); mx[k] = std::max(mx[k], o.mx[k]); } }
    void growPt(const float* p) { for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); } }
    float halfArea() const {
        float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

// AABB of a segment: min/max(p0,p1) -+ lineWidth/2 -- src/LineData/LineDataFlow.cpp:2230-2233
inline Box segmentBox(const Segment& s, float lineWidth) {
    Box b;
    float r = lineWidth * 0.5f;
    const float p0[3] = {s.p0.x, s.p0.y, s.p0.z}, p1[3] = {s.p1.x, s.p1.y, s.p1.z};
    for (int k = 0; k < 3; k++) { b.mn[k] = std::fmin(p0[k], p1[k]) - r; b.mx[k] = std::fmax(p0[k], p1[k]) + r; }
    return b;
}

struct Builder {
    Scene& sc;
    std::vector<Box> boxes;
    std::vector<float> cen;  // 3 per prim
    explicit Builder(Scene& s) : sc(s) {}

    void build() {
        size_t n = sc.segs.size();
        boxes.resize(n); cen.resize(3 * n); sc.primIdx.resize(n);
        for (size_t i = 0; i < n; i++) {
            boxes[i] = segmentBox(sc.segs[i], sc.lineWidth);
            for (int k = 0; k < 3; k++) cen[3 * i + k] = 0.5f * (boxes[i].mn[k] + boxes[i].mx[k]);
            sc.primIdx[i] = uint32_t(i);
        }
        sc.nodes.clear();
        sc.nodes.reserve(2 * n + 2);
        sc.nodes.push_back(Node{});
        if (n == 0) {
            Node& r = sc.nodes[0];
            for (int k = 0; k < 3; k++) { r.bmin[k] = 0; r.bmax[k] = 0; }
            r.left = 0; r.count = 0;
            return;
        }
        // iterative to survive deep trees
        struct Task { uint32_t node, begin, end; };
        std::vector<Task> stack;
        stack.push_back({0, 0, uint32_t(n)});
        while (!stack.empty()) {
            Task t = stack.back(); stack.pop_back();
            Box nb; nb.reset(); Box cb; cb.reset();
            for (uint32_t i = t.begin; i < t.end; i++) { uint32_t p = sc.primIdx[i]; nb.grow(boxes[p]); cb.growPt(&cen[3 * p]); }
            Node nd;
            for (int k = 0; k < 3; k++) { nd.bmin[k] = nb.mn[k]; nd.bmax[k] = nb.mx[k]; }
            uint32_t cnt = t.end - t.begin;
            int axis = 0; float ext = -1;
            for (int k = 0; k < 3; k++) { float e = cb.mx[k] - cb.mn[k]; if (e > ext) { ext = e; axis = k; } }
            if (cnt <= 4 || !(ext > 0.0f)) {
                nd.left = t.begin; nd.count = cnt; sc.nodes[t.node] = nd; continue;
            }
            // binned SAH over the widest centroid axis
            const int NB = 16;
            Box bb[NB]; uint32_t bc[NB];
            for (int b = 0; b < NB; b++) { bb[b].reset(); bc[b] = 0; }
            float k0 = cb.mn[axis], k1 = float(NB) / ext;
            auto binOf = [&](uint32_t p) { int b = int((cen[3 * p + axis] - k0) * k1); return b < 0 ? 0 : (b >= NB ? NB - 1 : b); };
            for (uint32_t i = t.begin; i < t.end; i++) { uint32_t p = sc.primIdx[i]; int b = binOf(p); bb[b].grow(boxes[p]); bc[b]++; }
            float rightA[NB]; uint32_t rightC[NB];
            Box acc; acc.reset(); uint32_t c = 0;
            for (int b = NB - 1; b > 0; b--) { acc.grow(bb[b]); c += bc[b]; rightA[b] = c ? acc.halfArea() : 0.f; rightC[b] = c; }
            acc.reset(); c = 0;
            float best = 3.4e38f; int bestSplit = -1;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bb[b]); c += bc[b];
                if (c == 0 || rightC[b + 1] == 0) continue;
                float cost = acc.halfArea() * float(c) + rightA[b + 1] * float(rightC[b + 1]);
                if (cost < best) { best = cost; bestSplit = b; }
            }
            uint32_t mid;
            if (bestSplit < 0) {
                mid = t.begin + cnt / 2;
                std::nth_element(sc.primIdx.begin() + t.begin, sc.primIdx.begin() + mid, sc.primIdx.begin() + t.end,
                                 [&](uint32_t a, uint32_t b) { return cen[3 * a + axis] < cen[3 * b + axis]; });
            } else {
                auto it = std::partition(sc.primIdx.begin() + t.begin, sc.primIdx.begin() + t.end,
                                         [&](uint32_t p) { return binOf(p) <= bestSplit; });
                mid = uint32_t(it - sc.primIdx.begin());
                if (mid == t.begin || mid == t.end) mid = t.begin + cnt / 2;
            }
            uint32_t l = uint32_t(sc.nodes.size());
            sc.nodes.push_back(Node{}); sc.nodes.push_back(Node{});
            nd.left = l; nd.count = 0;
            sc.nodes[t.node] = nd;
            stack.push_back({l, t.begin, mid});
            stack.push_back({l + 1, mid, t.end});
        }
    }
};

struct RayStats { uint64_t steps = 0, isect = 0, rays = 0; };

struct Ray {
    vec3 o, d; float tmin, tmax;
    RayInv ri;
    void prep() { ri = makeRayInv(o, d); }
};

inline bool slab(const Node& n, const Ray& r, float tmax, float& tnear) {
    return slabTest(r.ri, n.bmin, n.bmax, r.tmin, tmax, tnear);   // canonical test: conservative w.r.t. acceptCandidate
}

struct Hit { float t; uint32_t prim; int kind; };

// closest hit with hitT in [tmin, tmax]; ties -> lowest primitive index
inline bool traceClosest(const Scene& sc, Ray r, bool capped, Hit& best, RayStats& st, bool tieSafe = true) {
    const float margin = tieSafe ? sc.lineWidth : 0.0f;   // AO rays only need the distance: no tie rule, no margin
    st.rays++;
    best.t = r.tmax; best.prim = 0xFFFFFFFFu; best.kind = 0;
    bool found = false;
    if (sc.segs.empty()) return false;
    r.prep();
    const float radius = sc.lineWidth * 0.5f;
    uint32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        uint32_t ni = stack[--sp];
        const Node& n = sc.nodes[ni];
        float tn;
        // cull against best.t + one tube diameter so that tying candidates are always seen (closest-hit rule)
        if (!slab(n, r, best.t + margin, tn)) continue;
        if (n.count) {
            st.isect += n.count;
            for (uint32_t i = 0; i < n.count; i++) {
                uint32_t p = sc.primIdx[n.left + i];
                const Segment& s = sc.segs[p];
                float t; int kind;
                if (acceptCandidate(r.o, r.d, r.ri, s.p0, s.p1, radius, capped, r.tmin, r.tmax, t, kind)) {
                    if (!found || t < best.t || (t == best.t && p < best.prim)) { best.t = t; best.prim = p; best.kind = kind; found = true; }
                }
            }
        } else {
            st.steps++;
            float t0, t1;
            bool h0 = slab(sc.nodes[n.left], r, best.t + margin, t0), h1 = slab(sc.nodes[n.left + 1], r, best.t + margin, t1);
            if (h0 && h1) {
                if (t0 <= t1) { stack[sp++] = n.left + 1; stack[sp++] = n.left; } else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
            } else if (h0) stack[sp++] = n.left;
            else if (h1) stack[sp++] = n.left + 1;
        }
    }
    return found;
}

inline bool traceAny(const Scene& sc, Ray r, bool capped, RayStats& st) {
    st.rays++;
    if (sc.segs.empty()) return false;
    r.prep();
    const float radius = sc.lineWidth * 0.5f;
    uint32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Node& n = sc.nodes[stack[--sp]];
        float tn;
        if (!slab(n, r, r.tmax, tn)) continue;
        if (n.count) {
            st.isect += n.count;
            for (uint32_t i = 0; i < n.count; i++) {
                const Segment& s = sc.segs[sc.primIdx[n.left + i]];
                float t; int kind;
                if (acceptCandidate(r.o, r.d, r.ri, s.p0, s.p1, radius, capped, r.tmin, r.tmax, t, kind)) return true;
            }
        } else { st.steps++; stack[sp++] = n.left; stack[sp++] = n.left + 1; }
    }
    return false;
}

// all candidates whose reported hitT lies in [tmin, tmax] (PPLL fragment source)
template <class F>
void traceAll(const Scene& sc, Ray r, bool capped, RayStats& st, F&& f) {
    st.rays++;
    if (sc.segs.empty()) return;
    r.prep();
    const float radius = sc.lineWidth * 0.5f;
    uint32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Node& n = sc.nodes[stack[--sp]];
        float tn;
        if (!slab(n, r, r.tmax, tn)) continue;
        if (n.count) {
            st.isect += n.count;
            for (uint32_t i = 0; i < n.count; i++) {
                uint32_t p = sc.primIdx[n.left + i];
                const Segment& s = sc.segs[p];
                float t; int kind;
                if (acceptCandidate(r.o, r.d, r.ri, s.p0, s.p1, radius, capped, r.tmin, r.tmax, t, kind)) f(p, t, kind);
            }
        } else { st.steps++; stack[sp++] = n.left; stack[sp++] = n.left + 1; }
    }
}


inline void buildScene(Scene& sc) { Builder b(sc); b.build(); }
inline const char* backendName() { return "own-binned-sah"; }
inline uint64_t numNodes(const Scene& sc) { return sc.nodes.size(); }

}  // namespace lvo
#endif
