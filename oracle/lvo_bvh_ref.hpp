/*
 * lvo_bvh_ref.hpp -- BVH backend of the oracle that runs on the REFERENCE's own CPU BVH library,
 * madmann91/bvh, compiled from the sources where they lie: /root/reference/submodules/bvh/include
 * (pinned @2fd0db6 by the reference's .SUBMODULES.json).  Only oracle/Makefile's `ref` target uses this
 * header; its output goes to oracle/_ref/ (git-ignored).  No reference source is copied into this repo.
 *
 * TEST INFRASTRUCTURE ONLY (see lvo_shaders.hpp).  This is the "reference's own CPU path" for BVH
 * build + traversal: builders bvh::SweepSahBuilder (<= 1 M segments) /
 * bvh::BinnedSahBuilder<Bvh,16> (larger), traversal bvh::SingleRayTraverser
 * (include/bvh/single_ray_traverser.hpp:15-166) with its Statistics{traversal_steps, intersections}, and a
 * custom primitive in the pattern of submodules/bvh/test/custom_primitive.cpp:17-46 whose intersect() is the
 * restated IntersectionTube shader.
 */
#ifndef LVO_BVH_REF_HPP
#define LVO_BVH_REF_HPP
#include <cstdint>
#include <memory>
#include <optional>
#include <vector>

#include <bvh/bvh.hpp>
#include <bvh/vector.hpp>
#include <bvh/ray.hpp>
#include <bvh/sweep_sah_builder.hpp>
#include <bvh/binned_sah_builder.hpp>
#include <bvh/single_ray_traverser.hpp>
#include <bvh/primitive_intersectors.hpp>

#include "lvo_shaders.hpp"

namespace lvo {

using BScalar = float;
using BVec3 = bvh::Vector3<BScalar>;
using BBox = bvh::BoundingBox<BScalar>;
using BRay = bvh::Ray<BScalar>;
using BBvh = bvh::Bvh<BScalar>;

struct Segment { vec3 p0; float a0; vec3 p1; float a1; };

// custom primitive -- pattern: submodules/bvh/test/custom_primitive.cpp:17-46
struct TubePrimitive {
    struct Intersection {
        BScalar t; int kind;
        BScalar distance() const { return t; }
    };
    using ScalarType = BScalar;
    using IntersectionType = Intersection;

    vec3 p0, p1; float radius; bool capped;

    BVec3 center() const { return BVec3(0.5f * (p0.x + p1.x), 0.5f * (p0.y + p1.y), 0.5f * (p0.z + p1.z)); }
    // AABB = min/max(p0,p1) -+ lineWidth/2 -- src/LineData/LineDataFlow.cpp:2230-2233
    BBox bounding_box() const {
        return BBox(BVec3(std::fmin(p0.x, p1.x) - radius, std::fmin(p0.y, p1.y) - radius, std::fmin(p0.z, p1.z) - radius),
                    BVec3(std::fmax(p0.x, p1.x) + radius, std::fmax(p0.y, p1.y) + radius, std::fmax(p0.z, p1.z) + radius));
    }
    // acceptance rule of lvo_shaders.hpp (own-AABB slab test on the ORIGINAL interval + IntersectionTube + range check);
    // ray.tmax has meanwhile been shortened to the best hit by the traverser, which only removes non-improving candidates
    std::optional<Intersection> intersect(const BRay& ray, const RayInv& ri, float otmin, float otmax) const {
        float t; int kind;
        vec3 ro = V3(ray.origin[0], ray.origin[1], ray.origin[2]), rd = V3(ray.direction[0], ray.direction[1], ray.direction[2]);
        if (acceptCandidate(ro, rd, ri, p0, p1, radius, capped, otmin, otmax, t, kind))
            return std::make_optional(Intersection{t, kind});
        return std::nullopt;
    }
};

struct Scene {
    std::vector<Segment> segs;
    std::vector<TubePrimitive> prims;
    BBvh bvh;
    float lineWidth;
};

struct RayStats { uint64_t steps = 0, isect = 0, rays = 0; };
struct Ray { vec3 o, d; float tmin, tmax; };
struct Hit { float t; uint32_t prim; int kind; };

inline void buildScene(Scene& sc) {
    size_t n = sc.segs.size();
    sc.prims.resize(n);
    for (size_t i = 0; i < n; i++) sc.prims[i] = TubePrimitive{sc.segs[i].p0, sc.segs[i].p1, sc.lineWidth * 0.5f, true};
    if (n == 0) return;
    auto [bboxes, centers] = bvh::compute_bounding_boxes_and_centers(sc.prims.data(), n);
    auto global_bbox = bvh::compute_bounding_boxes_union(bboxes.get(), n);
    if (n <= 1000000) {
        bvh::SweepSahBuilder<BBvh> builder(sc.bvh);
        builder.build(global_bbox, bboxes.get(), centers.get(), n);
    } else {
        bvh::BinnedSahBuilder<BBvh, 16> builder(sc.bvh);
        builder.build(global_bbox, bboxes.get(), centers.get(), n);
    }
}
inline const char* backendName() { return "reference-madmann91-bvh"; }
inline uint64_t numNodes(const Scene& sc) { return sc.bvh.node_count; }

// Closest hit with the tie rule "lowest segment index": a later candidate replaces the best one only if it
// is strictly closer or equally close with a lower index.  (bvh::ClosestPrimitiveIntersector,
// include/bvh/primitive_intersectors.hpp:33-55, keeps whichever equal-t candidate it meets last.)
struct TieBreakClosestIntersector {
    struct Result {
        size_t primitive_index; TubePrimitive::Intersection intersection; BScalar margin;
        // what the traverser shortens ray.tmax to (single_ray_traverser.hpp:59): one tube diameter beyond the hit, so that
        // candidates tying the best hit are never culled by the library's node test (closest-hit rule)
        BScalar distance() const { return intersection.distance() + margin; }
    };
    static constexpr bool any_hit = false;
    const Scene& sc; bool capped; float otmin, otmax, margin; RayInv ri;
    bool have = false; float bestT = 0; size_t bestPrim = 0;
    TieBreakClosestIntersector(const Scene& s, bool c, float a, float b, float m, const RayInv& r) : sc(s), capped(c), otmin(a), otmax(b), margin(m), ri(r) {}
    std::optional<Result> intersect(size_t index, const BRay& ray) {
        size_t p = sc.bvh.primitive_indices[index];
        TubePrimitive prim = sc.prims[p]; prim.capped = capped;
        if (auto hit = prim.intersect(ray, ri, otmin, otmax)) {
            if (!have || hit->t < bestT || (hit->t == bestT && p < bestPrim)) {
                have = true; bestT = hit->t; bestPrim = p;
                return std::make_optional(Result{p, *hit, margin});
            }
        }
        return std::nullopt;
    }
};

struct AnyIntersector {
    struct Result { BScalar t; BScalar distance() const { return t; } };
    static constexpr bool any_hit = true;
    const Scene& sc; bool capped; float otmin, otmax; RayInv ri;
    AnyIntersector(const Scene& s, bool c, float a, float b, const RayInv& r) : sc(s), capped(c), otmin(a), otmax(b), ri(r) {}
    std::optional<Result> intersect(size_t index, const BRay& ray) {
        TubePrimitive prim = sc.prims[sc.bvh.primitive_indices[index]]; prim.capped = capped;
        if (auto hit = prim.intersect(ray, ri, otmin, otmax)) return std::make_optional(Result{hit->t});
        return std::nullopt;
    }
};

template <class F>
struct AllIntersector {
    struct Result { BScalar t; BScalar distance() const { return t; } };
    static constexpr bool any_hit = false;
    const Scene& sc; bool capped; float otmin, otmax; RayInv ri; F& f;
    AllIntersector(const Scene& s, bool c, float a, float b, const RayInv& r, F& fn) : sc(s), capped(c), otmin(a), otmax(b), ri(r), f(fn) {}
    std::optional<Result> intersect(size_t index, const BRay& ray) {
        size_t p = sc.bvh.primitive_indices[index];
        TubePrimitive prim = sc.prims[p]; prim.capped = capped;
        if (auto hit = prim.intersect(ray, ri, otmin, otmax)) f(uint32_t(p), hit->t, hit->kind);
        return std::nullopt;  // never shrink the interval: every candidate must be visited
    }
};

// RobustNodeIntersector (include/bvh/node_intersectors.hpp:56-78, T. Ize's robust BVH traversal): the library's box test
// must not reject a box whose segment the canonical acceptance rule admits.
using Traverser = bvh::SingleRayTraverser<BBvh, 64, bvh::RobustNodeIntersector<BBvh>>;

inline bool traceClosest(const Scene& sc, Ray r, bool capped, Hit& best, RayStats& st, bool tieSafe = true) {
    st.rays++;
    best.t = r.tmax; best.prim = 0xFFFFFFFFu; best.kind = 0;
    if (sc.segs.empty()) return false;
    BRay ray(BVec3(r.o.x, r.o.y, r.o.z), BVec3(r.d.x, r.d.y, r.d.z), r.tmin, r.tmax);
    TieBreakClosestIntersector isect(sc, capped, r.tmin, r.tmax, tieSafe ? sc.lineWidth : 0.0f, makeRayInv(r.o, r.d));
    Traverser trav(sc.bvh);
    Traverser::Statistics s;
    auto hit = trav.traverse(ray, isect, s);
    st.steps += s.traversal_steps; st.isect += s.intersections;
    if (!hit) return false;
    best.t = hit->intersection.t; best.prim = uint32_t(hit->primitive_index); best.kind = hit->intersection.kind;
    return true;
}

inline bool traceAny(const Scene& sc, Ray r, bool capped, RayStats& st) {
    st.rays++;
    if (sc.segs.empty()) return false;
    BRay ray(BVec3(r.o.x, r.o.y, r.o.z), BVec3(r.d.x, r.d.y, r.d.z), r.tmin, r.tmax);
    AnyIntersector isect(sc, capped, r.tmin, r.tmax, makeRayInv(r.o, r.d));
    Traverser trav(sc.bvh);
    Traverser::Statistics s;
    auto hit = trav.traverse(ray, isect, s);
    st.steps += s.traversal_steps; st.isect += s.intersections;
    return bool(hit);
}

template <class F>
inline void traceAll(const Scene& sc, Ray r, bool capped, RayStats& st, F&& f) {
    st.rays++;
    if (sc.segs.empty()) return;
    BRay ray(BVec3(r.o.x, r.o.y, r.o.z), BVec3(r.d.x, r.d.y, r.d.z), r.tmin, r.tmax);
    AllIntersector<F> isect(sc, capped, r.tmin, r.tmax, makeRayInv(r.o, r.d), f);
    Traverser trav(sc.bvh);
    Traverser::Statistics s;
    trav.traverse(ray, isect, s);
    st.steps += s.traversal_steps; st.isect += s.intersections;
}

}  // namespace lvo
#endif
