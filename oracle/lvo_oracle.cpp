/*
 * lvo_oracle.cpp -- CPU oracle for the LineVis hot path (tubes + RTAO, PPLL OIT).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for tests/, __graft_entry__.smoke() and bench.py's CPU legs.
 * The product (linevis_b200/) never links or calls this.  PARITY UNPINNED by the reference's own tests
 * (none exist for this path); see lvo_shaders.hpp for what is pinned instead.
 *
 * The shading / intersection / RNG / PPLL arithmetic is restated in lvo_shaders.hpp and lvo_sort.hpp.
 * This file adds what the reference delegates to the Vulkan driver (BVH build + traversal: the
 * reference's BVH is an opaque VkAccelerationStructureKHR, src/LineData/LineData.cpp:879-907) with a
 * plain binned-SAH BVH, and the per-pixel drivers of the ray-gen / compute / resolve shaders.
 * Result-defining rules the hardware leaves open are fixed here and mirrored by the CUDA side:
 *   - a candidate is accepted iff its reported hitT lies in [tMin, tMax] (reportIntersectionEXT);
 *   - closest hit = smallest hitT, ties broken by the lowest segment index.
 *
 * Build: g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <atomic>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/linevis_b200.h"
#include "lvo_shaders.hpp"
#include "lvo_sort.hpp"
#include "lvo_tritubes.hpp"
#ifdef LVO_USE_REFERENCE_BVH
#include "lvo_bvh_ref.hpp"   // traversal by the reference's submodules/bvh (built into oracle/_ref only)
#else
#include "lvo_bvh_own.hpp"
#endif

using namespace lvo;

extern "C" {

typedef struct lvo_options {
    int32_t use_capped_tubes;           // LineData::useCappedTubes (LineData.hpp:377)
    int32_t use_halos;                  // LineData::useHalos (LineData.hpp:378)
    float ao_strength;                  // ambient_occlusion_strength (LineRenderer.cpp:476)
    float ao_gamma;                     // ambient_occlusion_gamma
    float ao_radius;                    // ambient_occlusion_radius (VulkanRayTracedAmbientOcclusion.hpp:151)
    uint32_t ao_spp;                    // ambient_occlusion_samples_per_frame (:150)
    int32_t ao_use_distance;            // ambient_occlusion_distance_based (:152)
    int32_t ao_jitter_primary;          // use_jittered_primary_rays (:153)
    uint32_t tube_num_subdivisions;     // LineData::tubeNumSubdivisions (LineData.cpp:52)
    uint32_t num_samples_per_frame;     // num_samples_per_frame (VulkanRayTracer.hpp:137)
    int32_t use_jittered_rays;          // USE_JITTERED_RAYS (VulkanRayTracer.cpp:421)
    int32_t use_deterministic_sampling; // DETERMINISTIC_SAMPLING
    uint32_t max_depth_complexity;      // maxDepthComplexity (VulkanRayTracer.hpp:139)
    uint32_t tile_w, tile_h;            // LineRenderer::tileWidth/tileHeight (LineRenderer.cpp:739-740)
    float depth_cue_strength;           // depth_cue_strength (LineRenderer.cpp:449-460); 0 = USE_DEPTH_CUES off
    int32_t use_static_ao;              // ambient_occlusion_mode == "RTAO (Prebaker)": STATIC_AMBIENT_OCCLUSION_PREBAKING
} lvo_options;

// AmbientOcclusionComputeRenderPass settings (VulkanAmbientOcclusionBaker.hpp:163-168)
typedef struct lvo_bake_options {
    float ao_radius;                    // ambientOcclusionRadius
    uint32_t num_tube_subdivisions;     // numTubeSubdivisions
    uint32_t samples_per_frame;         // numAmbientOcclusionSamplesPerFrame
    int32_t use_distance;               // useDistance
} lvo_bake_options;

}  // extern "C"

namespace {

// BVH backend (Scene, buildScene, traceClosest/traceAny/traceAll) comes from the included header.

// Scene + the line-point data of the object-space AO prebaker (LinePointDataUnified: position / tangent / normal per point,
// src/LineData/LineRenderData.hpp:99-106) and its results (ambientOcclusionFactors / blending weights).
struct SceneX : Scene {
    std::vector<vec3> ptPos, ptTan, ptNrm;
    std::vector<SegmentLineData> segLine;          // per segment: point indices + line normals
    std::vector<float> aoFactors, aoBlendingWeights;
    uint32_t numParametrizationVertices = 0, numAoTubeSubdivisions = 0;
};
SceneX& sx(void* h) { return *static_cast<SceneX*>(static_cast<Scene*>(h)); }

// the reference's triangulated tubes + a BVH over the triangles (lvo_tritubes.hpp)
struct TubeMeshScene { TubeMesh mesh; TriBvh bvh; float lineWidth = 0.0f; };

Uniforms makeUniforms(const Scene& sc, const lv_camera& cam, const lvo_options& o, const float* tf, uint32_t K,
                      float amin, float amax, const float* aoTex) {
    Uniforms u{};
    u.cameraPosition = V3(cam.position[0], cam.position[1], cam.position[2]);
    u.fieldOfViewY = cam.fov_y;
    std::memcpy(u.viewMatrix, cam.view, 64); std::memcpy(u.projectionMatrix, cam.proj, 64);
    std::memcpy(u.inverseViewMatrix, cam.inv_view, 64); std::memcpy(u.inverseProjectionMatrix, cam.inv_proj, 64);
    u.backgroundColor = vec4{cam.background[0], cam.background[1], cam.background[2], cam.background[3]};
    // foregroundColor = vec4(1) - backgroundColor -- src/LineData/LineData.cpp:1284-1285
    u.foregroundColor = vec4{1.0f - cam.background[0], 1.0f - cam.background[1], 1.0f - cam.background[2], 1.0f - cam.background[3]};
    u.lineWidth = sc.lineWidth;
    u.ambientOcclusionStrength = o.ao_strength; u.ambientOcclusionGamma = o.ao_gamma;
    u.viewportW = cam.width; u.viewportH = cam.height;
    u.tfLut = tf; u.tfK = K; u.minAttributeValue = amin; u.maxAttributeValue = amax;
    u.useCappedTubes = o.use_capped_tubes != 0; u.useHalos = o.use_halos != 0;
    u.useAmbientOcclusion = (o.ao_strength > 0.0f) && aoTex != nullptr;
    u.aoTexture = aoTex;
    u.useDepthCues = o.depth_cue_strength > 0.0f;
    u.depthCueStrength = o.depth_cue_strength;
    u.staticAmbientOcclusionPrebaking = false;
    if (o.use_static_ao) {
        const SceneX& x = static_cast<const SceneX&>(sc);
        u.useAmbientOcclusion = (o.ao_strength > 0.0f) && !x.aoFactors.empty();
        u.staticAmbientOcclusionPrebaking = u.useAmbientOcclusion;
        u.ambientOcclusionFactors = x.aoFactors.data(); u.ambientOcclusionBlendingWeights = x.aoBlendingWeights.data();
        u.numAoTubeSubdivisions = x.numAoTubeSubdivisions; u.numLineVertices = uint32_t(x.aoBlendingWeights.size());
        u.numParametrizationVertices = x.numParametrizationVertices;
    }
    u.minDepth = 0.0f; u.maxDepth = 1.0f;                                              // LineRenderer.hpp:222-223
    if (u.useDepthCues) {
        // LineRenderer::computeDepthRange (LineRenderer.cpp:410-431) over the line vertices = the end points of all segments
        float dmin = cam.far_dist, dmax = cam.near_dist;
        for (const auto& s : sc.segs) {
            depthRangeOfVertex(cam.view, cam.proj, cam.near_dist, cam.far_dist, s.p0, dmin, dmax);
            depthRangeOfVertex(cam.view, cam.proj, cam.near_dist, cam.far_dist, s.p1, dmin, dmax);
        }
        u.minDepth = dmin; u.maxDepth = dmax;
    }
    return u;
}

void paddedSize(uint32_t W, uint32_t H, uint32_t tw, uint32_t th, uint32_t& pw, uint32_t& ph) {
    // LineRenderer::getScreenSizeWithTiling -- src/Renderers/LineRenderer.cpp:805-812
    pw = W; ph = H;
    if (pw % tw != 0) pw = (pw / tw + 1) * tw;
    if (ph % th != 0) ph = (ph / th + 1) * th;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------ unit helpers
uint32_t lvo_tea(uint32_t v0, uint32_t v1) { return tea(v0, v1); }
void lvo_rnd_stream(uint32_t seed, uint32_t n, uint32_t* lcg_out, float* rnd_out) {
    for (uint32_t i = 0; i < n; i++) {
        uint32_t s2 = seed;
        float f = rnd(s2);        // advances exactly like lcg(seed)
        lcg_out[i] = lcg(seed);
        rnd_out[i] = f;
    }
}
float lvo_det_pow(float x, float y) { return det_pow(x, y); }
void lvo_det_sincos2pi(float xi, float* c, float* s) { det_sincos2pi(xi, *c, *s); }
uint32_t lvo_addr_gen(uint32_t x, uint32_t y, uint32_t viewportW, uint32_t tw, uint32_t th) { return addrGen(x, y, viewportW, tw, th); }
uint32_t lvo_pack_unorm4x8(const float* c) { return packUnorm4x8(vec4{c[0], c[1], c[2], c[3]}); }
void lvo_sample_hemisphere(float xix, float xiy, float* out3) { vec3 v = sampleHemisphere(xix, xiy); out3[0] = v.x; out3[1] = v.y; out3[2] = v.z; }

// IntersectionTube for one ray / one segment.  Returns 1 if an intersection would be reported.
int lvo_intersect_tube(const float* ro, const float* rd, const float* p0, const float* p1, float radius, int capped,
                       float* t_out, int* kind_out) {
    float t; int k;
    bool h = intersectionTube(V3(ro[0], ro[1], ro[2]), V3(rd[0], rd[1], rd[2]), V3(p0[0], p0[1], p0[2]), V3(p1[0], p1[1], p1[2]),
                              radius, capped != 0, t, k);
    *t_out = t; *kind_out = k;
    return h ? 1 : 0;
}

// sort + blend of one pixel's fragment list.  canonical != 0 -> (depth, colour) key order (see lvo_sort.hpp)
void lvo_sort_blend(const uint32_t* colors, const float* depths, uint32_t n, uint32_t max_frags, int mode, int canonical,
                    float* rgba_out) {
    ResolveLists L;
    L.colorList.assign(colors, colors + n); L.depthList.assign(depths, depths + n);
    L.colorList.resize(std::max(n, max_frags)); L.depthList.resize(std::max(n, max_frags));
    vec4 c = canonical ? L.canonical(mode, n) : L.sortingAlgorithm(mode, n, max_frags);
    rgba_out[0] = c.x; rgba_out[1] = c.y; rgba_out[2] = c.z; rgba_out[3] = c.w;
}

// Segment list construction -- LineDataFlow::getLinePassTubeAabbRenderData, src/LineData/LineDataFlow.cpp:2140-2236.
// Input: polylines as (positions, attributes, line_offsets[n_lines+1]).  Output (caller-allocated, worst case sizes):
// filtered points, attributes, tangents, normals, seg_idx pairs.  Returns counts through n_pt_out / n_seg_out.
void lvo_segments_from_polylines(const float* pos, const float* attr, const uint64_t* line_offsets, uint64_t n_lines,
                                 float* pos_out, float* attr_out, float* tangent_out, float* normal_out, uint32_t* seg_idx_out,
                                 uint64_t* n_pt_out, uint64_t* n_seg_out) {
    uint64_t npt = 0, nseg = 0;
    uint32_t lineSegmentIndexCounter = 0;
    for (uint64_t li = 0; li < n_lines; li++) {
        uint64_t b = line_offsets[li], e = line_offsets[li + 1];
        uint64_t n = e - b;
        if (n < 2) continue;  // a one-point trajectory would read positions[i+1] out of bounds in the reference
        vec3 lastLineNormal = V3(1.0f, 0.0f, 0.0f);
        uint32_t numValidLinePoints = 0;
        auto P = [&](uint64_t i) { return V3(pos[3 * (b + i)], pos[3 * (b + i) + 1], pos[3 * (b + i) + 2]); };
        for (uint64_t i = 0; i < n; i++) {
            vec3 tangent;
            if (i == 0) tangent = P(i + 1) - P(i);
            else if (i + 1 == n) tangent = P(i) - P(i - 1);
            else tangent = P(i + 1) - P(i - 1);
            float tangentLength = length(tangent);
            if (tangentLength < 0.0001f) continue;
            tangent = normalize(tangent);
            vec3 helperAxis = lastLineNormal;
            if (length(cross(helperAxis, tangent)) < 0.01f) {
                helperAxis = V3(0.0f, 1.0f, 0.0f);
                if (length(cross(helperAxis, tangent)) < 0.01f) helperAxis = V3(0.0f, 0.0f, 1.0f);
            }
            vec3 normal = normalize(helperAxis - dot(helperAxis, tangent) * tangent);
            lastLineNormal = normal;
            vec3 p = P(i);
            pos_out[3 * npt] = p.x; pos_out[3 * npt + 1] = p.y; pos_out[3 * npt + 2] = p.z;
            attr_out[npt] = attr[b + i];
            if (tangent_out) { tangent_out[3 * npt] = tangent.x; tangent_out[3 * npt + 1] = tangent.y; tangent_out[3 * npt + 2] = tangent.z; }
            if (normal_out) { normal_out[3 * npt] = normal.x; normal_out[3 * npt + 1] = normal.y; normal_out[3 * npt + 2] = normal.z; }
            npt++;
            numValidLinePoints++;
        }
        if (numValidLinePoints == 1) npt--;
        if (numValidLinePoints <= 1) continue;
        for (uint32_t pointIdx = 1; pointIdx < numValidLinePoints; pointIdx++) {
            seg_idx_out[2 * nseg] = lineSegmentIndexCounter + pointIdx - 1;
            seg_idx_out[2 * nseg + 1] = lineSegmentIndexCounter + pointIdx;
            nseg++;
        }
        lineSegmentIndexCounter += numValidLinePoints;
    }
    *n_pt_out = npt; *n_seg_out = nseg;
}

// ------------------------------------------------------------------------------------------ scene
void* lvo_scene_create(const float* pos, const float* attr, const uint32_t* seg_idx, uint64_t n_pt, uint64_t n_seg, float line_width) {
    (void)n_pt;
    SceneX* sc = new SceneX();
    sc->lineWidth = line_width;
    sc->segs.resize(n_seg);
    for (uint64_t i = 0; i < n_seg; i++) {
        uint32_t a = seg_idx[2 * i], b = seg_idx[2 * i + 1];
        sc->segs[i] = Segment{V3(pos[3 * a], pos[3 * a + 1], pos[3 * a + 2]), attr[a], V3(pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]), attr[b]};
    }
    buildScene(*sc);
    sc->segLine.resize(n_seg);
    for (uint64_t i = 0; i < n_seg; i++) sc->segLine[i] = SegmentLineData{seg_idx[2 * i], seg_idx[2 * i + 1], V3(0, 0, 0), V3(0, 0, 0)};
    sc->ptPos.resize(n_pt);
    for (uint64_t i = 0; i < n_pt; i++) sc->ptPos[i] = V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    return static_cast<Scene*>(sc);
}
void lvo_scene_destroy(void* h) { delete &sx(h); }
uint64_t lvo_scene_num_nodes(void* h) { return numNodes(*static_cast<Scene*>(h)); }
const char* lvo_backend_name(void) { return backendName(); }

// Brute-force closest hit (no BVH) for validating the traversal itself on small scenes.
void lvo_trace_primary_bruteforce(void* h, const lv_camera* cam, const lvo_options* o, lv_hit* hits) {
    Scene& sc = *static_cast<Scene*>(h);
    Uniforms u = makeUniforms(sc, *cam, *o, nullptr, 0, 0, 1, nullptr);
    const float radius = sc.lineWidth * 0.5f;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < int64_t(cam->height); y++)
        for (uint32_t x = 0; x < cam->width; x++) {
            vec3 ro, rd; cameraRay(u, x, uint32_t(y), 0.5f, 0.5f, ro, rd);
            lv_hit best{0.0f, 0xFFFFFFFFu, 0, 0}; bool found = false;
            const RayInv ri = makeRayInv(ro, rd);
            for (size_t p = 0; p < sc.segs.size(); p++) {
                float t; int k;
                if (acceptCandidate(ro, rd, ri, sc.segs[p].p0, sc.segs[p].p1, radius, u.useCappedTubes, 0.0001f, 1000.0f, t, k)) {
                    if (!found || t < best.t) { best.t = t; best.prim = uint32_t(p); best.kind = uint32_t(k); found = true; }
                }
            }
            hits[size_t(y) * cam->width + x] = best;
        }
}

// S1+S2 closest-hit pass, pixel-centre rays, tMin 1e-4, tMax 1000 (TubeRayTracing.glsl:52-59).  stats = {T, I, rays}
void lvo_trace_primary(void* h, const lv_camera* cam, const lvo_options* o, lv_hit* hits, uint64_t* stats) {
    Scene& sc = *static_cast<Scene*>(h);
    Uniforms u = makeUniforms(sc, *cam, *o, nullptr, 0, 0, 1, nullptr);
    uint64_t T = 0, I = 0, R = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : T, I, R)
    for (int64_t y = 0; y < int64_t(cam->height); y++) {
        RayStats st;
        for (uint32_t x = 0; x < cam->width; x++) {
            Ray r; cameraRay(u, x, uint32_t(y), 0.5f, 0.5f, r.o, r.d); r.tmin = 0.0001f; r.tmax = 1000.0f;
            Hit hit;
            lv_hit out{0.0f, 0xFFFFFFFFu, 0, 0};
            if (traceClosest(sc, r, u.useCappedTubes, hit, st)) { out.t = hit.t; out.prim = hit.prim; out.kind = uint32_t(hit.kind); }
            hits[size_t(y) * cam->width + x] = out;
        }
        T += st.steps; I += st.isect; R += st.rays;
    }
    if (stats) { stats[0] = T; stats[1] = I; stats[2] = R; }
}

// S5: screen-space RTAO against the analytic capsules (VulkanRayTracedAmbientOcclusion.glsl:178-319).
// ao_inout: W*H floats (x channel of the rgba32f accumulation image).  stats = {T, I, rays_primary, rays_ao, pixels_hit}
void lvo_render_rtao(void* h, const lv_camera* cam, const lvo_options* o, uint32_t frame_number, float* ao_inout, uint64_t* stats) {
    Scene& sc = *static_cast<Scene*>(h);
    Uniforms u = makeUniforms(sc, *cam, *o, nullptr, 0, 0, 1, nullptr);
    const uint32_t W = cam->width, H = cam->height;
    const uint32_t globalFrameNumber = frame_number;  // useGlobalFrameNumber == false (VulkanRayTracedAmbientOcclusion.cpp:580-584)
    // subdivisionCorrectionFactor = cos(pi / tubeNumSubdivisions) -- VulkanRayTracedAmbientOcclusion.cpp:591
    const float subdivisionCorrectionFactor = float(std::cos(3.14159265358979323846 / double(o->tube_num_subdivisions)));
    uint64_t T = 0, I = 0, RP = 0, RA = 0, PH = 0, TA = 0, IA = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : T, I, RP, RA, PH, TA, IA)
    for (int64_t yy = 0; yy < int64_t(H); yy++) {
        RayStats sp, sa;
        uint32_t y = uint32_t(yy);
        for (uint32_t x = 0; x < W; x++) {
            uint32_t seed = tea(x + y * W, globalFrameNumber);                          // :187
            float xix = 0.5f, xiy = 0.5f;
            if (o->ao_jitter_primary) { xix = rnd(seed); xiy = rnd(seed); }            // :189-194
            Ray r; cameraRay(u, x, y, xix, xiy, r.o, r.d); r.tmin = 0.0001f; r.tmax = 1000.0f;  // :196-203
            float aoFactor = 1.0f;
            Hit hit;
            if (traceClosest(sc, r, u.useCappedTubes, hit, sp)) {
                PH++;
                const Segment& s = sc.segs[hit.prim];
                // analytic replacement of the barycentric vertex fetch (:222-267): position, normal, line centre, tangent
                vec3 vertexPositionWorld = r.o + r.d * hit.t;
                vec3 v = s.p1 - s.p0;
                vec3 linePosition;
                if (hit.kind == 0) { vec3 uu = vertexPositionWorld - s.p0; float t = dot(v, uu) / dot(v, v); linePosition = s.p0 + t * v; }
                else if (hit.kind == 1) linePosition = s.p0; else linePosition = s.p1;
                vec3 surfaceNormal = normalize(vertexPositionWorld - linePosition);
                vec3 surfaceTangent = normalize(v);
                vec3 surfaceBitangent = cross(surfaceNormal, surfaceTangent);          // :257
                const float offsetFactor = length(linePosition - vertexPositionWorld) / subdivisionCorrectionFactor;  // :276
                aoFactor = 0.0f;
                for (uint32_t sampleIdx = 0; sampleIdx < o->ao_spp; sampleIdx++) {     // :284-303
                    uint32_t seed2 = tea(x + y * W, globalFrameNumber * o->ao_spp + sampleIdx);
                    float a = rnd(seed2), b = rnd(seed2);
                    vec3 hs = sampleHemisphere(a, b);
                    vec3 dir = normalize((surfaceTangent * hs.x + surfaceBitangent * hs.y) + surfaceNormal * hs.z);
                    Ray ar; ar.o = vertexPositionWorld + dir * offsetFactor; ar.d = dir; ar.tmin = 0.0f; ar.tmax = o->ao_radius;
                    float occ = 1.0f;                                                   // traceAoRay :158-175
                    if (o->ao_use_distance) { Hit ah; if (traceClosest(sc, ar, u.useCappedTubes, ah, sa, false)) occ = ah.t / o->ao_radius; }
                    else { if (traceAny(sc, ar, u.useCappedTubes, sa)) occ = 0.0f; }
                    aoFactor += occ;
                }
                aoFactor /= float(o->ao_spp);
            }
            size_t idx = size_t(y) * W + x;
            if (frame_number != 0) { float prev = ao_inout[idx]; aoFactor = mix(prev, aoFactor, 1.0f / float(frame_number + 1)); }  // :313-317
            ao_inout[idx] = aoFactor;
        }
        T += sp.steps + sa.steps; I += sp.isect + sa.isect; RP += sp.rays; RA += sa.rays; TA += sa.steps; IA += sa.isect;
    }
    // stats[5], [6]: traversal steps / primitive tests of the AO rays alone (T and I per AO ray on this backend's tree)
    if (stats) { stats[0] = T; stats[1] = I; stats[2] = RP; stats[3] = RA; stats[4] = PH; stats[5] = TA; stats[6] = IA; }
}

// S1 ray-gen with S2/S3/S4: traceRayTransparent loop + running mean (TubeRayTracing.glsl:61-82,198-274).
// rgba_inout: W*H*4 floats (float stand-in for the rgba8 storage image; quantisation is the caller's).
// ao_tex: result of lvo_render_rtao or NULL.  stats = {T, I, rays}
void lvo_render_tubes(void* h, const lv_camera* cam, const lvo_options* o, const float* tf, uint32_t K, float amin, float amax,
                      const float* ao_tex, uint32_t frame_number, float* rgba_inout, uint64_t* stats) {
    Scene& sc = *static_cast<Scene*>(h);
    Uniforms u = makeUniforms(sc, *cam, *o, tf, K, amin, amax, ao_tex);
    const uint32_t W = cam->width, H = cam->height;
    uint64_t T = 0, I = 0, R = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : T, I, R)
    for (int64_t yy = 0; yy < int64_t(H); yy++) {
        RayStats st;
        uint32_t y = uint32_t(yy);
        for (uint32_t x = 0; x < W; x++) {
            vec4 fragmentColor{0, 0, 0, 0};
            uint32_t nspp = o->use_jittered_rays ? o->num_samples_per_frame : 1u;
            for (uint32_t sampleIdx = 0; sampleIdx < nspp; sampleIdx++) {
                float xix = 0.5f, xiy = 0.5f;
                if (o->use_jittered_rays) {
                    uint32_t seed = o->use_deterministic_sampling
                        ? tea(19u, frame_number * o->num_samples_per_frame + sampleIdx)
                        : tea(x + y * W, frame_number * o->num_samples_per_frame + sampleIdx);
                    xix = rnd(seed); xiy = rnd(seed);
                }
                vec3 ro, rd; cameraRay(u, x, y, xix, xiy, ro, rd);
                // traceRayTransparent :61-82
                vec4 fc{0, 0, 0, 0};
                float tMin = 0.0001f, tMax = 1000.0f;
                for (uint32_t hitIdx = 0; hitIdx < o->max_depth_complexity; hitIdx++) {
                    Ray r; r.o = ro; r.d = rd; r.tmin = tMin; r.tmax = tMax;
                    Hit hit; HitColor pl;
                    if (traceClosest(sc, r, u.useCappedTubes, hit, st)) {
                        const Segment& s = sc.segs[hit.prim];
                        pl = closestHitTubeAnalytic(u, ro, rd, hit.t, hit.kind, s.p0, s.a0, s.p1, s.a1, &static_cast<const SceneX&>(sc).segLine[hit.prim]);
                    } else pl = missShader(u);
                    tMin = pl.hitT + fmax_(pl.hitT * 1e-5f, 1e-7f);
                    fc.x = fc.x + (1.0f - fc.w) * pl.hitColor.w * pl.hitColor.x;
                    fc.y = fc.y + (1.0f - fc.w) * pl.hitColor.w * pl.hitColor.y;
                    fc.z = fc.z + (1.0f - fc.w) * pl.hitColor.w * pl.hitColor.z;
                    fc.w = fc.w + (1.0f - fc.w) * pl.hitColor.w;
                    if (!pl.hasHit || fc.w > 0.99f) break;
                }
                fragmentColor.x += fc.x; fragmentColor.y += fc.y; fragmentColor.z += fc.z; fragmentColor.w += fc.w;
            }
            if (o->use_jittered_rays) {
                float d = float(o->num_samples_per_frame);
                fragmentColor.x /= d; fragmentColor.y /= d; fragmentColor.z /= d; fragmentColor.w /= d;
            }
            float* px = rgba_inout + 4 * (size_t(y) * W + x);
            if (frame_number != 0) {                                                    // :269-272
                float a = 1.0f / float(frame_number + 1);
                fragmentColor = vec4{mix(px[0], fragmentColor.x, a), mix(px[1], fragmentColor.y, a), mix(px[2], fragmentColor.z, a), mix(px[3], fragmentColor.w, a)};
            }
            px[0] = fragmentColor.x; px[1] = fragmentColor.y; px[2] = fragmentColor.z; px[3] = fragmentColor.w;
        }
        T += st.steps; I += st.isect; R += st.rays;
    }
    if (stats) { stats[0] = T; stats[1] = I; stats[2] = R; }
}

// ------------------------------------------------------------------------------------------ PPLL
void lvo_ppll_padded_size(uint32_t W, uint32_t H, uint32_t tw, uint32_t th, uint32_t* pw, uint32_t* ph) { paddedSize(W, H, tw, th, *pw, *ph); }

// S8 + S9: clear + gather.  The fragment source (S11 is a hardware rasteriser) is redefined as: every capsule whose
// reported hit (S2) along the pixel-centre ray lies in [1e-4, 1000] yields one fragment shaded by S3.
// heads: paddedW*paddedH u32; nodes: capacity linked_list_size.  Returns fragCounter (may exceed linked_list_size).
// stats = {T, I, rays, frags_generated}
uint64_t lvo_ppll_gather(void* h, const lv_camera* cam, const lvo_options* o, const float* tf, uint32_t K, float amin, float amax,
                         const float* ao_tex, uint64_t linked_list_size, uint32_t* heads, lv_ppll_node* nodes, uint64_t* stats) {
    Scene& sc = *static_cast<Scene*>(h);
    Uniforms u = makeUniforms(sc, *cam, *o, tf, K, amin, amax, ao_tex);
    const uint32_t W = cam->width, H = cam->height;
    uint32_t pw, ph; paddedSize(W, H, o->tile_w, o->tile_h, pw, ph);
    for (size_t i = 0; i < size_t(pw) * ph; i++) heads[i] = 0xFFFFFFFFu;               // LinkedListClear.glsl:50
    uint64_t fragCounter = 0;
    uint64_t T = 0, I = 0, R = 0;
    struct Frag { uint32_t color; float depth; };
    std::vector<std::vector<Frag>> rowFrags(W);
    for (uint32_t y = 0; y < H; y++) {
        RayStats stRow;
#pragma omp parallel
        {
            RayStats st;
#pragma omp for schedule(dynamic, 16) nowait
            for (int64_t xx = 0; xx < int64_t(W); xx++) {
                uint32_t x = uint32_t(xx);
                rowFrags[x].clear();
                vec3 ro, rd; cameraRay(u, x, y, 0.5f, 0.5f, ro, rd);
                Ray r; r.o = ro; r.d = rd; r.tmin = 0.0001f; r.tmax = 1000.0f;
                traceAll(sc, r, u.useCappedTubes, st, [&](uint32_t p, float t, int kind) {
                    const Segment& s = sc.segs[p];
                    HitColor pl = closestHitTubeAnalytic(u, ro, rd, t, kind, s.p0, s.a0, s.p1, s.a1, &static_cast<const SceneX&>(sc).segLine[p]);
                    if (pl.hitColor.w < 0.001f) return;                                 // LinkedListGather.glsl:38
                    // frag.depth = length(fragmentPositionWorld - cameraPosition) (:52) == payload.hitT
                    rowFrags[x].push_back(Frag{packUnorm4x8(pl.hitColor), pl.hitT});
                });
            }
#pragma omp critical
            { stRow.steps += st.steps; stRow.isect += st.isect; stRow.rays += st.rays; }
        }
        T += stRow.steps; I += stRow.isect; R += stRow.rays;
        for (uint32_t x = 0; x < W; x++) {
            uint32_t pixelIndex = addrGen(x, y, pw, o->tile_w, o->tile_h);
            for (const Frag& f : rowFrags[x]) {
                uint64_t insertIndex = fragCounter++;                                   // atomicAdd(fragCounter, 1u) :55
                if (insertIndex < linked_list_size) {
                    uint32_t next = heads[pixelIndex]; heads[pixelIndex] = uint32_t(insertIndex);  // atomicExchange :59
                    nodes[insertIndex] = lv_ppll_node{f.color, f.depth, next};
                }
            }
        }
    }
    if (stats) { stats[0] = T; stats[1] = I; stats[2] = R; stats[3] = fragCounter; }
    return fragCounter;
}

// S10: resolve (LinkedListResolve.glsl:57-105) + final BACK_TO_FRONT_STRAIGHT_ALPHA blend over the clear colour
// (PerPixelLinkedListLineRenderer.cpp:70; dst = src.rgb*src.a + dst*(1-src.a), alpha = src.a + dst.a*(1-src.a)).
// stats = {frags_sorted, frags_truncated, max_depth_complexity}
void lvo_ppll_resolve(const lv_camera* cam, const lvo_options* o, const uint32_t* heads, const lv_ppll_node* nodes,
                      uint32_t max_frags, int sort_mode, int canonical, float* rgba_out, uint64_t* stats) {
    const uint32_t W = cam->width, H = cam->height;
    uint32_t pw, ph; paddedSize(W, H, o->tile_w, o->tile_h, pw, ph);
    uint64_t sorted = 0, trunc = 0, maxdc = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : sorted, trunc) reduction(max : maxdc)
    for (int64_t yy = 0; yy < int64_t(H); yy++) {
        ResolveLists L;
        L.colorList.resize(max_frags); L.depthList.resize(max_frags);
        uint32_t y = uint32_t(yy);
        for (uint32_t x = 0; x < W; x++) {
            uint32_t fragOffset = heads[addrGen(x, y, pw, o->tile_w, o->tile_h)];
            uint32_t numFrags = 0;
            for (uint32_t i = 0; i < max_frags; i++) {
                if (fragOffset == 0xFFFFFFFFu) break;
                const lv_ppll_node& f = nodes[fragOffset];
                fragOffset = f.next;
                L.colorList[i] = f.color; L.depthList[i] = f.depth;
                numFrags++;
            }
            uint64_t rest = 0;
            while (fragOffset != 0xFFFFFFFFu) { rest++; fragOffset = nodes[fragOffset].next; }
            sorted += numFrags; trunc += rest; maxdc = std::max<uint64_t>(maxdc, numFrags + rest);
            float* px = rgba_out + 4 * (size_t(y) * W + x);
            if (numFrags == 0) { for (int k = 0; k < 4; k++) px[k] = cam->background[k]; continue; }  // discard
            vec4 c = canonical ? L.canonical(sort_mode, numFrags) : L.sortingAlgorithm(sort_mode, numFrags, max_frags);
            px[0] = c.x * c.w + cam->background[0] * (1.0f - c.w);
            px[1] = c.y * c.w + cam->background[1] * (1.0f - c.w);
            px[2] = c.z * c.w + cam->background[2] * (1.0f - c.w);
            px[3] = c.w + cam->background[3] * (1.0f - c.w);
        }
    }
    if (stats) { stats[0] = sorted; stats[1] = trunc; stats[2] = maxdc; }
}

void lvo_depth_range(void* h, const lv_camera* cam, float* out2) {
    Scene& sc = *static_cast<Scene*>(h);
    lvo_options o{}; o.depth_cue_strength = 1.0f;
    Uniforms u = makeUniforms(sc, *cam, o, nullptr, 0, 0, 1, nullptr);
    out2[0] = u.minDepth; out2[1] = u.maxDepth;
}

// ------------------------------------------------------------------------------------------ object-space AO prebaker (S6)
float lvo_det_acos(float x) { return det_acos(x); }

// Line-point frames (tangent / normal per point) for the prebaker and its lookup.
void lvo_scene_set_lines(void* h, const float* tangent, const float* normal, uint64_t n_pt) {
    SceneX& sc = sx(h);
    sc.ptTan.resize(n_pt); sc.ptNrm.resize(n_pt);
    for (uint64_t i = 0; i < n_pt; i++) {
        sc.ptTan[i] = V3(tangent[3 * i], tangent[3 * i + 1], tangent[3 * i + 2]);
        sc.ptNrm[i] = V3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
    }
    for (auto& l : sc.segLine) { l.n0 = sc.ptNrm[l.idx0]; l.n1 = sc.ptNrm[l.idx1]; }
}

// AmbientOcclusionComputeRenderPass::generateBlendingWeightParametrization + recomputeStaticParametrization
// (src/Renderers/AmbientOcclusion/VulkanAmbientOcclusionBaker.cpp:513-655).  Polylines = (pos, line_offsets[n_lines + 1]).
// blending_weights: one float per line point; sampling_locations: capacity `cap`.  Returns numParametrizationVertices
// (which may exceed cap; then only the first cap entries were written).
uint64_t lvo_ao_parametrize(const float* pos, const uint64_t* line_offsets, uint64_t n_lines, float expectedParamSegmentLength,
                            float* blendingWeightParametrizationData, float* samplingLocations, uint64_t cap) {
    const float EPSILON = 1e-5f;
    uint64_t numSamplingLocations = 0;
    auto push = [&](float v) { if (numSamplingLocations < cap) samplingLocations[numSamplingLocations] = v; numSamplingLocations++; };
    size_t segmentVertexIdOffset = 0;
    size_t vertexIdx = 0;
    for (uint64_t lineIdx = 0; lineIdx < n_lines; lineIdx++) {
        const uint64_t b = line_offsets[lineIdx];
        const size_t n = size_t(line_offsets[lineIdx + 1] - b);
        auto line = [&](size_t i) { return V3(pos[3 * (b + i)], pos[3 * (b + i) + 1], pos[3 * (b + i) + 2]); };
        float polylineLength = 0.0f;                                                    // :540-544
        for (size_t i = 1; i < n; i++) polylineLength += length(line(i) - line(i - 1));

        uint32_t numLineSubdivs = std::max(1u, uint32_t(std::ceil(polylineLength / expectedParamSegmentLength)));  // :574
        float lineSubdivLength = polylineLength / float(numLineSubdivs);
        uint32_t numSubdivVertices = numLineSubdivs + 1;

        uint32_t startVertexIdx = uint32_t(vertexIdx);                                  // :580-582
        blendingWeightParametrizationData[vertexIdx] = float(segmentVertexIdOffset);
        vertexIdx++;

        float currentLength = 0.0f;                                                     // :585-592
        for (size_t i = 1; i < n; i++) {
            currentLength += length(line(i) - line(i - 1));
            float w = currentLength / lineSubdivLength;
            blendingWeightParametrizationData[vertexIdx] = float(segmentVertexIdOffset) + clamp(w, 0.0f, float(numLineSubdivs) - EPSILON);
            vertexIdx++;
        }

        float lastLength = 0.0f;                                                        // :594-612
        currentLength = length(line(1) - line(0));
        size_t currVertexIdx = 1;
        push(float(startVertexIdx));
        for (uint32_t i = 1; i < numSubdivVertices; i++) {
            uint32_t parametrizationIdx = uint32_t(currentLength / lineSubdivLength);
            while (i > parametrizationIdx && currVertexIdx < n - 1) {
                float segLength = length(line(currVertexIdx + 1) - line(currVertexIdx));
                lastLength = currentLength;
                currentLength += segLength;
                parametrizationIdx = uint32_t(currentLength / lineSubdivLength);
                currVertexIdx++;
            }
            float samplingLocation = float(currVertexIdx - 1) + (float(i) * lineSubdivLength - lastLength) / (currentLength - lastLength);
            samplingLocation = float(startVertexIdx) + std::min(samplingLocation, float(uint32_t(n) - 1u) - EPSILON);
            push(samplingLocation);
        }
        segmentVertexIdOffset += numSubdivVertices;
    }
    return numSamplingLocations;
}

// One dispatch of the baker compute shader (Data/Shaders/AO/RTAO/VulkanAmbientOcclusionBaker.glsl:190-282): for every
// parametrization vertex and tube subdivision, numAmbientOcclusionSamples hemisphere rays; running mean over frame_number.
// The reference traces against the triangulated tubes; like the screen-space RTAO pass this oracle traces the analytic
// capsules (DESIGN.md rule 5).  factors_inout: n_param * numTubeSubdivisions floats.  stats = {T, I, rays}
// rays_out (optional, tests): 6 floats (origin, direction) per ray, index (vertex * N + subdivision) * spp + ray.
void lvo_ao_bake_iteration(void* h, const lvo_bake_options* bo, int capped, const float* samplingLocations, uint64_t n_param,
                           uint32_t frameNumber, float* ambientOcclusionFactors, uint64_t* stats, float* rays_out, void* tube_mesh) {
    // tube_mesh != NULL: trace the AO rays against the reference's triangulated tubes (lvo_tubemesh_create) instead of the capsules
    SceneX& sc = sx(h);
    const TubeMeshScene* tms = static_cast<const TubeMeshScene*>(tube_mesh);
    const uint32_t numLinePoints = uint32_t(sc.ptPos.size());
    const uint32_t numTubeSubdivisions = bo->num_tube_subdivisions;
    const uint32_t numAmbientOcclusionSamples = bo->samples_per_frame;
    const float lineRadius = sc.lineWidth * 0.5f;                                       // VulkanAmbientOcclusionBaker.cpp:697
    const float ambientOcclusionRadius = bo->ao_radius;
    uint64_t T = 0, I = 0, R = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : T, I, R)
    for (int64_t gid = 0; gid < int64_t(n_param); gid++) {
        RayStats st;
        const uint32_t lineSamplingIdx = uint32_t(gid);
        uint32_t seed = tea(lineSamplingIdx, frameNumber);                              // :196
        // getInterpolatedLinePoint :104-124
        const float samplingLocation = samplingLocations[lineSamplingIdx];
        const uint32_t lowerIdx = uint32_t(samplingLocation);
        const uint32_t upperIdx = std::min(lowerIdx + 1u, numLinePoints - 1u);
        const float interpolationFactor = samplingLocation - floorf(samplingLocation);
        const vec3 tl = sc.ptTan[lowerIdx], tu = sc.ptTan[upperIdx], nl = sc.ptNrm[lowerIdx], nu = sc.ptNrm[upperIdx];
        const vec3 binormalLower = cross(tl, nl), binormalUpper = cross(tu, nu);
        auto mix3 = [&](vec3 a, vec3 b) { return V3(mix(a.x, b.x, interpolationFactor), mix(a.y, b.y, interpolationFactor), mix(a.z, b.z, interpolationFactor)); };
        const vec3 position = mix3(sc.ptPos[lowerIdx], sc.ptPos[upperIdx]);
        const vec3 tangent = normalize(mix3(tl, tu));
        const vec3 normal = normalize(mix3(nl, nu));
        const vec3 binormal = normalize(mix3(binormalLower, binormalUpper));
        for (uint32_t tubeSudivIdx = 0; tubeSudivIdx < numTubeSubdivisions; tubeSudivIdx++) {   // :238
            float cosAngle, sinAngle;
            det_sincos2pi(float(tubeSudivIdx) / float(numTubeSubdivisions), cosAngle, sinAngle);   // angle = idx / N * 2 pi
            const vec3 surfaceNormal = cosAngle * normal + sinAngle * binormal;          // :258
            const vec3 rayOrigin = position + (lineRadius + 1e-6f) * surfaceNormal;      // :259
            const vec3 surfaceBitangent = cross(surfaceNormal, tangent);                 // :262
            float occlusionFactorAccumulated = 0.0f;
            for (uint32_t rayIdx = 0; rayIdx < numAmbientOcclusionSamples; rayIdx++) {   // :266-273
                const float xix = rnd(seed), xiy = rnd(seed);
                const vec3 hs = sampleHemisphere(xix, xiy);
                const vec3 rayDirection = normalize((tangent * hs.x + surfaceBitangent * hs.y) + surfaceNormal * hs.z);   // frame * hs
                Ray ar; ar.o = rayOrigin; ar.d = rayDirection; ar.tmin = 0.0f; ar.tmax = ambientOcclusionRadius;
                if (rays_out) {
                    float* ro = rays_out + 6 * ((size_t(lineSamplingIdx) * numTubeSubdivisions + tubeSudivIdx) * numAmbientOcclusionSamples + rayIdx);
                    ro[0] = rayOrigin.x; ro[1] = rayOrigin.y; ro[2] = rayOrigin.z; ro[3] = rayDirection.x; ro[4] = rayDirection.y; ro[5] = rayDirection.z;
                }
                float occlusionFactor = 1.0f;                                            // traceAoRay :167-186
                if (tms) {
                    TriHit th; uint64_t s2 = 0, i2 = 0;
                    st.rays++;
                    if (traceTriangles(tms->bvh, ar.o, ar.d, 0.0f, ambientOcclusionRadius, bo->use_distance == 0, th, s2, i2))
                        occlusionFactor = bo->use_distance ? th.t / ambientOcclusionRadius : 0.0f;
                    st.steps += s2; st.isect += i2;
                } else if (bo->use_distance) { Hit ah; if (traceClosest(sc, ar, capped != 0, ah, st, false)) occlusionFactor = ah.t / ambientOcclusionRadius; }
                else { if (traceAny(sc, ar, capped != 0, st)) occlusionFactor = 0.0f; }
                occlusionFactorAccumulated += occlusionFactor;
            }
            occlusionFactorAccumulated /= float(numAmbientOcclusionSamples);
            float* dst = ambientOcclusionFactors + (tubeSudivIdx + size_t(numTubeSubdivisions) * lineSamplingIdx);
            if (frameNumber != 0) occlusionFactorAccumulated = mix(*dst, occlusionFactorAccumulated, 1.0f / float(frameNumber + 1));  // :276-279
            *dst = occlusionFactorAccumulated;
        }
        T += st.steps; I += st.isect; R += st.rays;
    }
    if (stats) { stats[0] = T; stats[1] = I; stats[2] = R; }
}

// Hands the baked factors + blending weights to the shading path (STATIC_AMBIENT_OCCLUSION_PREBAKING bindings,
// LineRenderer::setRenderDataBindings / Utils/AmbientOcclusion.glsl:30-37).
void lvo_scene_set_static_ao(void* h, const float* factors, uint64_t n_param, uint32_t n_subdiv, const float* blending_weights, uint64_t n_pt) {
    SceneX& sc = sx(h);
    sc.aoFactors.assign(factors, factors + n_param * n_subdiv);
    sc.aoBlendingWeights.assign(blending_weights, blending_weights + n_pt);
    sc.numParametrizationVertices = uint32_t(n_param); sc.numAoTubeSubdivisions = n_subdiv;
}

// ClosestHitTubeAnalytic for a list of given hits (tests: function-level parity of the device shading code).
// ro / rd: n*3; t: n; kind: n; prim: n (segment index).  out: n*5 floats (hitColor rgba, hitT)
void lvo_shade_hits(void* h, const lv_camera* cam, const lvo_options* o, const float* tf, uint32_t K, float amin, float amax,
                    const float* ao_tex, uint64_t n, const float* ro, const float* rd, const float* t, const uint32_t* kind,
                    const uint32_t* prim, float* out) {
    SceneX& sc = sx(h);
    Uniforms u = makeUniforms(sc, *cam, *o, tf, K, amin, amax, ao_tex);
    for (uint64_t i = 0; i < n; i++) {
        const Segment& s = sc.segs[prim[i]];
        HitColor pl = closestHitTubeAnalytic(u, V3(ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]), V3(rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]), t[i],
                                             int(kind[i]), s.p0, s.a0, s.p1, s.a1, &sc.segLine[prim[i]]);
        out[5 * i] = pl.hitColor.x; out[5 * i + 1] = pl.hitColor.y; out[5 * i + 2] = pl.hitColor.z; out[5 * i + 3] = pl.hitColor.w; out[5 * i + 4] = pl.hitT;
    }
}

// getAoFactor(fragmentVertexId, phi) alone, for unit tests
float lvo_static_ao_factor(void* h, float strength, float gamma, float fragmentVertexId, float phi) {
    SceneX& sc = sx(h);
    Uniforms u{};
    u.ambientOcclusionStrength = strength; u.ambientOcclusionGamma = gamma;
    u.ambientOcclusionFactors = sc.aoFactors.data(); u.ambientOcclusionBlendingWeights = sc.aoBlendingWeights.data();
    u.numAoTubeSubdivisions = sc.numAoTubeSubdivisions; u.numLineVertices = uint32_t(sc.aoBlendingWeights.size());
    u.numParametrizationVertices = sc.numParametrizationVertices;
    return getAoFactorStatic(u, fragmentVertexId, phi);
}

// ------------------------------------------------------------------------------------------ triangle-tube RTAO (reference geometry)
// createCappedTriangleTubesRenderDataCPU for polylines (pos, line_offsets[n_lines + 1]); tubeRadius = lineWidth / 2.
void* lvo_tubemesh_create(const float* pos, const uint64_t* line_offsets, uint64_t n_lines, float tube_radius, int num_circle_subdivisions) {
    TubeMeshScene* t = new TubeMeshScene();
    t->lineWidth = 2.0f * tube_radius;
    createCappedTriangleTubes(pos, line_offsets, n_lines, tube_radius, num_circle_subdivisions, t->mesh);
    t->bvh.build(t->mesh);
    return t;
}
void lvo_tubemesh_destroy(void* h) { delete static_cast<TubeMeshScene*>(h); }
void lvo_tubemesh_info(void* h, uint64_t* n_vertices, uint64_t* n_triangles, uint64_t* n_line_points) {
    TubeMeshScene& t = *static_cast<TubeMeshScene*>(h);
    *n_vertices = t.mesh.vertexDataList.size(); *n_triangles = t.mesh.triangleIndices.size() / 3; *n_line_points = t.mesh.linePositions.size();
}
// vertices: 8 floats each (position, as_float(vertexLinePointIndex), normal, phi); indices: 3 per triangle
void lvo_tubemesh_copy(void* h, float* vertices, uint32_t* indices) {
    TubeMeshScene& t = *static_cast<TubeMeshScene*>(h);
    static_assert(sizeof(TubeTriangleVertexData) == 32, "TubeTriangleVertexData is 32 bytes");
    if (vertices) std::memcpy(vertices, t.mesh.vertexDataList.data(), t.mesh.vertexDataList.size() * 32);
    if (indices) std::memcpy(indices, t.mesh.triangleIndices.data(), t.mesh.triangleIndices.size() * 4);
}

// The tube pass in the triangle-mesh geometry mode (RayTracingGeometryMode::TRIANGLE_MESH, VulkanRayTracer.hpp:54-63): ray-gen +
// traceRayTransparent as in lvo_render_tubes, closest hit against the tube mesh, ClosestHitTubeTriangles.  attr: per input point.
// stats = {T, I, rays}
void lvo_tubemesh_render_tubes(void* h, const float* attr, const lv_camera* cam, const lvo_options* o, const float* tf, uint32_t K, float amin, float amax,
                               const float* ao_tex, uint32_t frame_number, float* rgba_inout, uint64_t* stats) {
    TubeMeshScene& ts = *static_cast<TubeMeshScene*>(h);
    Scene dummy; dummy.lineWidth = ts.lineWidth;
    lvo_options o2 = *o; o2.use_static_ao = 0; o2.depth_cue_strength = 0.0f;
    Uniforms u = makeUniforms(dummy, *cam, o2, tf, K, amin, amax, ao_tex);
    const uint32_t W = cam->width, H = cam->height;
    uint64_t T = 0, I = 0, R = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : T, I, R)
    for (int64_t yy = 0; yy < int64_t(H); yy++) {
        uint64_t steps = 0, isect = 0, rays = 0;
        const uint32_t y = uint32_t(yy);
        for (uint32_t x = 0; x < W; x++) {
            vec4 fragmentColor{0, 0, 0, 0};
            const uint32_t nspp = o->use_jittered_rays ? o->num_samples_per_frame : 1u;
            for (uint32_t sampleIdx = 0; sampleIdx < nspp; sampleIdx++) {
                float xix = 0.5f, xiy = 0.5f;
                if (o->use_jittered_rays) {
                    uint32_t seed = o->use_deterministic_sampling ? tea(19u, frame_number * o->num_samples_per_frame + sampleIdx)
                                                                  : tea(x + y * W, frame_number * o->num_samples_per_frame + sampleIdx);
                    xix = rnd(seed); xiy = rnd(seed);
                }
                vec3 ro, rd; cameraRay(u, x, y, xix, xiy, ro, rd);
                vec4 fc{0, 0, 0, 0};
                float tMin = 0.0001f;
                for (uint32_t hitIdx = 0; hitIdx < o->max_depth_complexity; hitIdx++) {   // traceRayTransparent :61-82
                    TriHit hit; HitColor pl;
                    rays++;
                    if (traceTriangles(ts.bvh, ro, rd, tMin, 1000.0f, false, hit, steps, isect, ts.lineWidth)) pl = closestHitTubeTriangles(u, ts.mesh, attr, hit);
                    else pl = missShader(u);
                    tMin = pl.hitT + fmax_(pl.hitT * 1e-5f, 1e-7f);
                    fc.x = fc.x + (1.0f - fc.w) * pl.hitColor.w * pl.hitColor.x;
                    fc.y = fc.y + (1.0f - fc.w) * pl.hitColor.w * pl.hitColor.y;
                    fc.z = fc.z + (1.0f - fc.w) * pl.hitColor.w * pl.hitColor.z;
                    fc.w = fc.w + (1.0f - fc.w) * pl.hitColor.w;
                    if (!pl.hasHit || fc.w > 0.99f) break;
                }
                fragmentColor.x += fc.x; fragmentColor.y += fc.y; fragmentColor.z += fc.z; fragmentColor.w += fc.w;
            }
            if (o->use_jittered_rays) {
                const float d = float(o->num_samples_per_frame);
                fragmentColor.x /= d; fragmentColor.y /= d; fragmentColor.z /= d; fragmentColor.w /= d;
            }
            float* px = rgba_inout + 4 * (size_t(y) * W + x);
            if (frame_number != 0) {
                const float a = 1.0f / float(frame_number + 1);
                fragmentColor = vec4{mix(px[0], fragmentColor.x, a), mix(px[1], fragmentColor.y, a), mix(px[2], fragmentColor.z, a), mix(px[3], fragmentColor.w, a)};
            }
            px[0] = fragmentColor.x; px[1] = fragmentColor.y; px[2] = fragmentColor.z; px[3] = fragmentColor.w;
        }
        T += steps; I += isect; R += rays;
    }
    if (stats) { stats[0] = T; stats[1] = I; stats[2] = R; }
}

// The screen-space RTAO pass exactly as the reference runs it: against the TRIANGULATED tubes, with the barycentric vertex
// fetch (Data/Shaders/AO/RTAO/VulkanRayTracedAmbientOcclusion.glsl:178-319).  stats = {T, I, rays_primary, rays_ao, pixels_hit}
void lvo_render_rtao_triangles(void* h, const lv_camera* cam, const lvo_options* o, uint32_t frame_number, float* ao_inout, uint64_t* stats) {
    TubeMeshScene& ts = *static_cast<TubeMeshScene*>(h);
    const TubeMesh& m = ts.mesh;
    Uniforms u{};
    std::memcpy(u.inverseViewMatrix, cam->inv_view, 64); std::memcpy(u.inverseProjectionMatrix, cam->inv_proj, 64);
    u.viewportW = cam->width; u.viewportH = cam->height;
    const uint32_t W = cam->width, H = cam->height;
    const uint32_t globalFrameNumber = frame_number;
    const float subdivisionCorrectionFactor = float(std::cos(3.14159265358979323846 / double(o->tube_num_subdivisions)));   // .cpp:588
    uint64_t T = 0, I = 0, RP = 0, RA = 0, PH = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : T, I, RP, RA, PH)
    for (int64_t yy = 0; yy < int64_t(H); yy++) {
        uint64_t steps = 0, isect = 0;
        const uint32_t y = uint32_t(yy);
        for (uint32_t x = 0; x < W; x++) {
            uint32_t seed = tea(x + y * W, globalFrameNumber);                          // :187
            float xix = 0.5f, xiy = 0.5f;
            if (o->ao_jitter_primary) { xix = rnd(seed); xiy = rnd(seed); }            // :189-194
            vec3 ro, rd; cameraRay(u, x, y, xix, xiy, ro, rd);                          // :196-203
            float aoFactor = 1.0f;
            TriHit hit;
            RP++;
            if (traceTriangles(ts.bvh, ro, rd, 0.0001f, 1000.0f, false, hit, steps, isect, ts.lineWidth)) {
                PH++;
                const uint32_t* tri = &m.triangleIndices[3 * size_t(hit.tri)];          // :213-221
                const vec3 bary = V3(1.0f - hit.u - hit.v, hit.u, hit.v);
                const TubeTriangleVertexData& v0 = m.vertexDataList[tri[0]];
                const TubeTriangleVertexData& v1 = m.vertexDataList[tri[1]];
                const TubeTriangleVertexData& v2 = m.vertexDataList[tri[2]];
                const uint32_t l0 = v0.vertexLinePointIndex & 0x7FFFFFFFu, l1 = v1.vertexLinePointIndex & 0x7FFFFFFFu, l2 = v2.vertexLinePointIndex & 0x7FFFFFFFu;
                const vec3 vertexPositionWorld = interpolateVec3(v0.vertexPosition, v1.vertexPosition, v2.vertexPosition, bary);   // :251-255
                const vec3 surfaceNormal = normalize(interpolateVec3(v0.vertexNormal, v1.vertexNormal, v2.vertexNormal, bary));
                const vec3 linePosition = interpolateVec3(m.linePositions[l0], m.linePositions[l1], m.linePositions[l2], bary);    // :258-262
                const vec3 surfaceTangent = normalize(interpolateVec3(m.lineTangents[l0], m.lineTangents[l1], m.lineTangents[l2], bary));
                const vec3 surfaceBitangent = cross(surfaceNormal, surfaceTangent);
                const float offsetFactor = length(linePosition - vertexPositionWorld) / subdivisionCorrectionFactor;               // :276
                aoFactor = 0.0f;
                for (uint32_t sampleIdx = 0; sampleIdx < o->ao_spp; sampleIdx++) {     // :284-303
                    uint32_t seed2 = tea(x + y * W, globalFrameNumber * o->ao_spp + sampleIdx);
                    const float a = rnd(seed2), b = rnd(seed2);
                    const vec3 hs = sampleHemisphere(a, b);
                    const vec3 dir = normalize((surfaceTangent * hs.x + surfaceBitangent * hs.y) + surfaceNormal * hs.z);
                    const vec3 org = vertexPositionWorld + dir * offsetFactor;
                    TriHit ah;
                    RA++;
                    float occ = 1.0f;                                                   // traceAoRay :158-175
                    if (traceTriangles(ts.bvh, org, dir, 0.0f, o->ao_radius, o->ao_use_distance == 0, ah, steps, isect))
                        occ = o->ao_use_distance ? ah.t / o->ao_radius : 0.0f;
                    aoFactor += occ;
                }
                aoFactor /= float(o->ao_spp);
            }
            const size_t idx = size_t(y) * W + x;
            if (frame_number != 0) aoFactor = mix(ao_inout[idx], aoFactor, 1.0f / float(frame_number + 1));   // :313-317
            ao_inout[idx] = aoFactor;
        }
        T += steps; I += isect;
    }
    if (stats) { stats[0] = T; stats[1] = I; stats[2] = RP; stats[3] = RA; stats[4] = PH; }
}

int lvo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Size of the OpenMP pool of every driver above.  Launchers such as torchrun export OMP_NUM_THREADS=1 for nproc > 1; the benchmark's CPU
// legs set the pool explicitly (the reference library's own benchmark uses all hardware threads, submodules/bvh/test/benchmark.cpp:142-144).
void lvo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"
