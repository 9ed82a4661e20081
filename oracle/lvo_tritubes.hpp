/*
 * lvo_tritubes.hpp -- the reference's TRIANGULATED capped tubes and the RTAO pass traced against them.
 *
 * TEST INFRASTRUCTURE ONLY (see lvo_shaders.hpp).  The reference traces its RTAO rays (screen-space pass and prebaker)
 * against the triangle mesh of the tubes (N-gon cross sections, hemispherical caps), not against the analytic capsules
 * (src/Renderers/AmbientOcclusion/VulkanRayTracedAmbientOcclusion.cpp:444-445; DESIGN.md rule 5).  This header restates
 *   - the mesh generator: createCappedTriangleTubesRenderDataCPU + addHemisphereToMeshStart/Stop
 *     (src/Renderers/Tubes/CappedTriangleTubesCPU.cpp:33-385), initGlobalCircleVertexPositions + insertOrientedCirclePoints
 *     (src/Renderers/Tubes/Tubes.cpp:35-86), open tubes only (tubeClosed == false as in LineDataFlow.cpp:1975-1980);
 *   - the barycentric fetch of the RTAO shader (Data/Shaders/AO/RTAO/VulkanRayTracedAmbientOcclusion.glsl:205-276,
 *     BarycentricInterpolation.glsl:39-41)
 * so that the systematic difference between the analytic stand-in and the reference's geometry can be measured, and as
 * the oracle of a future triangle-tube mode of the CUDA path.  The hardware's ray/triangle test is not specified by
 * Vulkan beyond being watertight; it is fixed here as double-sided Moeller-Trumbore in float32 with hit acceptance
 * t in [tMin, tMax], closest hit = smallest t, ties -> lowest triangle index.  PARITY UNPINNED like the rest.
 */
#ifndef LVO_TRITUBES_HPP
#define LVO_TRITUBES_HPP
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "lvo_shaders.hpp"

namespace lvo {

struct TubeTriangleVertexData {   // src/LineData/LineRenderData.hpp:171-176
    vec3 vertexPosition;
    uint32_t vertexLinePointIndex;
    vec3 vertexNormal;
    float phi;
};

struct TubeMesh {
    std::vector<uint32_t> triangleIndices;
    std::vector<TubeTriangleVertexData> vertexDataList;
    std::vector<vec3> linePositions, lineTangents, lineNormals;   // tubeTriangleLinePointDataList (LineDataFlow.cpp:1997-2012)
    std::vector<uint32_t> lineSources;                            // index of the input point a mesh line point was made from (its lineAttribute)
};

namespace sglc {   // sgl/Math/Math.hpp:47-49
const float PI = 3.1415926535897932f;
const float TWO_PI = PI * 2.0f;
const float HALF_PI = PI / 2.0f;
}

// Tubes.cpp:35-52
inline std::vector<vec3> initCircleVertexPositions(int numCircleSubdivisions, float tubeRadius) {
    std::vector<vec3> globalCircleVertexPositions;
    const float theta = sglc::TWO_PI / numCircleSubdivisions;
    const float tangentialFactor = std::tan(theta);
    const float radialFactor = std::cos(theta);
    vec3 position = V3(tubeRadius, 0, 0);
    for (int i = 0; i < numCircleSubdivisions; i++) {
        globalCircleVertexPositions.push_back(position);
        vec3 tangent = V3(-position.y, position.x, 0);
        position = position + tangentialFactor * tangent;
        position = position * radialFactor;
    }
    return globalCircleVertexPositions;
}

// Tubes.cpp:54-86
inline void insertOrientedCirclePoints(const std::vector<vec3>& circle, vec3 center, vec3 tangent, vec3& lastNormal,
                                       uint32_t vertexLinePointIndex, std::vector<TubeTriangleVertexData>& vertexDataList) {
    vec3 helperAxis = lastNormal;
    if (length(cross(helperAxis, tangent)) < 0.01f) {
        helperAxis = V3(0.0f, 1.0f, 0.0f);
        if (length(cross(helperAxis, tangent)) < 0.01f) helperAxis = V3(0.0f, 0.0f, 1.0f);
    }
    vec3 normal = normalize(helperAxis - dot(helperAxis, tangent) * tangent);
    lastNormal = normal;
    vec3 binormal = cross(tangent, normal);
    for (size_t i = 0; i < circle.size(); i++) {
        vec3 pt = circle[i];
        vec3 transformedPoint = V3(pt.x * normal.x + pt.y * binormal.x + pt.z * tangent.x + center.x,
                                   pt.x * normal.y + pt.y * binormal.y + pt.z * tangent.y + center.y,
                                   pt.x * normal.z + pt.y * binormal.z + pt.z * tangent.z + center.z);
        TubeTriangleVertexData v{};
        v.vertexPosition = transformedPoint;
        v.vertexLinePointIndex = vertexLinePointIndex;
        v.vertexNormal = normalize(transformedPoint - center);
        v.phi = float(i) / float(circle.size()) * sglc::TWO_PI;
        vertexDataList.push_back(v);
    }
}

// one ring / pole vertex of a cap (CappedTriangleTubesCPU.cpp:52-76 and :141-165)
inline TubeTriangleVertexData capVertex(vec3 center, vec3 scaledNormal, vec3 scaledBinormal, vec3 scaledTangent, float theta, float phi,
                                        uint32_t vertexLinePointIndex, float storedPhi) {
    vec3 pt = V3(std::cos(theta) * std::sin(phi), std::sin(theta) * std::sin(phi), std::cos(phi));
    vec3 off = V3(pt.x * scaledNormal.x + pt.y * scaledBinormal.x + pt.z * scaledTangent.x,
                  pt.x * scaledNormal.y + pt.y * scaledBinormal.y + pt.z * scaledTangent.y,
                  pt.x * scaledNormal.z + pt.y * scaledBinormal.z + pt.z * scaledTangent.z);
    TubeTriangleVertexData v{};
    // the reference adds the centre inside the same expression (pt.x*n + pt.y*b + pt.z*t + c, left to right): off + center is that sum
    v.vertexPosition = V3(off.x + center.x, off.y + center.y, off.z + center.z);
    v.vertexLinePointIndex = vertexLinePointIndex | 0x80000000u;
    v.vertexNormal = normalize(off);
    v.phi = storedPhi;
    return v;
}

// CappedTriangleTubesCPU.cpp:33-119
inline void addHemisphereToMeshStart(vec3 center, vec3 tangent, vec3 normal, uint32_t indexOffsetCap, uint32_t triOffsetCap,
                                     uint32_t vertexLinePointIndex, float tubeRadius, int numLongitudeSubdivisions, int numLatitudeSubdivisions,
                                     std::vector<uint32_t>& triangleIndices, std::vector<TubeTriangleVertexData>& vertexDataList) {
    vec3 binormal = cross(normal, tangent);
    vec3 scaledTangent = tubeRadius * tangent, scaledNormal = tubeRadius * normal, scaledBinormal = tubeRadius * binormal;
    uint32_t vertexOffsetCap = indexOffsetCap;
    for (int lat = numLatitudeSubdivisions; lat >= 1; lat--) {
        float phi = sglc::HALF_PI * (1.0f - float(lat) / float(numLatitudeSubdivisions));
        for (int lon = 0; lon < numLongitudeSubdivisions; lon++) {
            float theta = sglc::TWO_PI * float(lon) / float(numLongitudeSubdivisions);
            vertexDataList.at(vertexOffsetCap++) = capVertex(center, scaledNormal, scaledBinormal, scaledTangent, theta, phi, vertexLinePointIndex, theta);
            if (lat == numLatitudeSubdivisions) break;
        }
    }
    const int L = numLongitudeSubdivisions;
    for (int lat = 0; lat < numLatitudeSubdivisions; lat++) {
        for (int lon = 0; lon < L; lon++) {
            if (lat > 0) {
                triangleIndices.at(triOffsetCap++) = indexOffsetCap + 1 + (lon) % L + (lat - 1) * L;
                triangleIndices.at(triOffsetCap++) = indexOffsetCap + 1 + (lon + 1) % L + (lat - 1) * L;
                triangleIndices.at(triOffsetCap++) = indexOffsetCap + 1 + (lon) % L + (lat) * L;
                triangleIndices.at(triOffsetCap++) = indexOffsetCap + 1 + (lon + 1) % L + (lat - 1) * L;
                triangleIndices.at(triOffsetCap++) = indexOffsetCap + 1 + (lon + 1) % L + (lat) * L;
                triangleIndices.at(triOffsetCap++) = indexOffsetCap + 1 + (lon) % L + (lat) * L;
            } else {
                triangleIndices.at(triOffsetCap++) = indexOffsetCap;
                triangleIndices.at(triOffsetCap++) = indexOffsetCap + 1 + (lon + 1) % L;
                triangleIndices.at(triOffsetCap++) = indexOffsetCap + 1 + (lon) % L;
            }
        }
    }
}

// CappedTriangleTubesCPU.cpp:121-212
inline void addHemisphereToMeshStop(vec3 center, vec3 tangent, vec3 normal, uint32_t indexOffset, uint32_t indexOffsetCap, uint32_t triOffsetCap,
                                    uint32_t vertexLinePointIndex, float tubeRadius, int numLongitudeSubdivisions, int numLatitudeSubdivisions,
                                    std::vector<uint32_t>& triangleIndices, std::vector<TubeTriangleVertexData>& vertexDataList) {
    vec3 binormal = cross(normal, tangent);
    vec3 scaledTangent = tubeRadius * tangent, scaledNormal = tubeRadius * normal, scaledBinormal = tubeRadius * binormal;
    uint32_t vertexIndexOffset = indexOffsetCap - indexOffset - numLongitudeSubdivisions;
    for (int lat = 1; lat <= numLatitudeSubdivisions; lat++) {
        float phi = sglc::HALF_PI * (1.0f - float(lat) / float(numLatitudeSubdivisions));
        for (int lon = 0; lon < numLongitudeSubdivisions; lon++) {
            float theta = -sglc::TWO_PI * float(lon) / float(numLongitudeSubdivisions);
            vertexDataList.at(indexOffsetCap++) = capVertex(center, scaledNormal, scaledBinormal, scaledTangent, theta, phi, vertexLinePointIndex, -theta);
            if (lat == numLatitudeSubdivisions) break;
        }
    }
    const int L = numLongitudeSubdivisions;
    for (int lat = 0; lat < numLatitudeSubdivisions; lat++) {
        for (int lon = 0; lon < L; lon++) {
            if (lat < numLatitudeSubdivisions - 1) {
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + (lon) % L + (lat) * L;
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + (lon + 1) % L + (lat) * L;
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + (lon) % L + (lat + 1) * L;
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + (lon + 1) % L + (lat) * L;
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + (lon + 1) % L + (lat + 1) * L;
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + (lon) % L + (lat + 1) * L;
            } else {
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + (lon) % L + (lat) * L;
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + (lon + 1) % L + (lat) * L;
                triangleIndices.at(triOffsetCap++) = indexOffset + vertexIndexOffset + 0 + (lat + 1) * L;
            }
        }
    }
}

// createCappedTriangleTubesRenderDataCPU, tubeClosed == false (CappedTriangleTubesCPU.cpp:214-385) + the line point list
// of LineDataFlow::getLinePassTubeTriangleMeshRenderData (LineDataFlow.cpp:1988-2012).  Polylines = (pos, line_offsets).
inline void createCappedTriangleTubes(const float* pos, const uint64_t* line_offsets, uint64_t n_lines, float tubeRadius,
                                      int numCircleSubdivisions, TubeMesh& m) {
    numCircleSubdivisions = std::max(numCircleSubdivisions, 4);
    const std::vector<vec3> circle = initCircleVertexPositions(numCircleSubdivisions, tubeRadius);
    int numLongitudeSubdivisions = numCircleSubdivisions;
    int numLatitudeSubdivisions = int(std::ceil(numCircleSubdivisions / 2));
    uint32_t numCapVertices = numLongitudeSubdivisions * (numLatitudeSubdivisions - 1) + 1;
    uint32_t numCapIndices = numLongitudeSubdivisions * (numLatitudeSubdivisions - 1) * 6 + numLongitudeSubdivisions * 3;
    auto& triangleIndices = m.triangleIndices;
    auto& vertexDataList = m.vertexDataList;

    for (uint64_t lineId = 0; lineId < n_lines; lineId++) {
        const uint64_t b = line_offsets[lineId];
        const size_t n = size_t(line_offsets[lineId + 1] - b);
        auto lineCenters = [&](size_t i) { return V3(pos[3 * (b + i)], pos[3 * (b + i) + 1], pos[3 * (b + i) + 2]); };
        auto lineIndexOffset = uint32_t(m.lineTangents.size());
        if (n < 2) continue;

        auto indexOffsetCapStart = uint32_t(vertexDataList.size());
        auto triOffsetCapStart = uint32_t(triangleIndices.size());
        vertexDataList.resize(vertexDataList.size() + numCapVertices);
        triangleIndices.resize(triangleIndices.size() + numCapIndices);
        auto indexOffset = uint32_t(vertexDataList.size());

        vec3 lastLineNormal = V3(1.0f, 0.0f, 0.0f);
        int firstIdx = int(n) - 2;
        int lastIdx = 1;
        int numValidLinePoints = 0;
        for (size_t i = 0; i < n; i++) {
            vec3 tangent;
            if (i == 0) tangent = lineCenters(i + 1) - lineCenters(i);
            else if (i == n - 1) tangent = lineCenters(i) - lineCenters(i - 1);
            else tangent = lineCenters((i + 1) % n) - lineCenters((i + n - 1) % n);
            float lineSegmentLength = length(tangent);
            if (lineSegmentLength < 0.0001f) continue;
            firstIdx = std::min(int(i), firstIdx);
            lastIdx = std::max(int(i), lastIdx);
            tangent = normalize(tangent);
            insertOrientedCirclePoints(circle, lineCenters(i), tangent, lastLineNormal, uint32_t(m.linePositions.size()), vertexDataList);
            m.lineTangents.push_back(tangent);
            m.lineNormals.push_back(lastLineNormal);
            m.linePositions.push_back(lineCenters(i));
            m.lineSources.push_back(uint32_t(b + i));
            numValidLinePoints++;
        }
        if (numValidLinePoints == 1) {
            vertexDataList.resize(indexOffsetCapStart);
            // (the reference leaves the reserved cap triangle indices in place here; they would reference removed vertices --
            //  a latent defect for polylines that degenerate to one point; dropped here)
            triangleIndices.resize(triOffsetCapStart);
            m.lineTangents.pop_back(); m.lineNormals.pop_back(); m.linePositions.pop_back(); m.lineSources.pop_back();
        }
        if (numValidLinePoints <= 1) {
            if (numValidLinePoints == 0) { vertexDataList.resize(indexOffsetCapStart); triangleIndices.resize(triOffsetCapStart); }
            continue;
        }
        const int N = numCircleSubdivisions;
        for (int i = 0; i < numValidLinePoints - 1; i++) {
            for (int j = 0; j < N; j++) {
                triangleIndices.push_back(indexOffset + i * N + j);
                triangleIndices.push_back(indexOffset + i * N + (j + 1) % N);
                triangleIndices.push_back(indexOffset + ((i + 1) % numValidLinePoints) * N + (j + 1) % N);
                triangleIndices.push_back(indexOffset + i * N + j);
                triangleIndices.push_back(indexOffset + ((i + 1) % numValidLinePoints) * N + (j + 1) % N);
                triangleIndices.push_back(indexOffset + ((i + 1) % numValidLinePoints) * N + j);
            }
        }
        auto indexOffsetCapEnd = uint32_t(vertexDataList.size());
        auto triOffsetCapEnd = uint32_t(triangleIndices.size());
        vertexDataList.resize(vertexDataList.size() + numCapVertices);
        triangleIndices.resize(triangleIndices.size() + numCapIndices);

        vec3 center0 = lineCenters(firstIdx);
        vec3 tangent0 = normalize(lineCenters(firstIdx) - lineCenters(firstIdx + 1));
        vec3 normal0 = m.lineNormals[lineIndexOffset];
        vec3 center1 = lineCenters(lastIdx);
        vec3 tangent1 = normalize(lineCenters(lastIdx) - lineCenters(lastIdx - 1));
        vec3 normal1 = m.lineNormals[lineIndexOffset + numValidLinePoints - 1];
        addHemisphereToMeshStart(center0, tangent0, normal0, indexOffsetCapStart, triOffsetCapStart, uint32_t(lineIndexOffset), tubeRadius,
                                 numLongitudeSubdivisions, numLatitudeSubdivisions, triangleIndices, vertexDataList);
        addHemisphereToMeshStop(center1, tangent1, normal1, indexOffset, indexOffsetCapEnd, triOffsetCapEnd, uint32_t(m.lineTangents.size() - 1),
                                tubeRadius, numLongitudeSubdivisions, numLatitudeSubdivisions, triangleIndices, vertexDataList);
    }
}

// ----------------------------------------------------------------------------------------------
// triangle BVH (median split over the longest centroid axis, leaves <= 4 triangles) + traversal
// ----------------------------------------------------------------------------------------------
struct TriBvh {
    struct Node { float bmin[3], bmax[3]; uint32_t left, count; };   // leaf: left = first triangle (in `order`), count > 0
    std::vector<Node> nodes;
    std::vector<uint32_t> order;
    const TubeMesh* mesh = nullptr;

    void triBox(uint32_t t, float* mn, float* mx) const {
        for (int k = 0; k < 3; k++) { mn[k] = 3.4e38f; mx[k] = -3.4e38f; }
        for (int c = 0; c < 3; c++) {
            const vec3 p = mesh->vertexDataList[mesh->triangleIndices[3 * t + c]].vertexPosition;
            const float q[3] = {p.x, p.y, p.z};
            for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], q[k]); mx[k] = std::max(mx[k], q[k]); }
        }
    }
    void build(const TubeMesh& m) {
        mesh = &m;
        const uint32_t nt = uint32_t(m.triangleIndices.size() / 3);
        order.resize(nt);
        std::vector<float> cen(3 * size_t(nt));
        for (uint32_t t = 0; t < nt; t++) {
            order[t] = t;
            float mn[3], mx[3]; triBox(t, mn, mx);
            for (int k = 0; k < 3; k++) cen[3 * size_t(t) + k] = 0.5f * (mn[k] + mx[k]);
        }
        nodes.clear(); nodes.reserve(2 * size_t(nt) / 2 + 2);
        nodes.push_back(Node{});
        struct Task { uint32_t node, b, e; };
        std::vector<Task> st; st.push_back({0, 0, nt});
        while (!st.empty()) {
            Task tk = st.back(); st.pop_back();
            Node nd{};
            for (int k = 0; k < 3; k++) { nd.bmin[k] = 3.4e38f; nd.bmax[k] = -3.4e38f; }
            float cmn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, cmx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
            for (uint32_t i = tk.b; i < tk.e; i++) {
                float mn[3], mx[3]; triBox(order[i], mn, mx);
                for (int k = 0; k < 3; k++) {
                    nd.bmin[k] = std::min(nd.bmin[k], mn[k]); nd.bmax[k] = std::max(nd.bmax[k], mx[k]);
                    cmn[k] = std::min(cmn[k], cen[3 * size_t(order[i]) + k]); cmx[k] = std::max(cmx[k], cen[3 * size_t(order[i]) + k]);
                }
            }
            const uint32_t cnt = tk.e - tk.b;
            int axis = 0;
            if (cmx[1] - cmn[1] > cmx[axis] - cmn[axis]) axis = 1;
            if (cmx[2] - cmn[2] > cmx[axis] - cmn[axis]) axis = 2;
            if (cnt <= 4 || !(cmx[axis] > cmn[axis])) { nd.left = tk.b; nd.count = cnt; nodes[tk.node] = nd; continue; }
            const uint32_t mid = tk.b + cnt / 2;
            std::nth_element(order.begin() + tk.b, order.begin() + mid, order.begin() + tk.e,
                             [&](uint32_t a, uint32_t c) { return cen[3 * size_t(a) + axis] < cen[3 * size_t(c) + axis]; });
            nd.left = uint32_t(nodes.size()); nd.count = 0;
            nodes[tk.node] = nd;
            nodes.push_back(Node{}); nodes.push_back(Node{});
            st.push_back({nd.left, tk.b, mid}); st.push_back({nd.left + 1, mid, tk.e});
        }
    }
};

struct TriHit { float t, u, v; uint32_t tri; };

// double-sided Moeller-Trumbore; (u, v) = barycentric weights of vertex 1 and 2 (rayQueryGetIntersectionBarycentricsEXT)
inline bool rayTriangle(vec3 o, vec3 d, vec3 p0, vec3 p1, vec3 p2, float tmin, float tmax, float& t, float& u, float& v) {
    const vec3 e1 = p1 - p0, e2 = p2 - p0;
    const vec3 pv = cross(d, e2);
    const float det = dot(e1, pv);
    if (det == 0.0f) return false;
    const float inv = 1.0f / det;
    const vec3 tv = o - p0;
    u = dot(tv, pv) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    const vec3 qv = cross(tv, e1);
    v = dot(d, qv) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot(e2, qv) * inv;
    return t >= tmin && t <= tmax;
}

// Acceptance (BVH independent, like the capsules'): the ray's [tmin, tmax] meets the triangle's own AABB under the canonical slab
// test AND Moeller-Trumbore reports t in [tmin, tmax].  cullMargin: closest-hit traversals that care about WHICH triangle ties
// (primary rays) cull boxes against best + margin so that every tying candidate is seen whatever the tree looks like.
inline bool traceTriangles(const TriBvh& bvh, vec3 o, vec3 d, float tmin, float tmax, bool anyHit, TriHit& best, uint64_t& steps, uint64_t& isect,
                           float cullMargin = 0.0f) {
    const RayInv ri = makeRayInv(o, d);
    bool found = false;
    best.t = tmax; best.tri = 0xFFFFFFFFu; best.u = best.v = 0.0f;
    if (bvh.nodes.empty() || bvh.order.empty()) return false;
    uint32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const TriBvh::Node& nd = bvh.nodes[stack[--sp]];
        float tn;
        steps++;
        if (!slabTest(ri, nd.bmin, nd.bmax, tmin, best.t + cullMargin, tn)) continue;
        if (nd.count) {
            for (uint32_t i = 0; i < nd.count; i++) {
                const uint32_t tri = bvh.order[nd.left + i];
                const uint32_t* ix = &bvh.mesh->triangleIndices[3 * size_t(tri)];
                float t, u, v, own_mn[3], own_mx[3], tn2;
                isect++;
                bvh.triBox(tri, own_mn, own_mx);
                if (slabTest(ri, own_mn, own_mx, tmin, tmax, tn2) &&
                    rayTriangle(o, d, bvh.mesh->vertexDataList[ix[0]].vertexPosition, bvh.mesh->vertexDataList[ix[1]].vertexPosition,
                                bvh.mesh->vertexDataList[ix[2]].vertexPosition, tmin, tmax, t, u, v)) {
                    if (!found || t < best.t || (t == best.t && tri < best.tri)) { best = TriHit{t, u, v, tri}; found = true; }
                    if (anyHit) return true;
                }
            }
        } else if (sp + 2 <= 128) { stack[sp++] = nd.left; stack[sp++] = nd.left + 1; }
    }
    return found;
}

static inline vec3 interpolateVec3(vec3 v0, vec3 v1, vec3 v2, vec3 b) {   // BarycentricInterpolation.glsl:39-41
    return (v0 * b.x + v1 * b.y) + v2 * b.z;
}
static inline float interpolateFloat(float v0, float v1, float v2, vec3 b) {   // BarycentricInterpolation.glsl:35-37
    return (v0 * b.x + v1 * b.y) + v2 * b.z;
}

// ClosestHitTubeTriangles main + LineAttributesBarycentric.glsl:1-39 (TubeRayTracing.glsl:301-351): the tube pass's hit shader in
// the triangle-mesh geometry mode.  attr = lineAttribute per INPUT point (mesh line points refer to them through lineSources).
// Variant as elsewhere: USE_CAPPED_TUBES, no bands / multi-var / helicity; the screen-space AO texture is supported, the prebaked
// AO lookup (phi / fragmentVertexId interpolation) is not part of this mode here.
static inline HitColor closestHitTubeTriangles(const Uniforms& u, const TubeMesh& m, const float* attr, const TriHit& hit) {
    const uint32_t* ix = &m.triangleIndices[3 * size_t(hit.tri)];
    const vec3 barycentricCoordinates = V3(1.0f - hit.u - hit.v, hit.u, hit.v);
    const TubeTriangleVertexData& v0 = m.vertexDataList[ix[0]];
    const TubeTriangleVertexData& v1 = m.vertexDataList[ix[1]];
    const TubeTriangleVertexData& v2 = m.vertexDataList[ix[2]];
    const vec3 fragmentPositionWorld = interpolateVec3(v0.vertexPosition, v1.vertexPosition, v2.vertexPosition, barycentricCoordinates);
    const uint32_t l0 = v0.vertexLinePointIndex & 0x7FFFFFFFu, l1 = v1.vertexLinePointIndex & 0x7FFFFFFFu, l2 = v2.vertexLinePointIndex & 0x7FFFFFFFu;
    const bool isCap = (v0.vertexLinePointIndex >> 31) != 0u || (v1.vertexLinePointIndex >> 31) != 0u || (v2.vertexLinePointIndex >> 31) != 0u;
    vec3 fragmentNormal = normalize(interpolateVec3(v0.vertexNormal, v1.vertexNormal, v2.vertexNormal, barycentricCoordinates));
    vec3 fragmentTangent = normalize(interpolateVec3(m.lineTangents[l0], m.lineTangents[l1], m.lineTangents[l2], barycentricCoordinates));
    const float fragmentAttribute = interpolateFloat(attr[m.lineSources[l0]], attr[m.lineSources[l1]], attr[m.lineSources[l2]], barycentricCoordinates);
    return computeFragmentColor(u, fragmentPositionWorld, fragmentNormal, fragmentTangent, isCap, 0.0f, 0.0f, fragmentAttribute);
}

}  // namespace lvo
#endif
