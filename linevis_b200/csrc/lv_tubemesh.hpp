// lv_tubemesh.hpp -- host-side generator of the reference's triangulated capped tubes (the geometry LineVis traces its RTAO
// passes against: src/Renderers/AmbientOcclusion/VulkanRayTracedAmbientOcclusion.cpp:444-445).
//
// Follows createCappedTriangleTubesRenderDataCPU with tubeClosed == false (reference src/Renderers/Tubes/CappedTriangleTubesCPU.cpp:
// 214-385), its hemisphere caps (:33-212) and the oriented circle rings (src/Renderers/Tubes/Tubes.cpp:35-86), on the host like
// the reference, in float32 and in the reference's operation order (the CPU oracle holds an independent restatement; the two
// meshes are compared bit for bit by the tests).  Output: vertex positions / normals / line-point indices, triangle indices and
// the list of line points (position, tangent) the vertices refer to.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace lvmesh {

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 mul(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { const float inv = 1.0f / std::sqrt(dot(a, a)); return {a.x * inv, a.y * inv, a.z * inv}; }

struct Vertex { V3 position; uint32_t line_point; V3 normal; float phi; };   // TubeTriangleVertexData, LineRenderData.hpp:171-176

struct TubeMesh {
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;              // 3 per triangle
    std::vector<V3> line_pos, line_tan, line_nrm;
    std::vector<uint32_t> line_src;             // index of the input point each mesh line point was made from
};

constexpr float kPi = 3.1415926535897932f, kTwoPi = kPi * 2.0f, kHalfPi = kPi / 2.0f;   // sgl/Math/Math.hpp:47-49

// frame of a line point: Gram-Schmidt of the previous normal against the tangent (Tubes.cpp:57-67)
inline V3 next_normal(V3 last, V3 tangent) {
    V3 helper = last;
    if (length(cross(helper, tangent)) < 0.01f) {
        helper = {0.0f, 1.0f, 0.0f};
        if (length(cross(helper, tangent)) < 0.01f) helper = {0.0f, 0.0f, 1.0f};
    }
    const float d = dot(helper, tangent);
    return normalize(sub(helper, mul(d, tangent)));
}

// point of a unit-frame (n, b, t) combination plus centre, evaluated left to right like the reference's expressions
inline V3 frame_point(V3 pt, V3 n, V3 b, V3 t, V3 c) {
    return {pt.x * n.x + pt.y * b.x + pt.z * t.x + c.x, pt.x * n.y + pt.y * b.y + pt.z * t.y + c.y, pt.x * n.z + pt.y * b.z + pt.z * t.z + c.z};
}

struct CapFrame { V3 center, n, b, t; uint32_t line_point; };

inline Vertex cap_vertex(const CapFrame& f, float theta, float phi, float stored_phi) {
    const V3 pt = {std::cos(theta) * std::sin(phi), std::sin(theta) * std::sin(phi), std::cos(phi)};
    const V3 off = {pt.x * f.n.x + pt.y * f.b.x + pt.z * f.t.x, pt.x * f.n.y + pt.y * f.b.y + pt.z * f.t.y, pt.x * f.n.z + pt.y * f.b.z + pt.z * f.t.z};
    Vertex v{};
    v.position = {off.x + f.center.x, off.y + f.center.y, off.z + f.center.z};
    v.line_point = f.line_point | 0x80000000u;
    v.normal = normalize(off);
    v.phi = stored_phi;
    return v;
}

// polylines = (pos, offsets[n_lines + 1]); radius = lineWidth / 2; subdivisions = tube_num_subdivisions (>= 4 as in the reference)
inline void build(const float* pos, const uint64_t* offsets, uint64_t n_lines, float radius, int subdivisions, TubeMesh& m) {
    const int N = std::max(subdivisions, 4);
    const int n_lat = int(std::ceil(N / 2));                                  // integer division first, as written in the reference
    const uint32_t cap_vertices = uint32_t(N * (n_lat - 1) + 1), cap_indices = uint32_t(N * (n_lat - 1) * 6 + N * 3);
    // unit circle ring by repeated tangent steps (Tubes.cpp:35-52)
    std::vector<V3> circle;
    {
        const float theta = kTwoPi / N, tangential = std::tan(theta), radial = std::cos(theta);
        V3 p = {radius, 0, 0};
        for (int i = 0; i < N; i++) {
            circle.push_back(p);
            const V3 tg = {-p.y, p.x, 0};
            p = {p.x + tangential * tg.x, p.y + tangential * tg.y, p.z + tangential * tg.z};
            p = {p.x * radial, p.y * radial, p.z * radial};
        }
    }
    for (uint64_t li = 0; li < n_lines; li++) {
        const uint64_t first = offsets[li];
        const size_t n = size_t(offsets[li + 1] - first);
        if (n < 2) continue;
        auto P = [&](size_t i) { return V3{pos[3 * (first + i)], pos[3 * (first + i) + 1], pos[3 * (first + i) + 2]}; };
        const uint32_t line_base = uint32_t(m.line_pos.size());
        const uint32_t cap0_v = uint32_t(m.vertices.size()), cap0_i = uint32_t(m.indices.size());
        m.vertices.resize(m.vertices.size() + cap_vertices);
        m.indices.resize(m.indices.size() + cap_indices);
        const uint32_t ring_base = uint32_t(m.vertices.size());
        V3 last_normal = {1.0f, 0.0f, 0.0f};
        int first_idx = int(n) - 2, last_idx = 1, valid = 0;
        for (size_t i = 0; i < n; i++) {
            V3 tangent = i == 0 ? sub(P(1), P(0)) : (i == n - 1 ? sub(P(i), P(i - 1)) : sub(P(i + 1), P(i - 1)));
            if (length(tangent) < 0.0001f) continue;                          // nearly identical neighbours: the point is skipped
            first_idx = std::min(int(i), first_idx); last_idx = std::max(int(i), last_idx);
            tangent = normalize(tangent);
            const V3 normal = next_normal(last_normal, tangent);
            last_normal = normal;
            const V3 binormal = cross(tangent, normal), c = P(i);
            for (int k = 0; k < N; k++) {
                Vertex v{};
                v.position = frame_point(circle[k], normal, binormal, tangent, c);
                v.line_point = uint32_t(m.line_pos.size());
                v.normal = normalize(sub(v.position, c));
                v.phi = float(k) / float(N) * kTwoPi;
                m.vertices.push_back(v);
            }
            m.line_pos.push_back(c); m.line_tan.push_back(tangent); m.line_nrm.push_back(normal); m.line_src.push_back(uint32_t(first + i));
            valid++;
        }
        if (valid <= 1) {                                                     // nothing (or a single point) left: the polyline vanishes
            m.vertices.resize(cap0_v); m.indices.resize(cap0_i);
            m.line_pos.resize(line_base); m.line_tan.resize(line_base); m.line_nrm.resize(line_base); m.line_src.resize(line_base);
            continue;
        }
        for (int i = 0; i < valid - 1; i++)
            for (int j = 0; j < N; j++) {                                     // two CCW triangles per side quad
                const uint32_t a = ring_base + i * N + j, b = ring_base + i * N + (j + 1) % N;
                const uint32_t c2 = ring_base + (i + 1) * N + (j + 1) % N, d = ring_base + (i + 1) * N + j;
                const uint32_t quad[6] = {a, b, c2, a, c2, d};
                m.indices.insert(m.indices.end(), quad, quad + 6);
            }
        const uint32_t cap1_v = uint32_t(m.vertices.size()), cap1_i = uint32_t(m.indices.size());
        m.vertices.resize(m.vertices.size() + cap_vertices);
        m.indices.resize(m.indices.size() + cap_indices);
        // start cap: pole first, then rings towards the tube's first ring, which its last band of quads shares (:33-119)
        {
            const V3 t0 = normalize(sub(P(first_idx), P(first_idx + 1))), n0 = m.line_nrm[line_base];
            const CapFrame f = {P(first_idx), mul(radius, n0), mul(radius, cross(n0, t0)), mul(radius, t0), line_base};
            uint32_t v = cap0_v, t = cap0_i;
            for (int lat = n_lat; lat >= 1; lat--) {
                const float phi = kHalfPi * (1.0f - float(lat) / float(n_lat));
                for (int lon = 0; lon < N; lon++) {
                    const float theta = kTwoPi * float(lon) / float(N);
                    m.vertices[v++] = cap_vertex(f, theta, phi, theta);
                    if (lat == n_lat) break;
                }
            }
            for (int lat = 0; lat < n_lat; lat++)
                for (int lon = 0; lon < N; lon++) {
                    const uint32_t r0 = cap0_v + 1 + (lat - 1) * N, r1 = cap0_v + 1 + lat * N, a = lon % N, b = (lon + 1) % N;
                    if (lat > 0) { const uint32_t q[6] = {r0 + a, r0 + b, r1 + a, r0 + b, r1 + b, r1 + a}; for (uint32_t x : q) m.indices[t++] = x; }
                    else { const uint32_t q[3] = {cap0_v, cap0_v + 1 + b, cap0_v + 1 + a}; for (uint32_t x : q) m.indices[t++] = x; }
                }
        }
        // end cap: continues from the tube's last ring, rings towards the pole (:121-212)
        {
            const V3 t1 = normalize(sub(P(last_idx), P(last_idx - 1))), n1 = m.line_nrm[line_base + valid - 1];
            const CapFrame f = {P(last_idx), mul(radius, n1), mul(radius, cross(n1, t1)), mul(radius, t1), uint32_t(m.line_pos.size() - 1)};
            uint32_t v = cap1_v, t = cap1_i;
            for (int lat = 1; lat <= n_lat; lat++) {
                const float phi = kHalfPi * (1.0f - float(lat) / float(n_lat));
                for (int lon = 0; lon < N; lon++) {
                    const float theta = -kTwoPi * float(lon) / float(N);
                    m.vertices[v++] = cap_vertex(f, theta, phi, -theta);
                    if (lat == n_lat) break;
                }
            }
            const uint32_t last_ring = cap1_v - N;
            for (int lat = 0; lat < n_lat; lat++)
                for (int lon = 0; lon < N; lon++) {
                    const uint32_t r0 = last_ring + lat * N, r1 = last_ring + (lat + 1) * N, a = lon % N, b = (lon + 1) % N;
                    if (lat < n_lat - 1) { const uint32_t q[6] = {r0 + a, r0 + b, r1 + a, r0 + b, r1 + b, r1 + a}; for (uint32_t x : q) m.indices[t++] = x; }
                    else { const uint32_t q[3] = {r0 + a, r0 + b, r1}; for (uint32_t x : q) m.indices[t++] = x; }
                }
        }
    }
}

}  // namespace lvmesh
