// lv_shade.cuh -- device implementation of the reference's intersection and hit shaders.
//
// Reference (paths under the LineVis tree):
//   S2  Data/Shaders/Renderers/RayTracing/RayIntersectionTestsVulkan.glsl:39-119 + TubeRayTracing.glsl:452-494
//   S3  TubeRayTracing.glsl:512-613 -> RayHitCommon.glsl:74-543 -> Utils/Lighting.glsl:100-191,
//       Utils/TransferFunction.glsl:66-71, Utils/AmbientOcclusion.glsl:84-99, Utils/Antialiasing.glsl:1-3
//   S4  TubeRayTracing.glsl:290-298
// Variant built: USE_CAPPED_TUBES, USE_HALOS, ANALYTIC_TUBE_INTERSECTIONS, optional
// USE_AMBIENT_OCCLUSION + GEOMETRY_PASS_TUBE; bands / multi-var / stress / MLAT / depth cues are out of scope.
#pragma once
#include "lv_math.cuh"
#include "lv_types.cuh"

namespace lv {

struct RayQ {
    Vec3 o, d;
    float dd;  // (d.x^2 + d.y^2) + d.z^2 -- the quadratic's A for both end spheres
};

LV_DEV RayQ make_rayq(Vec3 o, Vec3 d) {
    RayQ r; r.o = o; r.d = d; r.dd = (d.x * d.x + d.y * d.y) + d.z * d.z; return r;
}

// ray vs sphere of radius `rad` centred at c, first root >= 0 (RayIntersectionTestsVulkan.glsl:39-72).
// `oc` = rayOrigin - sphereCenter.
LV_DEV bool sphere_hit(const RayQ& r, Vec3 oc, float rad, float& t) {
    float A = r.dd;
    float B = 2.0f * ((r.d.x * oc.x + r.d.y * oc.y) + r.d.z * oc.z);
    float C = ((oc.x * oc.x + oc.y * oc.y) + oc.z * oc.z) - rad * rad;
    float disc = B * B - 4.0f * A * C;
    if (disc < 0.0f) return false;
    float sq = sqrtf(disc);
    float den = 2.0f * A;
    float t0 = (-B - sq) / den;
    if (t0 >= 0.0f) { t = t0; return true; }
    float t1 = (-B + sq) / den;
    if (t1 >= 0.0f) { t = t1; return true; }
    return false;
}

// ray vs open finite cylinder (RayIntersectionTestsVulkan.glsl:78-119)
// `axis` = normalize3(p1 - p0): a property of the segment, passed in so that callers that test one segment against many rays (the
// object-order PPLL gather) compute it once; the arithmetic is the same either way.
LV_DEV bool cylinder_hit(const RayQ& r, Vec3 p0, Vec3 p1, Vec3 axis, Vec3 op0, float rad, float& t) {
    Vec3 dperp = r.d - dot3(r.d, axis) * axis;
    Vec3 pperp = op0 - dot3(op0, axis) * axis;
    float A = (dperp.x * dperp.x + dperp.y * dperp.y) + dperp.z * dperp.z;
    float B = 2.0f * dot3(dperp, pperp);
    float C = ((pperp.x * pperp.x + pperp.y * pperp.y) + pperp.z * pperp.z) - rad * rad;
    float disc = B * B - 4.0f * A * C;
    if (disc < 0.0f) return false;
    float sq = sqrtf(disc);
    float den = 2.0f * A;
    float t0 = (-B - sq) / den;
    if (t0 >= 0.0f) {
        Vec3 ip = r.o + t0 * r.d;
        if (dot3(axis, ip - p0) > 0.0f && dot3(axis, ip - p1) < 0.0f) { t = t0; return true; }
    }
    float t1 = (-B + sq) / den;
    if (t1 >= 0.0f) {
        Vec3 ip = r.o + t1 * r.d;
        if (dot3(axis, ip - p0) > 0.0f && dot3(axis, ip - p1) < 0.0f) { t = t1; return true; }
    }
    return false;
}

// IntersectionTube main (TubeRayTracing.glsl:452-494): min over body / sphere(p0) / sphere(p1).
LV_DEV bool capsule_hit(const RayQ& r, const SegRec& s, Vec3 axis, float rad, bool capped, float& t, uint32_t& kind) {
    Vec3 p0 = v3(s.a.x, s.a.y, s.a.z), p1 = v3(s.b.x, s.b.y, s.b.z);
    Vec3 op0 = r.o - p0;
    bool has = false;
    float best = 1e7f;
    uint32_t k = 0;
    float tt;
    if (cylinder_hit(r, p0, p1, axis, op0, rad, tt)) { best = tt; has = true; }
    if (capped) {
        if (sphere_hit(r, op0, rad, tt) && tt < best) { best = tt; k = 1; has = true; }
        if (sphere_hit(r, r.o - p1, rad, tt) && tt < best) { best = tt; k = 2; has = true; }
    }
    t = best; kind = k;
    return has;
}
LV_DEV Vec3 seg_axis(const SegRec& s) { return normalize3(v3(s.b.x, s.b.y, s.b.z) - v3(s.a.x, s.a.y, s.a.z)); }
LV_DEV bool capsule_hit(const RayQ& r, const SegRec& s, float rad, bool capped, float& t, uint32_t& kind) {
    return capsule_hit(r, s, seg_axis(s), rad, capped, t, kind);
}

// ---------------------------------------------------------------------------------------------
LV_DEV Vec4 tf_lookup(const FrameParams& P, float attr) {  // TransferFunction.glsl:66-71
    float pos = clampf_((attr - P.amin) / (P.amax - P.amin), 0.0f, 1.0f);
    float x = pos * float(P.tfK) - 0.5f;
    float fl = floorf(x);
    float w = x - fl;
    int i0 = int(fl), i1 = i0 + 1, km = int(P.tfK) - 1;
    i0 = max(0, min(i0, km)); i1 = max(0, min(i1, km));
    float4 a = __ldg(P.tf + i0), b = __ldg(P.tf + i1);
    return v4(mixf_(a.x, b.x, w), mixf_(a.y, b.y, w), mixf_(a.z, b.z, w), mixf_(a.w, b.w, w));
}

LV_DEV float ao_tex_bilinear(const FrameParams& P, float u, float v) {
    float x = u * float(P.W) - 0.5f, y = v * float(P.H) - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float wx = x - fx, wy = y - fy;
    int x0 = int(fx), y0 = int(fy), x1 = x0 + 1, y1 = y0 + 1;
    int xm = int(P.W) - 1, ym = int(P.H) - 1;
    x0 = max(0, min(x0, xm)); x1 = max(0, min(x1, xm));
    y0 = max(0, min(y0, ym)); y1 = max(0, min(y1, ym));
    const float* t = P.ao_tex;
    float a = mixf_(__ldg(t + size_t(y0) * P.W + x0), __ldg(t + size_t(y0) * P.W + x1), wx);
    float b = mixf_(__ldg(t + size_t(y1) * P.W + x0), __ldg(t + size_t(y1) * P.W + x1), wx);
    return mixf_(a, b, wy);
}

LV_DEV float ao_factor(const FrameParams& P, Vec3 view_pos) {  // AmbientOcclusion.glsl:84-99
    Vec4 ndc = mat_mul(P.proj, v4(view_pos.x, view_pos.y, view_pos.z, 1.0f));
    float nx = ndc.x / ndc.w, ny = ndc.y / ndc.w;
    float ao = ao_tex_bilinear(P, nx * 0.5f + 0.5f, ny * 0.5f + 0.5f);
    ao = det_pow(ao, P.ao_gamma);
    return maxf_(0.0f, 1.0f - P.ao_strength + P.ao_strength * ao);
}

// getAoFactor(interpolatedVertexId, phi), STATIC_AMBIENT_OCCLUSION_PREBAKING variant (reference Utils/AmbientOcclusion.glsl:49-75):
// blending weight of the two neighbouring line points -> position on the baked parametrization; phi -> position on the
// circle of n_ao_subdiv baked directions; bilinear mix of the four factors.
LV_DEV float ao_factor_static(const FrameParams& P, float vertex_id, float phi) {
    const uint32_t last_pt = uint32_t(vertex_id);
    const uint32_t next_pt = last_pt + 1u < P.n_line_vertices - 1u ? last_pt + 1u : P.n_line_vertices - 1u;
    const float f_pt = vertex_id - floorf(vertex_id);
    const float w = mixf_(__ldg(P.sao_weights + last_pt), __ldg(P.sao_weights + next_pt), f_pt);
    const uint32_t last_v = uint32_t(w);
    const uint32_t next_v = last_v + 1u < P.n_param_vertices - 1u ? last_v + 1u : P.n_param_vertices - 1u;
    const float f_line = w - floorf(w);
    const uint32_t N = P.n_ao_subdiv;
    const float circle = clampf_(phi / 6.28318531f * float(N), 0.0f, float(N));
    const uint32_t c_last = (uint32_t(floorf(circle)) + N) % N;
    const uint32_t c_next = (c_last + 1u) % N;
    const float f_circle = circle - floorf(circle);
    const float a00 = __ldg(P.sao_factors + c_last + size_t(N) * last_v), a01 = __ldg(P.sao_factors + c_last + size_t(N) * next_v);
    const float a10 = __ldg(P.sao_factors + c_next + size_t(N) * last_v), a11 = __ldg(P.sao_factors + c_next + size_t(N) * next_v);
    float ao = mixf_(mixf_(a00, a01, f_line), mixf_(a10, a11, f_line), f_circle);
    ao = det_pow(ao, P.ao_gamma);
    return maxf_(0.0f, 1.0f - P.ao_strength + P.ao_strength * ao);
}

struct Shaded { Vec4 color; float hit_t; };

LV_DEV float aa_factor(const FrameParams& P, float distance) {  // Utils/Antialiasing.glsl:1-3
    return distance / float(P.H) * P.fov_y;
}

// ClosestHitTubeAnalytic + computeFragmentColor + blinnPhongShadingTube for one accepted hit.
// SAO = prebaked object-space AO ("RTAO (Prebaker)") instead of the screen-space AO texture; a template parameter so that the
// default instantiation carries none of its registers.  `aux`: the record's line-point data, read only with SAO.
// The terms of shade_hit that depend on the segment alone (one segment, many fragments: the object-order PPLL gather hoists them).
struct SegShade {
    Vec3 seg;        // p1 - p0
    float seg_dot;   // dot(seg, seg)
    Vec3 tan0;       // normalize(seg)            TubeRayTracing.glsl:544
    Vec3 tg;         // normalize(tan0)           RayHitCommon.glsl:144
    Vec3 t2;         // normalize(tg)             Lighting.glsl:139
};
LV_DEV SegShade seg_shade(const SegRec& s) {
    SegShade q;
    q.seg = v3(s.b.x, s.b.y, s.b.z) - v3(s.a.x, s.a.y, s.a.z);
    q.seg_dot = dot3(q.seg, q.seg);
    q.tan0 = normalize3(q.seg);
    q.tg = normalize3(q.tan0);
    q.t2 = normalize3(q.tg);
    return q;
}

// computeFragmentColor + blinnPhongShadingTube for a surface point (RayHitCommon.glsl:74-543): what both hit shaders -- the analytic
// tube's (shade_hit below) and the triangle mesh's (shade_tri_hit, lv_tri.cuh) -- end in.  tg = normalize(tan0), t2 = normalize(tg);
// u / aux: the analytic hit's segment parameter and line-point data for the prebaked AO lookup (SAO only).
template <bool SAO>
LV_DEV Shaded shade_surface(const FrameParams& P, Vec3 pos, Vec3 nrm0, Vec3 tan0, Vec3 tg, Vec3 t2, bool is_cap, float attr, float u, const SegAux* aux) {
    const Vec3 cam = v3(P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]);
    Vec4 base = tf_lookup(P, attr);                              // RayHitCommon.glsl:127
    const Vec3 n = normalize3(nrm0);
    const Vec3 v = normalize3(cam - pos);
    Vec3 helper = normalize3(cross3(tg, v));
    Vec3 new_v = normalize3(cross3(helper, tg));
    float ribbon = 0.0f;
    if (P.use_halos) {
        if (P.use_capped && is_cap) {                            // :193-229
            Vec3 cvn = cross3(v, n);
            ribbon = length3(cvn);
            float ribbon2 = length3(cross3(new_v, n));
            float w = dot3(tg, cvn);
            if (w < 0.0f) { ribbon2 = -ribbon2; ribbon = -ribbon; }
            ribbon2 = clampf_(ribbon2, -1.0f, 1.0f);
            if (fabsf(ribbon2) < fabsf(ribbon)) ribbon = ribbon2;
        } else {                                                 // :353-372
            Vec3 cvn = cross3(new_v, n);
            ribbon = length3(cvn);
            if (dot3(tg, cvn) < 0.0f) ribbon = -ribbon;
            ribbon = clampf_(ribbon, -1.0f, 1.0f);
        }
    }
    // blinnPhongShadingTube (Lighting.glsl:100-191)
    float aof = 1.0f, kA = 0.1f, kD = 0.9f;
    float view_z = 0.0f;
    if ((P.use_ao && !SAO) || P.use_depth_cues) {
        Vec4 vp = mat_mul(P.view, v4(pos.x, pos.y, pos.z, 1.0f));   // screenSpacePosition, RayHitCommon.glsl:389-391
        view_z = vp.z;
        if (P.use_ao && !SAO) aof = ao_factor(P, v3(vp.x, vp.y, vp.z));
    }
    if (SAO && P.use_ao) {                                       // TubeRayTracing.glsl:550-562, Lighting.glsl:118-119
        float phi = 0.0f, vertex_id = 0.0f;
        if (aux) {
            const float4 a0 = __ldg(&aux->n0), a1 = __ldg(&aux->n1);
            const Vec3 line_n = (1.0f - u) * v3(a0.x, a0.y, a0.z) + u * v3(a1.x, a1.y, a1.z);
            phi = det_acos(dot3(nrm0, line_n));
            if (dot3(line_n, cross3(nrm0, tan0)) < 0.0f) phi = 6.28318531f - phi;
            vertex_id = (1.0f - u) * float(__float_as_uint(a0.w)) + u * float(__float_as_uint(a1.w));
        }
        aof = ao_factor_static(P, vertex_id, phi);
    }
    if (P.use_ao) {
        kA = 0.2f + (1.0f - aof) * 0.5f;
        kD = 0.9f * aof;
    }
    const Vec3 n2 = normalize3(n);
    const Vec3 l = v;            // normalize(cameraPosition - fragmentPositionWorld), identical to v above
    const Vec3 h = normalize3(l + l);
    Vec3 helper_l = normalize3(cross3(t2, l));
    Vec3 new_l = normalize3(cross3(helper_l, t2));
    float c1 = det_pow(clampf_(fabsf(dot3(n2, l)), 0.0f, 1.0f), 1.7f);
    float c2 = det_pow(clampf_(fabsf(dot3(n2, new_l)), 0.0f, 1.0f), 1.7f);
    float cc = 0.3f * c1 + 0.7f * c2;
    float kdc = kD * cc;
    float spec = 0.3f * det_pow(clampf_(fabsf(dot3(n2, h)), 0.0f, 1.0f), 30.0f);
    Vec3 col = v3((kA * base.x + kdc * base.x) + spec, (kA * base.y + kdc * base.y) + spec, (kA * base.z + kdc * base.z) + spec);
    if (P.use_ao) col = col * aof;
    if (P.use_depth_cues) {                                      // Utils/Lighting.glsl:183-187
        const float dmin = __ldg(P.depth_min_max), dmax = __ldg(P.depth_min_max + 1);
        float f = clampf_((-view_z - dmin) / (dmax - dmin), 0.0f, 1.0f);
        f = f * f * P.depth_cue_strength;
        col = v3(mixf_(col.x, 0.5f, f), mixf_(col.y, 0.5f, f), mixf_(col.z, 0.5f, f));
    }
    // halo / outline (RayHitCommon.glsl:437-506)
    float abs_c = P.use_halos ? fabsf(ribbon) : 0.0f;
    float depth = length3(pos - cam);
    float eps_outline = clampf_(aa_factor(P, depth / P.line_width * 0.05f), 0.0f, 0.49f);   // :451
    float eps_white = clampf_(aa_factor(P, depth / P.line_width * 2.0f), 0.0f, 0.49f);      // :452
    float coverage = P.use_halos ? 1.0f - smoothstepf_(1.0f - eps_outline, 1.0f, abs_c) : 1.0f;
    float wmix = smoothstepf_(0.7f - eps_white, 0.7f + eps_white, abs_c);                   // WHITE_THRESHOLD 0.7
    Shaded out;
    out.color = v4(mixf_(col.x, P.fg[0], wmix), mixf_(col.y, P.fg[1], wmix), mixf_(col.z, P.fg[2], wmix), base.w * coverage);
    out.hit_t = depth;                                           // payload.hitT (:540)
    return out;
}
// ClosestHitTubeAnalytic (TubeRayTracing.glsl:512-613) for one accepted hit of the analytic tube.
template <bool SAO>
LV_DEV Shaded shade_hit(const FrameParams& P, Vec3 ro, Vec3 rd, float t_hit, uint32_t kind, const SegRec& s, const SegShade& q, const SegAux* aux) {
    Vec3 p0 = v3(s.a.x, s.a.y, s.a.z), p1 = v3(s.b.x, s.b.y, s.b.z);
    Vec3 pos = ro + rd * t_hit;                                  // TubeRayTracing.glsl:517
    const Vec3 seg = q.seg;
    Vec3 centre; float attr, u;
    if (kind == 0) {                                             // :524-530
        u = dot3(seg, pos - p0) / q.seg_dot;
        centre = p0 + u * seg;
        attr = (1.0f - u) * s.a.w + u * s.b.w;
    } else if (kind == 1) { centre = p0; attr = s.a.w; u = 0.0f; }
    else { centre = p1; attr = s.b.w; u = 1.0f; }
    // fragmentTangent / fragmentNormal are normalised at :544-545 and again inside computeFragmentColor (:141,:144)
    // and blinnPhongShadingTube (Lighting.glsl:138-139); the repeated normalisations are kept, they are not idempotent in float.
    return shade_surface<SAO>(P, pos, normalize3(pos - centre), q.tan0, q.tg, q.t2, kind != 0, attr, u, aux);
}
template <bool SAO>
LV_DEV Shaded shade_hit(const FrameParams& P, Vec3 ro, Vec3 rd, float t_hit, uint32_t kind, const SegRec& s, const SegAux* aux) {
    return shade_hit<SAO>(P, ro, rd, t_hit, kind, s, seg_shade(s), aux);
}

}  // namespace lv
