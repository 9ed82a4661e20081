// lv_emu.cpp -- HOST EMULATION of the per-thread device code (test infrastructure; CPU test suite only).
//
// The product's per-thread device functions (linevis_b200/csrc/*.cuh, declared LV_DEV) are compiled here for the host with
// -DLV_HOST_EMU and plain g++ (strict float: -ffp-contract=off, FMA only where the source spells std::fmaf), and exported
// with a C interface so that tests/test_emu.py can compare them bit for bit with the oracle WITHOUT a GPU.  Warp-collective
// code (packet traversal, the ray-stream loop, the PPLL sort) is not covered -- that is what the -m gpu tests are for.
// Nothing here is linked into the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../linevis_b200/csrc/lv_shade.cuh"
#include "../../linevis_b200/csrc/lv_trace.cuh"
#include "../../linevis_b200/csrc/lv_bake.cuh"

using namespace lv;

extern "C" {

float emu_det_acos(float x) { return det_acos(x); }
float emu_det_pow(float x, float y) { return det_pow(x, y); }
uint32_t emu_lcg_skip(uint32_t state, uint32_t n) { return lcg_skip(state, n); }
uint32_t emu_lcg_iterate(uint32_t state, uint32_t n) { for (uint32_t i = 0; i < n; i++) lcg(state); return state; }
uint32_t emu_tea(uint32_t a, uint32_t b) { return tea(a, b); }

// IntersectionTube + the acceptance rule's own-AABB slab test for one ray / one record
int emu_accept(const float* ro, const float* rd, const float* rec8, float radius, int capped, float tmin, float tmax, float* t_out, uint32_t* kind_out) {
    SegRec s; std::memcpy(&s, rec8, 32);
    const Vec3 o = v3(ro[0], ro[1], ro[2]), d = v3(rd[0], rd[1], rd[2]);
    const RayQ rq = make_rayq(o, d);
    const RayBox rb = make_raybox(o, d);
    float t; uint32_t k;
    const bool ok = seg_box_hit(rb, s, radius, tmin, tmax) && capsule_hit(rq, s, radius, capped != 0, t, k) && t >= tmin && t <= tmax;
    if (ok) { *t_out = t; *kind_out = k; }
    return ok ? 1 : 0;
}

struct EmuShade {
    lv_camera cam;
    float line_width;
    int use_capped, use_halos, use_ao, use_static_ao;
    float ao_strength, ao_gamma, depth_cue_strength;
    int use_depth_cues;
    float depth_min_max[2];
    const float* tf; uint32_t tfK; float amin, amax;
    const float* ao_tex;
    const float* sao_factors; const float* sao_weights;
    uint32_t n_ao_subdiv, n_line_vertices, n_param_vertices;
};

static FrameParams make_params(const EmuShade& e) {
    FrameParams P;
    std::memset(&P, 0, sizeof(P));
    std::memcpy(P.view, e.cam.view, 64); std::memcpy(P.proj, e.cam.proj, 64);
    std::memcpy(P.inv_view, e.cam.inv_view, 64); std::memcpy(P.inv_proj, e.cam.inv_proj, 64);
    std::memcpy(P.cam_pos, e.cam.position, 12);
    P.fov_y = e.cam.fov_y;
    for (int k = 0; k < 4; k++) { P.bg[k] = e.cam.background[k]; P.fg[k] = 1.0f - e.cam.background[k]; }
    P.W = e.cam.width; P.H = e.cam.height;
    P.line_width = e.line_width;
    P.use_capped = e.use_capped; P.use_halos = e.use_halos; P.use_ao = e.use_ao; P.use_static_ao = e.use_static_ao;
    P.ao_strength = e.ao_strength; P.ao_gamma = e.ao_gamma;
    P.use_depth_cues = e.use_depth_cues; P.depth_cue_strength = e.depth_cue_strength; P.depth_min_max = e.depth_min_max;
    P.near_dist = e.cam.near_dist; P.far_dist = e.cam.far_dist;
    P.tf = reinterpret_cast<const float4*>(e.tf); P.tfK = e.tfK; P.amin = e.amin; P.amax = e.amax;
    P.ao_tex = e.ao_tex;
    P.sao_factors = e.sao_factors; P.sao_weights = e.sao_weights;
    P.n_ao_subdiv = e.n_ao_subdiv; P.n_line_vertices = e.n_line_vertices; P.n_param_vertices = e.n_param_vertices;
    return P;
}

// shade_hit for n hits.  ro/rd: n*3; t: n; kind: n; recs: n*8 floats (SegRec); aux: n*8 floats (SegAux) or NULL.
// out: n*5 floats (rgba, hitT)
void emu_shade_hits(const EmuShade* e, uint64_t n, const float* ro, const float* rd, const float* t, const uint32_t* kind,
                    const float* recs, const float* aux, float* out) {
    const FrameParams P = make_params(*e);
    // the TF LUT must be 16-byte aligned for float4 access: copy
    std::vector<float4> tf(e->tfK);
    std::memcpy(tf.data(), e->tf, size_t(e->tfK) * 16);
    FrameParams Q = P; Q.tf = tf.data();
    for (uint64_t i = 0; i < n; i++) {
        SegRec s; std::memcpy(&s, recs + 8 * i, 32);
        SegAux a; if (aux) std::memcpy(&a, aux + 8 * i, 32);
        const Vec3 o = v3(ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]), d = v3(rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]);
        const Shaded sh = e->use_static_ao ? shade_hit<true>(Q, o, d, t[i], kind[i], s, aux ? &a : nullptr)
                                           : shade_hit<false>(Q, o, d, t[i], kind[i], s, nullptr);
        out[5 * i] = sh.color.x; out[5 * i + 1] = sh.color.y; out[5 * i + 2] = sh.color.z; out[5 * i + 3] = sh.color.w; out[5 * i + 4] = sh.hit_t;
    }
}

float emu_ao_factor_static(const EmuShade* e, float vertex_id, float phi) {
    const FrameParams P = make_params(*e);
    return ao_factor_static(P, vertex_id, phi);
}

// k_bake_setup, thread by thread.  pt_*: n_line_pts*4 floats (float4 layout); records out: n_param*n_subdiv*12 floats
void emu_bake_records(const float* pt_pos4, const float* pt_tan4, const float* pt_nrm4, const float* sampling, uint32_t n_line_pts,
                      uint32_t n_param, uint32_t n_subdiv, uint32_t spp, uint32_t frame_number, float line_radius, float* records) {
    std::vector<float4> p(n_line_pts), t(n_line_pts), nn(n_line_pts);
    std::memcpy(p.data(), pt_pos4, size_t(n_line_pts) * 16); std::memcpy(t.data(), pt_tan4, size_t(n_line_pts) * 16);
    std::memcpy(nn.data(), pt_nrm4, size_t(n_line_pts) * 16);
    BakeParams B;
    B.pt_pos = p.data(); B.pt_tan = t.data(); B.pt_nrm = nn.data(); B.sampling = sampling;
    B.n_line_pts = n_line_pts; B.n_param = n_param; B.n_subdiv = n_subdiv; B.spp = spp; B.frame_number = frame_number; B.line_radius = line_radius;
    for (uint32_t i = 0; i < n_param * n_subdiv; i++) {
        const AoHit r = bake_record(B, i / n_subdiv, i % n_subdiv);
        std::memcpy(records + 12 * size_t(i), &r, 48);
    }
}

// the ray k_rtao_rays builds for (record, sample): org[3], dir[3]
void emu_ao_ray(const float* record12, uint32_t sample, uint32_t spp, uint32_t frame_number, int bake, float* org, float* dir) {
    AoHit r; std::memcpy(&r, record12, 48);
    Vec3 o, d;
    if (bake) ao_ray_from_record<true>(&r, sample, spp, frame_number, o, d);
    else ao_ray_from_record<false>(&r, sample, spp, frame_number, o, d);
    org[0] = o.x; org[1] = o.y; org[2] = o.z; dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
}

// k_seg_aux for one record
void emu_seg_aux(uint32_t i0, uint32_t i1, const float* pt_nrm4, uint32_t n_line_pts, float* aux8) {
    std::vector<float4> nn(n_line_pts);
    std::memcpy(nn.data(), pt_nrm4, size_t(n_line_pts) * 16);
    const SegAux a = make_seg_aux(make_uint2(i0, i1), nn.data());
    std::memcpy(aux8, &a, 32);
}

}  // extern "C"
