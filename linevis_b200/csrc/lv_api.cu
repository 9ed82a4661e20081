// lv_api.cu -- implementation of the C ABI in include/linevis_b200.h on top of the sm_100a kernels.
// Host orchestration only: option parsing, buffer management, launches, timing, statistics.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/linevis_b200.h"
#include "lv_bvh.cuh"
#include "lv_kernels.cuh"
#include "lv_bake.cuh"
#include "lv_tubemesh.hpp"
#include "lv_sah_host.hpp"

using namespace lv;

namespace {

thread_local std::string g_global_error;

struct Options {
    // reference defaults: LineData.hpp:377-378, VulkanRayTracedAmbientOcclusion.hpp:108,150-153,
    // LineData.cpp:52, VulkanRayTracer.hpp:137-142, LineRenderer.cpp:739-740, PerPixelLinkedListLineRenderer.hpp:45-49
    float line_width = 0.002f;
    float band_width = 0.005f;
    float depth_cue_strength = 0.0f;
    bool use_capped_tubes = true, use_halos = true;
    float ao_strength = 0.0f, ao_gamma = 1.0f, ao_radius = 0.1f;
    uint32_t ao_iterations = 64, ao_spp = 4;
    bool ao_use_distance = true, ao_jitter_primary = true;
    uint32_t tube_num_subdivisions = 6;
    uint32_t num_samples_per_frame = 2, num_accumulated_frames = 32;
    bool use_deterministic_sampling = false;
    uint32_t max_depth_complexity = 1024;
    uint32_t tiling_w = 2, tiling_h = 8;
    uint32_t bvh_leaf_size = 1;
    bool bvh_cubic_morton = false;      // b200_bvh_morton = cubic: one scale for all axes in the Morton codes (default per_axis: each axis to [0, 1]; measured equal)
    bool bvh_sah_host = false;          // b200_bvh_builder = sah: top-down binned SAH on the host threads (lv_sah_host.hpp; a measurement of builder quality)
    bool bvh_ploc = false;              // b200_bvh_builder = ploc: parallel locally-ordered clustering instead of the Morton radix tree (one-record leaves only)
    uint32_t bvh_ploc_radius = 16;      // ... neighbours searched to either side per round
    uint32_t ao_refill_below = 0;       // 0 = the measured optimum of the kernel in use: 28 for k_rtao_rays_w, 30 for the other streams with b200_ao_raybuf (refilling is cheap), 24 without
    uint32_t ao_stack = 12;           // traversal stack of the AO ray kernel: 0 local 2x32-bit, 1 local packed 64-bit, K = 8 / 12 / 16 packed entries in shared memory + local spill
    bool ao_qnodes = false;           // experimental: AO ray stream over 32-byte quantised nodes (k_quantize_nodes, NodeQ); capsules + leaf queue only
    bool ao_triangles = false;        // b200_rtao_geometry = triangles: AO passes trace the reference's triangulated tubes (lv_tri.cuh)
    bool ao_queue = true;             // AO rays: leaf-queue kernel k_rtao_rays_q (one-record leaves), else the leaf-vote kernel k_rtao_rays
    uint32_t ao_min_blocks = 0;       // resident 128-thread blocks per SM the AO ray kernel is compiled for (8 / 9 / 10); 0 = best measured (queue 8, vote 9)
    bool tube_triangles = false;        // geometry_mode = "Triangle Mesh": the tube pass traces and shades the reference's triangulated tubes
    bool tube_prepass = true;           // b200_tube_prepass: the tube pass's first hits are traced on a second stream beside the RTAO pass
    bool frame_rgba8 = false;           // b200_frame_format = rgba8: rgba_out of the render calls is RGBA8 UNORM (uint32 per pixel), packed in the frame kernels' epilogue
    bool async_delivery = false;        // b200_async_delivery: an rgba8 frame for a HOST pointer is copied on a second stream from alternating staging buffers;
                                        // the call returns once the copy is enqueued, lv_synchronize waits for it (frame i's D2H overlaps frame i+1's render)
    bool ao_raybuf = true;              // b200_ao_raybuf: the AO stream generates rays 32 at a time by the whole warp into a shared batch
    bool ao_wide = true;                // b200_ao_wide: the AO ray stream traverses the 4-wide quantised tree (NodeW4)
    uint32_t ao_tq_bits = 0;            // b200_ao_tq_bits: bits of the packed stream's stack entry distances (0 = what the scene's index range allows: 7 or 4)
    int packet_carveout = -1;           // b200_packet_carveout (see lv_set_option)
    bool ao_packed = true;              // b200_ao_packed: ... with k_rtao_rays_w (lv_aostream.cuh: packed fp32x2 box tests, rays in shared memory)
    uint32_t ao_wide_reps = 1;          // ... node steps per pass of the traversal loop
    uint32_t ao_wide_top = 0;           // ... and serves the first levels (up to this many wide nodes) from shared memory (bulk-copied per block)
    bool ppll_raster_gather = true;     // b200_ppll_gather_mode = raster (default): object-order gather (one warp per segment); raycast = the BVH packet gather
    bool ppll_contiguous = false;       // ... = raster_contiguous: plus count -> scan -> fill, every list one contiguous run, pointer-free resolve
    float ppll_raster_slack = 0.1f;     // object-order gather: pixels added to the 2-D cull's bound on top of the float-error term (see the kernel)
    uint32_t ppll_raster_min_blocks = 6;  // ... blocks per SM the kernel is compiled for (4 / 5 / 6: 128 / 102 / 85 registers; config 4: 16.6 / 15.0 / 14.0 ms)
    uint32_t ppll_resolve_tile = 1024;  // plain resolve: keys per warp in the shared tile (256 / 512 / 1024; raised to hold max_frags)
    bool ppll_reg_sort = true;          // plain resolve: lists of 65..256 keys are sorted in registers (shuffles) instead of shared memory (config 4: 4.84 -> 2.87 ms)
    bool ppll_binned_resolve = false;   // count-binned resolve: faster on sparse scenes (config 2), slower on dense ones (config 4)
    uint32_t ao_leaf_vote = 12;
    uint32_t expected_avg_depth_complexity = 0;  // 0 = reference rule (20 / 120)
    std::string ao_mode = "RTAO (Screen Space)", denoiser = "None", geometry_mode = "AABBs (analytic)";
    // object-space AO prebaker (reference VulkanAmbientOcclusionBaker.hpp:108,163-168; GUI-only there, b200_prebaker_* keys here)
    bool ao_prebaker = false;
    uint32_t bake_iterations = 128, bake_spp = 4, bake_subdiv = 8;
    float bake_param_len = 0.001f, bake_radius = 0.1f;
    bool bake_use_distance = true;
};

template <class T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct lv_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    Options opt;
    std::string error;
    // transfer function
    DevBuf<float4> tf; uint32_t tfK = 0; float amin = 0.0f, amax = 1.0f;
    // sharding
    uint32_t rank = 0, world = 1, tile_size = 64;
    std::vector<uint2> tiles_host; DevBuf<uint2> tiles_dev; uint32_t tiles_w = 0, tiles_h = 0;
    std::vector<unsigned char> tile_owner; uint32_t owner_w = 0, owner_h = 0;   // lv_set_tile_owners: explicit owner per tile (Morton order) of an owner_w x owner_h frame
    DevBuf<unsigned int> tile_hist, small2;   // small2: record count + work counter of an lv_sao_trace launch
    DevBuf<unsigned char> owned_map; uint32_t tiles_x = 0;   // world > 1: 1 byte per tile of the frame, 1 = owned (object-order PPLL gather)
    DevBuf<uint2> tiles_tmp; std::vector<uint32_t> peer_off; uint32_t peer_w = 0, peer_h = 0, peer_world = 0, peer_tile = 0;
    // frame buffers
    DevBuf<unsigned int> apron_marks; unsigned int apron_stamp = 0;
    DevBuf<uint32_t> rgba8;
    // asynchronous delivery of rgba8 frames to host memory: two staging frames, a copy stream, events
    DevBuf<uint2> first_hits; cudaStream_t aux_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // b200_tube_prepass
    DevBuf<uint32_t> stage8[2]; cudaStream_t copy_stream = nullptr; cudaEvent_t ev_rendered[2] = {}, ev_copied[2] = {}; unsigned flip = 0; bool copies_pending = false;
    DevBuf<float4> image; DevBuf<float> ao, occ, depth_mm; DevBuf<lv_hit> hits; DevBuf<AoHit> ao_hits;
    uint32_t ao_w = 0, ao_h = 0;
    DevBuf<Counters> counters; DevBuf<unsigned int> small;  // small[0] = ao hit count, small[2..3] = 64-bit ao work counter
    // PPLL
    DevBuf<uint32_t> heads, counts, bin_order; DevBuf<unsigned int> bin_hist; DevBuf<lv_ppll_node> nodes; DevBuf<unsigned long long> frag_counter;
    // raster_contiguous gather: staged fragments, exclusive scan of the counts, fill cursors; lists_contiguous tells the resolve pass
    DevBuf<float4> pixel_rays;   // object-order gather: per-pixel camera rays of the frame (k_pixel_rays)
    DevBuf<lv_ppll_node> stage; DevBuf<uint32_t> list_offs; DevBuf<unsigned int> fill_cursor; DevBuf<char> scan_tmp; bool lists_contiguous = false;
    unsigned long long list_size = 0; uint32_t padded_w = 0, padded_h = 0;
    cudaEvent_t ev[8] = {};
    bool rtao_rays_timed = false, binned_attr_set = false;
};

struct lv_scene {
    lv_ctx* ctx = nullptr;
    DevBuf<SegRec> segs; DevBuf<float4> seg_axes; DevBuf<uint32_t> prim_ids; DevBuf<Node64> nodes;
    uint64_t n_seg = 0, n_nodes = 0, n_pt = 0;
    float line_width = 0.0f, build_ms = 0.0f;
    uint32_t depth = 0, leaf_size = 1;
    float bounds[6] = {0, 0, 0, 0, 0, 0};
    // line-point frames + object-space AO prebaker state (lv_scene_set_lines / lv_ao_bake)
    DevBuf<uint2> seg_idx;                         // caller's point index pairs, caller order
    DevBuf<float4> pt_pos, pt_tan, pt_nrm;         // per line point
    DevBuf<SegAux> seg_aux;                        // per record, BVH order
    std::vector<float> host_pos; std::vector<uint64_t> line_offsets;   // polylines for the (host-side) parametrization
    DevBuf<float> sampling, weights, factors;      // samplingLocations, blending weights, ambientOcclusionFactors
    DevBuf<NodeQ> qnodes; float q_origin[3] = {0, 0, 0}, q_scale[3] = {1, 1, 1};   // quantised copy of `nodes` (ensure_qnodes)
    // 4-wide quantised tree (ensure_wnodes): nodes in breadth-first order, level_end[l] = number of nodes in levels 0..l
    DevBuf<NodeW4> wnodes; float w_origin[3] = {0, 0, 0}, w_scale[3] = {1, 1, 1}; uint64_t n_wnodes = 0; uint32_t w_need = 0;
    std::vector<uint32_t> w_level_end; float w_build_ms = 0.0f; bool w_failed = false;
    uint32_t w_top = 0;   // nodes of the whole top levels that fit b200_ao_wide_top (set before every AO launch)
    void set_w_top(uint32_t cap) { w_top = 0; for (uint32_t e : w_level_end) { if (e <= cap && e <= n_wnodes) w_top = e; else break; } }
    // triangle-tube mode of the AO passes (ensure_tube_mesh): the reference's tube mesh + a BVH over its triangles
    DevBuf<TriRec> tris; DevBuf<uint32_t> tri_ids; DevBuf<Node64> tri_nodes; DevBuf<float4> tri_vattr, tri_line_pos, tri_line_tan;
    uint64_t n_tri = 0; uint32_t mesh_subdiv = 0; float tri_build_ms = 0.0f;
    uint32_t n_param = 0, param_subdiv = 0, bake_done = 0;
    uint64_t bake_first = 0, bake_count = 0;       // slice of parametrization vertices baked here (count 0 = all), lv_ao_set_vertex_range
    float param_len = 0.0f;
    bool has_lines = false;
    SceneDev dev() const {
        SceneDev s; s.segs = segs.p; s.prim_ids = prim_ids.p; s.seg_axes = seg_axes.p; s.nodes = nodes.p; s.n_seg = uint32_t(n_seg);
        s.seg_aux = has_lines ? seg_aux.p : nullptr;
        s.qnodes = qnodes.p;
        for (int k = 0; k < 3; k++) { s.q_origin[k] = q_origin[k]; s.q_scale[k] = q_scale[k]; s.w_origin[k] = w_origin[k]; s.w_scale[k] = w_scale[k]; }
        s.w_pk[0] = w_scale[0]; s.w_pk[1] = w_scale[1]; s.w_pk[2] = w_origin[0]; s.w_pk[3] = w_origin[1];
        s.w_pk[4] = s.w_pk[5] = w_scale[2]; s.w_pk[6] = s.w_pk[7] = w_origin[2];
        s.w_tq_bits = std::max<uint64_t>(n_wnodes, n_seg + 1) < (1ull << 24) ? 7u : 4u;
        s.wnodes = n_wnodes ? wnodes.p : nullptr; s.w_top = w_top;
        s.tris = tris.p; s.tri_ids = tri_ids.p; s.tri_nodes = tri_nodes.p; s.tri_vattr = tri_vattr.p;
        s.tri_line_pos = tri_line_pos.p; s.tri_line_tan = tri_line_tan.p; s.n_tri = uint32_t(n_tri);
        s.n_nodes = uint32_t(n_nodes); s.radius = line_width * 0.5f; s.line_width = line_width;
        return s;
    }
};

namespace {

int fail(lv_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->error = msg; else g_global_error = msg;
    return code;
}
#define LV_CUDA(ctx, expr)                                                                                      \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? LV_ERR_OUT_OF_MEMORY : LV_ERR_CUDA,             \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                                   \
    } while (0)

bool is_device_pointer(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

bool parse_bool(const char* v) { return !strcmp(v, "true") || !strcmp(v, "1"); }  // SettingsMap::getValueOpt(bool&), InternalState.hpp:67-74

uint32_t morton2(uint32_t x, uint32_t y) {
    auto part = [](uint32_t v) { v &= 0xffff; v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555; return v; };
    return part(x) | (part(y) << 1);
}

// tiles of a W x H frame in Morton order; tile i belongs to rank i % world
// tiles of a W x H frame in Morton order
void morton_tiles(uint32_t W, uint32_t H, uint32_t ts, std::vector<uint2>& out) {
    uint32_t tx = (W + ts - 1) / ts, ty = (H + ts - 1) / ts;
    std::vector<std::pair<uint32_t, uint2>> all;
    all.reserve(size_t(tx) * ty);
    for (uint32_t y = 0; y < ty; y++) for (uint32_t x = 0; x < tx; x++) all.push_back({morton2(x, y), make_uint2(x, y)});
    std::sort(all.begin(), all.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    out.clear();
    for (const auto& a : all) out.push_back(a.second);
}
// the tiles `rank` owns: tile i of the Morton order belongs to rank i % world, or to owners[i] when an explicit map is given
void enumerate_tiles(uint32_t W, uint32_t H, uint32_t ts, uint32_t rank, uint32_t world, std::vector<uint2>& out, const std::vector<unsigned char>* owners = nullptr) {
    std::vector<uint2> all;
    morton_tiles(W, H, ts, all);
    out.clear();
    const bool use = owners && owners->size() == all.size();
    for (size_t i = 0; i < all.size(); i++) if ((use ? uint32_t((*owners)[i]) : uint32_t(i % world)) == rank) out.push_back(all[i]);
}
const std::vector<unsigned char>* tile_owners_for(const lv_ctx* c, uint32_t W, uint32_t H) {
    return (!c->tile_owner.empty() && c->owner_w == W && c->owner_h == H) ? &c->tile_owner : nullptr;
}

int ensure_tiles(lv_ctx* c, uint32_t W, uint32_t H) {
    if (c->tiles_w == W && c->tiles_h == H && c->tiles_dev.p) return LV_OK;
    enumerate_tiles(W, H, c->tile_size, c->rank, c->world, c->tiles_host, tile_owners_for(c, W, H));
    LV_CUDA(c, c->tiles_dev.ensure(std::max<size_t>(1, c->tiles_host.size())));
    if (!c->tiles_host.empty())
        LV_CUDA(c, cudaMemcpyAsync(c->tiles_dev.p, c->tiles_host.data(), c->tiles_host.size() * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
    c->tiles_x = (W + c->tile_size - 1) / c->tile_size;
    if (c->world > 1) {
        const uint32_t tiles_y = (H + c->tile_size - 1) / c->tile_size;
        std::vector<unsigned char> map(size_t(c->tiles_x) * tiles_y, 0);
        for (const uint2& t : c->tiles_host) map[size_t(t.y) * c->tiles_x + t.x] = 1;
        LV_CUDA(c, c->owned_map.ensure(map.size()));
        LV_CUDA(c, cudaMemcpyAsync(c->owned_map.p, map.data(), map.size(), cudaMemcpyHostToDevice, c->stream));
    }
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    c->tiles_w = W; c->tiles_h = H;
    return LV_OK;
}

void padded_size(const lv_ctx* c, uint32_t W, uint32_t H, uint32_t& pw, uint32_t& ph) {
    // LineRenderer::getScreenSizeWithTiling (reference src/Renderers/LineRenderer.cpp:805-812)
    pw = W; ph = H;
    if (pw % c->opt.tiling_w) pw = (pw / c->opt.tiling_w + 1) * c->opt.tiling_w;
    if (ph % c->opt.tiling_h) ph = (ph / c->opt.tiling_h + 1) * c->opt.tiling_h;
}

int make_params(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, uint32_t frame_number, FrameParams& P) {
    if (!cam || cam->width == 0 || cam->height == 0) return fail(c, LV_ERR_INVALID_ARGUMENT, "camera: width/height must be > 0");
    if (uint64_t(cam->width) * cam->height > 0xFFFFFFFFull) return fail(c, LV_ERR_INVALID_ARGUMENT, "frame too large");
    // NaN / inf camera matrices would make NaN rays, and NaN passes every slab test of the traversals (also an absent child's): refuse them
    for (int k = 0; k < 16; k++)
        if (!std::isfinite(cam->view[k]) || !std::isfinite(cam->proj[k]) || !std::isfinite(cam->inv_view[k]) || !std::isfinite(cam->inv_proj[k]))
            return fail(c, LV_ERR_INVALID_ARGUMENT, "camera: matrices must be finite");
    int rc = ensure_tiles(c, cam->width, cam->height);
    if (rc) return rc;
    memset(&P, 0, sizeof(P));
    memcpy(P.view, cam->view, 64); memcpy(P.proj, cam->proj, 64);
    memcpy(P.inv_view, cam->inv_view, 64); memcpy(P.inv_proj, cam->inv_proj, 64);
    memcpy(P.cam_pos, cam->position, 12);
    P.fov_y = cam->fov_y;
    for (int k = 0; k < 4; k++) { P.bg[k] = cam->background[k]; P.fg[k] = 1.0f - cam->background[k]; }
    P.W = cam->width; P.H = cam->height;
    P.line_width = sc ? sc->line_width : c->opt.line_width;
    const Options& o = c->opt;
    P.use_capped = o.use_capped_tubes; P.use_halos = o.use_halos; P.use_ao = 0;
    P.use_depth_cues = 0; P.depth_cue_strength = o.depth_cue_strength; P.depth_min_max = nullptr;
    P.near_dist = cam->near_dist; P.far_dist = cam->far_dist;
    P.ao_strength = o.ao_strength; P.ao_gamma = o.ao_gamma; P.ao_radius = o.ao_radius;
    P.ao_spp = o.ao_spp; P.ao_use_distance = o.ao_use_distance; P.ao_jitter = o.ao_jitter_primary;
    P.ao_spp_local = o.ao_spp; P.ao_sample_first = 0;
    P.subdiv_corr = float(std::cos(3.14159265358979323846 / double(o.tube_num_subdivisions)));
    P.ao_refill_below = int(o.ao_refill_below ? o.ao_refill_below : (o.ao_raybuf ? (o.ao_packed && o.ao_wide ? 28u : 30u) : 24u));
    P.ao_leaf_vote = int(o.ao_leaf_vote);
    P.ao_wide_reps = int(o.ao_wide_reps);
    P.ao_tq_lo = (1.0f / P.ao_radius) * (1.0f - 1.0f / 4096.0f); P.ao_tq_hi = (1.0f / P.ao_radius) * (1.0f + 1.0f / 4096.0f);
    P.spp = o.num_samples_per_frame;
    // useJitteredSamples = maxNumFrames > 1 || numSamplesPerFrame > 1 (reference VulkanRayTracer.cpp:421)
    P.use_jitter = (o.num_accumulated_frames > 1 || o.num_samples_per_frame > 1) ? 1 : 0;
    P.det_sampling = o.use_deterministic_sampling;
    P.max_depth = o.max_depth_complexity;
    P.frame_number = frame_number;
    P.tf = c->tf.p; P.tfK = c->tfK; P.amin = c->amin; P.amax = c->amax;
    P.ao_tex = nullptr;
    P.use_static_ao = 0; P.sao_factors = nullptr; P.sao_weights = nullptr;
    P.tiles = c->tiles_dev.p; P.n_tiles = uint32_t(c->tiles_host.size()); P.tile_size = c->tile_size;
    padded_size(c, P.W, P.H, P.padded_w, P.padded_h);
    P.addr_tw = o.tiling_w; P.addr_th = o.tiling_h;
    for (P.addr_tw_log2 = 0; (1u << P.addr_tw_log2) < P.addr_tw; P.addr_tw_log2++) {}
    for (P.addr_th_log2 = 0; (1u << P.addr_th_log2) < P.addr_th; P.addr_th_log2++) {}
    P.raster_slack = o.ppll_raster_slack;
    return LV_OK;
}

uint32_t pixel_grid(const lv_ctx* c, const FrameParams& P) {
    return P.n_tiles * (c->tile_size / 16) * (c->tile_size / 8);
}

int reset_counters(lv_ctx* c) {
    LV_CUDA(c, c->counters.ensure(1));
    LV_CUDA(c, cudaMemsetAsync(c->counters.p, 0, sizeof(Counters), c->stream));
    return LV_OK;
}

int read_counters(lv_ctx* c, Counters& h) {
    LV_CUDA(c, cudaMemcpyAsync(&h, c->counters.p, sizeof(Counters), cudaMemcpyDeviceToHost, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    return LV_OK;
}

void fill_stats(lv_stats* s, const Counters& h) {
    s->rays_primary += h.rays_primary; s->rays_ao += h.rays_ao;
    s->traversal_steps += h.steps + h.ao_steps; s->intersections += h.isect + h.ao_isect;
    s->ao_traversal_steps += h.ao_steps; s->ao_intersections += h.ao_isect;
    s->pixels_hit += h.pixels_hit ? h.pixels_hit : h.ao_pixels_hit;
    s->frags_generated += h.frags_generated; s->frags_sorted += h.frags_sorted; s->frags_truncated += h.frags_truncated;
    s->max_depth_complexity = std::max(s->max_depth_complexity, h.max_depth_complexity);
}

// copy a device image (W*H elements of `elem` bytes) to the caller's buffer (device or host); sharded + host -> owned tiles only
int deliver(lv_ctx* c, const void* src, void* dst, uint32_t W, uint32_t H, size_t elem) {
    if (src == dst) return LV_OK;
    if (c->world == 1 || is_device_pointer(dst)) {
        if (c->world == 1) {
            LV_CUDA(c, cudaMemcpyAsync(dst, src, size_t(W) * H * elem, cudaMemcpyDefault, c->stream));
        } else {
            for (const uint2& t : c->tiles_host) {
                uint32_t x0 = t.x * c->tile_size, y0 = t.y * c->tile_size;
                uint32_t w = std::min(c->tile_size, W - x0), h = std::min(c->tile_size, H - y0);
                size_t off = (size_t(y0) * W + x0) * elem;
                LV_CUDA(c, cudaMemcpy2DAsync((char*)dst + off, size_t(W) * elem, (const char*)src + off, size_t(W) * elem, w * elem, h, cudaMemcpyDefault, c->stream));
            }
        }
    } else {
        for (const uint2& t : c->tiles_host) {
            uint32_t x0 = t.x * c->tile_size, y0 = t.y * c->tile_size;
            uint32_t w = std::min(c->tile_size, W - x0), h = std::min(c->tile_size, H - y0);
            size_t off = (size_t(y0) * W + x0) * elem;
            LV_CUDA(c, cudaMemcpy2DAsync((char*)dst + off, size_t(W) * elem, (const char*)src + off, size_t(W) * elem, w * elem, h, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    if (!is_device_pointer(dst)) LV_CUDA(c, cudaStreamSynchronize(c->stream));
    return LV_OK;
}

// Where a frame kernel puts its pixels (b200_frame_format), and how they reach the caller afterwards.
//   rgba32f: `image` = the caller's device frame, or the library's frame when the caller's pointer is host memory (then copied);
//   rgba8:   `image` = the library's float accumulation image (running mean over frames), `out8` = the caller's device / peer frame, or
//            a staging frame when the caller's pointer is host memory (then copied: synchronously, or -- b200_async_delivery -- on the
//            copy stream from one of two alternating staging frames, so that the copy overlaps the next frame's kernels).
struct FrameSink { float4* image = nullptr; uint32_t* out8 = nullptr; bool host = false; bool async = false; unsigned slot = 0; };

int open_sink(lv_ctx* c, void* rgba_out, uint32_t W, uint32_t H, FrameSink& s) {
    const size_t npx = size_t(W) * H;
    s = FrameSink();
    s.host = !is_device_pointer(rgba_out);
    if (!c->opt.frame_rgba8) {
        if (s.host) { LV_CUDA(c, c->image.ensure(npx)); s.image = c->image.p; }
        else if (reinterpret_cast<uintptr_t>(rgba_out) & 15) return fail(c, LV_ERR_INVALID_ARGUMENT, "device rgba_out must be 16-byte aligned");
        else s.image = reinterpret_cast<float4*>(rgba_out);
        return LV_OK;
    }
    LV_CUDA(c, c->image.ensure(npx));
    s.image = c->image.p;
    if (!s.host) {
        if (reinterpret_cast<uintptr_t>(rgba_out) & 3) return fail(c, LV_ERR_INVALID_ARGUMENT, "device rgba_out must be 4-byte aligned");
        s.out8 = reinterpret_cast<uint32_t*>(rgba_out);
        return LV_OK;
    }
    s.async = c->opt.async_delivery;
    if (!s.async) { LV_CUDA(c, c->rgba8.ensure(npx)); s.out8 = c->rgba8.p; return LV_OK; }
    if (!c->copy_stream) {
        LV_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) { LV_CUDA(c, cudaEventCreateWithFlags(&c->ev_rendered[k], cudaEventDisableTiming)); LV_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming)); }
    }
    s.slot = c->flip & 1u; c->flip++;
    const uint32_t* before = c->stage8[s.slot].p;
    if (c->stage8[s.slot].n < npx) { LV_CUDA(c, cudaStreamSynchronize(c->copy_stream)); LV_CUDA(c, c->stage8[s.slot].ensure(npx)); }
    if (before && before == c->stage8[s.slot].p) LV_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[s.slot], 0));   // the copy of two frames ago has left this buffer
    s.out8 = c->stage8[s.slot].p;
    return LV_OK;
}

int close_sink(lv_ctx* c, const FrameSink& s, void* rgba_out, uint32_t W, uint32_t H) {
    if (!s.host) return LV_OK;
    if (!c->opt.frame_rgba8) return deliver(c, s.image, rgba_out, W, H, 16);
    if (!s.async) return deliver(c, s.out8, rgba_out, W, H, 4);
    LV_CUDA(c, cudaEventRecord(c->ev_rendered[s.slot], c->stream));
    LV_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_rendered[s.slot], 0));
    if (c->world == 1) LV_CUDA(c, cudaMemcpyAsync(rgba_out, s.out8, size_t(W) * H * 4, cudaMemcpyDeviceToHost, c->copy_stream));
    else
        for (const uint2& t : c->tiles_host) {
            const uint32_t x0 = t.x * c->tile_size, y0 = t.y * c->tile_size, w = std::min(c->tile_size, W - x0), h = std::min(c->tile_size, H - y0);
            const size_t off = (size_t(y0) * W + x0) * 4;
            LV_CUDA(c, cudaMemcpy2DAsync((char*)rgba_out + off, size_t(W) * 4, (const char*)s.out8 + off, size_t(W) * 4, w * 4, h, cudaMemcpyDeviceToHost, c->copy_stream));
        }
    LV_CUDA(c, cudaEventRecord(c->ev_copied[s.slot], c->copy_stream));
    c->copies_pending = true;
    return LV_OK;
}

bool cam_ok(const FrameParams& P) { return P.far_dist > P.near_dist; }

float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms = 0.0f; cudaEventElapsedTime(&ms, a, b); return ms; }

// b200_ao_qnodes: the 16-bit quantised copy of the segment BVH's nodes, built on first use
int ensure_qnodes(lv_ctx* c, lv_scene* sc) {
    if (sc->qnodes.p || sc->n_nodes == 0) return LV_OK;
    for (int k = 0; k < 3; k++) {
        const float ext = sc->bounds[3 + k] - sc->bounds[k];
        sc->q_origin[k] = sc->bounds[k];
        sc->q_scale[k] = std::max(ext / 65535.0f * 1.0001f, 1e-30f);   // a little coarser than the exact grid: 65535 steps always reach past the upper bound
    }
    LV_CUDA(c, sc->qnodes.ensure(sc->n_nodes));
    k_quantize_nodes<<<uint32_t((sc->n_nodes + 255) / 256), 256, 0, c->stream>>>(sc->nodes.p, uint32_t(sc->n_nodes), sc->q_origin[0], sc->q_origin[1], sc->q_origin[2],
                                                                              sc->q_scale[0], sc->q_scale[1], sc->q_scale[2], sc->qnodes.p);
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// b200_ao_wide: the 4-wide quantised tree, collapsed from the child-pair nodes on first use (breadth-first rounds of k_w4_round).
// Leaves `n_wnodes` at 0 (the stream then stays on the child-pair nodes) when the scene has multi-record leaves or the tree would need
// more traversal stack than the kernel has.
int ensure_wnodes(lv_ctx* c, lv_scene* sc) {
    if (sc->n_wnodes || sc->w_failed || sc->n_nodes == 0) return LV_OK;
    if (sc->leaf_size != 1) { sc->w_failed = true; return LV_OK; }
    W4Grid G;
    for (int k = 0; k < 3; k++) {
        const float ext = sc->bounds[3 + k] - sc->bounds[k];
        float s = std::max(ext / 65535.0f * 1.0001f, 1e-30f), o = sc->bounds[k], om = 0.0f;
        auto dq = [&](uint32_t q) { uint32_t bits = 0x4B000000u | q; float x; memcpy(&x, &bits, 4); return std::fmaf(x, s, om); };   // = w4_dequant
        for (int it = 0; it < 64; it++) {   // the shifted origin is rounded at magnitude 2^23 * s: make sure the grid still spans the bounds
            om = std::fmaf(-8388608.0f, s, o);
            if (dq(0) > sc->bounds[k]) { o -= s; continue; }
            if (dq(65535) < sc->bounds[3 + k]) { s *= 1.001f; continue; }
            break;
        }
        if (dq(0) > sc->bounds[k] || dq(65535) < sc->bounds[3 + k]) { sc->w_failed = true; return LV_OK; }
        G.o[k] = o; G.s[k] = s; G.om[k] = om;
        sc->w_origin[k] = om; sc->w_scale[k] = s;
    }
    DevBuf<uint3> q0, q1; DevBuf<unsigned int> ctr;   // ctr: [0] next queue count, [1] node allocator, [2] stack need
    auto cleanup = [&]() { q0.release(); q1.release(); ctr.release(); };
#define LV_W4(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); sc->wnodes.release(); return fail(c, e_ == cudaErrorMemoryAllocation ? LV_ERR_OUT_OF_MEMORY : LV_ERR_CUDA, std::string("wide BVH: ") + cudaGetErrorString(e_)); } } while (0)
    const size_t cap = size_t(sc->n_nodes);          // every wide node is rooted at a distinct inner binary node
    LV_W4(sc->wnodes.ensure(cap)); LV_W4(q0.ensure(cap)); LV_W4(q1.ensure(cap)); LV_W4(ctr.ensure(4));
    cudaStream_t st = c->stream;
    LV_W4(cudaEventRecord(c->ev[6], st));
    const uint3 root = make_uint3(0u, 0u, 0u);
    const unsigned int init[4] = {0u, 1u, 0u, 0u};
    LV_W4(cudaMemcpyAsync(q0.p, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    LV_W4(cudaMemcpyAsync(ctr.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    uint32_t n_in = 1, total = 1;
    sc->w_level_end.clear();
    uint3 *in = q0.p, *out = q1.p;
    while (n_in) {
        sc->w_level_end.push_back(total);
        LV_W4(cudaMemsetAsync(ctr.p, 0, 4, st));
        k_w4_round<<<(n_in + 127) / 128, 128, 0, st>>>(sc->nodes.p, in, n_in, out, ctr.p, ctr.p + 1, sc->wnodes.p, ctr.p + 2, G, uint32_t(sc->n_seg));
        unsigned int h[3];
        LV_W4(cudaMemcpyAsync(h, ctr.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        LV_W4(cudaStreamSynchronize(st));
        n_in = h[0]; total = h[1]; sc->w_need = h[2];
        std::swap(in, out);
    }
    LV_W4(cudaEventRecord(c->ev[7], st));
    LV_W4(cudaStreamSynchronize(st));
    sc->w_build_ms = elapsed(c->ev[6], c->ev[7]);
    cleanup();
#undef LV_W4
    if (sc->w_need + 1 > uint32_t(kAoStackWide)) { sc->wnodes.release(); sc->w_failed = true; return LV_OK; }   // deeper than the stream's stack: stay on the child-pair nodes
    sc->n_wnodes = total;
    return LV_OK;
}

// Triangle-tube mode of the AO passes: generate the reference's tube mesh on the host (lv_tubemesh.hpp, like the reference's
// createCappedTriangleTubesRenderDataCPU), upload it and build an LBVH over its triangles (lv_tri.cuh), one triangle per leaf.
// Needs the polylines of lv_scene_set_lines.  Rebuilt when tube_num_subdivisions changes.
int ensure_tube_mesh(lv_ctx* c, lv_scene* sc) {
    if (!sc->has_lines) return fail(c, LV_ERR_STATE, "b200_rtao_geometry = triangles needs the polylines (lv_scene_set_lines)");
    if (sc->n_tri && sc->mesh_subdiv == c->opt.tube_num_subdivisions) return LV_OK;
    lvmesh::TubeMesh m;
    lvmesh::build(sc->host_pos.data(), sc->line_offsets.data(), sc->line_offsets.size() - 1, sc->line_width * 0.5f, int(c->opt.tube_num_subdivisions), m);
    const size_t nv = m.vertices.size(), nt = m.indices.size() / 3, nl = m.line_pos.size();
    if (nt == 0) return fail(c, LV_ERR_STATE, "triangle-tube mode: the polylines produce no tube geometry");
    if (nt >= (1ull << 27)) return fail(c, LV_ERR_INVALID_ARGUMENT, "triangle-tube mode: too many triangles (max 2^27-1)");
    std::vector<float> vpos(3 * nv), vnrm(3 * nv); std::vector<uint32_t> vline(nv); std::vector<float4> lpos(nl), ltan(nl);
    for (size_t i = 0; i < nv; i++) {
        vpos[3 * i] = m.vertices[i].position.x; vpos[3 * i + 1] = m.vertices[i].position.y; vpos[3 * i + 2] = m.vertices[i].position.z;
        vnrm[3 * i] = m.vertices[i].normal.x; vnrm[3 * i + 1] = m.vertices[i].normal.y; vnrm[3 * i + 2] = m.vertices[i].normal.z;
        vline[i] = m.vertices[i].line_point;
    }
    for (size_t i = 0; i < nl; i++) { lpos[i] = make_float4(m.line_pos[i].x, m.line_pos[i].y, m.line_pos[i].z, 0.0f); ltan[i] = make_float4(m.line_tan[i].x, m.line_tan[i].y, m.line_tan[i].z, 0.0f); }
    for (uint32_t src : m.line_src) if (src >= sc->n_pt) return fail(c, LV_ERR_STATE, "triangle-tube mode: the polylines of lv_scene_set_lines do not match the scene's points");
    cudaStream_t st = c->stream;
    const int n = int(nt);
    DevBuf<float> d_vpos, d_vnrm, bounds, boxes; DevBuf<uint32_t> d_idx, d_vline, vals; DevBuf<unsigned long long> keys, keys2;
    DevBuf<int2> children, ranges; DevBuf<int> parent; DevBuf<unsigned int> flags; DevBuf<char> cubtmp;
    auto cleanup = [&]() { d_vpos.release(); d_vnrm.release(); bounds.release(); boxes.release(); d_idx.release(); d_vline.release(); vals.release(); keys.release(); keys2.release();
                           children.release(); ranges.release(); parent.release(); flags.release(); cubtmp.release(); };
#define LV_TRI(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); sc->n_tri = 0; return fail(c, e__ == cudaErrorMemoryAllocation ? LV_ERR_OUT_OF_MEMORY : LV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    LV_TRI(cudaEventRecord(c->ev[6], st));
    LV_TRI(d_vpos.ensure(3 * nv)); LV_TRI(d_vnrm.ensure(3 * nv)); LV_TRI(d_vline.ensure(nv)); LV_TRI(d_idx.ensure(3 * nt));
    LV_TRI(cudaMemcpyAsync(d_vpos.p, vpos.data(), 12 * nv, cudaMemcpyHostToDevice, st));
    LV_TRI(cudaMemcpyAsync(d_vnrm.p, vnrm.data(), 12 * nv, cudaMemcpyHostToDevice, st));
    LV_TRI(cudaMemcpyAsync(d_vline.p, vline.data(), 4 * nv, cudaMemcpyHostToDevice, st));
    LV_TRI(cudaMemcpyAsync(d_idx.p, m.indices.data(), 12 * nt, cudaMemcpyHostToDevice, st));
    LV_TRI(sc->tri_vattr.ensure(nv)); LV_TRI(sc->tri_line_pos.ensure(nl)); LV_TRI(sc->tri_line_tan.ensure(nl));
    LV_TRI(cudaMemcpyAsync(sc->tri_line_pos.p, lpos.data(), 16 * nl, cudaMemcpyHostToDevice, st));
    LV_TRI(cudaMemcpyAsync(sc->tri_line_tan.p, ltan.data(), 16 * nl, cudaMemcpyHostToDevice, st));
    k_tri_vertex_attr<<<uint32_t(std::min<size_t>((nv + 255) / 256, 65535)), 256, 0, st>>>(d_vnrm.p, d_vline.p, uint32_t(nv), sc->tri_vattr.p);
    {   // lineAttribute of the mesh's line points -> tri_line_tan.w (the triangle geometry mode of the tube pass interpolates it)
        DevBuf<float> pt_attr; DevBuf<uint32_t> d_src;
        cudaError_t e = pt_attr.ensure(sc->n_pt);
        if (e == cudaSuccess) e = d_src.ensure(nl);
        if (e == cudaSuccess) e = cudaMemsetAsync(pt_attr.p, 0, sc->n_pt * 4, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_src.p, m.line_src.data(), 4 * nl, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            k_point_attr<<<uint32_t(std::min<size_t>((sc->n_seg + 255) / 256, 65535)), 256, 0, st>>>(sc->segs.p, sc->prim_ids.p, sc->seg_idx.p, uint32_t(sc->n_seg), pt_attr.p);
            k_tri_line_attr<<<uint32_t(std::min<size_t>((nl + 255) / 256, 65535)), 256, 0, st>>>(pt_attr.p, d_src.p, uint32_t(nl), sc->tri_line_tan.p);
            e = cudaStreamSynchronize(st);
        }
        pt_attr.release(); d_src.release();
        LV_TRI(e);
    }
    LV_TRI(bounds.ensure(6)); LV_TRI(keys.ensure(n)); LV_TRI(keys2.ensure(n)); LV_TRI(vals.ensure(n));
    LV_TRI(sc->tri_ids.ensure(size_t(n) + 1)); LV_TRI(sc->tris.ensure(size_t(n) + 1));   // + the dummy record of an absent child (k_emit_nodes)
    LV_TRI(cudaMemsetAsync(sc->tris.p + n, 0xFF, sizeof(TriRec), st));
    LV_TRI(cudaMemsetAsync(sc->tri_ids.p + n, 0xFF, 4, st));
    k_init_bounds<<<1, 32, 0, st>>>(bounds.p);
    k_tri_bounds<<<std::min(1024, (n + 255) / 256), 256, 0, st>>>(d_vpos.p, d_idx.p, uint32_t(n), bounds.p);
    k_tri_morton<<<(n + 255) / 256, 256, 0, st>>>(d_vpos.p, d_idx.p, uint32_t(n), bounds.p, keys.p, vals.p);
    size_t cub_bytes = 0;
    LV_TRI(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys.p, keys2.p, vals.p, sc->tri_ids.p, n, 0, 63, st));
    LV_TRI(cubtmp.ensure(cub_bytes + 16));
    LV_TRI(cub::DeviceRadixSort::SortPairs(cubtmp.p, cub_bytes, keys.p, keys2.p, vals.p, sc->tri_ids.p, n, 0, 63, st));
    k_pack_tris<<<(n + 255) / 256, 256, 0, st>>>(d_vpos.p, d_idx.p, sc->tri_ids.p, uint32_t(n), sc->tris.p);
    const int n_inner = std::max(1, n - 1);
    LV_TRI(children.ensure(n_inner)); LV_TRI(ranges.ensure(n_inner)); LV_TRI(parent.ensure(2 * size_t(n)));
    LV_TRI(boxes.ensure(6 * (2 * size_t(n)))); LV_TRI(flags.ensure(n_inner));
    LV_TRI(cudaMemsetAsync(flags.p, 0, size_t(n_inner) * 4, st));
    if (n > 1) k_radix_tree<<<(n - 1 + 255) / 256, 256, 0, st>>>(keys2.p, n, children.p, ranges.p, parent.p);
    k_tri_fit<<<(n + 255) / 256, 256, 0, st>>>(sc->tris.p, n, children.p, parent.p, boxes.p, flags.p);
    LV_TRI(sc->tri_nodes.ensure(n_inner));
    k_emit_nodes<<<(n_inner + 255) / 256, 256, 0, st>>>(n, 1, children.p, ranges.p, boxes.p, sc->tri_nodes.p);
    LV_TRI(cudaMemsetAsync(flags.p, 0, 4, st));
    k_tree_depth<<<(n + 255) / 256, 256, 0, st>>>(n, parent.p, flags.p);
    LV_TRI(cudaGetLastError());
    LV_TRI(cudaEventRecord(c->ev[7], st));
    uint32_t depth = 0;
    LV_TRI(cudaMemcpyAsync(&depth, flags.p, 4, cudaMemcpyDeviceToHost, st));
    LV_TRI(cudaStreamSynchronize(st));
    cleanup();
#undef LV_TRI
    if (depth + 1 > uint32_t(kStackSize) || depth + 1 > uint32_t(kAoStack)) {
        sc->n_tri = 0;
        return fail(c, LV_ERR_STATE, "triangle BVH depth " + std::to_string(depth) + " exceeds the traversal stack (" + std::to_string(kAoStack) + ")");
    }
    sc->tri_build_ms = elapsed(c->ev[6], c->ev[7]);
    sc->n_tri = nt; sc->mesh_subdiv = c->opt.tube_num_subdivisions;
    return LV_OK;
}

// persistent AO ray-stream kernel over the records in ctx->ao_hits (count in small[0], work counter in small[2..3]) into ctx->occ;
// timed with ev[4] / ev[5].  BAKE selects the prebaker's random stream / ray origin (lv_bake.cuh).
template <bool BAKE>
int launch_ao_rays(lv_ctx* c, const FrameParams& P, const SceneDev& S, bool one_record_leaves, unsigned long long max_rays, bool tri = false,
                   const AoHit* ext_hits = nullptr, unsigned int* ext_small = nullptr, float* ext_occ = nullptr) {
    const AoHit* hits = ext_hits ? ext_hits : c->ao_hits.p;
    unsigned int* small = ext_small ? ext_small : c->small.p;      // [0] record count, [2..3] the stream's work counter
    float* occ = ext_occ ? ext_occ : c->occ.p;
    if (P.ao_spp_local != P.ao_spp && !(c->opt.ao_queue && one_record_leaves && c->opt.ao_raybuf && c->opt.ao_packed && c->opt.ao_stack == 12 &&
                                        c->opt.ao_min_blocks == 0 && !c->opt.ao_qnodes && c->opt.ao_wide && S.wnodes && !S.w_top && !tri && max_rays < 0xFF000000ull))
        return fail(c, LV_ERR_STATE, "AO-sample-batch shards need the default AO ray stream (k_rtao_rays_w)");
    auto launch = [&](auto kern) -> int {
        int per_sm = 0;
        LV_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlockThreads, 0));
        const uint32_t pgrid = uint32_t(std::max(1, per_sm) * c->num_sms);
        LV_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
        kern<<<pgrid, kBlockThreads, 0, c->stream>>>(P, S, occ, hits, small, reinterpret_cast<unsigned long long*>(small + 2), c->counters.p);
        LV_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
        return LV_OK;
    };
    if (tri)   // triangle-tube mode: the leaf-queue kernel over the triangle BVH (always one triangle per leaf)
        return c->opt.ao_min_blocks >= 9 ? launch(k_rtao_rays_q<9, BAKE, 12, 1>) : launch(k_rtao_rays_q<8, BAKE, 12, 1>);
    const uint32_t stack = c->opt.ao_stack;
    const bool queue = c->opt.ao_queue && one_record_leaves && stack != 0;   // the leaf-queue kernel needs one-record leaves and a packed stack
    const bool default_tuning = stack == 12 && (c->opt.ao_min_blocks == 0 || c->opt.ao_min_blocks == 8) && !c->opt.ao_qnodes;
    // the packed-arithmetic stream (32-bit ray numbers: `max_rays` bounds records x samples)
    if (queue && c->opt.ao_raybuf && c->opt.ao_packed && default_tuning && c->opt.ao_wide && S.wnodes && !S.w_top && max_rays < 0xFF000000ull)
        {
        const bool seven = S.w_tq_bits == 7u && c->opt.ao_tq_bits != 4u;
        return seven ? launch(k_rtao_rays_w<8, BAKE, 7>) : launch(k_rtao_rays_w<8, BAKE, 4>);
    }
    if (queue && c->opt.ao_raybuf && default_tuning && !(c->opt.ao_wide && S.wnodes && S.w_top))   // warp-wide ray generation into a shared batch (default tuning only)
        return (c->opt.ao_wide && S.wnodes) ? launch(k_rtao_rays_q<8, BAKE, 12, 0, 2, 0, true>) : launch(k_rtao_rays_q<8, BAKE, 12, 0, 0, 0, true>);
    if (queue && c->opt.ao_wide && S.wnodes && stack == 12 && !c->opt.ao_qnodes) {   // 4-wide quantised tree; S.w_top: how many top-level nodes the kernel stages into shared memory
        if (S.w_top > 85) return launch(k_rtao_rays_q<8, BAKE, 12, 0, 2, 341>);
        if (S.w_top > 0) return c->opt.ao_min_blocks == 7 ? launch(k_rtao_rays_q<7, BAKE, 12, 0, 2, 85>) : launch(k_rtao_rays_q<8, BAKE, 12, 0, 2, 85>);
        return c->opt.ao_min_blocks == 9 ? launch(k_rtao_rays_q<9, BAKE, 12, 0, 2>) : c->opt.ao_min_blocks == 7 ? launch(k_rtao_rays_q<7, BAKE, 12, 0, 2>)
                                                                                                                : launch(k_rtao_rays_q<8, BAKE, 12, 0, 2>);
    }
    if (queue && c->opt.ao_qnodes && S.qnodes)   // experimental: quantised nodes (default register budget / stack only)
        return launch(k_rtao_rays_q<8, BAKE, 12, 0, 1>);
    const uint32_t mb = c->opt.ao_min_blocks ? c->opt.ao_min_blocks : (queue ? 8u : 9u);   // 0 = measured optimum of the variant
    if (queue) {
        if (mb >= 9) return stack == 8 ? launch(k_rtao_rays_q<9, BAKE, 8>) : stack == 1 ? launch(k_rtao_rays_q<9, BAKE, 1>) : launch(k_rtao_rays_q<9, BAKE, 12>);
        return stack == 8 ? launch(k_rtao_rays_q<8, BAKE, 8>) : stack == 1 ? launch(k_rtao_rays_q<8, BAKE, 1>) : launch(k_rtao_rays_q<8, BAKE, 12>);
    }
#define LV_AO_STACKS(MB)                                                  \
    switch (stack) {                                                      \
        case 0: return launch(k_rtao_rays<MB, BAKE, 0>);                  \
        case 1: return launch(k_rtao_rays<MB, BAKE, 1>);                  \
        case 8: return launch(k_rtao_rays<MB, BAKE, 8>);                  \
        case 16: return launch(k_rtao_rays<MB, BAKE, 16>);                \
        default: return launch(k_rtao_rays<MB, BAKE, 12>);                \
    }
    if (mb >= 10) { LV_AO_STACKS(10) }
    if (mb >= 9) { LV_AO_STACKS(9) }
    LV_AO_STACKS(8)
#undef LV_AO_STACKS
}

// ---- RTAO pass (S5) into ctx->ao --------------------------------------------------------------
// stages: kRtaoAll = the whole pass; kRtaoPrimaryOnly = up to the hit list (AO-sample-batch shards trace and reduce separately)
enum { kRtaoAll = 0, kRtaoPrimaryOnly = 1 };
int run_rtao(lv_ctx* c, const lv_scene* sc, FrameParams P, uint32_t frame_number, int stage = kRtaoAll) {
    const size_t npx = size_t(P.W) * P.H;
    c->rtao_rays_timed = false;
    if (c->ao_w != P.W || c->ao_h != P.H || !c->ao.p) {
        LV_CUDA(c, c->ao.ensure(npx));
        // untouched (not owned) texels must be finite for the bilinear lookup: initialise to "unoccluded"
        k_fill_f32<<<uint32_t(c->num_sms) * 8u, 256, 0, c->stream>>>(c->ao.p, npx, 1.0f);
        LV_CUDA(c, cudaGetLastError());
        c->ao_w = P.W; c->ao_h = P.H;
    }
    LV_CUDA(c, c->ao_hits.ensure(size_t(P.n_tiles) * (size_t(c->tile_size) * c->tile_size + 4 * c->tile_size + 4) + 1));
    LV_CUDA(c, c->small.ensure(4));
    LV_CUDA(c, cudaMemsetAsync(c->small.p, 0, 4 * sizeof(unsigned int), c->stream));
    P.frame_number = frame_number;
    const bool tri = c->opt.ao_triangles;
    if (tri) { int trc = ensure_tube_mesh(c, const_cast<lv_scene*>(sc)); if (trc) return trc; }
    else if (c->opt.ao_wide) { int wrc = ensure_wnodes(c, const_cast<lv_scene*>(sc)); if (wrc) return wrc; const_cast<lv_scene*>(sc)->set_w_top(c->opt.ao_wide_top); }
    else if (c->opt.ao_qnodes) { int qrc = ensure_qnodes(c, const_cast<lv_scene*>(sc)); if (qrc) return qrc; }
    const SceneDev S = sc->dev();
    const uint32_t grid = pixel_grid(c, P);
    if (grid == 0) return LV_OK;
    // tile-sharded AND jittered tube rays: the tube pass reads the AO image up to half a pixel away, so the one-pixel ring
    // around the owned tiles is computed too (second launch, ~6 % more pixels at 64x64 tiles)
    const bool apron = c->world > 1 && P.use_jitter;
    unsigned int stamp = 0;
    if (apron) {
        const unsigned int* before = c->apron_marks.p;
        LV_CUDA(c, c->apron_marks.ensure(npx));
        if (c->apron_marks.p != before) c->apron_stamp = 0;   // reallocated (the frame grew): the new words are uninitialised
        if (c->apron_stamp == 0) LV_CUDA(c, cudaMemsetAsync(c->apron_marks.p, 0, c->apron_marks.n * 4, c->stream));
        stamp = ++c->apron_stamp;
        P.apron_marks = c->apron_marks.p;
    }
    const uint32_t ring = 4 * c->tile_size + 4;
    if (tri) {
        k_rtao_primary<1><<<grid, kBlockThreads, 0, c->stream>>>(P, S, c->ao.p, c->ao_hits.p, c->small.p, c->counters.p, nullptr, stamp);
        if (apron)
            k_rtao_primary<1><<<P.n_tiles * ((ring + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, c->stream>>>(
                P, S, c->ao.p, c->ao_hits.p, c->small.p, c->counters.p, c->apron_marks.p, stamp);
    } else {
        k_rtao_primary<0><<<grid, kBlockThreads, 0, c->stream>>>(P, S, c->ao.p, c->ao_hits.p, c->small.p, c->counters.p, nullptr, stamp);
        if (apron)
            k_rtao_primary<0><<<P.n_tiles * ((ring + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, c->stream>>>(
                P, S, c->ao.p, c->ao_hits.p, c->small.p, c->counters.p, c->apron_marks.p, stamp);
    }
    if (stage == kRtaoAll) {
        // one float per (hit pixel, sample); worst case every owned pixel (and ring pixel) is hit
        const size_t max_hits = size_t(P.n_tiles) * (size_t(c->tile_size) * c->tile_size + (apron ? ring : 0));
        LV_CUDA(c, c->occ.ensure(max_hits * P.ao_spp));
        int lrc = launch_ao_rays<false>(c, P, S, sc->leaf_size == 1, (unsigned long long)max_hits * P.ao_spp, tri);
        if (lrc) return lrc;
        c->rtao_rays_timed = true;
        k_rtao_reduce<<<c->num_sms * 4, 256, 0, c->stream>>>(P, c->occ.p, c->ao_hits.p, c->small.p, c->ao.p);
    }
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// LineRenderer::computeDepthRange (reference src/Renderers/LineRenderer.cpp:410-431), once per frame when depth cues are on
int run_depth_range(lv_ctx* c, const lv_scene* sc, FrameParams& P) {
    if (!(c->opt.depth_cue_strength > 0.0f) || sc->n_seg == 0) return LV_OK;
    if (!(cam_ok(P))) return fail(c, LV_ERR_INVALID_ARGUMENT, "depth cues need lv_camera.near_dist < far_dist");
    LV_CUDA(c, c->depth_mm.ensure(2));
    const float init[2] = {P.far_dist, P.near_dist};
    LV_CUDA(c, cudaMemcpyAsync(c->depth_mm.p, init, 8, cudaMemcpyHostToDevice, c->stream));
    k_depth_range<<<c->num_sms * 8, 256, 0, c->stream>>>(P, sc->segs.p, uint32_t(sc->n_seg), c->depth_mm.p);
    LV_CUDA(c, cudaGetLastError());
    P.use_depth_cues = 1; P.depth_min_max = c->depth_mm.p;
    return LV_OK;
}

// ---- object-space AO prebaker (S6) -------------------------------------------------------------
// Arc-length parametrization of the polylines, on the host like the reference
// (AmbientOcclusionComputeRenderPass::generateBlendingWeightParametrization + recomputeStaticParametrization,
// src/Renderers/AmbientOcclusion/VulkanAmbientOcclusionBaker.cpp:513-655): every polyline is cut into
// ceil(length / expected) equal pieces; `weights` maps a line point to its (fractional) parametrization vertex,
// `sampling` maps a parametrization vertex to its (fractional) line point.
float seg_length(const float* pos, uint64_t a, uint64_t b) {
    const float dx = pos[3 * b] - pos[3 * a], dy = pos[3 * b + 1] - pos[3 * a + 1], dz = pos[3 * b + 2] - pos[3 * a + 2];
    return std::sqrt((dx * dx + dy * dy) + dz * dz);
}
void ao_parametrize_host(const float* pos, const uint64_t* offsets, uint64_t n_lines, float expected, std::vector<float>& weights,
                         std::vector<float>& sampling) {
    const float eps = 1e-5f;
    weights.assign(size_t(offsets[n_lines]), 0.0f);
    sampling.clear();
    size_t param_base = 0;
    for (uint64_t li = 0; li < n_lines; li++) {
        const uint64_t first = offsets[li], n = offsets[li + 1] - first;
        if (n == 0) continue;
        float total = 0.0f;
        for (uint64_t i = 1; i < n; i++) total += seg_length(pos, first + i - 1, first + i);
        const uint32_t pieces = std::max(1u, uint32_t(std::ceil(total / expected)));
        const float piece_len = total / float(pieces);
        // line point -> parametrization coordinate
        weights[first] = float(param_base);
        float walked = 0.0f;
        for (uint64_t i = 1; i < n; i++) {
            walked += seg_length(pos, first + i - 1, first + i);
            const float w = walked / piece_len;
            weights[first + i] = float(param_base) + std::min(std::max(w, 0.0f), float(pieces) - eps);
        }
        // parametrization vertex -> line point coordinate
        sampling.push_back(float(uint32_t(first)));
        if (n >= 2) {
            float seg_begin = 0.0f, seg_end = seg_length(pos, first, first + 1);
            uint64_t cur = 1;
            for (uint32_t k = 1; k <= pieces; k++) {
                uint32_t reached = uint32_t(seg_end / piece_len);
                while (k > reached && cur < n - 1) {
                    seg_begin = seg_end;
                    seg_end += seg_length(pos, first + cur, first + cur + 1);
                    reached = uint32_t(seg_end / piece_len);
                    cur++;
                }
                float loc = float(cur - 1) + (float(k) * piece_len - seg_begin) / (seg_end - seg_begin);
                loc = float(uint32_t(first)) + std::min(loc, float(uint32_t(n) - 1u) - eps);
                sampling.push_back(loc);
            }
        } else {
            for (uint32_t k = 1; k <= pieces; k++) sampling.push_back(float(uint32_t(first)));
        }
        param_base += size_t(pieces) + 1;
    }
}

int ensure_parametrization(lv_ctx* c, lv_scene* sc) {
    const Options& o = c->opt;
    if (sc->n_param && sc->param_len == o.bake_param_len && sc->param_subdiv == o.bake_subdiv) return LV_OK;
    std::vector<float> weights, sampling;
    ao_parametrize_host(sc->host_pos.data(), sc->line_offsets.data(), sc->line_offsets.size() - 1, o.bake_param_len, weights, sampling);
    if (sampling.empty()) return fail(c, LV_ERR_STATE, "AO prebaker: the scene has no line points");
    if (uint64_t(sampling.size()) * o.bake_subdiv >= 0xFFFFFFFFull) return fail(c, LV_ERR_INVALID_ARGUMENT, "AO prebaker: too many parametrization vertices x subdivisions");
    LV_CUDA(c, sc->sampling.ensure(sampling.size()));
    LV_CUDA(c, sc->weights.ensure(weights.size()));
    LV_CUDA(c, sc->factors.ensure(sampling.size() * o.bake_subdiv));
    LV_CUDA(c, cudaMemcpyAsync(sc->sampling.p, sampling.data(), sampling.size() * 4, cudaMemcpyHostToDevice, c->stream));
    LV_CUDA(c, cudaMemcpyAsync(sc->weights.p, weights.data(), weights.size() * 4, cudaMemcpyHostToDevice, c->stream));
    LV_CUDA(c, cudaMemsetAsync(sc->factors.p, 0, sampling.size() * o.bake_subdiv * 4, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    sc->n_param = uint32_t(sampling.size()); sc->param_len = o.bake_param_len; sc->param_subdiv = o.bake_subdiv; sc->bake_done = 0;
    return LV_OK;
}

// One dispatch of the baker compute shader with frameNumber = sc->bake_done (VulkanAmbientOcclusionBaker::updateIterative,
// VulkanAmbientOcclusionBaker.cpp:340-351).  Counters accumulate into ctx->counters (rays_ao, ao_steps, ao_isect).
int run_bake_iteration(lv_ctx* c, lv_scene* sc) {
    if (!sc->has_lines) return fail(c, LV_ERR_STATE, "AO prebaker: no line frames attached (lv_scene_set_lines)");
    if (sc->n_seg == 0) return fail(c, LV_ERR_STATE, "AO prebaker: the scene has no segments");
    int rc = ensure_parametrization(c, sc);
    if (rc) return rc;
    const Options& o = c->opt;
    const uint64_t first = std::min<uint64_t>(sc->bake_first, sc->n_param);
    const uint64_t count = sc->bake_count ? std::min<uint64_t>(sc->bake_count, sc->n_param - first) : sc->n_param - first;
    if (count == 0) { sc->bake_done++; return LV_OK; }   // an empty slice (more ranks than vertices) still takes part in the iteration count
    const size_t n_rec = size_t(count) * o.bake_subdiv;
    LV_CUDA(c, c->ao_hits.ensure(n_rec));
    LV_CUDA(c, c->occ.ensure(n_rec * o.bake_spp));
    LV_CUDA(c, c->small.ensure(4));
    const unsigned int init[4] = {unsigned(n_rec), 0u, 0u, 0u};
    LV_CUDA(c, cudaMemcpyAsync(c->small.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    BakeParams B;
    B.pt_pos = sc->pt_pos.p; B.pt_tan = sc->pt_tan.p; B.pt_nrm = sc->pt_nrm.p; B.sampling = sc->sampling.p;
    B.n_line_pts = uint32_t(sc->n_pt); B.n_param = sc->n_param; B.n_subdiv = o.bake_subdiv; B.spp = o.bake_spp;
    B.frame_number = sc->bake_done; B.line_radius = sc->line_width * 0.5f;
    B.first_vertex = uint32_t(first); B.n_vertices = uint32_t(count);
    FrameParams P;
    memset(&P, 0, sizeof(P));
    P.use_capped = o.use_capped_tubes; P.ao_radius = o.bake_radius; P.ao_spp = o.bake_spp; P.ao_spp_local = o.bake_spp; P.ao_sample_first = 0; P.ao_use_distance = o.bake_use_distance;
    P.ao_refill_below = int(o.ao_refill_below ? o.ao_refill_below : (o.ao_raybuf ? (o.ao_packed && o.ao_wide ? 28u : 30u) : 24u)); P.ao_leaf_vote = int(o.ao_leaf_vote); P.ao_wide_reps = int(o.ao_wide_reps); P.frame_number = sc->bake_done;
    P.ao_tq_lo = (1.0f / P.ao_radius) * (1.0f - 1.0f / 4096.0f); P.ao_tq_hi = (1.0f / P.ao_radius) * (1.0f + 1.0f / 4096.0f);
    k_bake_setup<<<c->num_sms * 8, 256, 0, c->stream>>>(B, c->ao_hits.p);
    c->rtao_rays_timed = false;
    if (o.ao_triangles && (rc = ensure_tube_mesh(c, sc))) return rc;
    if (!o.ao_triangles && o.ao_wide) { if ((rc = ensure_wnodes(c, sc))) return rc; sc->set_w_top(o.ao_wide_top); }
    if (!o.ao_triangles && !o.ao_wide && o.ao_qnodes && (rc = ensure_qnodes(c, sc))) return rc;
    if ((rc = launch_ao_rays<true>(c, P, sc->dev(), sc->leaf_size == 1, (unsigned long long)n_rec * o.bake_spp, o.ao_triangles))) return rc;
    c->rtao_rays_timed = true;
    k_rtao_reduce<<<c->num_sms * 4, 256, 0, c->stream>>>(P, c->occ.p, c->ao_hits.p, c->small.p, sc->factors.p);
    LV_CUDA(c, cudaGetLastError());
    sc->bake_done++;
    return LV_OK;
}

// "RTAO (Prebaker)" as the active AO mode of a frame: LineRenderer::renderBase runs one baking iteration per rendered frame
// while the baker is still collecting samples (BakingMode::ITERATIVE_UPDATE, reference LineRenderer.cpp:257-264), the hit
// shader then looks the factors up by (line vertex id, phi).  The scene is the baker's state holder, hence the const_cast.
int prepare_static_ao(lv_ctx* c, const lv_scene* sc_const, FrameParams& P) {
    if (!c->opt.ao_prebaker || !(c->opt.ao_strength > 0.0f)) return LV_OK;
    lv_scene* sc = const_cast<lv_scene*>(sc_const);
    if (!sc->has_lines) return fail(c, LV_ERR_STATE, "ambient_occlusion_mode 'RTAO (Prebaker)' needs line frames (lv_scene_set_lines)");
    int rc = ensure_parametrization(c, sc);
    if (rc) return rc;
    if (sc->bake_done < c->opt.bake_iterations && (rc = run_bake_iteration(c, sc))) return rc;
    P.use_ao = 1; P.use_static_ao = 1; P.ao_tex = nullptr;
    P.sao_factors = sc->factors.p; P.sao_weights = sc->weights.p;
    P.n_ao_subdiv = sc->param_subdiv; P.n_line_vertices = uint32_t(sc->n_pt); P.n_param_vertices = sc->n_param;
    return LV_OK;
}

int ppll_prepare(lv_ctx* c, const lv_scene* sc, const FrameParams& P, uint64_t linked_list_size) {
    const size_t npad = size_t(P.padded_w) * P.padded_h;
    if (linked_list_size == 0) {
        // expectedAvgDepthComplexity x paddedW x paddedH (reference PerPixelLinkedListLineRenderer.cpp:251-258), 20 / 120 at > 1 M
        // segments (.hpp:45-49, .cpp:109-126); a shard only stores its own share.
        uint64_t avg = c->opt.expected_avg_depth_complexity ? c->opt.expected_avg_depth_complexity
                                                            : ((sc && sc->n_seg > 1000000ull) ? 120ull : 20ull);
        linked_list_size = (avg * npad + c->world - 1) / c->world;
    }
    if (linked_list_size > 0xFFFFFFFEull) linked_list_size = 0xFFFFFFFEull;  // `next` is a u32 and 0xFFFFFFFF terminates
    LV_CUDA(c, c->heads.ensure(npad));
    LV_CUDA(c, c->counts.ensure(npad));
    LV_CUDA(c, c->nodes.ensure(linked_list_size));
    LV_CUDA(c, c->frag_counter.ensure(1));
    c->list_size = linked_list_size; c->padded_w = P.padded_w; c->padded_h = P.padded_h;
    return LV_OK;
}

}  // namespace

extern "C" {

int lv_abi_version(void) { return LV_ABI_VERSION; }
const char* lv_last_global_error(void) { return g_global_error.c_str(); }
const char* lv_last_error(const lv_ctx* ctx) { return ctx ? ctx->error.c_str() : g_global_error.c_str(); }

int lv_ctx_create(lv_ctx** out, int device, void* cuda_stream) {
    if (!out) return fail(nullptr, LV_ERR_INVALID_ARGUMENT, "out == NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(nullptr, LV_ERR_NO_DEVICE, "no CUDA device visible; linevis_b200 has no CPU fallback"); }
    if (device < 0 || device >= n) return fail(nullptr, LV_ERR_INVALID_ARGUMENT, "device index out of range");
    LV_CUDA(nullptr, cudaSetDevice(device));
    lv_ctx* c = new lv_ctx();
    c->device = device;
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    cudaDeviceProp prop;
    cudaError_t err = cudaGetDeviceProperties(&prop, device);
    c->num_sms = prop.multiProcessorCount;
    for (auto& e : c->ev) if (err == cudaSuccess) err = cudaEventCreate(&e);
    if (err != cudaSuccess) {   // nothing of the half-built context survives a failure
        for (auto& e : c->ev) if (e) cudaEventDestroy(e);
        delete c;
        return fail(nullptr, LV_ERR_CUDA, std::string("lv_ctx_create: ") + cudaGetErrorString(err));
    }
    *out = c;
    return LV_OK;
}

int lv_ctx_destroy(lv_ctx* c) {
    if (!c) return LV_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->rgba8.release(); c->pixel_rays.release(); c->stage8[0].release(); c->stage8[1].release();
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join); }
    c->first_hits.release();
    for (int k = 0; k < 2; k++) { if (c->ev_rendered[k]) cudaEventDestroy(c->ev_rendered[k]); if (c->ev_copied[k]) cudaEventDestroy(c->ev_copied[k]); }
    c->tf.release(); c->tile_hist.release(); c->small2.release(); c->tiles_dev.release(); c->owned_map.release(); c->stage.release(); c->list_offs.release(); c->fill_cursor.release(); c->scan_tmp.release(); c->tiles_tmp.release(); c->image.release(); c->ao.release(); c->apron_marks.release(); c->occ.release(); c->depth_mm.release(); c->hits.release(); c->ao_hits.release();
    c->counters.release(); c->small.release(); c->heads.release(); c->counts.release(); c->bin_order.release(); c->bin_hist.release(); c->nodes.release(); c->frag_counter.release();
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    delete c;
    return LV_OK;
}

int lv_synchronize(lv_ctx* c) {
    if (!c) return LV_ERR_INVALID_ARGUMENT;
    LV_CUDA(c, cudaSetDevice(c->device));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->copy_stream && c->copies_pending) { LV_CUDA(c, cudaStreamSynchronize(c->copy_stream)); c->copies_pending = false; }   // b200_async_delivery
    return LV_OK;
}

int lv_set_option(lv_ctx* c, const char* key, const char* value) {
    if (!c || !key || !value) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_set_option: NULL argument");
    Options& o = c->opt;
    const std::string k(key);
    auto f = [&]() { return float(atof(value)); };
    auto u = [&]() { return uint32_t(strtoul(value, nullptr, 10)); };
    if (k == "line_width") o.line_width = f();
    else if (k == "band_width") o.band_width = f();
    else if (k == "depth_cue_strength") o.depth_cue_strength = f() > 0.0f ? f() : 0.0f;   // <= 0 switches USE_DEPTH_CUES off (LineRenderer.cpp:449-460)
    else if (k == "ambient_occlusion_mode") {
        // AMBIENT_OCCLUSION_BAKER_TYPE_NAMES (reference AmbientOcclusionBaker.hpp:78-95); "RTAO" is kept as an alias of the screen-space mode
        if (!strcmp(value, "RTAO (Screen Space)") || !strcmp(value, "RTAO")) { o.ao_prebaker = false; o.ao_mode = "RTAO (Screen Space)"; }
        else if (!strcmp(value, "RTAO (Prebaker)")) { o.ao_prebaker = true; o.ao_mode = value; }
        else return fail(c, LV_ERR_INVALID_ARGUMENT, "ambient_occlusion_mode must be 'RTAO (Screen Space)' or 'RTAO (Prebaker)' (SSAO / GTAO are out of scope)");
    } else if (k == "ambient_occlusion_strength") o.ao_strength = f();
    else if (k == "ambient_occlusion_gamma") o.ao_gamma = f();
    else if (k == "ambient_occlusion_iterations") o.ao_iterations = u();
    else if (k == "ambient_occlusion_samples_per_frame") { if (u() == 0 || u() > 4096) return fail(c, LV_ERR_INVALID_ARGUMENT, "ambient_occlusion_samples_per_frame must be in [1, 4096]"); o.ao_spp = u(); }
    else if (k == "ambient_occlusion_radius") o.ao_radius = f();
    else if (k == "ambient_occlusion_distance_based") o.ao_use_distance = parse_bool(value);
    else if (k == "use_jittered_primary_rays") o.ao_jitter_primary = parse_bool(value);
    else if (k == "ambient_occlusion_denoiser") {
        if (strcmp(value, "None")) return fail(c, LV_ERR_INVALID_ARGUMENT, "denoisers are out of scope; use ambient_occlusion_denoiser = None");
    } else if (k == "geometry_mode") {   // RAY_TRACING_GEOMETRY_MODE_NAMES, VulkanRayTracer.hpp:58-63
        if (strcmp(value, "AABBs (analytic)") && strcmp(value, "Triangle Mesh"))
            return fail(c, LV_ERR_INVALID_ARGUMENT, "geometry_mode must be 'AABBs (analytic)' or 'Triangle Mesh' (linear swept spheres are not implemented)");
        o.geometry_mode = value; o.tube_triangles = !strcmp(value, "Triangle Mesh");
    } else if (k == "use_analytic_intersections") {   // the older spelling of the same switch (VulkanRayTracer.cpp:244-251)
        o.tube_triangles = !parse_bool(value); o.geometry_mode = o.tube_triangles ? "Triangle Mesh" : "AABBs (analytic)";
    } else if (k == "num_samples_per_frame") { if (u() == 0) return fail(c, LV_ERR_INVALID_ARGUMENT, "num_samples_per_frame must be >= 1"); o.num_samples_per_frame = u(); }
    else if (k == "num_accumulated_frames") o.num_accumulated_frames = u();
    else if (k == "use_deterministic_sampling") o.use_deterministic_sampling = parse_bool(value);
    else if (k == "use_mlat") { if (parse_bool(value)) return fail(c, LV_ERR_INVALID_ARGUMENT, "MLAT is out of scope"); }
    else if (k == "mlat_num_nodes") {}
    else if (k == "use_capped_tubes") o.use_capped_tubes = parse_bool(value);
    else if (k == "use_halos") o.use_halos = parse_bool(value);
    else if (k == "tube_num_subdivisions") { if (u() < 3) return fail(c, LV_ERR_INVALID_ARGUMENT, "tube_num_subdivisions must be >= 3"); o.tube_num_subdivisions = u(); }
    else if (k == "b200_prebaker_iterations") { if (u() == 0 || u() > 4096) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_prebaker_iterations must be in [1, 4096]"); o.bake_iterations = u(); }
    else if (k == "b200_prebaker_samples_per_frame") { if (u() == 0 || u() > 4096) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_prebaker_samples_per_frame must be in [1, 4096]"); o.bake_spp = u(); }
    else if (k == "b200_prebaker_subdivisions") { if (u() < 3 || u() > 64) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_prebaker_subdivisions must be in [3, 64]"); o.bake_subdiv = u(); }
    else if (k == "b200_prebaker_param_segment_length") { if (!(f() > 0.0f)) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_prebaker_param_segment_length must be > 0"); o.bake_param_len = f(); }
    else if (k == "b200_prebaker_radius") { if (!(f() > 0.0f)) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_prebaker_radius must be > 0"); o.bake_radius = f(); }
    else if (k == "b200_prebaker_distance_based") o.bake_use_distance = parse_bool(value);
    else if (k == "b200_max_depth_complexity") o.max_depth_complexity = u();
    else if (k == "b200_tiling_width" || k == "b200_tiling_height") {
        uint32_t v = u();
        if (v == 0 || (v & (v - 1))) return fail(c, LV_ERR_INVALID_ARGUMENT, "tiling sizes must be powers of two");
        (k == "b200_tiling_width" ? o.tiling_w : o.tiling_h) = v;
    } else if (k == "b200_bvh_builder") {
        if (strcmp(value, "lbvh") && strcmp(value, "ploc") && strcmp(value, "sah")) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_bvh_builder must be lbvh, ploc or sah");
        o.bvh_ploc = !strcmp(value, "ploc"); o.bvh_sah_host = !strcmp(value, "sah");
    } else if (k == "b200_bvh_morton") {
        if (strcmp(value, "cubic") && strcmp(value, "per_axis")) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_bvh_morton must be cubic or per_axis");
        o.bvh_cubic_morton = !strcmp(value, "cubic");
    } else if (k == "b200_bvh_ploc_radius") { if (u() == 0 || u() > 64) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_bvh_ploc_radius must be in [1, 64]"); o.bvh_ploc_radius = u();
    } else if (k == "b200_bvh_leaf_size") { if (u() == 0 || u() > 16) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_bvh_leaf_size must be in [1, 16]"); o.bvh_leaf_size = u(); }
    else if (k == "b200_expected_avg_depth_complexity") o.expected_avg_depth_complexity = u();
    else if (k == "b200_ppll_binned_resolve") o.ppll_binned_resolve = parse_bool(value);
    else if (k == "b200_ppll_reg_sort") o.ppll_reg_sort = parse_bool(value);
    else if (k == "b200_ppll_gather_mode") {
        if (strcmp(value, "raycast") && strcmp(value, "raster") && strcmp(value, "raster_contiguous"))
            return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ppll_gather_mode must be raycast, raster or raster_contiguous");
        o.ppll_raster_gather = strcmp(value, "raycast") != 0;
        o.ppll_contiguous = !strcmp(value, "raster_contiguous");
    }
    else if (k == "b200_ppll_raster_slack") { if (!(f() >= 0.0f)) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ppll_raster_slack must be >= 0"); o.ppll_raster_slack = f(); }
    else if (k == "b200_ppll_raster_min_blocks") { if (u() < 4 || u() > 6) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ppll_raster_min_blocks must be 4, 5 or 6"); o.ppll_raster_min_blocks = u(); }
    else if (k == "b200_ppll_resolve_tile") { if (u() != 256 && u() != 512 && u() != 1024) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ppll_resolve_tile must be 256, 512 or 1024"); o.ppll_resolve_tile = u(); }
    else if (k == "b200_ao_min_blocks") o.ao_min_blocks = u();
    else if (k == "b200_ao_queue") o.ao_queue = parse_bool(value);
    else if (k == "b200_ao_qnodes") o.ao_qnodes = parse_bool(value);
    else if (k == "b200_ao_wide") o.ao_wide = parse_bool(value);
    else if (k == "b200_ao_raybuf") o.ao_raybuf = parse_bool(value);
    else if (k == "b200_ao_packed") o.ao_packed = parse_bool(value);
    else if (k == "b200_ao_tq_bits") { if (u() != 0 && u() != 4 && u() != 7) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ao_tq_bits must be 0 (automatic), 4 or 7"); o.ao_tq_bits = u(); }
    else if (k == "b200_packet_carveout") {
        // shared-memory carveout (percent of the SM's L1 / shared array) the packet kernels of the tube + RTAO frame ask for.  An SM has
        // ONE carveout at a time: blocks of a kernel that wants another one cannot start on an SM until every block of the persistent AO
        // stream (which needs the largest carveout) has left it.  100 lets them share SMs with the draining stream; -1 = the driver's choice.
        const int pct = std::atoi(value);
        if (pct < -1 || pct > 100) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_packet_carveout must be -1 (default) or 0..100");
        LV_CUDA(c, cudaSetDevice(c->device));
        LV_CUDA(c, cudaFuncSetAttribute(k_rtao_primary<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        LV_CUDA(c, cudaFuncSetAttribute(k_tube_first, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        LV_CUDA(c, cudaFuncSetAttribute(k_tubes<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        LV_CUDA(c, cudaFuncSetAttribute(k_tubes<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        LV_CUDA(c, cudaFuncSetAttribute(k_rtao_reduce, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        o.packet_carveout = pct;
    }
    else if (k == "b200_frame_format") {
        if (strcmp(value, "rgba32f") && strcmp(value, "rgba8")) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_frame_format must be rgba32f or rgba8");
        o.frame_rgba8 = !strcmp(value, "rgba8");
    }
    else if (k == "b200_async_delivery") o.async_delivery = parse_bool(value);
    else if (k == "b200_tube_prepass") o.tube_prepass = parse_bool(value);
    else if (k == "b200_ao_wide_reps") { if (u() == 0 || u() > 4) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ao_wide_reps must be in [1, 4]"); o.ao_wide_reps = u(); }
    else if (k == "b200_ao_wide_top") { if (u() > 1365) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ao_wide_top must be <= 1365 nodes"); o.ao_wide_top = u(); }
    else if (k == "b200_rtao_geometry") {
        if (!strcmp(value, "triangles")) o.ao_triangles = true;
        else if (!strcmp(value, "capsules")) o.ao_triangles = false;
        else return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_rtao_geometry must be 'capsules' or 'triangles'");
    }
    else if (k == "b200_ao_stack") { if (u() != 0 && u() != 1 && u() != 8 && u() != 12 && u() != 16) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ao_stack must be 0, 1, 8, 12 or 16"); o.ao_stack = u(); }
    else if (k == "b200_ao_leaf_vote") { if (u() == 0 || u() > 32) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ao_leaf_vote must be in [1, 32]"); o.ao_leaf_vote = u(); }
    else if (k == "b200_ao_refill_below") { if (u() > 32) return fail(c, LV_ERR_INVALID_ARGUMENT, "b200_ao_refill_below must be in [0, 32] (0 = automatic)"); o.ao_refill_below = u(); }
    else return fail(c, LV_ERR_UNKNOWN_OPTION, "unknown option '" + k + "'");
    return LV_OK;
}

int lv_get_option(const lv_ctx* c, const char* key, char* buf, size_t cap) {
    if (!c || !key || !buf || cap == 0) return LV_ERR_INVALID_ARGUMENT;
    const Options& o = c->opt;
    const std::string k(key);
    std::string v;
    auto b = [](bool x) { return std::string(x ? "true" : "false"); };
    if (k == "line_width") v = std::to_string(o.line_width);
    else if (k == "band_width") v = std::to_string(o.band_width);
    else if (k == "depth_cue_strength") v = std::to_string(o.depth_cue_strength);
    else if (k == "ambient_occlusion_mode") v = o.ao_mode;
    else if (k == "ambient_occlusion_strength") v = std::to_string(o.ao_strength);
    else if (k == "ambient_occlusion_gamma") v = std::to_string(o.ao_gamma);
    else if (k == "ambient_occlusion_iterations") v = std::to_string(o.ao_iterations);
    else if (k == "ambient_occlusion_samples_per_frame") v = std::to_string(o.ao_spp);
    else if (k == "ambient_occlusion_radius") v = std::to_string(o.ao_radius);
    else if (k == "ambient_occlusion_distance_based") v = b(o.ao_use_distance);
    else if (k == "use_jittered_primary_rays") v = b(o.ao_jitter_primary);
    else if (k == "ambient_occlusion_denoiser") v = o.denoiser;
    else if (k == "geometry_mode") v = o.geometry_mode;
    else if (k == "use_analytic_intersections") v = b(!o.tube_triangles);
    else if (k == "num_samples_per_frame") v = std::to_string(o.num_samples_per_frame);
    else if (k == "num_accumulated_frames") v = std::to_string(o.num_accumulated_frames);
    else if (k == "use_deterministic_sampling") v = b(o.use_deterministic_sampling);
    else if (k == "use_mlat") v = "false";
    else if (k == "use_capped_tubes") v = b(o.use_capped_tubes);
    else if (k == "use_halos") v = b(o.use_halos);
    else if (k == "tube_num_subdivisions") v = std::to_string(o.tube_num_subdivisions);
    else if (k == "b200_prebaker_iterations") v = std::to_string(o.bake_iterations);
    else if (k == "b200_prebaker_samples_per_frame") v = std::to_string(o.bake_spp);
    else if (k == "b200_prebaker_subdivisions") v = std::to_string(o.bake_subdiv);
    else if (k == "b200_prebaker_param_segment_length") v = std::to_string(o.bake_param_len);
    else if (k == "b200_prebaker_radius") v = std::to_string(o.bake_radius);
    else if (k == "b200_prebaker_distance_based") v = b(o.bake_use_distance);
    else if (k == "b200_max_depth_complexity") v = std::to_string(o.max_depth_complexity);
    else if (k == "b200_tiling_width") v = std::to_string(o.tiling_w);
    else if (k == "b200_tiling_height") v = std::to_string(o.tiling_h);
    else if (k == "b200_bvh_leaf_size") v = std::to_string(o.bvh_leaf_size);
    else if (k == "b200_bvh_builder") v = o.bvh_sah_host ? "sah" : o.bvh_ploc ? "ploc" : "lbvh";
    else if (k == "b200_bvh_morton") v = o.bvh_cubic_morton ? "cubic" : "per_axis";
    else if (k == "b200_bvh_ploc_radius") v = std::to_string(o.bvh_ploc_radius);
    else if (k == "b200_expected_avg_depth_complexity") v = std::to_string(o.expected_avg_depth_complexity);
    else if (k == "b200_ao_refill_below") v = std::to_string(o.ao_refill_below);
    else if (k == "b200_ao_leaf_vote") v = std::to_string(o.ao_leaf_vote);
    else if (k == "b200_ao_min_blocks") v = std::to_string(o.ao_min_blocks);
    else if (k == "b200_ao_queue") v = b(o.ao_queue);
    else if (k == "b200_ao_qnodes") v = b(o.ao_qnodes);
    else if (k == "b200_ao_wide") v = b(o.ao_wide);
    else if (k == "b200_ao_raybuf") v = b(o.ao_raybuf);
    else if (k == "b200_ao_packed") v = b(o.ao_packed);
    else if (k == "b200_packet_carveout") v = std::to_string(o.packet_carveout);
    else if (k == "b200_ao_tq_bits") v = std::to_string(o.ao_tq_bits);
    else if (k == "b200_frame_format") v = o.frame_rgba8 ? "rgba8" : "rgba32f";
    else if (k == "b200_async_delivery") v = b(o.async_delivery);
    else if (k == "b200_tube_prepass") v = b(o.tube_prepass);
    else if (k == "b200_ao_wide_top") v = std::to_string(o.ao_wide_top);
    else if (k == "b200_ao_wide_reps") v = std::to_string(o.ao_wide_reps);
    else if (k == "b200_rtao_geometry") v = o.ao_triangles ? "triangles" : "capsules";
    else if (k == "b200_ao_stack") v = std::to_string(o.ao_stack);
    else if (k == "b200_ppll_binned_resolve") v = b(o.ppll_binned_resolve);
    else if (k == "b200_ppll_reg_sort") v = b(o.ppll_reg_sort);
    else if (k == "b200_ppll_gather_mode") v = o.ppll_contiguous ? "raster_contiguous" : (o.ppll_raster_gather ? "raster" : "raycast");
    else if (k == "b200_ppll_resolve_tile") v = std::to_string(o.ppll_resolve_tile);
    else if (k == "b200_ppll_raster_slack") v = std::to_string(o.ppll_raster_slack);
    else if (k == "b200_ppll_raster_min_blocks") v = std::to_string(o.ppll_raster_min_blocks);
    else return LV_ERR_UNKNOWN_OPTION;
    snprintf(buf, cap, "%s", v.c_str());
    return LV_OK;
}

int lv_set_transfer_function(lv_ctx* c, const float* rgba, uint32_t K, float attr_min, float attr_max) {
    if (!c || !rgba || K == 0) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_set_transfer_function: need K >= 1 entries");
    LV_CUDA(c, cudaSetDevice(c->device));
    LV_CUDA(c, c->tf.ensure(K));
    LV_CUDA(c, cudaMemcpyAsync(c->tf.p, rgba, size_t(K) * 16, cudaMemcpyDefault, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    c->tfK = K; c->amin = attr_min; c->amax = attr_max;
    return LV_OK;
}

int lv_set_tile_shard(lv_ctx* c, uint32_t rank, uint32_t world, uint32_t tile_size) {
    if (!c || world == 0 || rank >= world) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_set_tile_shard: need rank < world");
    if (tile_size == 0 || tile_size % 16) return fail(c, LV_ERR_INVALID_ARGUMENT, "tile_size must be a positive multiple of 16");
    c->rank = rank; c->world = world; c->tile_size = tile_size;
    c->tile_owner.clear(); c->owner_w = c->owner_h = 0; c->peer_w = c->peer_h = 0;
    c->tiles_w = c->tiles_h = 0;  // re-enumerate on next frame
    c->ao_w = c->ao_h = 0;
    c->apron_stamp = 0;           // the stamp array is cleared again before its next use
    return LV_OK;
}

int lv_set_tile_owners(lv_ctx* c, uint32_t width, uint32_t height, const unsigned char* owners, uint32_t n_tiles) {
    if (!c) return LV_ERR_INVALID_ARGUMENT;
    const uint32_t tx = (width + c->tile_size - 1) / c->tile_size, ty = (height + c->tile_size - 1) / c->tile_size;
    if (!owners || n_tiles != tx * ty) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_set_tile_owners: need one owner per tile of the frame (" + std::to_string(tx * ty) + ")");
    for (uint32_t i = 0; i < n_tiles; i++) if (owners[i] >= c->world) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_set_tile_owners: owner >= world");
    c->tile_owner.assign(owners, owners + n_tiles); c->owner_w = width; c->owner_h = height;
    c->tiles_w = c->tiles_h = 0; c->peer_w = c->peer_h = 0;   // re-enumerate on next frame
    c->ao_w = c->ao_h = 0;
    c->apron_stamp = 0;
    return LV_OK;
}

// per tile (Morton order, all tiles of the frame; 0 for tiles of other ranks): hit pixels of the last RTAO pass of this context
__global__ void k_tile_hist(const AoHit* hits, const unsigned int* n_hit, uint32_t W, uint32_t tile_size, uint32_t tiles_x, unsigned int* hist) {
    const uint32_t n = *n_hit;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t px = __float_as_uint(hits[i].nrm_px.w);
        atomicAdd(&hist[((px / W) / tile_size) * tiles_x + (px % W) / tile_size], 1u);
    }
}

int lv_get_tile_costs(lv_ctx* c, uint32_t width, uint32_t height, uint32_t* costs, uint32_t n_tiles) {
    if (!c || !costs) return LV_ERR_INVALID_ARGUMENT;
    const uint32_t tx = (width + c->tile_size - 1) / c->tile_size, ty = (height + c->tile_size - 1) / c->tile_size;
    if (n_tiles != tx * ty) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_get_tile_costs: need room for one cost per tile of the frame");
    if (!c->ao_hits.p || !c->small.p || c->ao_w != width || c->ao_h != height) return fail(c, LV_ERR_STATE, "lv_get_tile_costs: no RTAO pass of this frame size has run");
    LV_CUDA(c, cudaSetDevice(c->device));
    LV_CUDA(c, c->tile_hist.ensure(n_tiles));
    LV_CUDA(c, cudaMemsetAsync(c->tile_hist.p, 0, size_t(n_tiles) * 4, c->stream));
    k_tile_hist<<<c->num_sms * 4, 256, 0, c->stream>>>(c->ao_hits.p, c->small.p, width, c->tile_size, tx, c->tile_hist.p);
    std::vector<unsigned int> lin(n_tiles);
    LV_CUDA(c, cudaMemcpyAsync(lin.data(), c->tile_hist.p, size_t(n_tiles) * 4, cudaMemcpyDeviceToHost, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<uint2> all;
    morton_tiles(width, height, c->tile_size, all);
    for (uint32_t i = 0; i < n_tiles; i++) costs[i] = lin[size_t(all[i].y) * tx + all[i].x];
    return LV_OK;
}

int lv_get_owned_tiles(const lv_ctx* c, uint32_t width, uint32_t height, uint32_t* tiles_xy, uint32_t* n_owned) {
    if (!c || !n_owned) return LV_ERR_INVALID_ARGUMENT;
    std::vector<uint2> t;
    enumerate_tiles(width, height, c->tile_size, c->rank, c->world, t, tile_owners_for(c, width, height));
    *n_owned = uint32_t(t.size());
    if (tiles_xy) for (size_t i = 0; i < t.size(); i++) { tiles_xy[2 * i] = t[i].x; tiles_xy[2 * i + 1] = t[i].y; }
    return LV_OK;
}

int lv_pack_owned_tiles(lv_ctx* c, const float* image, uint32_t W, uint32_t H, float* packed) {
    if (!c || !image || !packed) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_pack_owned_tiles: NULL argument");
    LV_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_tiles(c, W, H);
    if (rc) return rc;
    if (c->tiles_host.empty()) return LV_OK;
    const uint32_t tt = c->tile_size * c->tile_size;
    dim3 grid((tt + 255) / 256, uint32_t(c->tiles_host.size()));
    k_pack_tiles<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(image), W, H, c->tiles_dev.p, c->tile_size, reinterpret_cast<float4*>(packed));
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int lv_unpack_tiles(lv_ctx* c, const float* packed, uint32_t src_rank, uint32_t world, uint32_t W, uint32_t H, float* image) {
    if (!c || !image || !packed || world == 0 || src_rank >= world) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_unpack_tiles: bad argument");
    LV_CUDA(c, cudaSetDevice(c->device));
    // device copies of every rank's tile list are cached per (W, H, world, tile size): no host work / sync per frame
    if (c->peer_w != W || c->peer_h != H || c->peer_world != world || c->peer_tile != c->tile_size) {
        std::vector<uint2> all;
        c->peer_off.assign(world + 1, 0);
        for (uint32_t r = 0; r < world; r++) {
            std::vector<uint2> t;
            enumerate_tiles(W, H, c->tile_size, r, world, t, tile_owners_for(c, W, H));
            all.insert(all.end(), t.begin(), t.end());
            c->peer_off[r + 1] = uint32_t(all.size());
        }
        LV_CUDA(c, c->tiles_tmp.ensure(std::max<size_t>(1, all.size())));
        LV_CUDA(c, cudaMemcpyAsync(c->tiles_tmp.p, all.data(), all.size() * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
        LV_CUDA(c, cudaStreamSynchronize(c->stream));
        c->peer_w = W; c->peer_h = H; c->peer_world = world; c->peer_tile = c->tile_size;
    }
    const uint32_t n = c->peer_off[src_rank + 1] - c->peer_off[src_rank];
    if (n == 0) return LV_OK;
    const uint32_t tt = c->tile_size * c->tile_size;
    dim3 grid((tt + 255) / 256, n);
    k_unpack_tiles<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(packed), W, H, c->tiles_tmp.p + c->peer_off[src_rank], c->tile_size,
                                                reinterpret_cast<float4*>(image));
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

// ---------------------------------------------------------------------------------------------- peer-memory frames
int lv_frame_alloc(lv_ctx* c, uint32_t width, uint32_t height, float** frame_out) {
    if (!c || !frame_out || width == 0 || height == 0) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_frame_alloc: bad argument");
    *frame_out = nullptr;
    LV_CUDA(c, cudaSetDevice(c->device));
    float* p = nullptr;
    const size_t bytes = size_t(width) * height * 16;
    LV_CUDA(c, cudaMalloc(&p, bytes));   // a plain cudaMalloc allocation: exportable with cudaIpcGetMemHandle at offset 0
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cudaFree(p); return fail(c, LV_ERR_CUDA, std::string("lv_frame_alloc: ") + cudaGetErrorString(e)); }
    *frame_out = p;
    return LV_OK;
}

int lv_frame_free(lv_ctx* c, float* frame) {
    if (!c) return LV_ERR_INVALID_ARGUMENT;
    if (!frame) return LV_OK;
    LV_CUDA(c, cudaSetDevice(c->device));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    LV_CUDA(c, cudaFree(frame));
    return LV_OK;
}

int lv_ipc_export(lv_ctx* c, const void* device_ptr, void* handle_out) {
    if (!c || !device_ptr || !handle_out) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == LV_IPC_HANDLE_BYTES, "LV_IPC_HANDLE_BYTES must match cudaIpcMemHandle_t");
    LV_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    LV_CUDA(c, cudaIpcGetMemHandle(&h, const_cast<void*>(device_ptr)));
    memcpy(handle_out, &h, sizeof(h));
    return LV_OK;
}

int lv_ipc_open(lv_ctx* c, const void* handle, void** peer_ptr_out) {
    if (!c || !handle || !peer_ptr_out) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_ipc_open: NULL argument");
    *peer_ptr_out = nullptr;
    LV_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    LV_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));   // also enables peer access to the exporting GPU
    *peer_ptr_out = p;
    return LV_OK;
}

int lv_ipc_close(lv_ctx* c, void* peer_ptr) {
    if (!c) return LV_ERR_INVALID_ARGUMENT;
    if (!peer_ptr) return LV_OK;
    LV_CUDA(c, cudaSetDevice(c->device));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    LV_CUDA(c, cudaIpcCloseMemHandle(peer_ptr));
    return LV_OK;
}

// ---------------------------------------------------------------------------------------------- scene
int lv_scene_create_device(lv_ctx* c, lv_scene** out, const float* d_pos, const float* d_attr, const uint32_t* d_idx,
                           uint64_t n_pt, uint64_t n_seg, float line_width) {
    if (!c || !out) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_create: NULL argument");
    *out = nullptr;
    if (n_seg > 0 && (!d_pos || !d_attr || !d_idx || n_pt == 0)) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_create: NULL array");
    if (n_seg >= (1ull << 27)) return fail(c, LV_ERR_INVALID_ARGUMENT, "too many segments (max 2^27-1: leaf references carry 27 index bits)");
    if (line_width <= 0.0f) line_width = c->opt.line_width;
    LV_CUDA(c, cudaSetDevice(c->device));
    lv_scene* s = new lv_scene();
    s->ctx = c; s->n_seg = n_seg; s->n_pt = n_pt; s->line_width = line_width; s->leaf_size = c->opt.bvh_leaf_size;
    const int n = int(n_seg);
    if (n == 0) { *out = s; return LV_OK; }
    const float r = line_width * 0.5f;
    cudaStream_t st = c->stream;
    DevBuf<float> bounds, boxes; DevBuf<unsigned long long> keys, keys2; DevBuf<uint32_t> vals;
    DevBuf<int2> children, ranges; DevBuf<int> parent; DevBuf<unsigned int> flags; DevBuf<char> cubtmp;
    auto cleanup = [&]() { bounds.release(); boxes.release(); keys.release(); keys2.release(); vals.release(); children.release(); ranges.release(); parent.release(); flags.release(); cubtmp.release(); };
#define LV_BUILD(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); delete s; return fail(c, e__ == cudaErrorMemoryAllocation ? LV_ERR_OUT_OF_MEMORY : LV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    LV_BUILD(cudaEventRecord(c->ev[0], st));
    LV_BUILD(bounds.ensure(6)); LV_BUILD(keys.ensure(n)); LV_BUILD(keys2.ensure(n)); LV_BUILD(vals.ensure(n));
    LV_BUILD(s->prim_ids.ensure(size_t(n) + 1)); LV_BUILD(s->segs.ensure(size_t(n) + 1));   // + the dummy record an absent child refers to (k_emit_nodes)
    LV_BUILD(cudaMemsetAsync(s->segs.p + n, 0xFF, sizeof(SegRec), st));            // all NaN: never hit
    LV_BUILD(cudaMemsetAsync(s->prim_ids.p + n, 0xFF, 4, st));
    LV_BUILD(s->seg_idx.ensure(n));   // caller's index pairs, kept for lv_scene_set_lines (8 B / segment)
    LV_BUILD(cudaMemcpyAsync(s->seg_idx.p, d_idx, size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
    k_init_bounds<<<1, 32, 0, st>>>(bounds.p);
    k_scene_bounds<<<std::min(1024, (n + 255) / 256), 256, 0, st>>>(d_pos, d_idx, uint32_t(n), r, bounds.p);
    k_morton<<<(n + 255) / 256, 256, 0, st>>>(d_pos, d_idx, uint32_t(n), r, bounds.p, keys.p, vals.p, c->opt.bvh_cubic_morton);
    size_t cub_bytes = 0;
    LV_BUILD(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys.p, keys2.p, vals.p, s->prim_ids.p, n, 0, 63, st));
    LV_BUILD(cubtmp.ensure(cub_bytes + 16));
    LV_BUILD(cub::DeviceRadixSort::SortPairs(cubtmp.p, cub_bytes, keys.p, keys2.p, vals.p, s->prim_ids.p, n, 0, 63, st));
    k_pack_segments<<<(n + 255) / 256, 256, 0, st>>>(d_pos, d_attr, d_idx, s->prim_ids.p, uint32_t(n), s->segs.p);
    LV_BUILD(s->seg_axes.ensure(size_t(n) + 1));
    k_seg_axes<<<(n + 1 + 255) / 256, 256, 0, st>>>(s->segs.p, uint32_t(n) + 1u, s->seg_axes.p);   // incl. the dummy record (NaN)
    const int n_inner = std::max(1, n - 1);
    LV_BUILD(children.ensure(n_inner)); LV_BUILD(ranges.ensure(n_inner)); LV_BUILD(parent.ensure(2 * size_t(n)));
    LV_BUILD(boxes.ensure(6 * (2 * size_t(n)))); LV_BUILD(flags.ensure(n_inner));
    LV_BUILD(cudaMemsetAsync(flags.p, 0, size_t(n_inner) * 4, st));
    const bool ploc = c->opt.bvh_ploc && c->opt.bvh_leaf_size == 1 && n >= 2;
    LV_BUILD(s->nodes.ensure(n_inner));
    if (ploc) {
        // PLOC (lv_bvh.cuh): rounds of nearest-neighbour search in Morton order + merge + ordered compaction; one host read per round
        DevBuf<PlocCluster> ca, cb; DevBuf<uint32_t> nn; DevBuf<unsigned long long> pf, ps;
        auto pcleanup = [&]() { ca.release(); cb.release(); nn.release(); pf.release(); ps.release(); };
#define LV_PLOC(expr) do { cudaError_t e2__ = (expr); if (e2__ != cudaSuccess) { pcleanup(); LV_BUILD(e2__); } } while (0)
        LV_PLOC(ca.ensure(n)); LV_PLOC(cb.ensure(n)); LV_PLOC(nn.ensure(n)); LV_PLOC(pf.ensure(size_t(n) + 1)); LV_PLOC(ps.ensure(size_t(n) + 1));
        size_t scan_bytes = 0;
        LV_PLOC(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, pf.p, ps.p, n + 1, st));
        LV_PLOC(cubtmp.ensure(std::max(scan_bytes, cub_bytes) + 16));
        k_ploc_init<<<(n + 255) / 256, 256, 0, st>>>(s->segs.p, uint32_t(n), r, ca.p);
        uint32_t m = uint32_t(n), created = 0;
        PlocCluster *cin = ca.p, *cout = cb.p;
        const int radius = int(c->opt.bvh_ploc_radius);
        while (m > 1) {
            const uint32_t g = (m + 255) / 256;
            k_ploc_nearest<<<g, 256, 0, st>>>(cin, m, radius, nn.p);
            k_ploc_flags<<<g, 256, 0, st>>>(nn.p, m, pf.p);
            LV_PLOC(cudaMemsetAsync(pf.p + m, 0, 8, st));       // the scan's last element = the totals
            LV_PLOC(cub::DeviceScan::ExclusiveSum(cubtmp.p, scan_bytes, pf.p, ps.p, int(m) + 1, st));
            k_ploc_apply<<<g, 256, 0, st>>>(cin, nn.p, pf.p, ps.p, m, created, uint32_t(n_inner), s->nodes.p, cout);
            unsigned long long tot = 0;
            LV_PLOC(cudaMemcpyAsync(&tot, ps.p + m, 8, cudaMemcpyDeviceToHost, st));
            LV_PLOC(cudaStreamSynchronize(st));
            const uint32_t merged = uint32_t(tot >> 32);
            if (merged == 0) { pcleanup(); LV_BUILD(cudaErrorUnknown); }   // cannot happen: the pair with the globally smallest area is always mutual
            created += merged; m = uint32_t(tot);
            std::swap(cin, cout);
        }
        PlocCluster root;
        LV_PLOC(cudaMemcpyAsync(&root, cin, sizeof(root), cudaMemcpyDeviceToHost, st));
        LV_PLOC(cudaStreamSynchronize(st));
        uint32_t height; memcpy(&height, &root.hi.w, 4);
        LV_PLOC(cudaMemcpyAsync(flags.p, &height, 4, cudaMemcpyHostToDevice, st));    // where the depth is read from below
        LV_PLOC(cudaStreamSynchronize(st));
        pcleanup();
#undef LV_PLOC
    }
    bool sah_done = false;
    if (!ploc && c->opt.bvh_sah_host && c->opt.bvh_leaf_size == 1 && n >= 2) {
        // binned SAH on the host (lv_sah_host.hpp): the Morton-ordered records come back, the finished nodes go up.  A SAH tree can be
        // deep where the Morton tree is not; one that would not fit the traversal stacks is dropped for the Morton tree below.
        std::vector<float> h_segs; h_segs.resize(size_t(n) * 8);
        std::vector<Node64> h_nodes; h_nodes.resize(size_t(n_inner));
        LV_BUILD(cudaMemcpyAsync(h_segs.data(), s->segs.p, size_t(n) * sizeof(SegRec), cudaMemcpyDeviceToHost, st));
        LV_BUILD(cudaStreamSynchronize(st));
        const uint32_t depth = lvsah::build(h_segs.data(), uint32_t(n), r, h_nodes.data());
        if (depth + 1 <= uint32_t(kStackSize) && depth + 1 <= uint32_t(kAoStack)) {
            LV_BUILD(cudaMemcpyAsync(s->nodes.p, h_nodes.data(), size_t(n_inner) * sizeof(Node64), cudaMemcpyHostToDevice, st));
            LV_BUILD(cudaMemcpyAsync(flags.p, &depth, 4, cudaMemcpyHostToDevice, st));
            LV_BUILD(cudaStreamSynchronize(st));
            sah_done = true;
        }
    }
    if (!ploc && !sah_done) {
        if (n > 1) k_radix_tree<<<(n - 1 + 255) / 256, 256, 0, st>>>(keys2.p, n, children.p, ranges.p, parent.p);
        k_fit<<<(n + 255) / 256, 256, 0, st>>>(s->segs.p, n, r, children.p, parent.p, boxes.p, flags.p);
        k_emit_nodes<<<(n_inner + 255) / 256, 256, 0, st>>>(n, int(c->opt.bvh_leaf_size), children.p, ranges.p, boxes.p, s->nodes.p);
        LV_BUILD(cudaMemsetAsync(flags.p, 0, 4, st));   // reuse flags[0] as the depth accumulator
        k_tree_depth<<<(n + 255) / 256, 256, 0, st>>>(n, parent.p, flags.p);
    }
    LV_BUILD(cudaGetLastError());
    LV_BUILD(cudaEventRecord(c->ev[1], st));
    LV_BUILD(cudaMemcpyAsync(s->bounds, bounds.p, 24, cudaMemcpyDeviceToHost, st));
    LV_BUILD(cudaMemcpyAsync(&s->depth, flags.p, 4, cudaMemcpyDeviceToHost, st));
    LV_BUILD(cudaStreamSynchronize(st));
    if (s->depth + 1 > uint32_t(kStackSize) || s->depth + 1 > uint32_t(kAoStack)) {
        const uint32_t depth = s->depth;
        cleanup(); s->segs.release(); s->seg_axes.release(); s->prim_ids.release(); s->nodes.release(); s->seg_idx.release(); delete s;
        return fail(c, LV_ERR_STATE, "BVH depth " + std::to_string(depth) + " exceeds the traversal stack (" + std::to_string(kAoStack) + ")");
    }
    s->build_ms = elapsed(c->ev[0], c->ev[1]);
    s->n_nodes = uint64_t(n_inner);
    cleanup();
#undef LV_BUILD
    // the AO stream's 4-wide tree is part of the scene build when it is going to be used (b200_ao_wide at creation time), so that
    // its cost shows in the build time and not in the first frame; switched on later, it is built on first use
    if (c->opt.ao_wide && s->leaf_size == 1) {
        const int wrc = ensure_wnodes(c, s);
        if (wrc) { lv_scene_destroy(s); return wrc; }
        s->build_ms += s->w_build_ms;
    }
    *out = s;
    return LV_OK;
}

int lv_scene_create(lv_ctx* c, lv_scene** out, const float* pos, const float* attr, const uint32_t* idx, uint64_t n_pt, uint64_t n_seg, float line_width) {
    if (!c || !out) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_create: NULL argument");
    *out = nullptr;
    if (n_seg > 0 && (!pos || !attr || !idx || n_pt == 0)) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_create: NULL array");
    for (uint64_t i = 0; i < 2 * n_seg; i++) if (idx[i] >= n_pt) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_create: segment index out of range");
    LV_CUDA(c, cudaSetDevice(c->device));
    DevBuf<float> dpos, dattr; DevBuf<uint32_t> didx;
    int rc = LV_OK;
    auto up = [&]() -> cudaError_t {
        cudaError_t e;
        if ((e = dpos.ensure(std::max<uint64_t>(1, 3 * n_pt))) != cudaSuccess) return e;
        if ((e = dattr.ensure(std::max<uint64_t>(1, n_pt))) != cudaSuccess) return e;
        if ((e = didx.ensure(std::max<uint64_t>(1, 2 * n_seg))) != cudaSuccess) return e;
        if (n_seg == 0) return cudaSuccess;
        if ((e = cudaMemcpyAsync(dpos.p, pos, 12 * n_pt, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(dattr.p, attr, 4 * n_pt, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;
        return cudaMemcpyAsync(didx.p, idx, 8 * n_seg, cudaMemcpyHostToDevice, c->stream);
    };
    cudaError_t e = up();
    if (e != cudaSuccess) rc = fail(c, e == cudaErrorMemoryAllocation ? LV_ERR_OUT_OF_MEMORY : LV_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(e));
    else rc = lv_scene_create_device(c, out, dpos.p, dattr.p, didx.p, n_pt, n_seg, line_width);
    cudaStreamSynchronize(c->stream);
    dpos.release(); dattr.release(); didx.release();
    return rc;
}

int lv_scene_destroy(lv_scene* s) {
    if (!s) return LV_OK;
    if (s->ctx) { cudaSetDevice(s->ctx->device); cudaStreamSynchronize(s->ctx->stream); }
    s->segs.release(); s->seg_axes.release(); s->prim_ids.release(); s->nodes.release(); s->seg_idx.release();
    s->pt_pos.release(); s->pt_tan.release(); s->pt_nrm.release(); s->seg_aux.release();
    s->sampling.release(); s->weights.release(); s->factors.release();
    s->qnodes.release(); s->wnodes.release();
    s->tris.release(); s->tri_ids.release(); s->tri_nodes.release(); s->tri_vattr.release(); s->tri_line_pos.release(); s->tri_line_tan.release();
    delete s;
    return LV_OK;
}

// ---- line frames + object-space AO prebaker
int lv_scene_set_lines(lv_scene* s, const float* pos_xyz, const float* tangent_xyz, const float* normal_xyz, uint64_t n_pt,
                       const uint64_t* line_offsets, uint64_t n_lines) {
    if (!s || !s->ctx) return LV_ERR_INVALID_ARGUMENT;
    lv_ctx* c = s->ctx;
    if (!pos_xyz || !tangent_xyz || !normal_xyz || !line_offsets || n_lines == 0) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_set_lines: NULL argument");
    if (n_pt != s->n_pt || n_pt == 0) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_set_lines: n_pt differs from the scene's point count");
    if (line_offsets[0] != 0 || line_offsets[n_lines] != n_pt) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_set_lines: line_offsets must run from 0 to n_pt");
    for (uint64_t i = 0; i < n_lines; i++)
        if (line_offsets[i + 1] < line_offsets[i] + 2) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_scene_set_lines: every polyline needs >= 2 points");
    LV_CUDA(c, cudaSetDevice(c->device));
    s->has_lines = false;
    s->host_pos.assign(pos_xyz, pos_xyz + 3 * n_pt);
    s->line_offsets.assign(line_offsets, line_offsets + n_lines + 1);
    DevBuf<float> tmp;
    LV_CUDA(c, tmp.ensure(3 * n_pt));
    LV_CUDA(c, s->pt_pos.ensure(n_pt)); LV_CUDA(c, s->pt_tan.ensure(n_pt)); LV_CUDA(c, s->pt_nrm.ensure(n_pt));
    const float* src[3] = {pos_xyz, tangent_xyz, normal_xyz};
    float4* dst[3] = {s->pt_pos.p, s->pt_tan.p, s->pt_nrm.p};
    const uint32_t grid = uint32_t(std::min<uint64_t>((n_pt + 255) / 256, 65535));
    for (int k = 0; k < 3; k++) {
        LV_CUDA(c, cudaMemcpyAsync(tmp.p, src[k], 12 * n_pt, cudaMemcpyHostToDevice, c->stream));
        k_expand_xyz<<<grid, 256, 0, c->stream>>>(tmp.p, uint32_t(n_pt), dst[k]);
    }
    if (s->n_seg) {
        LV_CUDA(c, s->seg_aux.ensure(s->n_seg));
        k_seg_aux<<<uint32_t(std::min<uint64_t>((s->n_seg + 255) / 256, 65535)), 256, 0, c->stream>>>(s->prim_ids.p, s->seg_idx.p, s->pt_nrm.p, uint32_t(s->n_seg), s->seg_aux.p);
    }
    LV_CUDA(c, cudaGetLastError());
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    tmp.release();
    s->has_lines = true;
    s->n_param = 0; s->bake_done = 0;   // new frames invalidate parametrization and baked factors
    s->n_tri = 0; s->mesh_subdiv = 0;   // ... and the tube mesh
    return LV_OK;
}

int lv_ao_parametrize(const float* pos_xyz, const uint64_t* line_offsets, uint64_t n_lines, float expected_param_segment_length,
                      float* blending_weights, float* sampling_locations, uint64_t cap, uint64_t* n_param_vertices) {
    if (!pos_xyz || !line_offsets || n_lines == 0 || !(expected_param_segment_length > 0.0f) || !n_param_vertices) return LV_ERR_INVALID_ARGUMENT;
    std::vector<float> w, sl;
    ao_parametrize_host(pos_xyz, line_offsets, n_lines, expected_param_segment_length, w, sl);
    if (blending_weights) memcpy(blending_weights, w.data(), w.size() * 4);
    if (sampling_locations) memcpy(sampling_locations, sl.data(), std::min<uint64_t>(cap, sl.size()) * 4);
    *n_param_vertices = sl.size();
    return LV_OK;
}

int lv_tube_mesh(const float* pos_xyz, const uint64_t* line_offsets, uint64_t n_lines, float tube_radius, uint32_t num_subdivisions,
                 float* vertices, uint64_t vertices_cap, uint32_t* indices, uint64_t triangles_cap,
                 uint64_t* n_vertices, uint64_t* n_triangles, uint64_t* n_line_points) {
    if (!pos_xyz || !line_offsets || n_lines == 0 || !(tube_radius > 0.0f)) return LV_ERR_INVALID_ARGUMENT;
    lvmesh::TubeMesh m;
    lvmesh::build(pos_xyz, line_offsets, n_lines, tube_radius, int(num_subdivisions), m);
    static_assert(sizeof(lvmesh::Vertex) == 32, "tube mesh vertex is 32 bytes");
    if (vertices) memcpy(vertices, m.vertices.data(), std::min<uint64_t>(vertices_cap, m.vertices.size()) * 32);
    if (indices) memcpy(indices, m.indices.data(), std::min<uint64_t>(triangles_cap, m.indices.size() / 3) * 12);
    if (n_vertices) *n_vertices = m.vertices.size();
    if (n_triangles) *n_triangles = m.indices.size() / 3;
    if (n_line_points) *n_line_points = m.line_pos.size();
    return LV_OK;
}

int lv_ao_bake(lv_ctx* c, lv_scene* s, uint32_t n_iterations, lv_stats* stats) {
    if (!c || !s) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_ao_bake: NULL argument");
    LV_CUDA(c, cudaSetDevice(c->device));
    int rc = reset_counters(c);
    if (rc) return rc;
    if (!s->has_lines) return fail(c, LV_ERR_STATE, "lv_ao_bake: no line frames attached (lv_scene_set_lines)");
    if ((rc = ensure_parametrization(c, s))) return rc;
    const uint32_t left = s->bake_done < c->opt.bake_iterations ? c->opt.bake_iterations - s->bake_done : 0u;
    const uint32_t todo = n_iterations ? std::min(n_iterations, left) : left;
    LV_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    float ms_rays = 0.0f;
    for (uint32_t i = 0; i < todo; i++) {
        if ((rc = run_bake_iteration(c, s))) return rc;
        if (stats) { LV_CUDA(c, cudaEventSynchronize(c->ev[5])); ms_rays += elapsed(c->ev[4], c->ev[5]); }
    }
    LV_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        Counters h;
        if ((rc = read_counters(c, h))) return rc;
        fill_stats(stats, h);
        stats->ms_rtao = elapsed(c->ev[0], c->ev[1]);
        stats->ms_rtao_rays = ms_rays;
        stats->ms_total = stats->ms_rtao;
    }
    return LV_OK;
}

int lv_ao_set_vertex_range(lv_scene* s, uint64_t first_vertex, uint64_t n_vertices) {
    if (!s) return LV_ERR_INVALID_ARGUMENT;
    s->bake_first = first_vertex; s->bake_count = n_vertices;
    return LV_OK;
}

int lv_ao_factors(lv_scene* s, float** device_factors, uint64_t* n_floats) {
    if (!s || !s->ctx || !device_factors) return LV_ERR_INVALID_ARGUMENT;
    lv_ctx* c = s->ctx;
    if (!s->has_lines) return fail(c, LV_ERR_STATE, "lv_ao_factors: no line frames attached (lv_scene_set_lines)");
    LV_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_parametrization(c, s);
    if (rc) return rc;
    *device_factors = s->factors.p;
    if (n_floats) *n_floats = uint64_t(s->n_param) * s->param_subdiv;
    return LV_OK;
}

int lv_ao_bake_reset(lv_scene* s) {
    if (!s) return LV_ERR_INVALID_ARGUMENT;
    s->bake_done = 0;
    return LV_OK;
}

int lv_ao_read(lv_scene* s, float* factors, size_t factors_cap, float* blending_weights, size_t weights_cap,
               float* sampling_locations, size_t sampling_cap, uint32_t* n_param_vertices, uint32_t* n_subdivisions,
               uint32_t* iterations_done) {
    if (!s || !s->ctx) return LV_ERR_INVALID_ARGUMENT;
    lv_ctx* c = s->ctx;
    LV_CUDA(c, cudaSetDevice(c->device));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    if (n_param_vertices) *n_param_vertices = s->n_param;
    if (n_subdivisions) *n_subdivisions = s->param_subdiv;
    if (iterations_done) *iterations_done = s->bake_done;
    if (factors && s->n_param) LV_CUDA(c, cudaMemcpy(factors, s->factors.p, std::min(factors_cap, size_t(s->n_param) * s->param_subdiv) * 4, cudaMemcpyDeviceToHost));
    if (blending_weights && s->n_param) LV_CUDA(c, cudaMemcpy(blending_weights, s->weights.p, std::min(weights_cap, size_t(s->n_pt)) * 4, cudaMemcpyDeviceToHost));
    if (sampling_locations && s->n_param) LV_CUDA(c, cudaMemcpy(sampling_locations, s->sampling.p, std::min(sampling_cap, size_t(s->n_param)) * 4, cudaMemcpyDeviceToHost));
    return LV_OK;
}

int lv_scene_info(const lv_scene* s, uint64_t* n_seg, uint64_t* n_nodes, float* build_ms, float* aabb) {
    if (!s) return LV_ERR_INVALID_ARGUMENT;
    if (n_seg) *n_seg = s->n_seg;
    if (n_nodes) *n_nodes = s->n_nodes;
    if (build_ms) *build_ms = s->build_ms;
    if (aabb) memcpy(aabb, s->bounds, 24);
    return LV_OK;
}

int lv_scene_copy_bvh(const lv_scene* s, void* nodes_out, size_t cap_bytes) {
    if (!s || !nodes_out) return LV_ERR_INVALID_ARGUMENT;
    size_t bytes = std::min(cap_bytes, size_t(s->n_nodes) * sizeof(Node64));
    if (bytes == 0) return LV_OK;
    LV_CUDA(s->ctx, cudaSetDevice(s->ctx->device));
    LV_CUDA(s->ctx, cudaMemcpy(nodes_out, s->nodes.p, bytes, cudaMemcpyDeviceToHost));
    return LV_OK;
}

// ---------------------------------------------------------------------------------------------- frames
int lv_trace_primary(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, lv_hit* hits_out, lv_stats* stats) {
    if (!c || !sc || !hits_out) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_trace_primary: NULL argument");
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, sc, cam, 0, P);
    if (rc) return rc;
    if ((rc = reset_counters(c))) return rc;
    const size_t npx = size_t(P.W) * P.H;
    lv_hit* dst = hits_out;
    const bool dev = is_device_pointer(hits_out);
    if (!dev) { LV_CUDA(c, c->hits.ensure(npx)); dst = c->hits.p; }
    LV_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    if (P.n_tiles) k_primary<<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, sc->dev(), dst, c->counters.p);
    LV_CUDA(c, cudaGetLastError());
    LV_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if (!dev && (rc = deliver(c, dst, hits_out, P.W, P.H, sizeof(lv_hit)))) return rc;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        Counters h;
        if ((rc = read_counters(c, h))) return rc;
        fill_stats(stats, h);
        stats->ms_trace = elapsed(c->ev[0], c->ev[1]);
        stats->ms_total = stats->ms_trace;
    }
    return LV_OK;
}

int lv_render_rtao(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, uint32_t frame_number, float* ao_out, lv_stats* stats) {
    if (!c || !sc) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_render_rtao: NULL argument");
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, sc, cam, frame_number, P);
    if (rc) return rc;
    if ((rc = reset_counters(c))) return rc;
    LV_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    if ((rc = run_rtao(c, sc, P, frame_number))) return rc;
    LV_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if (ao_out && (rc = deliver(c, c->ao.p, ao_out, P.W, P.H, sizeof(float)))) return rc;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        Counters h;
        if ((rc = read_counters(c, h))) return rc;
        fill_stats(stats, h);
        stats->ms_rtao = elapsed(c->ev[0], c->ev[1]);
        if (c->rtao_rays_timed) stats->ms_rtao_rays = elapsed(c->ev[4], c->ev[5]);
        stats->ms_total = stats->ms_rtao;
    }
    return LV_OK;
}

// the tube pass proper (k_tubes into the sink) and the frame's statistics: the tail of lv_render_tubes / lv_sao_finish
int tube_pass(lv_ctx* c, const lv_scene* sc, FrameParams& P, FrameSink& sink, void* rgba_out, const uint2* first, bool use_ao, lv_stats* stats) {
    int rc;
    float4* img = sink.image;
    if ((rc = run_depth_range(c, sc, P))) return rc;
    const bool tri_tubes = c->opt.tube_triangles && sc->n_seg;
    if (tri_tubes) {
        if (P.use_static_ao) return fail(c, LV_ERR_INVALID_ARGUMENT, "geometry_mode = 'Triangle Mesh' does not support ambient_occlusion_mode = 'RTAO (Prebaker)'");
        if ((rc = ensure_tube_mesh(c, const_cast<lv_scene*>(sc)))) return rc;
    }
    LV_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if (P.n_tiles && tri_tubes) k_tubes<false, 1><<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, sc->dev(), img, c->counters.p, sink.out8, nullptr);
    else if (P.n_tiles) {
        if (P.use_static_ao) k_tubes<true><<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, sc->dev(), img, c->counters.p, sink.out8, first);
        else k_tubes<false><<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, sc->dev(), img, c->counters.p, sink.out8, first);
    }
    LV_CUDA(c, cudaGetLastError());
    LV_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    if ((rc = close_sink(c, sink, rgba_out, P.W, P.H))) return rc;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        Counters h;
        if ((rc = read_counters(c, h))) return rc;
        fill_stats(stats, h);
        stats->ms_rtao = elapsed(c->ev[0], c->ev[1]);
        if (use_ao && c->rtao_rays_timed) stats->ms_rtao_rays = elapsed(c->ev[4], c->ev[5]);
        stats->ms_trace = elapsed(c->ev[1], c->ev[2]);
        stats->ms_total = elapsed(c->ev[0], c->ev[2]);
    }
    return LV_OK;
}

int lv_render_tubes(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, uint32_t frame_number, float* rgba_out, lv_stats* stats) {
    if (!c || !sc || !rgba_out) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_render_tubes: NULL argument");
    if (!c->tf.p) return fail(c, LV_ERR_STATE, "lv_render_tubes: no transfer function set (lv_set_transfer_function)");
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, sc, cam, frame_number, P);
    if (rc) return rc;
    if ((rc = reset_counters(c))) return rc;
    FrameSink sink;
    if ((rc = open_sink(c, rgba_out, P.W, P.H, sink))) return rc;
    float4* img = sink.image;
    LV_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    const bool use_ao = c->opt.ao_strength > 0.0f && !c->opt.ao_prebaker;
    if ((rc = prepare_static_ao(c, sc, P))) return rc;
    const uint2* first = nullptr;
    if (use_ao) {
        // LineRenderer::renderBase -> ambientOcclusionBaker->updateIterative (reference LineRenderer.cpp:259-265): one RTAO
        // iteration per rendered frame until maxNumAccumulatedFrames (VulkanRayTracedAmbientOcclusion.cpp:89-107).
        if (frame_number < c->opt.ao_iterations || !c->ao.p || c->ao_w != P.W || c->ao_h != P.H) {
            if (c->opt.tube_prepass && P.n_tiles && sc->n_seg && !c->opt.tube_triangles) {
                // fork: the tube pass's first-hit traversal on the second stream, beside the RTAO pass (joined in front of k_tubes)
                if (!c->aux_stream) {
                    LV_CUDA(c, cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
                    LV_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)); LV_CUDA(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
                }
                LV_CUDA(c, c->first_hits.ensure(size_t(P.W) * P.H));
                LV_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
                LV_CUDA(c, cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
                k_tube_first<<<pixel_grid(c, P), kBlockThreads, 0, c->aux_stream>>>(P, sc->dev(), c->first_hits.p, c->counters.p);
                LV_CUDA(c, cudaGetLastError());
                LV_CUDA(c, cudaEventRecord(c->ev_join, c->aux_stream));
                first = c->first_hits.p;
            }
            if ((rc = run_rtao(c, sc, P, frame_number))) return rc;
            if (first) LV_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        }
        P.use_ao = 1; P.ao_tex = c->ao.p;
    }
    return tube_pass(c, sc, P, sink, rgba_out, first, use_ao, stats);
}

// ---- AO-sample-batch shards (north_star's second shard axis) ----------------------------------------------------------
// The RTAO pass of a frame in three stages, so that N ranks can split its SAMPLES instead of (only) its pixels: every rank finds the
// hit pixels of its own tiles (stage 1), the hit lists are all-gathered by the caller (torch.distributed / NCCL, sharding.py::
// SampleShards), every rank traces samples [r spp / N, (r + 1) spp / N) of EVERY hit pixel of the frame (stage 2: perfectly balanced by
// construction -- the seeds are sample-indexed, VulkanRayTracedAmbientOcclusion.glsl:289-292), the per-sample results return to the
// pixel's owner (all-to-all), which sums them in sample order -- bit for bit the sum of a one-GPU frame -- and renders its tiles (stage 3).
int lv_sao_primary(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, uint32_t frame_number, const void** hits_device, uint32_t* n_hits) {
    if (!c || !sc || !hits_device || !n_hits) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_sao_primary: NULL argument");
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, sc, cam, frame_number, P);
    if (rc) return rc;
    if (!(c->opt.ao_strength > 0.0f) || c->opt.ao_prebaker) return fail(c, LV_ERR_STATE, "lv_sao_primary: screen-space RTAO is off");
    if ((rc = reset_counters(c))) return rc;
    LV_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    if ((rc = run_rtao(c, sc, P, frame_number, kRtaoPrimaryOnly))) return rc;
    unsigned int n = 0;
    LV_CUDA(c, cudaMemcpyAsync(&n, c->small.p, 4, cudaMemcpyDeviceToHost, c->stream));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    *hits_device = c->ao_hits.p; *n_hits = n;
    return LV_OK;
}

int lv_sao_trace(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, uint32_t frame_number, const void* hits_device, uint32_t n_hits,
                 uint32_t sample_first, uint32_t sample_count, float* occ_device) {
    if (!c || !sc || !occ_device || (!hits_device && n_hits)) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_sao_trace: NULL argument");
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, sc, cam, frame_number, P);
    if (rc) return rc;
    if (sample_count == 0 || sample_first + sample_count > P.ao_spp || P.ao_spp % sample_count || sample_first % sample_count)
        return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_sao_trace: the sample range must be one of spp / count equal batches");
    P.frame_number = frame_number; P.ao_spp_local = sample_count; P.ao_sample_first = sample_first;
    if (c->opt.ao_wide) { if ((rc = ensure_wnodes(c, const_cast<lv_scene*>(sc)))) return rc; const_cast<lv_scene*>(sc)->set_w_top(c->opt.ao_wide_top); }
    LV_CUDA(c, c->small2.ensure(4));
    const unsigned int init[4] = {n_hits, 0u, 0u, 0u};
    LV_CUDA(c, cudaMemcpyAsync(c->small2.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    if (n_hits == 0) return LV_OK;
    if ((rc = launch_ao_rays<false>(c, P, sc->dev(), sc->leaf_size == 1, (unsigned long long)n_hits * sample_count, false,
                                    static_cast<const AoHit*>(hits_device), c->small2.p, occ_device))) return rc;
    c->rtao_rays_timed = true;
    LV_CUDA(c, cudaGetLastError());
    return LV_OK;
}

int lv_sao_finish(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, uint32_t frame_number, const float* occ_parts, uint32_t n_parts,
                  float* rgba_out, lv_stats* stats) {
    if (!c || !sc || !rgba_out || !occ_parts) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_sao_finish: NULL argument");
    if (!c->tf.p) return fail(c, LV_ERR_STATE, "lv_sao_finish: no transfer function set (lv_set_transfer_function)");
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, sc, cam, frame_number, P);
    if (rc) return rc;
    if (n_parts == 0 || P.ao_spp % n_parts) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_sao_finish: spp must be a multiple of the number of parts");
    if (!c->ao.p || c->ao_w != P.W || c->ao_h != P.H || !c->small.p) return fail(c, LV_ERR_STATE, "lv_sao_finish: lv_sao_primary has not run for this frame size");
    FrameSink sink;
    if ((rc = open_sink(c, rgba_out, P.W, P.H, sink))) return rc;
    if ((rc = prepare_static_ao(c, sc, P))) return rc;
    P.frame_number = frame_number; P.ao_spp_local = P.ao_spp / n_parts;
    k_rtao_reduce<<<c->num_sms * 4, 256, 0, c->stream>>>(P, occ_parts, c->ao_hits.p, c->small.p, c->ao.p);
    LV_CUDA(c, cudaGetLastError());
    P.ao_spp_local = P.ao_spp;
    P.use_ao = 1; P.ao_tex = c->ao.p;
    return tube_pass(c, sc, P, sink, rgba_out, nullptr, true, stats);
}

int lv_ppll_clear(lv_ctx* c, const lv_camera* cam, uint64_t linked_list_size) {
    if (!c) return LV_ERR_INVALID_ARGUMENT;
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, nullptr, cam, 0, P);
    if (rc) return rc;
    if ((rc = ppll_prepare(c, nullptr, P, linked_list_size))) return rc;
    const size_t npad = size_t(P.padded_w) * P.padded_h;
    // LinkedListClear.glsl:50 (startOffset = -1) + fragmentCounterBuffer->fill(0) (PerPixelLinkedListLineRenderer.cpp:431)
    LV_CUDA(c, cudaMemsetAsync(c->heads.p, 0xFF, npad * 4, c->stream));
    LV_CUDA(c, cudaMemsetAsync(c->counts.p, 0, npad * 4, c->stream));
    LV_CUDA(c, cudaMemsetAsync(c->frag_counter.p, 0, 8, c->stream));
    return LV_OK;
}

int lv_ppll_gather(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, lv_stats* stats) {
    if (!c || !sc) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_ppll_gather: NULL argument");
    if (!c->tf.p) return fail(c, LV_ERR_STATE, "lv_ppll_gather: no transfer function set");
    if (!c->heads.p || (!c->nodes.p && c->list_size)) return fail(c, LV_ERR_STATE, "lv_ppll_gather: call lv_ppll_clear first");
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, sc, cam, 0, P);
    if (rc) return rc;
    if (P.padded_w != c->padded_w || P.padded_h != c->padded_h) return fail(c, LV_ERR_STATE, "lv_ppll_gather: resolution changed since lv_ppll_clear");
    if ((rc = reset_counters(c))) return rc;
    if ((rc = prepare_static_ao(c, sc, P))) return rc;
    if ((rc = run_depth_range(c, sc, P))) return rc;
    LV_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    if (P.n_tiles && c->opt.ppll_raster_gather && sc->n_seg && P.W < 65536u && P.H < 65536u) {   // object order: one warp per segment (candidate pixels are queued as x | y << 16)
        const unsigned char* owned = c->world > 1 ? c->owned_map.p : nullptr;
        unsigned long long n_pixels = 0;
        for (const uint2& t : c->tiles_host)
            n_pixels += (unsigned long long)std::min(c->tile_size, P.W - t.x * c->tile_size) * std::min(c->tile_size, P.H - t.y * c->tile_size);
        LV_CUDA(c, c->small.ensure(4));
        LV_CUDA(c, cudaMemsetAsync(c->small.p, 0, 4 * sizeof(unsigned int), c->stream));
        unsigned long long* work = reinterpret_cast<unsigned long long*>(c->small.p + 2);
        int per_sm = 0;
        const bool contig = c->opt.ppll_contiguous;
        if (contig) LV_CUDA(c, c->stage.ensure(c->list_size));
        lv_ppll_node* dst = contig ? c->stage.p : c->nodes.p;
        // the pixel-centre camera rays of this frame, once per pixel (the gather tests every pixel against many segments)
        LV_CUDA(c, c->pixel_rays.ensure(2 * size_t(P.W) * P.H));
        k_pixel_rays<<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, c->pixel_rays.p);
#define LV_RASTER(SAOV, STAGEV, MB) do { \
            LV_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ppll_gather_raster<SAOV, STAGEV, MB>, kBlockThreads, 0)); \
            k_ppll_gather_raster<SAOV, STAGEV, MB><<<uint32_t(std::max(1, per_sm) * c->num_sms), kBlockThreads, 0, c->stream>>>( \
                P, sc->dev(), c->heads.p, c->counts.p, dst, c->frag_counter.p, c->list_size, c->counters.p, work, owned, c->tiles_x, n_pixels, c->pixel_rays.p); } while (0)
        if (P.use_static_ao) { if (contig) LV_RASTER(true, true, 4); else LV_RASTER(true, false, 4); }
        else if (contig) LV_RASTER(false, true, 4);
        else if (c->opt.ppll_raster_min_blocks == 6) LV_RASTER(false, false, 6);
        else if (c->opt.ppll_raster_min_blocks == 5) LV_RASTER(false, false, 5);
        else LV_RASTER(false, false, 4);
#undef LV_RASTER
        if (contig) {   // counts -> offsets (exclusive scan) -> every pixel's fragments into one contiguous run of the node buffer
            const size_t npad = size_t(P.padded_w) * P.padded_h;
            LV_CUDA(c, c->list_offs.ensure(npad)); LV_CUDA(c, c->fill_cursor.ensure(npad));
            LV_CUDA(c, cudaMemsetAsync(c->fill_cursor.p, 0, npad * 4, c->stream));
            size_t tmp_bytes = 0;
            LV_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->counts.p, c->list_offs.p, int(npad), c->stream));
            LV_CUDA(c, c->scan_tmp.ensure(tmp_bytes + 16));
            LV_CUDA(c, cub::DeviceScan::ExclusiveSum(c->scan_tmp.p, tmp_bytes, c->counts.p, c->list_offs.p, int(npad), c->stream));
            k_ppll_fill<<<uint32_t(c->num_sms) * 8u, 256, 0, c->stream>>>(c->stage.p, c->frag_counter.p, c->list_size, c->list_offs.p, c->fill_cursor.p,
                                                                     c->counts.p, c->heads.p, c->nodes.p);
        }
        c->lists_contiguous = contig;
    } else if (P.n_tiles) {
        c->lists_contiguous = false;
        if (P.use_static_ao)
            k_ppll_gather<true><<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, sc->dev(), c->heads.p, c->counts.p, c->nodes.p, c->frag_counter.p, c->list_size, c->counters.p);
        else
            k_ppll_gather<false><<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, sc->dev(), c->heads.p, c->counts.p, c->nodes.p, c->frag_counter.p, c->list_size, c->counters.p);
    }
    LV_CUDA(c, cudaGetLastError());
    LV_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        Counters h;
        if ((rc = read_counters(c, h))) return rc;
        fill_stats(stats, h);
        stats->frags_stored = std::min<uint64_t>(h.frags_generated, c->list_size);
        stats->frags_dropped = h.frags_generated - stats->frags_stored;
        stats->ms_gather = elapsed(c->ev[0], c->ev[1]);
        stats->ms_total = stats->ms_gather;
    }
    return LV_OK;
}

int lv_ppll_resolve(lv_ctx* c, const lv_camera* cam, uint32_t max_frags, uint32_t sort_mode, float* rgba_out, lv_stats* stats) {
    if (!c || !rgba_out) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_ppll_resolve: NULL argument");
    if (!c->heads.p) return fail(c, LV_ERR_STATE, "lv_ppll_resolve: no gathered fragments");
    if (max_frags == 0 || max_frags > uint32_t(kResolveCap)) return fail(c, LV_ERR_INVALID_ARGUMENT, "max_frags must be in [1, " + std::to_string(kResolveCap) + "]");
    if (sort_mode > LV_SORT_QUICKSORT_HYBRID) return fail(c, LV_ERR_INVALID_ARGUMENT, "unknown sort_mode");
    LV_CUDA(c, cudaSetDevice(c->device));
    FrameParams P;
    int rc = make_params(c, nullptr, cam, 0, P);
    if (rc) return rc;
    if (P.padded_w != c->padded_w || P.padded_h != c->padded_h) return fail(c, LV_ERR_STATE, "lv_ppll_resolve: resolution changed since lv_ppll_clear");
    if ((rc = reset_counters(c))) return rc;
    FrameSink sink;
    if ((rc = open_sink(c, rgba_out, P.W, P.H, sink))) return rc;
    float4* img = sink.image;
    uint32_t* out8 = c->opt.ppll_binned_resolve ? nullptr : sink.out8;   // the binned kernels write floats; converted below
    LV_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    // All eight modes produce the depth-sorted order; only the priority queue stops blending at alpha >= 0.99
    // (reference LinkedListSort.glsl:217).  See DESIGN.md for the reference's bitonicSort defect.
    const int early_out = (sort_mode == LV_SORT_PRIORITY_QUEUE) ? 1 : 0;
    if (P.n_tiles && !c->opt.ppll_binned_resolve) {
        // shared key tile per warp: the smallest of 256 / 512 / 1024 that is >= the option and can hold the longest list
        const uint32_t cap = std::max(c->opt.ppll_resolve_tile, max_frags) <= 256u ? 256u : (std::max(c->opt.ppll_resolve_tile, max_frags) <= 512u ? 512u : 1024u);
        const uint32_t grid = pixel_grid(c, P);
#define LV_RESOLVE(RS, CAPV, CT) k_ppll_resolve<RS, CAPV, CT><<<grid, kBlockThreads, 0, c->stream>>>(P, c->heads.p, c->counts.p, c->nodes.p, max_frags, early_out, img, c->counters.p, nullptr, nullptr, out8)
#define LV_RESOLVE_CAP(RS, CT) do { if (cap == 256u) LV_RESOLVE(RS, 256, CT); else if (cap == 512u) LV_RESOLVE(RS, 512, CT); else LV_RESOLVE(RS, 1024, CT); } while (0)
        if (c->lists_contiguous) { if (c->opt.ppll_reg_sort) LV_RESOLVE_CAP(true, true); else LV_RESOLVE_CAP(false, true); }
        else { if (c->opt.ppll_reg_sort) LV_RESOLVE_CAP(true, false); else LV_RESOLVE_CAP(false, false); }
#undef LV_RESOLVE_CAP
#undef LV_RESOLVE
    } else if (P.n_tiles) {
        const size_t n_own = size_t(P.n_tiles) * c->tile_size * c->tile_size;
        LV_CUDA(c, c->bin_hist.ensure(2 * (size_t(kResolveCap) + 2) + 8));   // hist | offsets | n_sorted[kBinClasses]
        LV_CUDA(c, c->bin_order.ensure(n_own));
        unsigned int* hist = c->bin_hist.p; unsigned int* offs = hist + kResolveCap + 2; unsigned int* nsort = offs + kResolveCap + 2;
        LV_CUDA(c, cudaMemsetAsync(hist, 0, (2 * (size_t(kResolveCap) + 2) + 8) * sizeof(unsigned int), c->stream));
        const uint32_t grid = pixel_grid(c, P);
        k_ppll_bin_count<<<grid, kBlockThreads, 0, c->stream>>>(P, c->counts.p, max_frags, hist);
        k_ppll_bin_scan<<<1, 32, 0, c->stream>>>(max_frags, hist, offs, nsort);
        k_ppll_bin_scatter<<<grid, kBlockThreads, 0, c->stream>>>(P, c->counts.p, max_frags, offs, c->bin_order.p, img);
        // lists longer than 256 keys: cooperative kernel over the head of the order array; the launch covers the worst case,
        // surplus blocks exit on n_sorted[0]
        if (max_frags > 256u)
            k_ppll_resolve<false><<<uint32_t((n_own + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, c->stream>>>(
                P, c->heads.p, c->counts.p, c->nodes.p, max_frags, early_out, img, c->counters.p, c->bin_order.p, nsort, nullptr);
        auto smem = [](int maxn, int warps) { return size_t(warps) * maxn * 32 * 8 + 256 * 4; };
        if (!c->binned_attr_set) {
            LV_CUDA(c, cudaFuncSetAttribute(k_ppll_resolve_binned<256, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem(256, 1))));
            c->binned_attr_set = true;
        }
        const uint32_t pg = uint32_t(c->num_sms) * 6u;
        if (max_frags > 128u) k_ppll_resolve_binned<256, 1, 1><<<pg, 32, smem(256, 1), c->stream>>>(P, c->heads.p, c->counts.p, c->nodes.p, c->bin_order.p, nsort, max_frags, early_out, img, c->counters.p);
        if (max_frags > 64u) k_ppll_resolve_binned<128, 1, 2><<<pg, 32, smem(128, 1), c->stream>>>(P, c->heads.p, c->counts.p, c->nodes.p, c->bin_order.p, nsort, max_frags, early_out, img, c->counters.p);
        if (max_frags > 32u) k_ppll_resolve_binned<64, 2, 3><<<pg, 64, smem(64, 2), c->stream>>>(P, c->heads.p, c->counts.p, c->nodes.p, c->bin_order.p, nsort, max_frags, early_out, img, c->counters.p);
        k_ppll_resolve_binned<32, 4, 4><<<pg, 128, smem(32, 4), c->stream>>>(P, c->heads.p, c->counts.p, c->nodes.p, c->bin_order.p, nsort, max_frags, early_out, img, c->counters.p);
    }
    if (P.n_tiles && sink.out8 && !out8) k_frame_to_rgba8<<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, img, sink.out8);   // binned resolve + rgba8
    LV_CUDA(c, cudaGetLastError());
    LV_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if ((rc = close_sink(c, sink, rgba_out, P.W, P.H))) return rc;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        Counters h;
        if ((rc = read_counters(c, h))) return rc;
        fill_stats(stats, h);
        stats->ms_resolve = elapsed(c->ev[0], c->ev[1]);
        stats->ms_total = stats->ms_resolve;
    }
    return LV_OK;
}

int lv_render_ppll(lv_ctx* c, const lv_scene* sc, const lv_camera* cam, uint32_t max_frags, uint32_t sort_mode,
                   uint64_t linked_list_size, float* rgba_out, lv_stats* stats) {
    if (!c || !sc || !rgba_out) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_render_ppll: NULL argument");
    LV_CUDA(c, cudaSetDevice(c->device));
    lv_stats g{}, r{};
    LV_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    FrameParams P;
    int rc = make_params(c, sc, cam, 0, P);
    if (rc) return rc;
    if ((rc = ppll_prepare(c, sc, P, linked_list_size))) return rc;
    if ((rc = lv_ppll_clear(c, cam, c->list_size))) return rc;
    LV_CUDA(c, cudaEventRecord(c->ev[7], c->stream));
    if ((rc = lv_ppll_gather(c, sc, cam, stats ? &g : nullptr))) return rc;
    if ((rc = lv_ppll_resolve(c, cam, max_frags, sort_mode, rgba_out, stats ? &r : nullptr))) return rc;
    if (stats) {
        *stats = g;
        stats->frags_sorted = r.frags_sorted; stats->frags_truncated = r.frags_truncated; stats->max_depth_complexity = r.max_depth_complexity;
        stats->ms_resolve = r.ms_resolve;
        stats->ms_clear = elapsed(c->ev[6], c->ev[7]);
        stats->ms_total = stats->ms_clear + stats->ms_gather + stats->ms_resolve;
    }
    return LV_OK;
}

int lv_frame_to_rgba8(lv_ctx* c, const float* rgba_device, uint32_t W, uint32_t H, uint32_t* rgba8_out) {
    if (!c || !rgba_device || !rgba8_out || W == 0 || H == 0) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_frame_to_rgba8: bad argument");
    if (!is_device_pointer(rgba_device) || (reinterpret_cast<uintptr_t>(rgba_device) & 15)) return fail(c, LV_ERR_INVALID_ARGUMENT, "lv_frame_to_rgba8: the float frame must be a 16-byte aligned device pointer");
    LV_CUDA(c, cudaSetDevice(c->device));
    lv_camera cam;
    memset(&cam, 0, sizeof(cam));
    cam.width = W; cam.height = H;
    FrameParams P;
    int rc = make_params(c, nullptr, &cam, 0, P);
    if (rc) return rc;
    const bool dev = is_device_pointer(rgba8_out);
    uint32_t* dst = rgba8_out;
    if (!dev) { LV_CUDA(c, c->rgba8.ensure(size_t(W) * H)); dst = c->rgba8.p; }
    if (P.n_tiles) k_frame_to_rgba8<<<pixel_grid(c, P), kBlockThreads, 0, c->stream>>>(P, reinterpret_cast<const float4*>(rgba_device), dst);
    LV_CUDA(c, cudaGetLastError());
    if (!dev && (rc = deliver(c, dst, rgba8_out, W, H, 4))) return rc;
    return LV_OK;
}

int lv_ppll_read(lv_ctx* c, uint32_t* frag_counter, uint32_t* start_offset, size_t start_offset_cap, lv_ppll_node* nodes,
                 size_t nodes_cap, uint32_t* padded_w, uint32_t* padded_h) {
    if (!c || !c->heads.p) return fail(c, LV_ERR_STATE, "lv_ppll_read: nothing gathered");
    LV_CUDA(c, cudaSetDevice(c->device));
    LV_CUDA(c, cudaStreamSynchronize(c->stream));
    unsigned long long cnt = 0;
    LV_CUDA(c, cudaMemcpy(&cnt, c->frag_counter.p, 8, cudaMemcpyDeviceToHost));
    if (frag_counter) *frag_counter = uint32_t(std::min<unsigned long long>(cnt, 0xFFFFFFFFull));
    if (padded_w) *padded_w = c->padded_w;
    if (padded_h) *padded_h = c->padded_h;
    if (start_offset) {
        size_t n = std::min(start_offset_cap, size_t(c->padded_w) * c->padded_h);
        LV_CUDA(c, cudaMemcpy(start_offset, c->heads.p, n * 4, cudaMemcpyDeviceToHost));
    }
    if (nodes) {
        size_t n = std::min<size_t>(nodes_cap, size_t(std::min<unsigned long long>(cnt, c->list_size)));
        if (n) LV_CUDA(c, cudaMemcpy(nodes, c->nodes.p, n * sizeof(lv_ppll_node), cudaMemcpyDeviceToHost));
    }
    return LV_OK;
}

}  // extern "C"
