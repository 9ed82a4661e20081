// lv_types.cuh -- device data layout (see DESIGN.md "Data layout in HBM").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/linevis_b200.h"

namespace lv {

// 32-byte packed segment record, stored in BVH (Morton) order.  Replaces two dependent 48-byte
// LinePointDataUnified gathers (reference src/LineData/LineRenderData.hpp:99-106) by one 32-byte load.
struct __align__(32) SegRec {
    float4 a;  // p0.xyz, attr0
    float4 b;  // p1.xyz, attr1
};

// 64-byte child-pair node: both children's boxes + references in one 64-byte (2-sector) fetch.
// child word: leaf  = bit31 | (count-1) << 27 | first record (BVH order), 1 <= count <= 16, record index < 2^27;
//             inner = node index.  `count` is repeated in the max.w lane (0 for inner).  An absent child is the point
//             box (+inf, +inf, +inf), which the canonical slab test can never hit.
struct __align__(64) Node64 {
    float4 l0;  // lmin.xyz, as_float(left child word)
    float4 l1;  // lmax.xyz, as_float(lcount)
    float4 r0;  // rmin.xyz, as_float(right child word)
    float4 r1;  // rmax.xyz, as_float(rcount)
};

struct SceneDev {
    const SegRec* segs;        // [n_seg] BVH order
    const uint32_t* prim_ids;  // [n_seg] BVH order -> caller's segment index
    const Node64* nodes;       // [n_nodes], root = 0
    uint32_t n_seg;
    uint32_t n_nodes;
    float radius;              // lineWidth * 0.5
    float line_width;
};

// Everything a frame kernel needs; passed by value as a __grid_constant__ parameter.
struct FrameParams {
    // LineUniformData (reference src/LineData/LineData.hpp:428-464)
    float view[16], proj[16], inv_view[16], inv_proj[16];
    float cam_pos[3];
    float fov_y;
    float bg[4], fg[4];
    uint32_t W, H;
    float line_width;
    // shader defines / settings
    int use_capped, use_halos, use_ao;
    float ao_strength, ao_gamma, ao_radius;
    int use_depth_cues;            // USE_DEPTH_CUES (depth_cue_strength > 0)
    float depth_cue_strength;
    const float* depth_min_max;    // device: {minDepth, maxDepth} of this frame (k_depth_range)
    float near_dist, far_dist;
    uint32_t ao_spp;
    int ao_use_distance, ao_jitter;
    float subdiv_corr;        // cos(pi / tubeNumSubdivisions)
    int ao_refill_below;      // k_rtao_rays refills a warp once fewer lanes than this are live
    int ao_leaf_vote;         // ... and intersects postponed leaves once this many lanes hold one
    uint32_t spp;             // numSamplesPerFrame
    int use_jitter, det_sampling;
    uint32_t max_depth;       // maxDepthComplexity
    uint32_t frame_number;
    // transfer function
    const float4* tf;
    uint32_t tfK;
    float amin, amax;
    // AO texture (W*H floats) or nullptr
    const float* ao_tex;
    // owned image tiles (Morton order); tile_size x tile_size pixels each
    const uint2* tiles;
    uint32_t n_tiles, tile_size;
    unsigned int* apron_marks;   // W*H stamps, only in tile-sharded + jittered mode (see k_rtao_primary)
    // PPLL addressing (reference Data/Shaders/Utils/TiledAddress.glsl)
    uint32_t padded_w, padded_h, addr_tw, addr_th;
};

// per-launch counters accumulated with one atomic per warp
struct Counters {
    unsigned long long rays_primary, rays_ao, steps, isect, ao_steps, ao_isect, pixels_hit, ao_pixels_hit;
    unsigned long long frags_generated, frags_sorted, frags_truncated;
    unsigned int max_depth_complexity, pad;
};

}  // namespace lv
