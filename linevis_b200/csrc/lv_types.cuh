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

// Line-point data a hit needs besides position / attribute when the prebaked (object-space) AO is looked up: the two
// line-point indices (linePointIndices, reference TubeRayTracing.glsl:513) and the line normals at both points (:551).
// One 32-byte record per segment, BVH order like SegRec; only fetched while shading in "RTAO (Prebaker)" mode.
struct __align__(32) SegAux {
    float4 n0;  // lineNormal(point 0).xyz, as_float(point index 0)
    float4 n1;  // lineNormal(point 1).xyz, as_float(point index 1)
};

// Start frame of a batch of AO rays: what the screen-space RTAO pass keeps per hit pixel and the prebaker per
// (parametrization vertex, tube subdivision).  k_rtao_rays turns (record, sample) into a ray.
struct __align__(16) AoHit {
    float4 pos_off;   // position; RTAO: AO ray origin offset |linePos - pos| / cos(pi/N) (baker: origin is the position itself)
    float4 nrm_px;    // surface normal, as_float(output index: pixel y*W + x, or subdivision + N * vertex)
    float4 tng;       // surface tangent; baker: as_float(LCG state of this record's first random number)
};

// 48-byte triangle record of the triangle-tube mode (lv_tri.cuh): three vertex positions, the mesh's vertex indices in the w
// lanes (as_float), stored in BVH (Morton) order.
struct __align__(16) TriRec {
    float4 a, b, c;
};

// 32-byte QUANTISED child-pair node for the AO ray stream (b200_ao_qnodes, experimental): both child boxes on a 16-bit grid over the
// scene bounds, so that one 256-bit load brings a whole node (the stream is bound by L1 wavefronts: two scattered 32-byte
// requests per lane-step with Node64).  w[0] = lmin.x | lmin.y << 16, w[1] = lmin.z | lmax.x << 16, w[2] = lmax.y | lmax.z << 16,
// w[3] = left child word (kAbsentChild if there is none); w[4..7] the same for the right child.  Bounds are dequantised as
// fma(q, scale, origin) and rounded OUTWARD at build time against exactly that expression, so a quantised box always encloses
// the exact one: with the canonical slab test an enclosing box is never missed when the enclosed one is hit (rule 2), and the
// accepted set -- decided by the record's own exact AABB at the leaf -- does not change.
struct __align__(32) NodeQ {
    uint32_t w[8];
};
constexpr uint32_t kAbsentChild = 0x7FFFFFFDu;

// 64-byte 4-WIDE quantised node for the AO ray stream (b200_ao_wide): a binary node collapsed with the children of its largest
// children (lv_bvh.cuh: k_w4_round), four child boxes on the 16-bit grid of the scene bounds + four child words -- ONE 64-byte fetch
// (two 256-bit loads) tests four boxes, the number of dependent node fetches per ray roughly halves.  Child k: w[3 k + a] = lo_a |
// hi_a << 16 for the axes a = 0, 1, 2 (the two bounds of an axis share a word: one byte-permute both picks the bound that is NEAR for
// the ray's direction sign and builds the float 2^23 + q, see w4_dequant); w[12 + k] = child word (leaf bit | record, wide-node index,
// or kAbsentChild).  Bounds are rounded OUTWARD at build time against exactly the traversal's dequantisation, like NodeQ: a quantised
// box always encloses the exact one, the accepted set is decided by the record's own exact AABB in the leaf batch (rule 2).
struct __align__(64) NodeW4 {
    uint32_t w[16];
};

struct SceneDev {
    const SegRec* segs;        // [n_seg] BVH order
    const uint32_t* prim_ids;  // [n_seg] BVH order -> caller's segment index
    const float4* seg_axes;    // [n_seg + 1] BVH order: normalize(p1 - p0) of the record (cylinder_hit's axis), computed once at scene creation
    const Node64* nodes;       // [n_nodes], root = 0
    const SegAux* seg_aux;     // [n_seg] BVH order, or nullptr (no line frames attached)
    const NodeQ* qnodes;       // [n_nodes] quantised copy of `nodes` (b200_ao_qnodes), or nullptr
    float q_origin[3], q_scale[3];
    const NodeW4* wnodes;      // 4-wide quantised tree (b200_ao_wide), root = 0, breadth-first order; or nullptr
    float w_origin[3], w_scale[3];   // w4_dequant's constants (origin already shifted by the conversion's magic number)
    alignas(8) float w_pk[8];        // the same as aligned pairs for the packed fp32x2 box tests: {sx, sy}, {ox, oy}, {sz, sz}, {oz, oz}
    uint32_t w_tq_bits;        // k_rtao_rays_w: 7 if every wide node / record index (incl. the dummy record) is below 2^24, else 4
    uint32_t w_top;            // the first w_top wide nodes are whole top levels (staged into shared memory by the AO ray stream)
    // triangle-tube mode of the AO passes (lv_tri.cuh); all nullptr / 0 unless the tube mesh has been built
    const TriRec* tris;        // [n_tri] BVH order
    const uint32_t* tri_ids;   // [n_tri] BVH order -> triangle index of the mesh
    const Node64* tri_nodes;   // BVH over the triangles, root = 0
    const float4* tri_vattr;   // per mesh vertex: normal.xyz, as_float(line point index)
    const float4* tri_line_pos;  // per line point of the mesh: linePosition
    const float4* tri_line_tan;  // ... lineTangent
    uint32_t n_tri;
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
    uint32_t ao_spp_local, ao_sample_first;   // AO-sample-batch shards: this launch traces samples [first, first + local) of every record (default: all)
    int ao_use_distance, ao_jitter;
    float subdiv_corr;        // cos(pi / tubeNumSubdivisions)
    int ao_refill_below;      // k_rtao_rays refills a warp once fewer lanes than this are live
    int ao_leaf_vote;         // ... and intersects postponed leaves once this many lanes hold one
    int ao_wide_reps;         // wide-tree stream: node steps per pass of the traversal loop
    float ao_tq_lo, ao_tq_hi; // k_rtao_rays_w: (1 / ao_radius) (1 -+ 2^-12), the scales of its 4-bit stack entry distances
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
    // prebaked object-space AO (STATIC_AMBIENT_OCCLUSION_PREBAKING, reference Utils/AmbientOcclusion.glsl:30-75)
    int use_static_ao;
    const float* sao_factors;          // [n_param_vertices * n_ao_subdiv]
    const float* sao_weights;          // [n_line_vertices] blending weight parametrization
    uint32_t n_ao_subdiv, n_line_vertices, n_param_vertices;
    // owned image tiles (Morton order); tile_size x tile_size pixels each
    const uint2* tiles;
    uint32_t n_tiles, tile_size;
    unsigned int* apron_marks;   // W*H stamps, only in tile-sharded + jittered mode (see k_rtao_primary)
    // PPLL addressing (reference Data/Shaders/Utils/TiledAddress.glsl)
    uint32_t padded_w, padded_h, addr_tw, addr_th;
    uint32_t addr_tw_log2, addr_th_log2;   // the tile sizes are powers of two
    float raster_slack;                    // object-order gather: pixels added to the projected radius bound in the 2-D cull
};

// per-launch counters accumulated with one atomic per warp
struct Counters {
    unsigned long long rays_primary, rays_ao, steps, isect, ao_steps, ao_isect, pixels_hit, ao_pixels_hit;
    unsigned long long frags_generated, frags_sorted, frags_truncated;
    unsigned int max_depth_complexity, pad;
};

}  // namespace lv
