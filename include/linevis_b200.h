/*
 * linevis_b200.h -- C ABI of the B200-native dense-line renderer (drop-in for LineVis's hot path).
 *
 * This is the only boundary between host code (LineVis's C++ LineRenderer adapter, the Python
 * test/bench driver, or any other FFI) and the sm_100a CUDA implementation.  Plain pointers and
 * sizes only; no C++/torch types.  Every entry point returns an lv_status (0 == LV_OK, < 0 error);
 * the message of the last error on a context is available from lv_last_error().
 *
 * Each entry point cites the reference interface (path:line under the LineVis tree) it replaces.
 *
 * Threading: a context is NOT re-entrant.  One context per GPU.  All work is enqueued on the
 * CUDA stream handed to lv_ctx_create(); calls that return data to HOST pointers synchronise that
 * stream before returning, calls that write DEVICE pointers do not.
 *
 * Conventions
 *   - matrices are column-major float[16] exactly like glm::mat4 (m[col*4+row]);
 *   - images are row-major, pixel (x, y) at index y*W + x, row 0 == launch row 0 of the
 *     reference's ray-gen shader (ndc.y == -1 side);
 *   - colours are linear float RGBA unless a function says "rgba8".
 */
#ifndef LINEVIS_B200_H
#define LINEVIS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LV_ABI_VERSION 1

typedef enum lv_status {
    LV_OK = 0,
    LV_ERR_INVALID_ARGUMENT = -1,
    LV_ERR_CUDA = -2,
    LV_ERR_OUT_OF_MEMORY = -3,
    LV_ERR_NO_DEVICE = -4,
    LV_ERR_UNKNOWN_OPTION = -5,
    LV_ERR_STATE = -6
} lv_status;

typedef struct lv_ctx lv_ctx;     /* per-GPU renderer state (settings, TF LUT, frame buffers) */
typedef struct lv_scene lv_scene; /* device-resident segment soup + BVH */

/*
 * Camera + frame description.  Mirrors the camera part of LineData::LineUniformData
 * (src/LineData/LineData.hpp:428-464, filled at src/LineData/LineData.cpp:1275-1319) and the
 * RTAO uniform block (Data/Shaders/AO/RTAO/VulkanRayTracedAmbientOcclusion.glsl:43-69).
 */
typedef struct lv_camera {
    float view[16];             /* viewMatrix */
    float proj[16];             /* projectionMatrix */
    float inv_view[16];         /* inverseViewMatrix */
    float inv_proj[16];         /* inverseProjectionMatrix */
    float position[3];          /* cameraPosition (world) */
    float fov_y;                /* fieldOfViewY, radians */
    float background[4];        /* backgroundColor; foregroundColor = 1 - background (LineData.cpp:1284-1285) */
    uint32_t width, height;     /* viewportSize */
    float near_dist, far_dist;  /* camera near / far clip distance (depth cues: DepthCues/ComputeDepthValues.glsl:39-40) */
} lv_camera;

/*
 * Per-frame counters.  Overflow / truncation never aborts (reference drops silently,
 * Data/Shaders/Renderers/PPLL/LinkedListGather.glsl:53); it is reported here instead.
 */
typedef struct lv_stats {
    uint64_t rays_primary;          /* BVH traversals started for camera rays */
    uint64_t rays_ao;               /* BVH traversals started for AO rays */
    uint64_t traversal_steps;       /* T: child-pair node visits (64 B each), all rays of the call */
    uint64_t intersections;         /* I: segment records tested (32 B each), all rays of the call */
    uint64_t ao_traversal_steps;    /* T of the AO rays alone (the RTAO ray kernel, the dominant one) */
    uint64_t ao_intersections;      /* I of the AO rays alone */
    uint64_t pixels_hit;            /* pixels whose primary ray hit a tube */
    uint64_t frags_generated;       /* fragments offered to the PPLL gather (alpha >= 0.001) */
    uint64_t frags_stored;          /* min(generated, linkedListSize) */
    uint64_t frags_dropped;         /* generated - stored (fragment buffer overflow) */
    uint64_t frags_sorted;          /* nodes consumed by resolve: sum_px min(len_px, max_frags) */
    uint64_t frags_truncated;       /* nodes beyond max_frags skipped by resolve */
    uint32_t max_depth_complexity;  /* longest per-pixel list */
    uint32_t reserved;
    float ms_trace;                 /* CUDA-event time of the tube primary+shade kernel */
    float ms_rtao;                  /* ... of the RTAO pass (primary + AO ray kernels) */
    float ms_rtao_rays;             /* ... of the AO ray kernel alone */
    float ms_clear, ms_gather, ms_resolve; /* PPLL stages (PerPixelLinkedListLineRenderer.cpp:411-420) */
    float ms_total;
} lv_stats;

/* Sorting algorithms of the resolve pass, same order as SortingAlgorithmMode (src/Renderers/PPLL.hpp:39-48). */
typedef enum lv_sort_mode {
    LV_SORT_PRIORITY_QUEUE = 0,
    LV_SORT_BUBBLE = 1,
    LV_SORT_INSERTION = 2,
    LV_SORT_SHELL = 3,
    LV_SORT_MAX_HEAP = 4,
    LV_SORT_BITONIC = 5,
    LV_SORT_QUICKSORT = 6,
    LV_SORT_QUICKSORT_HYBRID = 7
} lv_sort_mode;

/* One closest-hit record per pixel (debug / parity output of the primary pass). */
typedef struct lv_hit {
    float t;            /* hit distance along the (unit) camera ray; 0 on miss */
    uint32_t prim;      /* segment index, 0xFFFFFFFF on miss */
    uint32_t kind;      /* 0 body, 1 sphere at p0, 2 sphere at p1 (TubeRayTracing.glsl:461-488) */
    uint32_t pad;
} lv_hit;

/* PPLL fragment node, bit-identical to LinkedListFragmentNode (LinkedListHeader.glsl:36-43). */
typedef struct lv_ppll_node {
    uint32_t color;     /* packUnorm4x8(rgba) */
    float depth;        /* length(fragmentPositionWorld - cameraPosition) */
    uint32_t next;      /* 0xFFFFFFFF terminates the list */
} lv_ppll_node;

/* ------------------------------------------------------------------ context ------------------ */

/* Replaces renderer construction in MainApp::setRenderer (src/MainApp.cpp:765-831).
 * `cuda_stream` is a cudaStream_t (NULL = the legacy default stream). */
int lv_ctx_create(lv_ctx** out, int device, void* cuda_stream);
int lv_ctx_destroy(lv_ctx* ctx);
/* sgl::Logfile::writeError equivalent: last error text for this context ("" if none). */
const char* lv_last_error(const lv_ctx* ctx);
/* Error text for failures of lv_ctx_create itself (no context yet). */
const char* lv_last_global_error(void);
int lv_abi_version(void);

/* Replaces LineRenderer::setNewSettings(const SettingsMap&) and its overrides
 * (src/Renderers/LineRenderer.cpp:433-498, RayTracing/VulkanRayTracer.cpp:226-278,
 * AmbientOcclusion/VulkanRayTracedAmbientOcclusion.cpp:115-144) -- same snake_case keys, values as
 * strings exactly as ReplayWidget hands them over.  Extra keys of this library are prefixed "b200_".
 * Unknown keys return LV_ERR_UNKNOWN_OPTION and change nothing. */
int lv_set_option(lv_ctx* ctx, const char* key, const char* value);
/* Reads back the current value of an option into buf (NUL-terminated, truncated to cap). */
int lv_get_option(const lv_ctx* ctx, const char* key, char* buf, size_t cap);

/* Replaces the transferFunctionTexture + MinMaxUniformBuffer bindings (src/LineData/LineData.cpp:1258-1273;
 * lookup semantics Data/Shaders/Utils/TransferFunction.glsl:66-71).  `rgba` = K linear RGBA entries,
 * sampled like a 1-D texture with linear filtering and clamp-to-edge (texel centres at (i+0.5)/K). */
int lv_set_transfer_function(lv_ctx* ctx, const float* rgba, uint32_t K, float attr_min, float attr_max);

/* Image-tile sharding (new).  The image is cut into tile_size x tile_size tiles, tiles are
 * enumerated in Morton order and tile i belongs to rank (i % world).  With world == 1 (default) the
 * context renders the whole frame.  With world > 1 render calls only touch owned pixels; all other
 * pixels of the output buffer are left untouched. */
int lv_set_tile_shard(lv_ctx* ctx, uint32_t rank, uint32_t world, uint32_t tile_size);
/* Cost-balanced tile ownership (new).  lv_get_tile_costs: per tile of the width x height frame, in the Morton order of the
 * enumeration, the hit pixels of this context's last RTAO pass (0 for tiles of other ranks) -- summed over the ranks this is the
 * AO-ray cost map of the frame.  lv_set_tile_owners replaces "tile i -> rank i % world" by an explicit owner per tile (same order,
 * same map on every rank) for frames of exactly this size, until the next lv_set_tile_shard; the frame kernels, the peer-frame
 * stores and the tile pack / unpack follow it.  linevis_b200/sharding.py::balance_tiles builds the map (longest processing time
 * first).  The reference has no counterpart: its frame is rendered by one GPU. */
int lv_get_tile_costs(lv_ctx* ctx, uint32_t width, uint32_t height, uint32_t* costs, uint32_t n_tiles);
int lv_set_tile_owners(lv_ctx* ctx, uint32_t width, uint32_t height, const unsigned char* owners, uint32_t n_tiles);
/* Number of tiles this context owns for a W x H frame and their (tile_x, tile_y) coordinates. */
int lv_get_owned_tiles(const lv_ctx* ctx, uint32_t width, uint32_t height,
                       uint32_t* tiles_xy /* n_owned*2 or NULL */, uint32_t* n_owned);
/* Copies the owned tiles of `image` (W*H*4 floats, device) into a compact block
 * [n_owned][tile*tile][4] (device) and back.  Used around the NCCL gather. */
int lv_pack_owned_tiles(lv_ctx* ctx, const float* image, uint32_t width, uint32_t height, float* packed);
int lv_unpack_tiles(lv_ctx* ctx, const float* packed, uint32_t src_rank, uint32_t world,
                    uint32_t width, uint32_t height, float* image);

/* Peer-memory frame assembly (new): instead of pack -> NCCL all_gather -> unpack, every rank's frame kernels store
 * their owned tiles straight into ONE frame buffer that lives on the assembling rank's GPU (NVLink peer stores, overlapped
 * with the kernel's own work); a single tiny collective (or any other cross-rank fence) after the render call is the frame
 * fence.  The assembling rank allocates the frame with lv_frame_alloc and exports it; the other ranks open the handle and pass
 * the returned pointer as `rgba_out` of lv_render_tubes / lv_render_ppll / lv_ppll_resolve.
 * handle: LV_IPC_HANDLE_BYTES opaque bytes (a cudaIpcMemHandle_t) to be shipped to the other processes by any means. */
#define LV_IPC_HANDLE_BYTES 64
int lv_frame_alloc(lv_ctx* ctx, uint32_t width, uint32_t height, float** frame_out /* device, W*H*4 floats, zeroed */);
int lv_frame_free(lv_ctx* ctx, float* frame);
int lv_ipc_export(lv_ctx* ctx, const void* device_ptr, void* handle_out);
int lv_ipc_open(lv_ctx* ctx, const void* handle, void** peer_ptr_out);
int lv_ipc_close(lv_ctx* ctx, void* peer_ptr);

/* ------------------------------------------------------------------ scene -------------------- */

/* Replaces LineDataFlow::getLinePassTubeAabbRenderData (src/LineData/LineDataFlow.cpp:2112-2277) +
 * LineData::getRayTracingTubeAabbTopLevelAS (src/LineData/LineData.cpp:879-907,1057-1075):
 * uploads the segment soup, packs 32-byte segment records, builds the BVH on the GPU.
 * pos_xyz: n_pt*3, attr: n_pt, seg_idx: n_seg*2 point indices (host pointers; copied). */
int lv_scene_create(lv_ctx* ctx, lv_scene** out, const float* pos_xyz, const float* attr,
                    const uint32_t* seg_idx, uint64_t n_pt, uint64_t n_seg, float line_width);
/* Same, with all three arrays already resident on the context's device. */
int lv_scene_create_device(lv_ctx* ctx, lv_scene** out, const float* d_pos_xyz, const float* d_attr,
                           const uint32_t* d_seg_idx, uint64_t n_pt, uint64_t n_seg, float line_width);
int lv_scene_destroy(lv_scene* scene);
/* BVH introspection for tests: node count, build time and a host copy of the 64-byte child-pair nodes
 * (see DESIGN.md for the layout).  nodes_out may be NULL. */
int lv_scene_info(const lv_scene* scene, uint64_t* n_seg, uint64_t* n_nodes, float* build_ms,
                  float* aabb_min_max /* 6 floats */);
int lv_scene_copy_bvh(const lv_scene* scene, void* nodes_out, size_t cap_bytes);

/* ------------------------------------------------------------------ line frames + AO prebaker */

/* Attaches the per-point line frames and the polyline structure to a scene.  Replaces the lineTangent / lineNormal members of
 * the linePointDataBuffer (LinePointDataUnified, src/LineData/LineRenderData.hpp:99-106, filled at
 * src/LineData/LineDataFlow.cpp:2140-2236) and `lines = lineData->getFilteredLines()` of the AO baker
 * (src/Renderers/AmbientOcclusion/VulkanAmbientOcclusionBaker.cpp:476-497).  Needed only for ambient_occlusion_mode =
 * "RTAO (Prebaker)".  pos_xyz / tangent_xyz / normal_xyz: n_pt*3 floats (pos_xyz = the array given to lv_scene_create);
 * line_offsets: n_lines+1 point offsets, polyline l = points [line_offsets[l], line_offsets[l+1]), each >= 2 points,
 * together covering all n_pt points in order.  Host pointers; copied.  Resets the baked factors. */
int lv_scene_set_lines(lv_scene* scene, const float* pos_xyz, const float* tangent_xyz, const float* normal_xyz, uint64_t n_pt,
                       const uint64_t* line_offsets, uint64_t n_lines);

/* Host-only (no GPU needed).  Replaces AmbientOcclusionComputeRenderPass::generateBlendingWeightParametrization +
 * recomputeStaticParametrization (VulkanAmbientOcclusionBaker.cpp:513-655): blending_weights[n_pt] (line point ->
 * parametrization coordinate; may be NULL), sampling_locations[cap] (parametrization vertex -> line point coordinate; may be
 * NULL), *n_param_vertices = numParametrizationVertices (may exceed cap). */
int lv_ao_parametrize(const float* pos_xyz, const uint64_t* line_offsets, uint64_t n_lines, float expected_param_segment_length,
                      float* blending_weights, float* sampling_locations, uint64_t cap, uint64_t* n_param_vertices);

/* Host-only (no GPU needed).  The reference's triangulated capped tubes: createCappedTriangleTubesRenderDataCPU with tubeClosed == false
 * (src/Renderers/Tubes/CappedTriangleTubesCPU.cpp:33-385, src/Renderers/Tubes/Tubes.cpp:35-86) -- the geometry the reference's
 * RTAO passes are traced against and what `b200_rtao_geometry = triangles` makes lv_render_rtao / lv_render_tubes / lv_ao_bake
 * trace instead of the analytic capsules (the scene then needs lv_scene_set_lines).  vertices: 8 floats each -- position,
 * as_float(line point index | 0x80000000 on cap vertices), normal, phi (TubeTriangleVertexData, LineRenderData.hpp:171-176);
 * indices: 3 per triangle.  Either array may be NULL; caps in vertices / triangles; the counts are always returned. */
int lv_tube_mesh(const float* pos_xyz, const uint64_t* line_offsets, uint64_t n_lines, float tube_radius, uint32_t num_subdivisions,
                 float* vertices, uint64_t vertices_cap, uint32_t* indices, uint64_t triangles_cap,
                 uint64_t* n_vertices, uint64_t* n_triangles, uint64_t* n_line_points);

/* Object-space RTAO prebaker: runs baking iterations (one dispatch of Data/Shaders/AO/RTAO/VulkanAmbientOcclusionBaker.glsl:190-282
 * each, frameNumber = iterations done so far) until b200_prebaker_iterations are reached; n_iterations = 0 runs all that are
 * left (BakingMode::IMMEDIATE, VulkanAmbientOcclusionBaker::startAmbientOcclusionBaking :161-192), n_iterations = 1 is one
 * VulkanAmbientOcclusionBaker::updateIterative step (:340-351).  Render calls in "RTAO (Prebaker)" mode run one iteration per
 * frame themselves while iterations are left (LineRenderer::renderBase, src/Renderers/LineRenderer.cpp:257-264).
 * Settings (GUI-only in the reference, VulkanAmbientOcclusionBaker.hpp:108,163-168): b200_prebaker_iterations (128),
 * b200_prebaker_samples_per_frame (4), b200_prebaker_subdivisions (8), b200_prebaker_param_segment_length (0.001),
 * b200_prebaker_radius (0.1), b200_prebaker_distance_based (true).  stats: rays_ao / ao_traversal_steps / ao_intersections. */
int lv_ao_bake(lv_ctx* ctx, lv_scene* scene, uint32_t n_iterations, lv_stats* stats);
/* Multi-GPU baking (new; the reference bakes on one GPU): the parametrization vertices are independent, so rank r bakes the slice
 * [first_vertex, first_vertex + n_vertices) (n_vertices == 0: everything from first_vertex) with the same settings and the same
 * iteration count as the other ranks and the ranks then exchange their slices of the factor buffer (lv_ao_factors gives the
 * device buffer, index subdivision + n_subdiv * vertex, n_floats = n_param_vertices * n_subdiv): any collective or peer copy will
 * do, linevis_b200/sharding.py::exchange_baked_slices broadcasts each rank's slice.  The result is bit-identical to a one-GPU bake. */
int lv_ao_set_vertex_range(lv_scene* scene, uint64_t first_vertex, uint64_t n_vertices);
int lv_ao_factors(lv_scene* scene, float** device_factors, uint64_t* n_floats);
/* Restart baking from iteration 0 (startAmbientOcclusionBaking: numIterations = 0). */
int lv_ao_bake_reset(lv_scene* scene);
/* Host copies of the baker's buffers: ambientOcclusionFactors[n_param * n_subdiv] (index subdivision + n_subdiv * vertex),
 * blending weights [n_pt], sampling locations [n_param].  Any pointer may be NULL; caps in elements. */
int lv_ao_read(lv_scene* scene, float* factors, size_t factors_cap, float* blending_weights, size_t weights_cap,
               float* sampling_locations, size_t sampling_cap, uint32_t* n_param_vertices, uint32_t* n_subdivisions,
               uint32_t* iterations_done);

/* ------------------------------------------------------------------ frames ------------------- */

/* Replaces VulkanRayTracer::render() (src/Renderers/RayTracing/VulkanRayTracer.cpp:131-154):
 * LineRenderer::renderBase -> RTAO updateIterative (if ambient_occlusion_strength > 0) -> traceRays(W,H,1).
 * frame_number is RayTracerSettings::frameNumber (running mean over frames, TubeRayTracing.glsl:268-273);
 * the accumulation image is `rgba_out` itself, exactly like the reference's storage image.
 * rgba_out: W*H*4 floats, device or host (detected).  stats may be NULL.
 * Frame format (option b200_frame_format, the reference's sceneTexture is RGBA8 / RGBA16 UNORM, src/Widgets/DataView.cpp:100-108):
 *   "rgba32f" (default)  as above;
 *   "rgba8"              rgba_out is a uint32_t[W*H] RGBA8 UNORM frame (packUnorm4x8 of the float pixel, written in the frame kernel's
 *                        epilogue); the float accumulation image then lives inside the context.  The same holds for lv_render_ppll /
 *                        lv_ppll_resolve.  With b200_async_delivery = true a HOST rgba8 frame is copied on a second stream from one of
 *                        two staging frames: the call returns once the copy is enqueued and lv_synchronize() waits for it, so that the
 *                        read-back of frame i overlaps the kernels of frame i + 1 (pinned memory; at most two frames in flight).
 * geometry_mode = "Triangle Mesh" (RayTracingGeometryMode::TRIANGLE_MESH, VulkanRayTracer.hpp:54-63; needs lv_scene_set_lines):
 * the pass traces the reference's triangulated tubes and shades with ClosestHitTubeTriangles (TubeRayTracing.glsl:301-351). */
int lv_render_tubes(lv_ctx* ctx, const lv_scene* scene, const lv_camera* cam, uint32_t frame_number,
                    float* rgba_out, lv_stats* stats);

/* AO-sample-batch shards (new; SURVEY 8e, the second shard axis): lv_render_tubes in three stages, so that N ranks split the SAMPLES of
 * the screen-space RTAO pass -- the samples of a pixel are independent, their seeds are tea(pixel, frameNumber * spp + sample)
 * (Data/Shaders/AO/RTAO/VulkanRayTracedAmbientOcclusion.glsl:289-292).  With lv_set_tile_shard(rank, world) in force:
 *   lv_sao_primary  the RTAO pass's camera rays on the rank's own tiles (.glsl:178-276); returns the device address and length of the
 *                   rank's hit list (48-byte start frames: position + ray origin offset, normal + pixel, tangent);
 *   -- the caller concatenates the ranks' lists in rank order (all-gather), linevis_b200/sharding.py::SampleShards --
 *   lv_sao_trace    AO rays [sample_first, sample_first + sample_count) of EVERY record of `hits_device` (.glsl:158-175,289-305) into
 *                   occ_device[record * sample_count + k]: the numerator of traceAoRay's t / radius.  sample_count divides spp;
 *   -- the caller returns each record's values to its owner (all-to-all): occ_parts[part][own record][k] --
 *   lv_sao_finish   sums every own record's spp values IN SAMPLE ORDER (.glsl:301-317: bit for bit the one-GPU sum), then the tube pass
 *                   of lv_render_tubes on the rank's tiles into rgba_out (peer frame, device or host).
 * Needs the default AO ray stream; frames are identical to lv_render_tubes' whatever N is. */
int lv_sao_primary(lv_ctx* ctx, const lv_scene* scene, const lv_camera* cam, uint32_t frame_number, const void** hits_device, uint32_t* n_hits);
int lv_sao_trace(lv_ctx* ctx, const lv_scene* scene, const lv_camera* cam, uint32_t frame_number, const void* hits_device, uint32_t n_hits,
                 uint32_t sample_first, uint32_t sample_count, float* occ_device);
int lv_sao_finish(lv_ctx* ctx, const lv_scene* scene, const lv_camera* cam, uint32_t frame_number, const float* occ_parts, uint32_t n_parts,
                  float* rgba_out, lv_stats* stats);

/* Replaces PerPixelLinkedListLineRenderer::render() (src/Renderers/OIT/PerPixelLinkedListLineRenderer.cpp:399-427):
 * clear() -> gather() -> resolve().  max_frags == MAX_NUM_FRAGS (expectedMaxDepthComplexity),
 * linked_list_size == fragmentBufferSize (0 = expectedAvgDepthComplexity * paddedW * paddedH like
 * reallocateFragmentBuffer, :251-258). */
int lv_render_ppll(lv_ctx* ctx, const lv_scene* scene, const lv_camera* cam, uint32_t max_frags,
                   uint32_t sort_mode, uint64_t linked_list_size, float* rgba_out, lv_stats* stats);

/* ---- single stages (parity tests and benchmarks time these individually) ---- */

/* Closest-hit pass only (S1+S2): one lv_hit per pixel.  hits_out device or host. */
int lv_trace_primary(lv_ctx* ctx, const lv_scene* scene, const lv_camera* cam, lv_hit* hits_out,
                     lv_stats* stats);
/* Screen-space RTAO pass (S5): ao_out = W*H floats (the .x channel of the reference's rgba32f
 * accumulation image), running mean over frame_number like VulkanRayTracedAmbientOcclusion.glsl:313-319. */
int lv_render_rtao(lv_ctx* ctx, const lv_scene* scene, const lv_camera* cam, uint32_t frame_number,
                   float* ao_out, lv_stats* stats);
/* PPLL stages on the context's internal buffers. */
int lv_ppll_clear(lv_ctx* ctx, const lv_camera* cam, uint64_t linked_list_size);
int lv_ppll_gather(lv_ctx* ctx, const lv_scene* scene, const lv_camera* cam, lv_stats* stats);
int lv_ppll_resolve(lv_ctx* ctx, const lv_camera* cam, uint32_t max_frags, uint32_t sort_mode,
                    float* rgba_out, lv_stats* stats);
/* Host copies of the PPLL buffers after gather: fragCounter, startOffset[paddedW*paddedH],
 * nodes[min(counter, linkedListSize)].  Any pointer may be NULL.  Sizes in elements. */
int lv_ppll_read(lv_ctx* ctx, uint32_t* frag_counter, uint32_t* start_offset, size_t start_offset_cap,
                 lv_ppll_node* nodes, size_t nodes_cap, uint32_t* padded_w, uint32_t* padded_h);

/* The frame in the reference's own output format: sceneTexture is an RGBA8 UNORM storage image (VulkanRayTracer.cpp:677-679), written
 * by imageStore of the float colour, i.e. packUnorm4x8 per pixel (round(clamp(c, 0, 1) * 255), x in the lowest byte).  Converts a
 * device RGBA32F frame (W*H*4 floats, as written by lv_render_tubes / lv_render_ppll; with tile sharding only the owned tiles are
 * converted) into rgba8_out (W*H uint32, device or host): a quarter of the float frame's read-back bytes for callers that display. */
int lv_frame_to_rgba8(lv_ctx* ctx, const float* rgba_device, uint32_t width, uint32_t height, uint32_t* rgba8_out);

/* Device synchronisation helper for FFI callers that do not link the CUDA runtime. */
int lv_synchronize(lv_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LINEVIS_B200_H */
