#include "LineRenderer.hpp"

#include <cstring>
#include <stdexcept>

float LineRenderer::lineWidth = 0.002f;  // STANDARD_LINE_WIDTH, src/Loaders/DataSetList.hpp:46

namespace {
void invert4x4(const float* m, float* out) {  // glm::inverse (cofactor expansion), column-major
    double a[16], inv[16];
    for (int i = 0; i < 16; i++) a[i] = m[i];
    inv[0] = a[5]*a[10]*a[15] - a[5]*a[11]*a[14] - a[9]*a[6]*a[15] + a[9]*a[7]*a[14] + a[13]*a[6]*a[11] - a[13]*a[7]*a[10];
    inv[4] = -a[4]*a[10]*a[15] + a[4]*a[11]*a[14] + a[8]*a[6]*a[15] - a[8]*a[7]*a[14] - a[12]*a[6]*a[11] + a[12]*a[7]*a[10];
    inv[8] = a[4]*a[9]*a[15] - a[4]*a[11]*a[13] - a[8]*a[5]*a[15] + a[8]*a[7]*a[13] + a[12]*a[5]*a[11] - a[12]*a[7]*a[9];
    inv[12] = -a[4]*a[9]*a[14] + a[4]*a[10]*a[13] + a[8]*a[5]*a[14] - a[8]*a[6]*a[13] - a[12]*a[5]*a[10] + a[12]*a[6]*a[9];
    inv[1] = -a[1]*a[10]*a[15] + a[1]*a[11]*a[14] + a[9]*a[2]*a[15] - a[9]*a[3]*a[14] - a[13]*a[2]*a[11] + a[13]*a[3]*a[10];
    inv[5] = a[0]*a[10]*a[15] - a[0]*a[11]*a[14] - a[8]*a[2]*a[15] + a[8]*a[3]*a[14] + a[12]*a[2]*a[11] - a[12]*a[3]*a[10];
    inv[9] = -a[0]*a[9]*a[15] + a[0]*a[11]*a[13] + a[8]*a[1]*a[15] - a[8]*a[3]*a[13] - a[12]*a[1]*a[11] + a[12]*a[3]*a[9];
    inv[13] = a[0]*a[9]*a[14] - a[0]*a[10]*a[13] - a[8]*a[1]*a[14] + a[8]*a[2]*a[13] + a[12]*a[1]*a[10] - a[12]*a[2]*a[9];
    inv[2] = a[1]*a[6]*a[15] - a[1]*a[7]*a[14] - a[5]*a[2]*a[15] + a[5]*a[3]*a[14] + a[13]*a[2]*a[7] - a[13]*a[3]*a[6];
    inv[6] = -a[0]*a[6]*a[15] + a[0]*a[7]*a[14] + a[4]*a[2]*a[15] - a[4]*a[3]*a[14] - a[12]*a[2]*a[7] + a[12]*a[3]*a[6];
    inv[10] = a[0]*a[5]*a[15] - a[0]*a[7]*a[13] - a[4]*a[1]*a[15] + a[4]*a[3]*a[13] + a[12]*a[1]*a[7] - a[12]*a[3]*a[5];
    inv[14] = -a[0]*a[5]*a[14] + a[0]*a[6]*a[13] + a[4]*a[1]*a[14] - a[4]*a[2]*a[13] - a[12]*a[1]*a[6] + a[12]*a[2]*a[5];
    inv[3] = -a[1]*a[6]*a[11] + a[1]*a[7]*a[10] + a[5]*a[2]*a[11] - a[5]*a[3]*a[10] - a[9]*a[2]*a[7] + a[9]*a[3]*a[6];
    inv[7] = a[0]*a[6]*a[11] - a[0]*a[7]*a[10] - a[4]*a[2]*a[11] + a[4]*a[3]*a[10] + a[8]*a[2]*a[7] - a[8]*a[3]*a[6];
    inv[11] = -a[0]*a[5]*a[11] + a[0]*a[7]*a[9] + a[4]*a[1]*a[11] - a[4]*a[3]*a[9] - a[8]*a[1]*a[7] + a[8]*a[3]*a[5];
    inv[15] = a[0]*a[5]*a[10] - a[0]*a[6]*a[9] - a[4]*a[1]*a[10] + a[4]*a[2]*a[9] + a[8]*a[1]*a[6] - a[8]*a[2]*a[5];
    double det = a[0]*inv[0] + a[1]*inv[4] + a[2]*inv[8] + a[3]*inv[12];
    for (int i = 0; i < 16; i++) out[i] = float(inv[i] / det);
}
}  // namespace

LineRenderer::LineRenderer(std::string name, SceneData* sd, TransferFunction& tf, int device, void* cudaStream)
    : windowName(std::move(name)), sceneData(sd), transferFunctionWindow(tf) {
    int rc = lv_ctx_create(&ctx, device, cudaStream);
    if (rc != LV_OK) throw std::runtime_error(std::string("lv_ctx_create: ") + lv_last_global_error());  // no CPU fallback
    onTransferFunctionMapRebuilt();
}

LineRenderer::~LineRenderer() {
    if (scene) lv_scene_destroy(scene);
    if (ctx) lv_ctx_destroy(ctx);
}

const char* LineRenderer::getLastError() const { return lv_last_error(ctx); }

void LineRenderer::check(int status, const char* what) {
    if (status != LV_OK) throw std::runtime_error(std::string(what) + ": " + lv_last_error(ctx));
}

void LineRenderer::onTransferFunctionMapRebuilt() {
    if (!transferFunctionWindow.rgba.empty())
        check(lv_set_transfer_function(ctx, transferFunctionWindow.rgba.data(), uint32_t(transferFunctionWindow.rgba.size() / 4),
                                       transferFunctionWindow.attrMin, transferFunctionWindow.attrMax), "lv_set_transfer_function");
    reRender = true;
}

void LineRenderer::setLineData(LineDataPtr& data, bool isNewData) {
    // LineRenderer::updateNewLineData (LineRenderer.cpp:676-708) + RayTracingRenderPass::setLineData (VulkanRayTracer.cpp:370-392)
    lineData = data;
    const TubeAabbRenderData& rd = lineData->getLinePassTubeAabbRenderData(lineWidth);
    std::vector<float> pos(rd.linePointDataBuffer.size() * 3), attr(rd.linePointDataBuffer.size());
    for (size_t i = 0; i < rd.linePointDataBuffer.size(); i++) {
        pos[3 * i] = rd.linePointDataBuffer[i].linePosition.x; pos[3 * i + 1] = rd.linePointDataBuffer[i].linePosition.y;
        pos[3 * i + 2] = rd.linePointDataBuffer[i].linePosition.z; attr[i] = rd.linePointDataBuffer[i].lineAttribute;
    }
    if (scene) { lv_scene_destroy(scene); scene = nullptr; }
    check(lv_set_option(ctx, "use_capped_tubes", lineData->useCappedTubes ? "true" : "false"), "use_capped_tubes");
    check(lv_set_option(ctx, "use_halos", lineData->useHalos ? "true" : "false"), "use_halos");
    check(lv_set_option(ctx, "tube_num_subdivisions", std::to_string(lineData->tubeNumSubdivisions).c_str()), "tube_num_subdivisions");
    check(lv_scene_create(ctx, &scene, pos.data(), attr.data(), rd.indexBuffer.data(), attr.size(), rd.indexBuffer.size() / 2, lineWidth),
          "lv_scene_create");
    if (rd.lineOffsets.size() >= 2) {
        // line frames for ambient_occlusion_mode = "RTAO (Prebaker)" (AmbientOcclusionComputeRenderPass::setLineData,
        // VulkanAmbientOcclusionBaker.cpp:476-497): cheap, so always attached
        std::vector<float> tangent(pos.size()), normal(pos.size());
        for (size_t i = 0; i < rd.linePointDataBuffer.size(); i++) {
            const LinePointDataUnified& p = rd.linePointDataBuffer[i];
            tangent[3 * i] = p.lineTangent.x; tangent[3 * i + 1] = p.lineTangent.y; tangent[3 * i + 2] = p.lineTangent.z;
            normal[3 * i] = p.lineNormal.x; normal[3 * i + 1] = p.lineNormal.y; normal[3 * i + 2] = p.lineNormal.z;
        }
        check(lv_scene_set_lines(scene, pos.data(), tangent.data(), normal.data(), attr.size(), rd.lineOffsets.data(), rd.lineOffsets.size() - 1),
              "lv_scene_set_lines");
    }
    (void)isNewData;
    lineData->resetDirty();
    dirty = false;
    reRender = true;
}

void LineRenderer::onResolutionChanged() {
    sceneData->sceneTexture.assign(size_t(sceneData->viewportWidth) * sceneData->viewportHeight * 4, 0.0f);
    reRender = true;
}

bool LineRenderer::setNewSettings(const SettingsMap& settings) {
    // every key goes to the library unchanged; line_width additionally invalidates the scene like setTriangleRepresentationDirty
    bool shallReloadGatherShader = false;
    for (const auto& kv : settings.getMap()) {
        int rc = lv_set_option(ctx, kv.first.c_str(), kv.second.c_str());
        if (rc == LV_ERR_UNKNOWN_OPTION) continue;   // keys of other subsystems (camera, dataset, ...) are not ours
        check(rc, kv.first.c_str());
    }
    float newLineWidth = lineWidth;
    if (settings.getValueOpt("line_width", newLineWidth) && newLineWidth != lineWidth) {
        lineWidth = newLineWidth;
        if (lineData) { LineDataPtr d = lineData; setLineData(d, false); }
    }
    if (settings.getValueOpt("depth_cue_strength", depthCueStrength)) shallReloadGatherShader = true;
    if (settings.getValueOpt("ambient_occlusion_strength", ambientOcclusionStrength)) shallReloadGatherShader = true;
    settings.getValueOpt("ambient_occlusion_gamma", ambientOcclusionGamma);
    reRender = true;
    return shallReloadGatherShader;
}

void LineRenderer::setNewState(const InternalState& newState) {
    // MainApp::setNewState -> LineRenderer::setNewTilingMode (LineRenderer.cpp:739-760): the PPLL start-offset addressing tile
    if (newState.tilingWidth > 0 && newState.tilingHeight > 0) {
        check(lv_set_option(ctx, "b200_tiling_width", std::to_string(newState.tilingWidth).c_str()), "tilingWidth");
        check(lv_set_option(ctx, "b200_tiling_height", std::to_string(newState.tilingHeight).c_str()), "tilingHeight");
    }
    reRender = true;
}

void LineRenderer::fillCamera(lv_camera& cam) const {
    // LineData::updateVulkanUniformBuffers, src/LineData/LineData.cpp:1275-1319
    std::memcpy(cam.view, sceneData->viewMatrix, 64);
    std::memcpy(cam.proj, sceneData->projectionMatrix, 64);
    invert4x4(cam.view, cam.inv_view);
    invert4x4(cam.proj, cam.inv_proj);
    std::memcpy(cam.position, sceneData->cameraPosition, 12);
    cam.fov_y = sceneData->fovY;
    std::memcpy(cam.background, sceneData->clearColor, 16);
    cam.width = sceneData->viewportWidth;
    cam.height = sceneData->viewportHeight;
    cam.near_dist = sceneData->nearClip;
    cam.far_dist = sceneData->farClip;
}

// ---------------------------------------------------------------------------------------------- B200RayTracer
B200RayTracer::B200RayTracer(SceneData* sd, TransferFunction& tf, int device, void* stream)
    : LineRenderer("Vulkan Ray Tracer (B200)", sd, tf, device, stream) {
    check(lv_set_option(ctx, "num_samples_per_frame", "2"), "num_samples_per_frame");
    check(lv_set_option(ctx, "num_accumulated_frames", "32"), "num_accumulated_frames");
}

void B200RayTracer::setLineData(LineDataPtr& data, bool isNewData) {
    LineRenderer::setLineData(data, isNewData);
    accumulatedFramesCounter = 0;
}

void B200RayTracer::onResolutionChanged() {
    LineRenderer::onResolutionChanged();
    accumulatedFramesCounter = 0;
}

bool B200RayTracer::needsReRender() {
    if (accumulatedFramesCounter < maxNumAccumulatedFrames) return true;
    // the AO pass is still collecting iterations: frames keep coming (and the frame counter keeps counting, VulkanRayTracer.cpp:153)
    if (ambientOcclusionStrength > 0.0f && accumulatedFramesCounter < ambientOcclusionIterations) return true;
    return LineRenderer::needsReRender();
}

void B200RayTracer::setNewState(const InternalState& newState) {
    LineRenderer::setNewState(newState);
    const SettingsMap& rs = newState.rendererSettings;
    std::string geometryMode;
    bool useAnalyticIntersections = true;
    // RAY_TRACING_GEOMETRY_MODE_NAMES (VulkanRayTracer.hpp:58-63): "Triangle Mesh" and "AABBs (analytic)" exist here; anything else
    // (linear swept spheres) is refused loudly by the library instead of being measured under the wrong name
    if (rs.getValueOpt("geometryMode", geometryMode)) {
        check(lv_set_option(ctx, "geometry_mode", geometryMode.c_str()), "geometryMode");
        accumulatedFramesCounter = 0;
    } else if (rs.getValueOpt("useAnalyticIntersections", useAnalyticIntersections)) {
        check(lv_set_option(ctx, "use_analytic_intersections", useAnalyticIntersections ? "true" : "false"), "useAnalyticIntersections");
        accumulatedFramesCounter = 0;
    }
    if (rs.getValueOpt("numSamplesPerFrame", numSamplesPerFrame)) {
        check(lv_set_option(ctx, "num_samples_per_frame", std::to_string(numSamplesPerFrame).c_str()), "numSamplesPerFrame");
        accumulatedFramesCounter = 0;
    }
    if (rs.getValueOpt("maxNumAccumulatedFrames", maxNumAccumulatedFrames)) {
        check(lv_set_option(ctx, "num_accumulated_frames", std::to_string(maxNumAccumulatedFrames).c_str()), "maxNumAccumulatedFrames");
        accumulatedFramesCounter = 0;
    }
    bool b = false;
    if (rs.getValueOpt("useDeterministicSampling", b)) {
        check(lv_set_option(ctx, "use_deterministic_sampling", b ? "true" : "false"), "useDeterministicSampling");
        accumulatedFramesCounter = 0;
    }
    if (rs.getValueOpt("useMlat", b)) {
        if (b) throw std::runtime_error("setNewState: useMlat is not available");
        accumulatedFramesCounter = 0;
    }
}

bool B200RayTracer::setNewSettings(const SettingsMap& settings) {
    bool r = LineRenderer::setNewSettings(settings);
    if (settings.getValueOpt("num_samples_per_frame", numSamplesPerFrame)) accumulatedFramesCounter = 0;
    if (settings.getValueOpt("num_accumulated_frames", maxNumAccumulatedFrames)) accumulatedFramesCounter = 0;
    bool b;
    if (settings.getValueOpt("use_deterministic_sampling", b)) accumulatedFramesCounter = 0;
    if (settings.getValueOpt("ambient_occlusion_iterations", ambientOcclusionIterations)) accumulatedFramesCounter = 0;
    return r;
}

void B200RayTracer::render() {
    if (!scene) {  // empty acceleration structure: clear to the clear colour (VulkanRayTracer.cpp:143-148)
        for (size_t i = 0; i < sceneData->sceneTexture.size(); i++) sceneData->sceneTexture[i] = sceneData->clearColor[i & 3];
        return;
    }
    lv_camera cam;
    fillCamera(cam);
    check(lv_render_tubes(ctx, scene, &cam, accumulatedFramesCounter, sceneData->sceneTexture.data(), &lastStats), "lv_render_tubes");
    accumulatedFramesCounter++;
}

// ---------------------------------------------------------------------------------------------- PPLL
B200PerPixelLinkedListLineRenderer::B200PerPixelLinkedListLineRenderer(SceneData* sd, TransferFunction& tf, int device, void* stream)
    : LineRenderer("Per-Pixel Linked List Renderer (B200)", sd, tf, device, stream) {}

void B200PerPixelLinkedListLineRenderer::setNewState(const InternalState& newState) {
    LineRenderer::setNewState(newState);
    currentStateName = newState.name;   // the name the per-state timings are filed under (PerPixelLinkedListLineRenderer.cpp:99)
}

void B200PerPixelLinkedListLineRenderer::updateLargeMeshMode() {
    const bool large = lineData && lineData->getNumLineSegments() > size_t(1e6);
    expectedAvgDepthComplexity = large ? 120 : 20;
    expectedMaxDepthComplexity = large ? 380 : 100;
    reallocateFragmentBuffer();
}

void B200PerPixelLinkedListLineRenderer::reallocateFragmentBuffer() {
    // padded to the 2x8 addressing tile like getScreenSizeWithTiling (LineRenderer.cpp:805-812)
    uint64_t pw = sceneData->viewportWidth, ph = sceneData->viewportHeight;
    if (pw % 2) pw = (pw / 2 + 1) * 2;
    if (ph % 8) ph = (ph / 8 + 1) * 8;
    fragmentBufferSize = uint64_t(expectedAvgDepthComplexity) * pw * ph;
}

void B200PerPixelLinkedListLineRenderer::setLineData(LineDataPtr& data, bool isNewData) {
    LineRenderer::setLineData(data, isNewData);
    updateLargeMeshMode();
}

void B200PerPixelLinkedListLineRenderer::onResolutionChanged() {
    LineRenderer::onResolutionChanged();
    reallocateFragmentBuffer();
}

void B200PerPixelLinkedListLineRenderer::render() {
    if (!scene) {
        for (size_t i = 0; i < sceneData->sceneTexture.size(); i++) sceneData->sceneTexture[i] = sceneData->clearColor[i & 3];
        return;
    }
    lv_camera cam;
    fillCamera(cam);
    check(lv_render_ppll(ctx, scene, &cam, uint32_t(expectedMaxDepthComplexity), uint32_t(sortingAlgorithmMode), fragmentBufferSize,
                         sceneData->sceneTexture.data(), &lastStats), "lv_render_ppll");
}
