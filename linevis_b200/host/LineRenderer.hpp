// LineRenderer.hpp -- the renderer interface of the reference (src/Renderers/LineRenderer.hpp:66-277) reduced to what
// MainApp calls on the hot path, with sgl types replaced by plain structs.  B200RayTracer and
// B200PerPixelLinkedListLineRenderer derive from it exactly like VulkanRayTracer / PerPixelLinkedListLineRenderer do, and
// forward to the C ABI in include/linevis_b200.h.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "../../include/linevis_b200.h"
#include "LineData.hpp"
#include "SettingsMap.hpp"

// Stand-in for sgl's TransferFunctionWindow (not in the reference tree): the 1-D RGBA LUT it uploads and the attribute range
// it puts into MinMaxUniformBuffer (src/LineData/LineData.cpp:1258-1273).
struct TransferFunction {
    std::vector<float> rgba;  // K x 4, linear RGB + opacity
    float attrMin = 0.0f, attrMax = 1.0f;
};

// Stand-in for SceneData: camera (sgl::Camera), clear colour, viewport and the scene texture the renderer writes
// (RGBA8/RGBA16 UNORM storage image in the reference, src/Widgets/DataView.cpp:100-111; float RGBA here).
struct SceneData {
    float viewMatrix[16], projectionMatrix[16];
    float cameraPosition[3];
    float fovY = 1.0f;
    float nearClip = 0.01f, farClip = 100.0f;   // sgl::Camera::getNearClipDistance / getFarClipDistance (depth cues)
    float clearColor[4] = {1, 1, 1, 1};
    uint32_t viewportWidth = 0, viewportHeight = 0;
    std::vector<float> sceneTexture;  // viewportWidth * viewportHeight * 4
};

// Stand-in for InternalState (src/Utils/InternalState.hpp:171-199): what a --perf state hands to LineRenderer::setNewState.
// rendererSettings carries the camelCase keys of the performance-measurement modes (InternalState.cpp), not the snake_case
// keys of the replay scripts.
struct InternalState {
    std::string name;
    int renderingMode = 0;
    SettingsMap rendererSettings;
    int tilingWidth = 2, tilingHeight = 8;        // setNewTilingMode (LineRenderer.cpp:739-760)
    bool useMortonCodeForTiling = false;
    std::string transferFunctionName;
    int windowResolution[2] = {0, 0};
};

enum RenderingMode { RENDERING_MODE_PER_PIXEL_LINKED_LIST = 1, RENDERING_MODE_VULKAN_RAY_TRACER = 9 };  // src/Renderers/RenderingModes.hpp:32-53

class LineRenderer {
public:
    LineRenderer(std::string windowName, SceneData* sceneData, TransferFunction& transferFunctionWindow, int device = 0, void* cudaStream = nullptr);
    virtual ~LineRenderer();
    virtual RenderingMode getRenderingMode() const = 0;
    virtual bool getIsTransparencyUsed() { return true; }
    bool isDirty() const { return dirty; }
    virtual bool needsReRender() { bool tmp = reRender; reRender = false; return tmp; }
    virtual bool getIsTriangleRepresentationUsed() const { return false; }

    /// Re-generates the visualization mapping (uploads the segment soup, builds the BVH).
    virtual void setLineData(LineDataPtr& lineData, bool isNewData);
    /// Renders the object to sceneData->sceneTexture.
    virtual void render() = 0;
    virtual void onResolutionChanged();
    virtual void onClearColorChanged() { reRender = true; }
    virtual void onTransferFunctionMapRebuilt();
    virtual void onHasMoved() { reRender = true; }
    virtual void notifyReRenderTriggeredExternally() { internalReRender = false; }
    /// Same keys as the reference (src/Renderers/LineRenderer.cpp:433-498); returns whether a gather-shader reload would be needed.
    virtual bool setNewSettings(const SettingsMap& settings);
    /// Called when a new performance-measurement state is set (LineRenderer.hpp:109-110; MainApp::setNewState).  The base applies the
    /// state's PPLL addressing tile (setNewTilingMode); renderers override it for their rendererSettings keys.
    virtual void setNewState(const InternalState& newState);

    static void setLineWidth(float width) { lineWidth = width; }
    static float getLineWidth() { return lineWidth; }
    const std::string& getWindowName() const { return windowName; }
    const lv_stats& getLastStats() const { return lastStats; }
    const char* getLastError() const;

protected:
    void fillCamera(lv_camera& cam) const;
    void check(int status, const char* what);
    std::string windowName;
    SceneData* sceneData;
    TransferFunction& transferFunctionWindow;
    LineDataPtr lineData;
    lv_ctx* ctx = nullptr;
    lv_scene* scene = nullptr;
    bool dirty = true, reRender = true, internalReRender = true;
    float depthCueStrength = 0.0f, ambientOcclusionStrength = 0.0f, ambientOcclusionGamma = 1.0f;
    lv_stats lastStats{};
    static float lineWidth;
};

// Counterpart of VulkanRayTracer (src/Renderers/RayTracing/VulkanRayTracer.{hpp,cpp})
class B200RayTracer : public LineRenderer {
public:
    B200RayTracer(SceneData* sceneData, TransferFunction& tf, int device = 0, void* cudaStream = nullptr);
    RenderingMode getRenderingMode() const override { return RENDERING_MODE_VULKAN_RAY_TRACER; }
    bool getIsTransparencyUsed() override { return false; }
    void setLineData(LineDataPtr& lineData, bool isNewData) override;
    void onResolutionChanged() override;
    void onHasMoved() override { accumulatedFramesCounter = 0; LineRenderer::onHasMoved(); }
    bool needsReRender() override;                       // VulkanRayTracer.cpp:330-336
    void notifyReRenderTriggeredExternally() override { internalReRender = false; accumulatedFramesCounter = 0; }
    bool setNewSettings(const SettingsMap& settings) override;  // VulkanRayTracer.cpp:226-278
    void setNewState(const InternalState& newState) override;   // VulkanRayTracer.cpp:280-328
    void render() override;                              // VulkanRayTracer.cpp:131-154
    uint32_t getAccumulatedFramesCounter() const { return accumulatedFramesCounter; }

private:
    uint32_t numSamplesPerFrame = 2, maxNumAccumulatedFrames = 32, accumulatedFramesCounter = 0;  // VulkanRayTracer.hpp:137-142
    // the screen-space RTAO pass accumulates its own iterations (VulkanRayTracedAmbientOcclusion.hpp:150: maxNumAccumulatedFrames 64);
    // LineRenderer::renderBase keeps forcing frames while it is still running (LineRenderer.cpp:257-264)
    uint32_t ambientOcclusionIterations = 64;
};

// Counterpart of PerPixelLinkedListLineRenderer (src/Renderers/OIT/PerPixelLinkedListLineRenderer.{hpp,cpp})
class B200PerPixelLinkedListLineRenderer : public LineRenderer {
public:
    B200PerPixelLinkedListLineRenderer(SceneData* sceneData, TransferFunction& tf, int device = 0, void* cudaStream = nullptr);
    RenderingMode getRenderingMode() const override { return RENDERING_MODE_PER_PIXEL_LINKED_LIST; }
    void setLineData(LineDataPtr& lineData, bool isNewData) override;
    void onResolutionChanged() override;
    void render() override;                              // PerPixelLinkedListLineRenderer.cpp:399-427
    void setNewState(const InternalState& newState) override;   // PerPixelLinkedListLineRenderer.cpp:98-107
    const std::string& getCurrentStateName() const { return currentStateName; }
    void setSortingAlgorithmMode(lv_sort_mode mode) { sortingAlgorithmMode = mode; reRender = true; }

private:
    void updateLargeMeshMode();                          // PerPixelLinkedListLineRenderer.cpp:109-126
    void reallocateFragmentBuffer();                     // :251-258
    lv_sort_mode sortingAlgorithmMode = LV_SORT_PRIORITY_QUEUE;   // .hpp:113
    int expectedAvgDepthComplexity = 20, expectedMaxDepthComplexity = 100;  // MESH_MODE_DEPTH_COMPLEXITIES_PPLL, .hpp:45-49
    uint64_t fragmentBufferSize = 0;
    std::string currentStateName;
};
