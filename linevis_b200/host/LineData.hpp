// LineData.hpp -- host-side line data model, the part of the reference's LineData / LineDataFlow the hot path needs.
//
// Mirrors: Trajectory / Trajectories (src/Loaders/TrajectoryFile.hpp:38-43), LinePointDataUnified
// (src/LineData/LineRenderData.hpp:99-106), TubeAabbRenderData (:203-210) with host vectors instead of Vulkan buffers, and
// LineDataFlow::getLinePassTubeAabbRenderData (src/LineData/LineDataFlow.cpp:2112-2277).  LineVis's loaders fill
// `trajectories` exactly as they do today; nothing here touches the GPU.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace lvh {
struct vec3 { float x = 0, y = 0, z = 0; };
}

struct Trajectory {
    std::vector<lvh::vec3> positions;
    std::vector<std::vector<float>> attributes;  // [attribute][point]
};
typedef std::vector<Trajectory> Trajectories;

struct LinePointDataUnified {  // 48 bytes, src/LineData/LineRenderData.hpp:99-106
    lvh::vec3 linePosition; float lineAttribute;
    lvh::vec3 lineTangent; float lineRotation;
    lvh::vec3 lineNormal; uint32_t lineStartIndex;
};
static_assert(sizeof(LinePointDataUnified) == 48, "LinePointDataUnified must stay 48 bytes");

struct AABB3 { lvh::vec3 min, max; };

struct TubeAabbRenderData {
    std::vector<uint32_t> indexBuffer;                  // 2 point indices per segment
    std::vector<AABB3> aabbBuffer;                      // one per segment
    std::vector<LinePointDataUnified> linePointDataBuffer;
    // first point of every surviving polyline + one-past-the-end (what getFilteredLines() gives the AO prebaker as separate
    // polylines, VulkanAmbientOcclusionBaker.cpp:482); host-side addition, not part of the reference struct
    std::vector<uint64_t> lineOffsets;
};

class LineData {
public:
    virtual ~LineData() = default;
    void setTrajectoryData(const Trajectories& t) { trajectories = t; dirty = true; cachedValid = false; }
    const Trajectories& getTrajectories() const { return trajectories; }
    void setSelectedAttributeIndex(int i) { selectedAttributeIndex = i; cachedValid = false; }
    size_t getNumLines() const { return trajectories.size(); }
    size_t getNumLinePoints() const { size_t n = 0; for (auto& t : trajectories) n += t.positions.size(); return n; }
    size_t getNumLineSegments() const { size_t n = 0; for (auto& t : trajectories) if (!t.positions.empty()) n += t.positions.size() - 1; return n; }
    bool getUseCappedTubes() const { return useCappedTubes; }
    int getTubeNumSubdivisions() const { return tubeNumSubdivisions; }
    void getMinMaxAttributeValues(float& mn, float& mx) const;
    bool isDirty() const { return dirty; }
    void resetDirty() { dirty = false; }

    // src/LineData/LineDataFlow.cpp:2112-2277 (flow lines, no ribbons / multi-var / helicity)
    const TubeAabbRenderData& getLinePassTubeAabbRenderData(float lineWidth);

    bool useCappedTubes = true;   // src/LineData/LineData.hpp:377
    bool useHalos = true;         // :378
    int tubeNumSubdivisions = 6;  // src/LineData/LineData.cpp:52

protected:
    Trajectories trajectories;
    int selectedAttributeIndex = 0;
    bool dirty = true;
    TubeAabbRenderData cachedTubeAabbRenderData;
    bool cachedValid = false;
    float cachedLineWidth = 0.0f;
};
typedef std::shared_ptr<LineData> LineDataPtr;
