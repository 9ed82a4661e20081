#include "LineData.hpp"

#include <algorithm>
#include <cmath>

using lvh::vec3;
namespace {
vec3 sub(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
float length(vec3 a) { return std::sqrt(dot(a, a)); }
vec3 normalize(vec3 a) { float inv = 1.0f / std::sqrt(dot(a, a)); return {a.x * inv, a.y * inv, a.z * inv}; }
}  // namespace

void LineData::getMinMaxAttributeValues(float& mn, float& mx) const {
    mn = 3.4e38f; mx = -3.4e38f;
    for (auto& t : trajectories)
        if (size_t(selectedAttributeIndex) < t.attributes.size())
            for (float v : t.attributes[selectedAttributeIndex]) { mn = std::min(mn, v); mx = std::max(mx, v); }
    if (mn > mx) { mn = 0.0f; mx = 1.0f; }
}

const TubeAabbRenderData& LineData::getLinePassTubeAabbRenderData(float lineWidth) {
    if (cachedValid && cachedLineWidth == lineWidth) return cachedTubeAabbRenderData;
    TubeAabbRenderData& out = cachedTubeAabbRenderData;
    out = TubeAabbRenderData();
    const float r = lineWidth * 0.5f;
    uint32_t lineSegmentIndexCounter = 0;
    for (const Trajectory& trajectory : trajectories) {
        const size_t n = trajectory.positions.size();
        if (n < 2) continue;
        vec3 lastLineNormal{1.0f, 0.0f, 0.0f};
        uint32_t numValidLinePoints = 0;
        for (size_t i = 0; i < n; i++) {
            vec3 tangent;
            if (i == 0) tangent = sub(trajectory.positions[i + 1], trajectory.positions[i]);
            else if (i + 1 == n) tangent = sub(trajectory.positions[i], trajectory.positions[i - 1]);
            else tangent = sub(trajectory.positions[i + 1], trajectory.positions[i - 1]);
            if (length(tangent) < 0.0001f) continue;  // almost identical vertices: skip (LineDataFlow.cpp:2160-2163)
            tangent = normalize(tangent);
            vec3 helperAxis = lastLineNormal;
            if (length(cross(helperAxis, tangent)) < 0.01f) {
                helperAxis = {0.0f, 1.0f, 0.0f};
                if (length(cross(helperAxis, tangent)) < 0.01f) helperAxis = {0.0f, 0.0f, 1.0f};
            }
            float d = dot(helperAxis, tangent);
            vec3 normal = normalize({helperAxis.x - d * tangent.x, helperAxis.y - d * tangent.y, helperAxis.z - d * tangent.z});  // Gram-Schmidt
            lastLineNormal = normal;
            LinePointDataUnified p{};
            p.linePosition = trajectory.positions[i];
            p.lineAttribute = size_t(selectedAttributeIndex) < trajectory.attributes.size() ? trajectory.attributes[selectedAttributeIndex][i] : 0.0f;
            p.lineTangent = tangent;
            p.lineNormal = normal;
            out.linePointDataBuffer.push_back(p);
            numValidLinePoints++;
        }
        if (numValidLinePoints == 1) out.linePointDataBuffer.pop_back();
        if (numValidLinePoints <= 1) continue;
        out.lineOffsets.push_back(lineSegmentIndexCounter);
        for (uint32_t pointIdx = 1; pointIdx < numValidLinePoints; pointIdx++) {
            out.indexBuffer.push_back(lineSegmentIndexCounter + pointIdx - 1);
            out.indexBuffer.push_back(lineSegmentIndexCounter + pointIdx);
            const vec3& a = out.linePointDataBuffer[lineSegmentIndexCounter + pointIdx - 1].linePosition;
            const vec3& b = out.linePointDataBuffer[lineSegmentIndexCounter + pointIdx].linePosition;
            AABB3 box;
            box.min = {std::min(a.x, b.x) - r, std::min(a.y, b.y) - r, std::min(a.z, b.z) - r};
            box.max = {std::max(a.x, b.x) + r, std::max(a.y, b.y) + r, std::max(a.z, b.z) + r};
            out.aabbBuffer.push_back(box);
        }
        lineSegmentIndexCounter += numValidLinePoints;
    }
    out.lineOffsets.push_back(lineSegmentIndexCounter);
    cachedValid = true;
    cachedLineWidth = lineWidth;
    return out;
}
