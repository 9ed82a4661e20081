// host_selftest.cpp -- drives the C++ adapter the way MainApp does: construct renderer, setLineData, setNewSettings,
// onResolutionChanged, render.  Without a GPU it verifies that construction fails loudly (no CPU fallback) and that the
// host-side data model (segment builder, SettingsMap) behaves; with a GPU it renders one frame per renderer.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "LineRenderer.hpp"

static void lookAt(const float* eye, float* m) {  // camera at eye looking at the origin, up +y (glm::lookAt, column-major)
    float f[3] = {-eye[0], -eye[1], -eye[2]};
    float fl = std::sqrt(f[0]*f[0] + f[1]*f[1] + f[2]*f[2]);
    for (float& v : f) v /= fl;
    float up[3] = {0, 1, 0};
    float s[3] = {f[1]*up[2] - f[2]*up[1], f[2]*up[0] - f[0]*up[2], f[0]*up[1] - f[1]*up[0]};
    float sl = std::sqrt(s[0]*s[0] + s[1]*s[1] + s[2]*s[2]);
    for (float& v : s) v /= sl;
    float u[3] = {s[1]*f[2] - s[2]*f[1], s[2]*f[0] - s[0]*f[2], s[0]*f[1] - s[1]*f[0]};
    float r[16] = {s[0], u[0], -f[0], 0, s[1], u[1], -f[1], 0, s[2], u[2], -f[2], 0,
                   -(s[0]*eye[0] + s[1]*eye[1] + s[2]*eye[2]), -(u[0]*eye[0] + u[1]*eye[1] + u[2]*eye[2]), f[0]*eye[0] + f[1]*eye[1] + f[2]*eye[2], 1};
    std::memcpy(m, r, 64);
}
static void perspective(float fovy, float aspect, float zn, float zf, float* m) {
    std::memset(m, 0, 64);
    float t = std::tan(fovy / 2);
    m[0] = 1 / (aspect * t); m[5] = 1 / t; m[10] = -(zf + zn) / (zf - zn); m[11] = -1; m[14] = -(2 * zf * zn) / (zf - zn);
}

int main() {
    // data model: one helix + a degenerate trajectory that must vanish
    auto data = std::make_shared<LineData>();
    Trajectories tr(2);
    for (int i = 0; i < 200; i++) {
        float a = 0.1f * i;
        tr[0].positions.push_back({0.2f * std::cos(a), -0.2f + 0.002f * i, 0.2f * std::sin(a)});
    }
    tr[0].attributes.push_back(std::vector<float>(200));
    for (int i = 0; i < 200; i++) tr[0].attributes[0][i] = i / 199.0f;
    tr[1].positions = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    tr[1].attributes.push_back({0, 0, 0});
    data->setTrajectoryData(tr);
    const TubeAabbRenderData& rd = data->getLinePassTubeAabbRenderData(0.01f);
    if (rd.indexBuffer.size() != 2 * 199 || rd.linePointDataBuffer.size() != 200 || rd.aabbBuffer.size() != 199) {
        std::printf("FAIL: segment builder produced %zu indices / %zu points\n", rd.indexBuffer.size(), rd.linePointDataBuffer.size());
        return 1;
    }
    SettingsMap s;
    s.addKeyValue("ambient_occlusion_strength", 1.0f);
    s.addKeyValue("ambient_occlusion_samples_per_frame", 4);
    s.addKeyValue("use_jittered_primary_rays", true);
    s.addKeyValue("depth_cue_strength", 0.0f);
    bool b = false; float f = 0;
    if (!s.getValueOpt("use_jittered_primary_rays", b) || !b || !s.getValueOpt("ambient_occlusion_strength", f) || f != 1.0f) { std::printf("FAIL: SettingsMap\n"); return 1; }

    SceneData sd;
    float eye[3] = {0, 0, 0.8f};
    lookAt(eye, sd.viewMatrix);
    sd.fovY = 2 * std::atan(0.5f);
    perspective(sd.fovY, 160.0f / 96.0f, 0.01f, 100.0f, sd.projectionMatrix);
    std::memcpy(sd.cameraPosition, eye, 12);
    sd.viewportWidth = 160; sd.viewportHeight = 96;
    TransferFunction tf;
    tf.rgba = {0.05f, 0.07f, 0.5f, 1.0f, 0.7f, 0.7f, 0.7f, 1.0f, 0.45f, 0.01f, 0.02f, 1.0f};
    LineRenderer::setLineWidth(0.01f);
    try {
        B200RayTracer rt(&sd, tf);
        rt.onResolutionChanged();
        LineDataPtr d = data;
        rt.setLineData(d, true);
        rt.setNewSettings(s);
        SettingsMap once; once.addKeyValue("num_samples_per_frame", 1); once.addKeyValue("num_accumulated_frames", 1);
        rt.setNewSettings(once);
        rt.render();
        double sum = 0; size_t nonbg = 0;
        for (size_t i = 0; i < sd.sceneTexture.size(); i += 4) { sum += sd.sceneTexture[i]; if (sd.sceneTexture[i] < 0.999f) nonbg++; }
        std::printf("ray tracer: %llu rays, %zu non-background pixels, mean R %.4f\n",
                    (unsigned long long)(rt.getLastStats().rays_primary + rt.getLastStats().rays_ao), nonbg, sum / (sd.sceneTexture.size() / 4));
        if (nonbg == 0 || rt.getLastStats().rays_ao == 0) { std::printf("FAIL: empty frame\n"); return 1; }
        // a --perf state (camelCase keys, VulkanRayTracer.cpp:280-328): resets the accumulation, applies samples / frames; AO keeps frames coming
        InternalState st; st.name = "RT 1spp"; st.rendererSettings.addKeyValue("numSamplesPerFrame", 1); st.rendererSettings.addKeyValue("maxNumAccumulatedFrames", 2);
        st.rendererSettings.addKeyValue("useAnalyticIntersections", true);
        rt.setNewState(st);
        if (rt.getAccumulatedFramesCounter() != 0 || !rt.needsReRender()) { std::printf("FAIL: setNewState\n"); return 1; }
        SettingsMap aoit; aoit.addKeyValue("ambient_occlusion_iterations", 3);
        rt.setNewSettings(aoit);
        rt.LineRenderer::needsReRender();   // consume the "settings changed" flag: only the accumulation logic is left
        int frames = 0;
        while (rt.needsReRender() && frames < 10) { rt.render(); frames++; }
        if (frames != 3) { std::printf("FAIL: %d frames rendered, expected 3 (2 accumulated frames, 3 AO iterations)\n", frames); return 1; }
        bool refused = false;
        InternalState bad; bad.rendererSettings.addKeyValue("geometryMode", std::string("Linear Swept Spheres"));
        try { rt.setNewState(bad); } catch (const std::runtime_error&) { refused = true; }
        if (!refused) { std::printf("FAIL: an unknown geometry mode was not refused\n"); return 1; }
        // the triangle-mesh geometry mode: same scene, the tube pass traces and shades the reference's tube mesh
        InternalState tri; tri.rendererSettings.addKeyValue("geometryMode", std::string("Triangle Mesh"));
        rt.setNewState(tri);
        rt.render();
        size_t nonbgTri = 0;
        for (size_t i = 0; i < sd.sceneTexture.size(); i += 4) if (sd.sceneTexture[i] < 0.999f) nonbgTri++;
        std::printf("triangle-mesh mode: %zu non-background pixels (analytic: %zu)\n", nonbgTri, nonbg);
        if (nonbgTri == 0 || nonbgTri > 2 * nonbg) { std::printf("FAIL: triangle-mesh frame\n"); return 1; }
        InternalState ana; ana.rendererSettings.addKeyValue("geometryMode", std::string("AABBs (analytic)"));
        rt.setNewState(ana);
        rt.setNewSettings(once);
        // object-space AO prebaker: the first frames each run one baking iteration, then only look the factors up
        SettingsMap pre; pre.addKeyValue("ambient_occlusion_mode", std::string("RTAO (Prebaker)")); pre.addKeyValue("b200_prebaker_iterations", 2);
        pre.addKeyValue("b200_prebaker_param_segment_length", 0.01f);
        rt.setNewSettings(pre);
        unsigned long long bakeRays[3];
        for (int i = 0; i < 3; i++) { rt.render(); bakeRays[i] = rt.getLastStats().rays_ao; }
        std::printf("prebaker: AO rays per frame %llu %llu %llu\n", bakeRays[0], bakeRays[1], bakeRays[2]);
        if (bakeRays[0] == 0 || bakeRays[1] != bakeRays[0] || bakeRays[2] != 0) { std::printf("FAIL: prebaker iteration schedule\n"); return 1; }
        for (size_t i = 0; i < sd.sceneTexture.size(); i++) if (!(sd.sceneTexture[i] >= 0.0f && sd.sceneTexture[i] <= 1.0001f)) { std::printf("FAIL: prebaker frame\n"); return 1; }
        B200PerPixelLinkedListLineRenderer pp(&sd, tf);
        pp.onResolutionChanged();
        pp.setLineData(d, true);
        pp.render();
        std::printf("ppll: %llu fragments sorted, max depth complexity %u\n", (unsigned long long)pp.getLastStats().frags_sorted, pp.getLastStats().max_depth_complexity);
        if (pp.getLastStats().frags_sorted == 0) { std::printf("FAIL: no fragments\n"); return 1; }
        std::printf("host adapter OK (GPU)\n");
    } catch (const std::runtime_error& e) {
        if (std::strstr(e.what(), "no CPU fallback")) { std::printf("host adapter OK (no GPU: %s)\n", e.what()); return 0; }
        std::printf("FAIL: %s\n", e.what());
        return 1;
    }
    return 0;
}
