// fs3d_input.hpp — the reference engine's input state and camera integration, plus a paint brush, for driving a
// VoxelWorld from the frame loop.  Header-only C++, engine style.
//
// Mirrors:
//   /root/reference/src/engine/window.h:12-19     the eight public key flags of engine::Window
//   /root/reference/src/engine/window.cpp:34-107  SDL KEYDOWN/KEYUP -> flag (W A S D LCTRL SPACE LEFT RIGHT)
//   /root/reference/src/engine/rendering/renderer.cpp:438-467  camera integration at the top of Renderer::draw:
//        W/S: camPos.z +/-,  A/D: camPos.x +/-,  Space: camPos.y -,  Ctrl: camPos.y +   at 1.5 units/s,
//        Right/Left: camRot.y +/- at 10 degrees/s, each `else if` pair exclusive, FIXED dt = 0.016 per frame
//   /root/reference/src/engine/rendering/renderer.h:148-149    camPos {0, 0, -5}, camRot {0, 0, 0}
// tests/test_input.py replays a key script through this header (compiled with g++), through the Python mirror
// (fallingsand3d_b200/input.py) and through the reference's own lines compiled into oracle/_ref (camera_ref.cpp):
// the three camera tracks must be bit-identical floats.
//
// Builder-defined (the reference has no paint input): paint / erase flags and a material selector; the brush sits
// `brushDistance` world units in front of the camera along the view direction of the centre pixel and is applied
// with fs3d_paint_sphere.
#pragma once
#include <cmath>
#include <cstdint>
#include <sstream>
#include <string>
#include <vector>

#include "fs3d.h"

namespace engine {
namespace sim {

struct KeyFlags {                    // window.h:12-19, same names
    bool holdingW{false}, holdingA{false}, holdingS{false}, holdingD{false};
    bool holdingCTRL{false}, holdingSpace{false}, holdingLeft{false}, holdingRight{false};
    // builder-defined paint input
    bool holdingPaint{false}, holdingErase{false};
    uint8_t material{FS3D_SAND};

    // one key event, named like the SDL scancode without its prefix: "W", "A", "S", "D", "LCTRL", "SPACE", "LEFT",
    // "RIGHT" (window.cpp:51-104), plus "PAINT", "ERASE" and "1".."7" (select material).  Unknown keys are ignored,
    // as the reference's switch ignores them.
    void onKey(const std::string &key, bool down) {
        if (key == "W") holdingW = down; else if (key == "A") holdingA = down; else if (key == "S") holdingS = down;
        else if (key == "D") holdingD = down; else if (key == "LCTRL") holdingCTRL = down;
        else if (key == "SPACE") holdingSpace = down; else if (key == "LEFT") holdingLeft = down;
        else if (key == "RIGHT") holdingRight = down; else if (key == "PAINT") holdingPaint = down;
        else if (key == "ERASE") holdingErase = down;
        else if (down && key.size() == 1 && key[0] >= '1' && key[0] <= '7') material = (uint8_t)(key[0] - '0');
    }
};

struct CameraController {
    float camPos[3] = {0.0f, 0.0f, -5.0f};   // renderer.h:148
    float camRot[3] = {0.0f, 0.0f, 0.0f};    // renderer.h:149 (degrees; only .y is ever changed)
    float camMoveSpeed = 1.5f;                // renderer.cpp:438
    float camRotSpeed = 10.0f;                // renderer.cpp:439

    // renderer.cpp:441-467 with the reference's fixed time step
    void integrate(const KeyFlags &k) { integrate(k, 0.016f); }
    // ... or with a measured frame time ("real delta-time instead of the fixed 0.016")
    void integrate(const KeyFlags &k, float dt) {
        if (k.holdingW) camPos[2] += camMoveSpeed * dt; else if (k.holdingS) camPos[2] -= camMoveSpeed * dt;
        if (k.holdingA) camPos[0] += camMoveSpeed * dt; else if (k.holdingD) camPos[0] -= camMoveSpeed * dt;
        if (k.holdingSpace) camPos[1] -= camMoveSpeed * dt; else if (k.holdingCTRL) camPos[1] += camMoveSpeed * dt;
        if (k.holdingRight) camRot[1] += camRotSpeed * dt; else if (k.holdingLeft) camRot[1] -= camRotSpeed * dt;
    }

    fs3d_camera camera(float aspect = 1700.0f / 900.0f) const {      // materials.cpp:540
        fs3d_camera c{};
        c.pos[0] = camPos[0]; c.pos[1] = camPos[1]; c.pos[2] = camPos[2];
        c.yaw_deg = camRot[1];
        c.aspect = aspect;
        return c;
    }
};

// Voxel under the point `distance` world units in front of the camera (view direction of the centre pixel: (0,0,1)
// turned by the yaw, raymarch.cuh).  The volume is the box the ray-marcher draws: voxel edge h = 1 / max(nx, ny, nz),
// centred at the origin, grid +y pointing to world -y.  Evaluated in double so that the C++ and Python sides agree.
inline void brushCentre(const CameraController &c, double distance, uint32_t nx, uint32_t ny, uint32_t nz, int32_t out[3]) {
    const double yaw = (double)c.camRot[1] * 3.14159265358979323846 / 180.0;
    const double px = (double)c.camPos[0] + distance * std::sin(yaw);
    const double py = (double)c.camPos[1];
    const double pz = (double)c.camPos[2] + distance * std::cos(yaw);
    const uint32_t nmax = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
    const double h = 1.0 / (double)nmax;
    out[0] = (int32_t)std::floor((px + 0.5 * nx * h) / h);
    out[1] = (int32_t)ny - 1 - (int32_t)std::floor((py + 0.5 * ny * h) / h);
    out[2] = (int32_t)std::floor((pz + 0.5 * nz * h) / h);
}

// A recorded input session: "frame key down|up" per line ('#' starts a comment) — what Window::handleEvents would
// have seen, frame by frame.  Replaying it drives the camera and the brush without a window.
class KeyScript {
public:
    struct Event { uint32_t frame; std::string key; bool down; };
    explicit KeyScript(const std::string &text) {
        std::istringstream in(text);
        std::string line;
        while (std::getline(in, line)) {
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line.resize(hash);
            std::istringstream ls(line);
            Event e; std::string what;
            if (ls >> e.frame >> e.key >> what) { e.down = what == "down"; mEvents.push_back(e); }
        }
    }
    // applies the events of `frame` (call once per frame, frames ascending) — the stand-in for handleEvents()
    void handleEvents(uint32_t frame, KeyFlags &flags) {
        while (mNext < mEvents.size() && mEvents[mNext].frame <= frame) { flags.onKey(mEvents[mNext].key, mEvents[mNext].down); ++mNext; }
    }
    uint32_t lastFrame() const { return mEvents.empty() ? 0u : mEvents.back().frame; }
private:
    std::vector<Event> mEvents;
    size_t mNext = 0;
};

}  // namespace sim
}  // namespace engine
