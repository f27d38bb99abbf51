// camera_ref.cpp — runs the REFERENCE's own camera integration on the CPU.
//
// TEST INFRASTRUCTURE ONLY (oracle/): the Makefile cuts the statements between `void Renderer::draw() {` and the first
// `// Wait until the GPU` comment out of /root/reference/src/engine/rendering/renderer.cpp (they are lines 438-467:
// W/S/A/D/Space/Ctrl at 1.5 u/s, Left/Right at 10 deg/s, fixed dt 0.016) into oracle/_ref/renderer_draw_camera.inc
// (git-ignored build output; no reference source is copied into the repo) and this file supplies what they use:
// `mWindow` with the eight key flags of /root/reference/src/engine/window.h:12-19 and the camPos / camRot members
// of /root/reference/src/engine/rendering/renderer.h:148-149 (glm::vec3, the reference's vendored glm).
#include <glm/glm.hpp>
#include <cstdint>

namespace ref_camera {
struct Window {
    bool holdingW{false}, holdingA{false}, holdingS{false}, holdingD{false};
    bool holdingCTRL{false}, holdingSpace{false}, holdingLeft{false}, holdingRight{false};
};
struct Renderer {
    Window win;
    Window *mWindow = &win;
    glm::vec3 camPos{0, 0, -5};
    glm::vec3 camRot{0, 0, 0};
    void draw() {
#include "renderer_draw_camera.inc"
    }
};
}  // namespace ref_camera

extern "C" {

// flags: W, A, S, D, CTRL, Space, Left, Right (window.h order).  One Renderer::draw() worth of camera motion.
void fs_camera_ref_step(const uint8_t flags[8], float pos[3], float rot[3]) {
    ref_camera::Renderer r;
    r.win.holdingW = flags[0]; r.win.holdingA = flags[1]; r.win.holdingS = flags[2]; r.win.holdingD = flags[3];
    r.win.holdingCTRL = flags[4]; r.win.holdingSpace = flags[5]; r.win.holdingLeft = flags[6]; r.win.holdingRight = flags[7];
    r.camPos = glm::vec3(pos[0], pos[1], pos[2]);
    r.camRot = glm::vec3(rot[0], rot[1], rot[2]);
    r.draw();
    pos[0] = r.camPos.x; pos[1] = r.camPos.y; pos[2] = r.camPos.z;
    rot[0] = r.camRot.x; rot[1] = r.camRot.y; rot[2] = r.camRot.z;
}

// the defaults a freshly constructed Renderer holds (renderer.h:148-149)
void fs_camera_ref_defaults(float pos[3], float rot[3]) {
    ref_camera::Renderer r;
    pos[0] = r.camPos.x; pos[1] = r.camPos.y; pos[2] = r.camPos.z;
    rot[0] = r.camRot.x; rot[1] = r.camRot.y; rot[2] = r.camRot.z;
}

}  // extern "C"
