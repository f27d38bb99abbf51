// fs_raymarch_ref.cpp — runs the REFERENCE's own fragment shader on the CPU.
//
// TEST INFRASTRUCTURE ONLY (oracle/): this wrapper compiles /root/reference/shaders/fs_raymarch.frag itself — not a
// restatement — as C++ against the reference's vendored glm (/root/reference/third_party/glm, 0.9.9.7).  The
// Makefile strips the two kinds of line C++ cannot parse (`#version`, `layout (...) in/out ...;`) into
// oracle/_ref/fs_raymarch_frag.inc (git-ignored build output; no reference source is copied into the repo) and
// this file supplies what those lines declared: the stage inputs/outputs as globals, GLSL's `in` parameter
// qualifier as an empty macro, and the two mixed int/float vector operators GLSL allows and glm does not.
// Built with -fsingle-precision-constant (GLSL literals are float) and -ffp-contract=off.
//
// The frame loop reproduces what the engine's full-screen quad feeds the shader
// (/root/reference/src/engine/rendering/renderer.cpp:1253-1267 and Vulkan's y-down NDC): pixel centre (px, py)
// -> inUV = (1 - (px + .5) / W, (py + .5) / H); fs_raymarch.vert:30-37 passes camPos.xyz and aspect.x through.
#define GLM_FORCE_INTRINSICS   // puts glm 0.9.9 in the language mode where GLM_FORCE_SWIZZLE gives `.xyy` members
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cstdint>

namespace ref_shader {
using namespace glm;
static thread_local vec3 inFragOrigin;
static thread_local vec2 inUV;
static thread_local float inAspect;
static thread_local vec4 outFragColor;
inline vec2 operator*(const vec2 &v, int s) { return v * float(s); }   // `inUV * 2`
inline vec2 operator-(const vec2 &v, int s) { return v - float(s); }   // `... - 1`
#define in
#define main fs_raymarch_main
#include "fs_raymarch_frag.inc"
#undef main
#undef in
}  // namespace ref_shader

extern "C" {

// linear RGBA float of one pixel
void fs_raymarch_ref_pixel(const float origin[3], float aspect, uint32_t W, uint32_t H, uint32_t px, uint32_t py, float rgba[4]) {
    using namespace ref_shader;
    inFragOrigin = glm::vec3(origin[0], origin[1], origin[2]);
    inAspect = aspect;
    inUV = glm::vec2(1.0f - ((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
    fs_raymarch_main();
    rgba[0] = outFragColor.x; rgba[1] = outFragColor.y; rgba[2] = outFragColor.z; rgba[3] = outFragColor.w;
}

// whole frame, linear RGBA float, row-major (H, W, 4)
void fs_raymarch_ref_frame(const float origin[3], float aspect, uint32_t W, uint32_t H, float *rgba) {
#pragma omp parallel for schedule(static)
    for (int64_t py = 0; py < (int64_t)H; ++py)
        for (uint32_t px = 0; px < W; ++px)
            fs_raymarch_ref_pixel(origin, aspect, W, H, px, (uint32_t)py, rgba + 4 * ((size_t)py * W + px));
}

}  // extern "C"
