// engine_loop.cpp — how the reference's frame loop drives libfs3d (headless stand-in).
//
// Mirrors /root/reference/src/engine/engine.cpp:59-70:
//     while (!mWindow.shouldQuit()) { mWindow.handleEvents(); ...; mRenderer.draw(); }
// with world.step() between event handling and drawing, and the offscreen CUDA ray-march standing
// in for Renderer::draw (no Vulkan/SDL in this image).  Camera defaults are the reference's
// (renderer.h:148-149, window.h:41, materials.cpp:540).
//
//   engine_loop N FRAMES                 BASELINE config 1: a sand block dropped into an N^3 box
//   engine_loop N FRAMES SCRIPT [DIST]   the same loop driven by a recorded key session (include/fs3d_input.hpp):
//                                        handleEvents() = the script's events of this frame, the camera moves as in
//                                        Renderer::draw (renderer.cpp:438-467), PAINT / ERASE apply the brush
//
// build: g++ -std=c++17 -Iinclude examples/engine_loop.cpp -Lfallingsand3d_b200 -lfs3d -Wl,-rpath,$PWD/fallingsand3d_b200 -o engine_loop
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include "fs3d.hpp"
#include "fs3d_input.hpp"

static void writePpm(const std::string &path, const std::vector<uint8_t> &img, uint32_t w, uint32_t h) {
    std::ofstream f(path, std::ios::binary);
    f << "P6\n" << w << " " << h << "\n255\n";
    for (size_t i = 0; i < img.size(); i += 4) f.write((const char *)&img[i], 3);
}

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? (uint32_t)std::stoul(argv[1]) : 64;
    const int frames = argc > 2 ? std::stoi(argv[2]) : 500;
    const bool scripted = argc > 3;
    const double brushDistance = argc > 4 ? std::stod(argv[4]) : 3.7;
    try {
        std::string text;
        if (scripted) {
            std::ifstream f(argv[3]);
            if (!f) engine::sim::displayError(std::string("cannot open key script ") + argv[3]);
            text.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        }
        engine::sim::KeyScript script(text);
        engine::sim::KeyFlags flags;
        engine::sim::CameraController cam;
        const uint32_t ny = scripted ? n * 3 / 4 : n, nz = scripted ? n * 5 / 8 : n;
        engine::sim::VoxelWorld world(n, ny, nz, /*seed=*/scripted ? 5 : 1);
        world.generate(scripted ? FS3D_SCENE_MIXED : FS3D_SCENE_SAND_BLOCK, 1);
        const auto h0 = world.histogram();
        if (!scripted) cam.camPos[2] = -2.0f;
        for (int frame = 0; frame < frames; ++frame) {
            script.handleEvents((uint32_t)frame, flags);            // mWindow.handleEvents();
            cam.integrate(flags);                                   // top of mRenderer.draw(), renderer.cpp:438-467
            if (flags.holdingPaint || flags.holdingErase) {         // builder-defined paint input
                int32_t c[3];
                engine::sim::brushCentre(cam, brushDistance, n, ny, nz, c);
                world.paintSphere(c[0], c[1], c[2], 3, flags.holdingErase ? (uint8_t)FS3D_EMPTY : flags.material, !flags.holdingErase);
            }
            world.step();                                           // <- the inserted call
            // mRenderer.draw();   (hand-off: world.volumeView() -> device pointer for the ray-march)
            if (scripted ? frame % 40 == 39 : frame % 100 == 99) {
                const uint32_t W = scripted ? 170 : 850, H = scripted ? 90 : 450;
                auto img = world.raymarch(cam.camera(), W, H, scripted ? FS3D_RM_VOXELS : (FS3D_RM_VOXELS | FS3D_RM_SRGB));
                writePpm("frame_" + std::to_string(frame + 1) + ".ppm", img, W, H);
            }
        }
        world.waitForSimulation();
        const auto h1 = world.histogram();
        uint32_t bits[4];
        const float pose[4] = {cam.camPos[0], cam.camPos[1], cam.camPos[2], cam.camRot[1]};
        std::memcpy(bits, pose, sizeof(bits));
        std::printf("steps %llu  sand %llu -> %llu  digest %016llx  camera %08x %08x %08x %08x\n", (unsigned long long)world.stepIndex(),
                    (unsigned long long)h0[FS3D_SAND], (unsigned long long)h1[FS3D_SAND], (unsigned long long)world.digest(),
                    bits[0], bits[1], bits[2], bits[3]);
        return (scripted || h0 == h1) ? 0 : 1;
    } catch (const std::runtime_error &) {
        return 2;                                                   // already logged as "ERROR: ..."
    }
}
