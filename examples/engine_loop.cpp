// engine_loop.cpp — how the reference's frame loop drives libfs3d (headless stand-in).
//
// Mirrors /root/reference/src/engine/engine.cpp:59-70:
//     while (!mWindow.shouldQuit()) { mWindow.handleEvents(); ...; mRenderer.draw(); }
// with world.step() between event handling and drawing, and the offscreen CUDA ray-march standing
// in for Renderer::draw (no Vulkan/SDL in this image).  Camera defaults are the reference's
// (renderer.h:148-149, window.h:41, materials.cpp:540).
//
// build: g++ -std=c++17 -Iinclude examples/engine_loop.cpp -Lfallingsand3d_b200 -lfs3d -Wl,-rpath,$PWD/fallingsand3d_b200 -o engine_loop
#include <cstdio>
#include <fstream>
#include "fs3d.hpp"

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? (uint32_t)std::stoul(argv[1]) : 64;
    const int frames = argc > 2 ? std::stoi(argv[2]) : 500;
    try {
        engine::sim::VoxelWorld world(n, n, n, /*seed=*/1);
        world.generate(FS3D_SCENE_SAND_BLOCK, 1);                  // BASELINE config 1
        const auto h0 = world.histogram();
        fs3d_camera cam{{0.0f, 0.0f, -2.0f}, 0.0f, 1700.0f / 900.0f};
        for (int frame = 0; frame < frames; ++frame) {
            // mWindow.handleEvents();   (paint/erase cells here with world.setCell)
            world.step();                                           // <- the inserted call
            // mRenderer.draw();         (hand-off: world.volumeView() -> device pointer for the ray-march)
            if (frame % 100 == 99) {
                auto img = world.raymarch(cam, 850, 450, FS3D_RM_VOXELS | FS3D_RM_SRGB);
                std::ofstream f("frame_" + std::to_string(frame + 1) + ".ppm", std::ios::binary);
                f << "P6\n850 450\n255\n";
                for (size_t i = 0; i < img.size(); i += 4) f.write((const char *)&img[i], 3);
            }
        }
        world.waitForSimulation();
        const auto h1 = world.histogram();
        std::printf("steps %llu  sand %llu -> %llu  digest %016llx\n", (unsigned long long)world.stepIndex(),
                    (unsigned long long)h0[FS3D_SAND], (unsigned long long)h1[FS3D_SAND], (unsigned long long)world.digest());
        return h0 == h1 ? 0 : 1;
    } catch (const std::runtime_error &) {
        return 2;                                                   // already logged as "ERROR: ..."
    }
}
