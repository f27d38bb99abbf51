// input_host_test.cpp — replays a key script through include/fs3d_input.hpp (no GPU needed) and prints the camera
// pose after every frame as float32 bit patterns, plus the brush voxel; tests/test_input.py compares with the
// reference-generated golden track.
// build: g++ -std=c++17 -ffp-contract=off -Iinclude tests/host/input_host_test.cpp -o input_host_test
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include "fs3d_input.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    std::ifstream f(argv[1]);
    const std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const int frames = std::atoi(argv[2]);
    engine::sim::KeyScript script(text);
    engine::sim::KeyFlags flags;
    engine::sim::CameraController cam;
    for (int frame = 0; frame < frames; ++frame) {
        script.handleEvents((uint32_t)frame, flags);
        cam.integrate(flags);
        uint32_t b[4];
        const float v[4] = {cam.camPos[0], cam.camPos[1], cam.camPos[2], cam.camRot[1]};
        std::memcpy(b, v, sizeof(b));
        int32_t c[3];
        engine::sim::brushCentre(cam, 4.5, 64, 48, 40, c);
        std::printf("%08x %08x %08x %08x  %d %d %d  %d %d %u\n", b[0], b[1], b[2], b[3], c[0], c[1], c[2],
                    (int)flags.holdingPaint, (int)flags.holdingErase, (unsigned)flags.material);
    }
    return 0;
}
