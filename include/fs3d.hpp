// fs3d.hpp — header-only C++ wrapper over the C ABI (include/fs3d.h), in the reference engine's
// style: namespace engine::..., RAII, and errors reported the way util::displayError does —
// print "ERROR: <msg>" to stderr, then throw std::runtime_error
// (/root/reference/src/util/debug.cpp:6-27).
//
// What holds one of these in the reference: VulkanEngine, next to mWindow and mRenderer
// (/root/reference/src/engine/engine.h:16-35); what calls step(): VulkanEngine::run between
// handleEvents() and draw() (/root/reference/src/engine/engine.cpp:59-70).  See INTEGRATION.md.
#pragma once
#include <array>
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "fs3d.h"

namespace engine {
namespace sim {

inline void displayError(const std::string &err) {   // mirrors util::displayError (debug.cpp:23-27)
    std::cerr << "ERROR: " << err << std::endl;
    throw std::runtime_error(err);
}

class VoxelWorld {
public:
    enum Material : uint8_t { Empty = FS3D_EMPTY, Sand = FS3D_SAND, Water = FS3D_WATER, Stone = FS3D_STONE };

    VoxelWorld(uint32_t nx, uint32_t ny, uint32_t nz, uint64_t seed = 1, int nGpus = 1, uint32_t flags = 0) {
        fs3d_desc d{};
        d.nx = nx; d.ny = ny; d.nz = nz; d.seed = seed; d.n_gpus = nGpus; d.devices = nullptr; d.flags = flags;
        check(fs3d_create(&d, &mWorld));
        mNx = nx; mNy = ny; mNz = nz;
    }
    ~VoxelWorld() { cleanup(); }
    VoxelWorld(const VoxelWorld &) = delete;
    VoxelWorld &operator=(const VoxelWorld &) = delete;

    void cleanup() { if (mWorld) { fs3d_destroy(mWorld); mWorld = nullptr; } }

    void setCell(uint32_t x, uint32_t y, uint32_t z, uint8_t m) { check(fs3d_set_cell(mWorld, x, y, z, m)); }
    uint8_t getCell(uint32_t x, uint32_t y, uint32_t z) { uint8_t m = 0; check(fs3d_get_cell(mWorld, x, y, z, &m)); return m; }
    void fillBox(std::array<uint32_t, 3> lo, std::array<uint32_t, 3> hi, uint8_t m) { check(fs3d_fill_box(mWorld, lo.data(), hi.data(), m)); }
    void paintSphere(int32_t cx, int32_t cy, int32_t cz, uint32_t radius, uint8_t m, bool onlyEmpty = false) {
        check(fs3d_paint_sphere(mWorld, cx, cy, cz, radius, m, onlyEmpty ? 1 : 0));
    }
    void generate(int sceneId, uint64_t seed) { check(fs3d_generate(mWorld, sceneId, seed)); }
    void upload(const std::vector<uint8_t> &grid) {
        if (grid.size() != (size_t)mNx * mNy * mNz) displayError("VoxelWorld::upload: wrong grid size");
        check(fs3d_upload(mWorld, grid.data()));
    }
    std::vector<uint8_t> download() { std::vector<uint8_t> g((size_t)mNx * mNy * mNz); check(fs3d_download(mWorld, g.data())); return g; }

    void step(uint32_t n = 1) { check(fs3d_step(mWorld, n)); }     // asynchronous, like a queue submit
    void waitForSimulation() { check(fs3d_sync(mWorld)); }          // cf. Renderer::waitForGraphics
    uint64_t stepIndex() { uint64_t s = 0; check(fs3d_step_index(mWorld, &s)); return s; }

    // checkpoint: cells + step index + seed (format in fs3d.h); load() continues the run bit-identically
    void save(const std::string &path) { check(fs3d_save(mWorld, path.c_str())); }
    void load(const std::string &path) { check(fs3d_load(mWorld, path.c_str())); }

    std::array<uint64_t, 256> histogram() { std::array<uint64_t, 256> h{}; check(fs3d_histogram(mWorld, h.data())); return h; }
    uint64_t digest() { uint64_t d = 0; check(fs3d_digest(mWorld, &d)); return d; }

    int numSlabs() { int32_t n = 0; check(fs3d_num_slabs(mWorld, &n)); return n; }
    fs3d_view volumeView(int slab = 0) { fs3d_view v{}; check(fs3d_volume_view(mWorld, slab, &v)); return v; }
    void setPalette(const float *rgba256x4) { check(fs3d_set_palette(mWorld, rgba256x4)); }
    std::vector<uint8_t> raymarch(const fs3d_camera &cam, uint32_t width, uint32_t height, uint32_t mode = FS3D_RM_VOXELS) {
        std::vector<uint8_t> img((size_t)width * height * 4);
        check(fs3d_raymarch(mWorld, &cam, width, height, mode, img.data()));
        return img;
    }

    fs3d_world *handle() { return mWorld; }

private:
    static void check(int rc) { if (rc != FS3D_OK) displayError(std::string("fs3d: ") + fs3d_last_error()); }
    fs3d_world *mWorld = nullptr;
    uint32_t mNx = 0, mNy = 0, mNz = 0;
};

}  // namespace sim
}  // namespace engine
