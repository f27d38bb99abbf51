// common.cuh — shared host/device helpers of libfs3d (no reference counterpart).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda_runtime.h>
#include "../../include/fs3d.h"

namespace fs3d {

// ---- SCHEDULE.md §3 ---------------------------------------------------------------------------
__host__ __device__ inline uint64_t mix64(uint64_t v) {
    v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ull;
    v ^= v >> 27; v *= 0x94D049BB133111EBull;
    v ^= v >> 31;
    return v;
}
__host__ __device__ inline uint32_t step_key(uint64_t seed, uint64_t t, uint32_t axis) {
    uint64_t v = seed ^ (t * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(axis + 1) * 0xD1B54A32D192ED03ull);
    v = mix64(v);
    return (uint32_t)(v ^ (v >> 32));
}
__host__ __device__ inline uint32_t hash3(uint32_t key, uint32_t a, uint32_t y, uint32_t z) {
    uint32_t v = key + a * 0x9E3779B1u + y * 0x85EBCA77u + z * 0xC2B2AE3Du;
    v ^= v >> 16; v *= 0x7FEB352Du;
    v ^= v >> 15; v *= 0x846CA68Bu;
    v ^= v >> 16;
    return v;
}

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define FS3D_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::fs3d::fail(_e == cudaErrorMemoryAllocation ? FS3D_ERR_OOM : FS3D_ERR_CUDA, \
                                std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    } while (0)

}  // namespace fs3d
