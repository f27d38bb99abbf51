// aux_kernels.cuh — scene generation, validation, histogram, digest, box fill (SCHEDULE.md §4-5).
// Not on the per-step hot path.  No reference counterpart.
#pragma once
#include "common.cuh"

namespace fs3d {

// ---- scenes (must match oracle/fs3d_oracle.c: fs3d_oracle_scene_cell) -------------------------
__device__ __constant__ int c_stone_boxes[8][6] = {
    {  8, 28, 20, 22,  8, 28 }, { 36, 56, 20, 22, 36, 56 }, { 30, 34,  1, 30, 30, 34 },
    {  8, 28, 32, 34, 36, 56 }, { 36, 56, 32, 34,  8, 28 }, { 20, 22,  1, 12,  4, 60 },
    {  4, 60,  1, 10, 42, 44 }, { 44, 52,  1,  6, 12, 20 },
};
__device__ __constant__ int c_sand_box[6]  = { 10, 30, 44, 60, 10, 54 };
__device__ __constant__ int c_water_box[6] = { 34, 54, 44, 60, 10, 54 };

__device__ inline bool in_box(int64_t x, int64_t y, int64_t z, int64_t nx, int64_t ny, int64_t nz, const int *b) {
    return x >= b[0] * nx / 64 && x < b[1] * nx / 64 && y >= b[2] * ny / 64 && y < b[3] * ny / 64 &&
           z >= b[4] * nz / 64 && z < b[5] * nz / 64;
}
__device__ inline uint8_t random_cell(uint32_t key, int64_t x, int64_t y, int64_t z) {
    uint32_t u = hash3(key, (uint32_t)x, (uint32_t)y, (uint32_t)z) & 3u;
    return u == 0 ? FS3D_SAND : (u == 1 ? FS3D_WATER : FS3D_EMPTY);
}
__device__ __constant__ int c_gas_box[6]    = { 10, 26, 12, 19, 10, 26 };
__device__ __constant__ int c_oil_box[6]    = { 34, 54, 36, 43, 10, 30 };
__device__ __constant__ int c_honey_box[6]  = { 38, 50, 24, 31, 38, 54 };
__device__ __constant__ int c_gravel_box[6] = { 12, 24, 36, 43, 38, 54 };
__device__ __constant__ uint8_t c_pick8[16] = { FS3D_SAND, FS3D_SAND, FS3D_WATER, FS3D_WATER, FS3D_OIL, FS3D_GAS, FS3D_HONEY, FS3D_GRAVEL,
                                                0, 0, 0, 0, 0, 0, 0, 0 };
__device__ inline uint8_t random8_cell(uint32_t key, int64_t x, int64_t y, int64_t z) {
    return c_pick8[(hash3(key, (uint32_t)x, (uint32_t)y, (uint32_t)z) >> 4) & 15u];
}
__device__ inline uint8_t mixed_cell(int64_t nx, int64_t ny, int64_t nz, int64_t x, int64_t y, int64_t z) {
    int64_t floor_h = ny / 64 > 1 ? ny / 64 : 1;
    if (y < floor_h) return FS3D_STONE;
    for (int i = 0; i < 8; ++i) if (in_box(x, y, z, nx, ny, nz, c_stone_boxes[i])) return FS3D_STONE;
    if (in_box(x, y, z, nx, ny, nz, c_sand_box)) return FS3D_SAND;
    if (in_box(x, y, z, nx, ny, nz, c_water_box)) return FS3D_WATER;
    return FS3D_EMPTY;
}
__device__ inline uint8_t scene_cell(int scene, uint32_t key, int64_t nx, int64_t ny, int64_t nz,
                                     int64_t x, int64_t y, int64_t z) {
    switch (scene) {
    case FS3D_SCENE_SAND_BLOCK:
        return (x >= 3 * nx / 8 && x < 5 * nx / 8 && z >= 3 * nz / 8 && z < 5 * nz / 8 &&
                y >= 5 * ny / 8 && y < 7 * ny / 8) ? FS3D_SAND : FS3D_EMPTY;
    case FS3D_SCENE_MIXED: return mixed_cell(nx, ny, nz, x, y, z);
    case FS3D_SCENE_RANDOM: return random_cell(key, x, y, z);
    case FS3D_SCENE_MIXED_NOISE: {
        uint8_t m = mixed_cell(nx, ny, nz, x, y, z);
        if (m == FS3D_EMPTY && y >= ny / 2) m = random_cell(key, x, y, z);
        return m;
    }
    case FS3D_SCENE_RANDOM8: return random8_cell(key, x, y, z);
    case FS3D_SCENE_MIXED8: {
        uint8_t m = mixed_cell(nx, ny, nz, x, y, z);
        if (m != FS3D_EMPTY) return m;
        if (in_box(x, y, z, nx, ny, nz, c_gas_box)) return FS3D_GAS;
        if (in_box(x, y, z, nx, ny, nz, c_oil_box)) return FS3D_OIL;
        if (in_box(x, y, z, nx, ny, nz, c_honey_box)) return FS3D_HONEY;
        if (in_box(x, y, z, nx, ny, nz, c_gravel_box)) return FS3D_GRAVEL;
        if (y >= ny / 2) m = random8_cell(key, x, y, z);
        return m;
    }
    default: return FS3D_EMPTY;
    }
}

// owned: pointer to local plane 1; one thread writes 16 cells
__global__ void generate_kernel(uint8_t *owned, uint32_t nx, uint32_t ny, uint32_t nzg, uint32_t z0, uint32_t nzl,
                                int scene, uint32_t key) {
    const uint64_t nvec = (uint64_t)nx / 16 * ny * nzl;
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t rowv = nx / 16;
        const uint32_t xv = (uint32_t)(v % rowv);
        const uint64_t ry = v / rowv;
        const uint32_t y = (uint32_t)(ry % ny);
        const uint32_t lz = (uint32_t)(ry / ny);
        uint32_t out[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t wv = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                wv |= (uint32_t)scene_cell(scene, key, nx, ny, nzg, xv * 16 + q * 4 + b, y, z0 + lz) << (8 * b);
            out[q] = wv;
        }
        reinterpret_cast<uint4 *>(owned)[v] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

__global__ void fill_kernel(uint8_t *p, uint64_t nbytes16, uint32_t pattern) {
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < nbytes16; v += (uint64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4 *>(p)[v] = make_uint4(pattern, pattern, pattern, pattern);
}

// flag[0] |= 1 if any byte has a bit of `bad_bits` set (0xFC..: codes > 3, schedule version 1; 0xF8..: codes > 7, version 2)
__global__ void validate_kernel(const uint8_t *p, uint64_t n16, uint32_t bad_bits, uint32_t *flag) {
    uint32_t bad = 0;
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n16; v += (uint64_t)gridDim.x * blockDim.x) {
        uint4 q = reinterpret_cast<const uint4 *>(p)[v];
        bad |= (q.x | q.y | q.z | q.w) & bad_bits;
    }
    if (__any_sync(0xFFFFFFFFu, bad != 0) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

// counts[256] += histogram of n16*16 bytes
__global__ void histogram_kernel(const uint8_t *p, uint64_t n16, unsigned long long *counts) {
    __shared__ unsigned int sh[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t iter = 0;
    for (; v < n16; v += (uint64_t)gridDim.x * blockDim.x) {
        uint4 q = reinterpret_cast<const uint4 *>(p)[v];
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((w[k] & 0xF8F8F8F8u) == 0) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t m = (w[k] >> (8 * b)) & 7u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) c[q] += m == (uint32_t)q ? 1u : 0u;      // registers, not local memory
                }
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) atomicAdd(&sh[(w[k] >> (8 * b)) & 0xFFu], 1u);
            }
        }
        if (++iter == (1u << 20)) {  // keep 32-bit counters from overflowing on huge grids
#pragma unroll
            for (int m = 0; m < 8; ++m) { atomicAdd(&counts[m], (unsigned long long)c[m]); c[m] = 0; }
            iter = 0;
        }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        unsigned int s = c[m];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&counts[m], (unsigned long long)s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&counts[i], (unsigned long long)sh[i]);
}

// *out += Σ_{m != 0} mix64(8·(base + i) + m)   (SCHEDULE.md §4)
__global__ void digest_kernel(const uint8_t *p, uint64_t n16, uint64_t base, unsigned long long *out) {
    uint64_t sum = 0;
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n16; v += (uint64_t)gridDim.x * blockDim.x) {
        uint4 q = reinterpret_cast<const uint4 *>(p)[v];
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
        if ((q.x | q.y | q.z | q.w) == 0) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint32_t m = (w[k] >> (8 * b)) & 0xFFu;
                if (m) sum += mix64(8ull * (base + v * 16 + k * 4 + b) + m);
            }
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(out, (unsigned long long)sum);
}

// checkpoint payload (fs3d.h "checkpoint"): 4 voxels per byte, voxel i of the stream in bits 2(i & 3) of byte i >> 2.
// One thread packs 16 voxels (one uint4) into one uint32 / unpacks one uint32 into a uint4.
__global__ void pack2_kernel(const uint8_t *cells, uint64_t n16, uint32_t *packed) {
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n16; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 q = reinterpret_cast<const uint4 *>(cells)[v];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = (w[k] & 3u) | (((w[k] >> 8) & 3u) << 2) | (((w[k] >> 16) & 3u) << 4) | (((w[k] >> 24) & 3u) << 6);
            out |= b << (8 * k);
        }
        packed[v] = out;
    }
}
__global__ void unpack2_kernel(const uint32_t *packed, uint64_t n16, uint8_t *cells) {
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n16; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t in = packed[v];
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = (in >> (8 * k)) & 0xFFu;
            w[k] = (b & 3u) | (((b >> 2) & 3u) << 8) | (((b >> 4) & 3u) << 16) | (((b >> 6) & 3u) << 24);
        }
        reinterpret_cast<uint4 *>(cells)[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// schedule version 2 checkpoints (encoding 2): 2 voxels per byte, voxel i of the stream in bits 4(i & 1) of byte i >> 1.
// One thread packs 16 voxels (one uint4) into one uint2 / unpacks one uint2 into a uint4.
__global__ void pack4_kernel(const uint8_t *cells, uint64_t n16, uint2 *packed) {
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n16; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 q = reinterpret_cast<const uint4 *>(cells)[v];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)     // bytes b0 b1 b2 b3 -> 16 bits: b0 | b1 << 4 | b2 << 8 | b3 << 12
            h[k] = (w[k] & 15u) | (((w[k] >> 8) & 15u) << 4) | (((w[k] >> 16) & 15u) << 8) | (((w[k] >> 24) & 15u) << 12);
        packed[v] = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
    }
}
__global__ void unpack4_kernel(const uint2 *packed, uint64_t n16, uint8_t *cells) {
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n16; v += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 in = packed[v];
        const uint32_t h[4] = {in.x & 0xFFFFu, in.x >> 16, in.y & 0xFFFFu, in.y >> 16};
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = (h[k] & 15u) | (((h[k] >> 4) & 15u) << 8) | (((h[k] >> 8) & 15u) << 16) | (((h[k] >> 12) & 15u) << 24);
        reinterpret_cast<uint4 *>(cells)[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// fill the part of a global box that lies in this slab; owned = local plane 1
__global__ void fill_box_kernel(uint8_t *owned, uint32_t nx, uint32_t ny, uint32_t z0, uint32_t nzl,
                                uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, uint32_t zb0, uint32_t zb1, uint8_t m) {
    const uint64_t bx = x1 - x0, by = y1 - y0, bz = zb1 - zb0;
    const uint64_t n = bx * by * bz;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x = x0 + (uint32_t)(i % bx);
        uint32_t y = y0 + (uint32_t)((i / bx) % by);
        uint32_t z = zb0 + (uint32_t)(i / (bx * by));
        if (z >= z0 && z < z0 + nzl) owned[x + (uint64_t)nx * (y + (uint64_t)ny * (z - z0))] = m;
    }
}

// brush for the engine's paint/erase input: every cell of this slab whose centre lies within `radius` cells of
// (cx, cy, cz) becomes m; with only_empty, cells that already hold a material are left alone
__global__ void paint_sphere_kernel(uint8_t *owned, uint32_t nx, uint32_t ny, uint32_t nzg, uint32_t z0, uint32_t nzl,
                                    int64_t cx, int64_t cy, int64_t cz, int64_t radius, uint8_t m, int only_empty) {
    const int64_t x0 = cx - radius < 0 ? 0 : cx - radius, x1 = cx + radius >= nx ? (int64_t)nx - 1 : cx + radius;
    const int64_t y0 = cy - radius < 0 ? 0 : cy - radius, y1 = cy + radius >= ny ? (int64_t)ny - 1 : cy + radius;
    const int64_t zb0 = cz - radius < 0 ? 0 : cz - radius, zb1 = cz + radius >= nzg ? (int64_t)nzg - 1 : cz + radius;
    if (x1 < x0 || y1 < y0 || zb1 < zb0) return;
    const uint64_t bx = x1 - x0 + 1, by = y1 - y0 + 1, bz = zb1 - zb0 + 1, n = bx * by * bz;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t x = x0 + (int64_t)(i % bx), y = y0 + (int64_t)((i / bx) % by), z = zb0 + (int64_t)(i / (bx * by));
        const int64_t dx = x - cx, dy = y - cy, dz = z - cz;
        if (dx * dx + dy * dy + dz * dz > radius * radius) continue;
        if (z < (int64_t)z0 || z >= (int64_t)z0 + nzl) continue;
        uint8_t *c = owned + x + (uint64_t)nx * (y + (uint64_t)ny * (z - z0));
        if (!only_empty || *c == FS3D_EMPTY) *c = m;
    }
}

}  // namespace fs3d
