// step_kernel.cuh — the fused sm_100a voxel step (SCHEDULE.md, schedule version 1).
//
// No reference counterpart exists; the call site this serves is the frame loop
// /root/reference/src/engine/engine.cpp:59-70.
//
// Design:
//  * One WARP is an autonomous unit.  It owns one z-pair of rows (z_l, z_l+1) — exactly the
//    ZY-block pair of this step — across the full x extent, and marches along y.  Because the
//    rows it owns are the ZY block and the XY blocks live inside a row, a step needs NO halo in
//    x or z and none in y beyond a 2-plane lead-in per march segment: every byte of the source
//    buffer is read once and every byte of the destination written once (2 B / voxel-update).
//  * Each lane loads one 32-byte sector (32 voxels) per row per plane with a 256-bit LDG, so a
//    warp request is 1 KiB contiguous; stores are 256-bit STG of whole sectors.
//  * The 32 bytes are bit-sliced into two 32-bit planes (bit0, bit1 of the material code) with
//    voxel x at bit 8·(x&3) + ((x>>2)&7) — the transpose that makes byte<->plane conversion a
//    few shifts/ands.  All rule evaluation is then 32 voxels per logic instruction.
//  * XY blocks pair x with x±1: inside a word that is a byte permute (PRMT); across words the
//    single edge bit comes from the neighbouring lane by warp shuffle.
//  * Loads for the next plane pair are issued before the current pair is evaluated
//    (register double-buffering), so ~J·4 KiB per warp is always in flight.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "bitslice.cuh"

namespace fs3d {

struct StepParams {
    const uint8_t *src;      // slab buffer incl. ghost planes; local plane 0 = ghost-low
    uint8_t *dst;
    uint32_t nx, ny;
    uint32_t wpr;            // 32-voxel words per row (nx / 32)
    uint32_t lpr;            // lanes per row group (J == 1: min(wpr, 32)); else 32
    uint32_t groups;         // z-pairs handled side by side in one warp (J == 1 only), else 1
    int32_t  z0;             // global z of local plane 1 (first owned plane)
    uint32_t nzl;            // owned planes
    uint32_t lz_first;       // local plane index of the left row of pair 0 (0 or 1)
    uint32_t pair_begin, pair_end;   // pairs [begin, end) handled by this launch
    uint32_t nit;            // march iterations per pair: ny / 2 + 1
    uint32_t key_xy, key_zy; // SCHEDULE.md §3 key(seed, t, axis)
    // settled-tile skipping (nullptr = off)
    const uint8_t *skip;     // [ztiles][ytiles] 1 = tile provably static this step
    uint32_t *last_active;   // [ztiles][ytiles] step+1 of last enabled block
    uint32_t ytile_log2;     // y-tile height = 1 << ytile_log2 (even, >= 4)
    uint32_t ztile_log2;     // z-tile depth (local planes, owned index) = 1 << ztile_log2
    uint32_t nytiles;
    uint32_t step_plus1;     // (uint32)(t + 1)
};

__device__ __forceinline__ void ld256(const uint8_t *p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void st256(uint8_t *p, const uint32_t (&r)[8]) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

template <int J>
struct Raw { uint32_t w[J][2][2][8]; };   // [word][row l/r][plane lo/hi][8 x u32]

// ---------------------------------------------------------------------------------------------
template <int J, int OX, int TODD, int THREADS>
__global__ void __launch_bounds__(THREADS) step_kernel(const StepParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;

    uint32_t g = 0, xw0 = lane;
    bool lane_ok = true;
    if (J == 1) { g = lane / p.lpr; xw0 = lane - g * p.lpr; lane_ok = g < p.groups; }

    const uint32_t npairs = p.pair_end - p.pair_begin;
    const uint32_t npg = (npairs + p.groups - 1) / p.groups;
    const uint64_t total = (uint64_t)npg * p.nit;
    uint64_t pos = total * gw / nw;
    const uint64_t end = total * (gw + 1) / nw;

    const size_t row_bytes = p.nx;
    const size_t plane_rows = p.ny;

    while (pos < end) {
        const uint32_t pg = (uint32_t)(pos / p.nit);
        const uint32_t it_a = (uint32_t)(pos - (uint64_t)pg * p.nit);
        const uint64_t left = end - pos;
        const uint32_t it_b = (left < (uint64_t)(p.nit - it_a)) ? it_a + (uint32_t)left : p.nit;
        pos += it_b - it_a;

        const uint32_t pair = p.pair_begin + pg * p.groups + g;
        const bool pair_ok = lane_ok && pair < p.pair_end;
        const uint32_t lzl = p.lz_first + 2u * pair;          // local plane of the left row
        const int32_t zgl = p.z0 + (int32_t)lzl - 1;          // its global z
        bool own[2];
        own[0] = pair_ok && lzl >= 1u && lzl <= p.nzl;
        own[1] = pair_ok && lzl + 1u >= 1u && lzl + 1u <= p.nzl;
        // rows beyond the allocation (lzl + 1 > nzl + 1) cannot occur: pairs are enumerated so
        // that lzl <= nzl, hence lzl + 1 <= nzl + 1 = ghost-high.

        bool wok[J];                 // this lane's word j exists
        uint32_t xw[J];
        bool hasp[J], hasn[J];
        uint32_t hxy[J][2], hzy[J];  // linear hash parts
        const uint8_t *srow[J][2];
        uint8_t *drow[J][2];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            xw[j] = xw0 + 32u * j;
            wok[j] = pair_ok && xw[j] < p.wpr;
            hasp[j] = xw[j] > 0u;
            hasn[j] = xw[j] + 1u < p.wpr;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const size_t off = (size_t)(lzl + r) * plane_rows * row_bytes + (size_t)xw[j] * 32u;
                srow[j][r] = p.src + off;
                drow[j][r] = p.dst + off;
                hxy[j][r] = p.key_xy + xw[j] * HC1 + (uint32_t)(zgl + r) * HC3;
            }
            hzy[j] = p.key_zy + xw[j] * HC1 + (uint32_t)zgl * HC3;
        }

        Raw<J> raw;
        auto load_pair = [&](uint32_t it) {     // planes y1 = 2·it (lo) and y1 + 1 (hi)
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t y = 2u * it + h;
                        if (wok[j] && y < p.ny) {
                            ld256(srow[j][r] + (size_t)y * row_bytes, raw.w[j][r][h]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) raw.w[j][r][h][q] = 0x03030303u;   // STONE
                        }
                    }
        };

        P2 prev[J][2], lo[J][2], hi[J][2];
#pragma unroll
        for (int j = 0; j < J; ++j) { prev[j][0] = {ONES, ONES}; prev[j][1] = {ONES, ONES}; }

        // XY sub-step on (upper, lower) for both rows, upper row is plane yu
        auto do_xy = [&](P2 (&up)[J][2], P2 (&lw)[J][2], uint32_t yu) {
            uint32_t en = 0;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint32_t rw[J], e[J], ep[J], enx[J];
#pragma unroll
                for (int j = 0; j < J; ++j) rw[j] = hash_word(hxy[j][r] + yu * HC2);
                if (OX == 1) {
#pragma unroll
                    for (int j = 0; j < J; ++j) e[j] = edge_pack(up[j][r], lw[j][r], rw[j]);
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        uint32_t a = __shfl_up_sync(ONES, e[j], 1);
                        uint32_t b = __shfl_down_sync(ONES, e[j], 1);
                        if (J > 1) {
                            if (j > 0)     { uint32_t t = __shfl_sync(ONES, e[j - 1], 31); if (lane == 0)  a = t; }
                            if (j < J - 1) { uint32_t t = __shfl_sync(ONES, e[j + 1], 0);  if (lane == 31) b = t; }
                        }
                        ep[j]  = hasp[j] ? a : EDGE_STONE;
                        enx[j] = hasn[j] ? b : EDGE_STONE;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < J; ++j) { ep[j] = EDGE_STONE; enx[j] = EDGE_STONE; }
                }
#pragma unroll
                for (int j = 0; j < J; ++j) en |= xy_substep<OX>(up[j][r], lw[j][r], rw[j], ep[j], enx[j]);
            }
            return en;
        };
        // ZY sub-step across the two rows, upper row is plane yu
        auto do_zy = [&](P2 (&up)[J][2], P2 (&lw)[J][2], uint32_t yu) {
            uint32_t en = 0;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                uint32_t rw = hash_word(hzy[j] + yu * HC2);
                en |= block_rule(up[j][0], up[j][1], lw[j][0], lw[j][1], rw);
            }
            return en;
        };

        // lead-in: sub-step 1 of the pair below the segment gives `prev` (plane 2·it_a − 1)
        if (it_a > 0) {
            load_pair(it_a - 1);
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) { lo[j][r] = pack(raw.w[j][r][0]); hi[j][r] = pack(raw.w[j][r][1]); }
            const uint32_t y1 = 2u * (it_a - 1);
            if (TODD == 0) do_xy(hi, lo, y1 + 1); else do_zy(hi, lo, y1 + 1);
#pragma unroll
            for (int j = 0; j < J; ++j) { prev[j][0] = hi[j][0]; prev[j][1] = hi[j][1]; }
        }

        load_pair(it_a);
        for (uint32_t it = it_a; it < it_b; ++it) {
            const uint32_t y1 = 2u * it;
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) { lo[j][r] = pack(raw.w[j][r][0]); hi[j][r] = pack(raw.w[j][r][1]); }
            if (it + 1 < it_b) load_pair(it + 1);          // in flight while we evaluate this pair

            uint32_t en;
            if (TODD == 0) { en = do_xy(hi, lo, y1 + 1); en |= do_zy(lo, prev, y1); }
            else           { en = do_zy(hi, lo, y1 + 1); en |= do_xy(lo, prev, y1); }
            (void)en;

            // planes y1 − 1 (prev) and y1 (lo) are final
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (wok[j] && own[r]) {
                        uint32_t o[8];
                        if (y1 >= 1u) { unpack(prev[j][r], o); st256(drow[j][r] + (size_t)(y1 - 1) * row_bytes, o); }
                        if (y1 < p.ny) { unpack(lo[j][r], o);   st256(drow[j][r] + (size_t)y1 * row_bytes, o); }
                    }
                }
#pragma unroll
            for (int j = 0; j < J; ++j) { prev[j][0] = hi[j][0]; prev[j][1] = hi[j][1]; }
        }
    }
}

}  // namespace fs3d
