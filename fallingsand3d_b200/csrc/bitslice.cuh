// bitslice.cuh — the bit-sliced block rule of SCHEDULE.md on 32 voxels per 32-bit word.
//
// Pure functions only (no memory access, no intrinsics that need a GPU) so the same source
// compiles under nvcc for the kernels and under g++ for tests/host/bitslice_host_test.cpp.
// No reference counterpart.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FS3D_HD __host__ __device__ __forceinline__
#else
#define FS3D_HD inline
#endif

namespace fs3d {

FS3D_HD uint32_t byte_perm_2301(uint32_t P) {   // swap bytes 0<->1 and 2<->3
#if defined(__CUDA_ARCH__)
    return __byte_perm(P, 0, 0x2301);
#else
    return ((P & 0x00FF00FFu) << 8) | ((P >> 8) & 0x00FF00FFu);
#endif
}

constexpr uint32_t ONES = 0xFFFFFFFFu;
constexpr uint32_t HC1 = 0x9E3779B1u, HC2 = 0x85EBCA77u, HC3 = 0xC2B2AE3Du;

struct P2 { uint32_t p0, p1; };   // two bit-planes of 32 voxels

// 32 bytes (codes 0..3) -> two bit-planes, voxel x = 4k + i (word k, byte i) at bit 8i + k.
FS3D_HD P2 pack(const uint32_t (&w)[8]) {
    uint32_t te = w[0] + (w[2] << 2) + (w[4] << 4) + (w[6] << 6);   // even k: 2-bit codes at 2·(k/2)
    uint32_t to = w[1] + (w[3] << 2) + (w[5] << 4) + (w[7] << 6);   // odd k
    P2 c;
    c.p0 = (te & 0x55555555u) | ((to << 1) & 0xAAAAAAAAu);
    c.p1 = ((te >> 1) & 0x55555555u) | (to & 0xAAAAAAAAu);
    return c;
}
FS3D_HD void unpack(P2 c, uint32_t (&w)[8]) {
    uint32_t te = (c.p0 & 0x55555555u) | ((c.p1 << 1) & 0xAAAAAAAAu);
    uint32_t to = ((c.p0 >> 1) & 0x55555555u) | (c.p1 & 0xAAAAAAAAu);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        w[2 * a]     = (te >> (2 * a)) & 0x03030303u;
        w[2 * a + 1] = (to >> (2 * a)) & 0x03030303u;
    }
}

FS3D_HD uint32_t hash_word(uint32_t v) {   // SCHEDULE.md §3 H, after the linear part
    v ^= v >> 16; v *= 0x7FEB352Du;
    v ^= v >> 15; v *= 0x846CA68Bu;
    v ^= v >> 16;
    return v;
}

// heavier(u, l): u in {SAND, WATER}, l in {EMPTY, WATER}, density(u) > density(l)
FS3D_HD uint32_t heavier(P2 u, P2 l) {
    return (u.p0 ^ u.p1) & (u.p0 | ~l.p1) & ~l.p0;
}
FS3D_HD void cswap(uint32_t m, P2 &x, P2 &y) {
    uint32_t t0 = (x.p0 ^ y.p0) & m, t1 = (x.p1 ^ y.p1) & m;
    x.p0 ^= t0; y.p0 ^= t0; x.p1 ^= t1; y.p1 ^= t1;
}

// The block rule F, D, L on 32 blocks at once. a,b upper row; c,d lower row; r = coin bits.
// Returns the mask of enabled blocks (coin ignored).
FS3D_HD uint32_t block_rule(P2 &a, P2 &b, P2 &c, P2 &d, uint32_t r) {
    uint32_t fa = heavier(a, c); cswap(fa, a, c);
    uint32_t fb = heavier(b, d); cswap(fb, b, d);
    uint32_t da = heavier(a, d) & ~(b.p0 & b.p1);
    uint32_t db = heavier(b, c) & ~(a.p0 & a.p1);
    cswap(da, a, d); cswap(db, b, c);
    uint32_t la = ~a.p0 & a.p1 & ~b.p0 & ~b.p1;
    uint32_t lb = ~b.p0 & b.p1 & ~a.p0 & ~a.p1;
    uint32_t l = la | lb;
    cswap(l & r, a, b);
    return fa | fb | da | db | l;
}

// x-partner of every voxel of a word for XY blocks with x-origin parity OX.
//  OX = 0: pairs (4k, 4k+1), (4k+2, 4k+3): swap bytes 0<->1, 2<->3.
//  OX = 1: pairs (4k+1, 4k+2) [bytes 1<->2] and (4k+3, 4k+4) [byte 3 bit k <-> byte 0 bit k+1];
//          pbit = bit 31 of the previous word, nbit = bit 0 of the next word.
template <int OX>
FS3D_HD uint32_t xpartner(uint32_t P, uint32_t pbit, uint32_t nbit) {
    if (OX == 0) return byte_perm_2301(P);
    uint32_t q0 = ((P >> 23) & 0xFEu) | pbit;
    uint32_t q3 = ((P & 0xFFu) >> 1) | (nbit << 7);
    return q0 | ((P >> 8) & 0x0000FF00u) | ((P << 8) & 0x00FF0000u) | (q3 << 24);
}
template <int OX> FS3D_HD constexpr uint32_t leftmask() { return OX == 0 ? 0x00FF00FFu : 0xFF00FF00u; }

// edge word exchanged between neighbouring lanes for OX = 1:
//  bits 0..3 = bit 0 of (U.p0, U.p1, L.p0, L.p1); bits 4..7 = bit 31 of the same; bit 8 = coin bit 31
constexpr uint32_t EDGE_STONE = 0xFFu;
FS3D_HD uint32_t edge_pack(P2 U, P2 L, uint32_t rw) {
    return (U.p0 & 1u) | ((U.p1 & 1u) << 1) | ((L.p0 & 1u) << 2) | ((L.p1 & 1u) << 3) |
           ((U.p0 >> 31) << 4) | ((U.p1 >> 31) << 5) | ((L.p0 >> 31) << 6) | ((L.p1 >> 31) << 7) |
           ((rw >> 31) << 8);
}

// XY sub-step on one row word: U = upper plane (y0+1), L = lower plane (y0); rw = coin word of the
// upper row; eprev/enext = edge words of the neighbouring words (EDGE_STONE at the walls).
template <int OX>
FS3D_HD uint32_t xy_substep(P2 &U, P2 &L, uint32_t rw, uint32_t eprev, uint32_t enext) {
    P2 hU, hL;
    hU.p0 = xpartner<OX>(U.p0, (eprev >> 4) & 1u, enext & 1u);
    hU.p1 = xpartner<OX>(U.p1, (eprev >> 5) & 1u, (enext >> 1) & 1u);
    hL.p0 = xpartner<OX>(L.p0, (eprev >> 6) & 1u, (enext >> 2) & 1u);
    hL.p1 = xpartner<OX>(L.p1, (eprev >> 7) & 1u, (enext >> 3) & 1u);
    uint32_t rm = rw & leftmask<OX>();                       // coin lives at the block's left cell
    uint32_t rs = rm | xpartner<OX>(rm, (eprev >> 8) & 1u, 0u);
    // evaluate every block from both of its columns (the rule is mirror symmetric): a = me
    return block_rule(U, hU, L, hL, rs);
}

}  // namespace fs3d
