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

// PRMT: byte i of the result = byte ((sel >> 4i) & 7) of the 8-byte value {b, a} (a = bytes 0-3, b = bytes 4-7)
FS3D_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
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

// ---------------------------------------------------------------------------------------------------
// XY sub-step on BOTH rows of a z-pair with every block evaluated once.
//
// xy_substep above evaluates each block from both of its columns (32 voxels per word = 16 blocks, each
// computed twice).  Here the left cells of the two rows' blocks are gathered into one word and the
// right cells into another (two PRMTs per plane), so one block_rule call handles the 32 distinct
// blocks of the two rows, and the words are interleaved back afterwards.
//   OX = 0: blocks (4k, 4k+1), (4k+2, 4k+3): left cells = bytes 0, 2, right cells = bytes 1, 3.
//   OX = 1: blocks (4k+1, 4k+2): left = byte 1, right = byte 2 (same bit k);
//           blocks (4k+3, 4k+4): left = byte 3 bit k, right = byte 0 bit k+1 — for k = 7 the right cell is
//           voxel 0 of the NEXT word.  A lane therefore evaluates the blocks whose LEFT cell it owns: it
//           needs the next word's voxel-0 bits before the rule (`nb`) and hands the new values of those
//           cells back afterwards (`carry`); its own voxel 0 comes from the previous word's carry (`pb`),
//           or from xy_wall_first at the grid wall.
// Packed edge bits: bit 2p + r, p = 0: U.p0, 1: U.p1, 2: L.p0, 3: L.p1, r = row.  0xFF = STONE.
FS3D_HD uint32_t xy_pair_substep0(P2 &U0, P2 &L0, P2 &U1, P2 &L1, uint32_t rw0, uint32_t rw1) {
    P2 a{prmt(U0.p0, U1.p0, 0x6240), prmt(U0.p1, U1.p1, 0x6240)}, b{prmt(U0.p0, U1.p0, 0x7351), prmt(U0.p1, U1.p1, 0x7351)};
    P2 c{prmt(L0.p0, L1.p0, 0x6240), prmt(L0.p1, L1.p1, 0x6240)}, d{prmt(L0.p0, L1.p0, 0x7351), prmt(L0.p1, L1.p1, 0x7351)};
    const uint32_t en = block_rule(a, b, c, d, prmt(rw0, rw1, 0x6240));   // coin lives at the block's left cell
    U0.p0 = prmt(a.p0, b.p0, 0x6240); U1.p0 = prmt(a.p0, b.p0, 0x7351);
    U0.p1 = prmt(a.p1, b.p1, 0x6240); U1.p1 = prmt(a.p1, b.p1, 0x7351);
    L0.p0 = prmt(c.p0, d.p0, 0x6240); L1.p0 = prmt(c.p0, d.p0, 0x7351);
    L0.p1 = prmt(c.p1, d.p1, 0x6240); L1.p1 = prmt(c.p1, d.p1, 0x7351);
    return en;
}

// voxel-0 bits of the eight plane words, packed
FS3D_HD uint32_t xy_first_bits(P2 U0, P2 L0, P2 U1, P2 L1) {
    return (U0.p0 & 1u) | ((U1.p0 & 1u) << 1) | ((U0.p1 & 1u) << 2) | ((U1.p1 & 1u) << 3) |
           ((L0.p0 & 1u) << 4) | ((L1.p0 & 1u) << 5) | ((L0.p1 & 1u) << 6) | ((L1.p1 & 1u) << 7);
}
// right-cell word of one plane for OX = 1: [r0.b2, r1.b2, r0.b0 >> 1 | nb_r0 << 7, r1.b0 >> 1 | nb_r1 << 7]
FS3D_HD uint32_t xy_right1(uint32_t w0, uint32_t w1, uint32_t nb2 /* bit 0: row 0, bit 1: row 1 */) {
    const uint32_t g = prmt(w0, w1, 0x4062);
    return (g & 0x0000FFFFu) | ((g >> 1) & 0x7F7F0000u) | (((nb2 & 3u) * 0x40800000u) & 0x80800000u);
}
// one row's plane word back from the left/right words (OX = 1); bit 0 (voxel 0) is left clear for xy_pair_post1
FS3D_HD uint32_t xy_merge1(uint32_t l, uint32_t r, int row) {
    const uint32_t s = prmt(l, r, row == 0 ? 0x2406u : 0x3517u);   // [R.byte(2+row), L.byte(row), R.byte(row), L.byte(2+row)]
    return prmt(s, s << 1, 0x3214);                                 // byte 0 <- (byte 0 << 1) & 0xFF
}
FS3D_HD uint32_t xy_pair_substep1(P2 &U0, P2 &L0, P2 &U1, P2 &L1, uint32_t rw0, uint32_t rw1, uint32_t nb, uint32_t &carry) {
    P2 a{prmt(U0.p0, U1.p0, 0x7351), prmt(U0.p1, U1.p1, 0x7351)}, c{prmt(L0.p0, L1.p0, 0x7351), prmt(L0.p1, L1.p1, 0x7351)};
    P2 b{xy_right1(U0.p0, U1.p0, nb), xy_right1(U0.p1, U1.p1, nb >> 2)}, d{xy_right1(L0.p0, L1.p0, nb >> 4), xy_right1(L0.p1, L1.p1, nb >> 6)};
    const uint32_t en = block_rule(a, b, c, d, prmt(rw0, rw1, 0x7351));
    // new values of the next word's voxel 0: bit 7 of the right words' bytes 2 (row 0) and 3 (row 1)
    carry = (((b.p0 >> 23) & 1u) | ((b.p0 >> 30) & 2u)) | ((((b.p1 >> 23) & 1u) | ((b.p1 >> 30) & 2u)) << 2) |
            ((((d.p0 >> 23) & 1u) | ((d.p0 >> 30) & 2u)) << 4) | ((((d.p1 >> 23) & 1u) | ((d.p1 >> 30) & 2u)) << 6);
    U0.p0 = xy_merge1(a.p0, b.p0, 0); U1.p0 = xy_merge1(a.p0, b.p0, 1);
    U0.p1 = xy_merge1(a.p1, b.p1, 0); U1.p1 = xy_merge1(a.p1, b.p1, 1);
    L0.p0 = xy_merge1(c.p0, d.p0, 0); L1.p0 = xy_merge1(c.p0, d.p0, 1);
    L0.p1 = xy_merge1(c.p1, d.p1, 0); L1.p1 = xy_merge1(c.p1, d.p1, 1);
    return en;
}
// voxel 0 of every plane word from the previous word's carry
FS3D_HD void xy_pair_post1(P2 &U0, P2 &L0, P2 &U1, P2 &L1, uint32_t pb) {
    U0.p0 |= pb & 1u;        U1.p0 |= (pb >> 1) & 1u; U0.p1 |= (pb >> 2) & 1u; U1.p1 |= (pb >> 3) & 1u;
    L0.p0 |= (pb >> 4) & 1u; L1.p0 |= (pb >> 5) & 1u; L0.p1 |= (pb >> 6) & 1u; L1.p1 |= (pb >> 7) & 1u;
}
// voxel 0 of the first word of a row (global x = 0) under OX = 1: its block's left column is the wall, so only
// F applies (STONE never moves, blocks D and L need a movable left cell).  `first` = the cells' bits BEFORE the
// sub-step (xy_first_bits); returns their new values in the same packing; en |= blocks enabled.
FS3D_HD uint32_t xy_wall_first(uint32_t first, uint32_t &en) {
    const uint32_t u0 = first & 3u, u1 = (first >> 2) & 3u, l0 = (first >> 4) & 3u, l1 = (first >> 6) & 3u;   // bit r = row r
    const uint32_t h = (u0 ^ u1) & (u0 | ~l1) & ~l0 & 3u;      // heavier(U, L) on the two rows
    const uint32_t t0 = (u0 ^ l0) & h, t1 = (u1 ^ l1) & h;
    en |= h;
    return (u0 ^ t0) | ((u1 ^ t1) << 2) | ((l0 ^ t0) << 4) | ((l1 ^ t1) << 6);
}

}  // namespace fs3d
