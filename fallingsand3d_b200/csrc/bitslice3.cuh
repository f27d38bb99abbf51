// bitslice3.cuh — the bit-sliced block rule of schedule version 2 (SCHEDULE.md §7: eight materials) on 32 voxels per
// word: THREE bit-planes per cell.
//
// Inside the kernels a cell is held in RANK encoding, not in its material code: rank = density order
//     GAS 0 < EMPTY 1 < OIL 2 < WATER 3 < HONEY 4 < SAND 5 < GRAVEL 6, STONE 7,
// because every predicate of the rule is a comparison or a range test of ranks (three-input logic ops), and STONE
// = 7 = all ones keeps "out of grid = all-ones word" true as in version 1.  Codes <-> ranks, per plane word:
//     code (c2 c1 c0):  EMPTY 000  SAND 001  WATER 010  STONE 011  GAS 100  OIL 101  HONEY 110  GRAVEL 111
//     rank (q2 q1 q0):  q0 = ~c2,  q1 = c2 ? c0 : c1,  q2 = c2 ? c1 : c0        (and back: c2 = ~q0, c1 = q0 ? q1 : q2,
//                                                                                 c0 = q0 ? q2 : q1)
// Pure functions only, so the same source compiles under nvcc for the kernels and under g++ for
// tests/host/bitslice_host_test.cpp.  No reference counterpart.
#pragma once
#include "bitslice.cuh"

namespace fs3d {

struct P3 { uint32_t p0, p1, p2; };   // rank bit-planes of 32 voxels, voxel x = 4k + i at bit 8i + k (as in P2)

FS3D_HD uint32_t sel(uint32_t m, uint32_t x, uint32_t y) { return (x & m) | (y & ~m); }   // m ? x : y, one LOP3

// 32 bytes (codes 0..7) -> three rank planes.  Words k and k + 4 are first merged into nibbles (codes < 8 leave
// bit 3 of a nibble clear), then bit j of every nibble is gathered: word k lands at bit k, word k + 4 at bit k + 4.
FS3D_HD P3 pack3(const uint32_t (&w)[8]) {
    uint32_t a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = w[k] + (w[k + 4] << 4);
    uint32_t c[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        c[j] = ((a[0] >> j) & 0x11111111u) + (((a[1] >> j) & 0x11111111u) << 1) + (((a[2] >> j) & 0x11111111u) << 2) +
               (((a[3] >> j) & 0x11111111u) << 3);
    P3 q;
    q.p0 = ~c[2];
    q.p1 = sel(c[2], c[0], c[1]);
    q.p2 = sel(c[2], c[1], c[0]);
    return q;
}
FS3D_HD void unpack3(P3 q, uint32_t (&w)[8]) {
    const uint32_t c2 = ~q.p0, c1 = sel(q.p0, q.p1, q.p2), c0 = sel(q.p0, q.p2, q.p1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // nibble (c2 c1 c0) of words k (low) and k + 4 (high) for every byte lane: bit j of the nibble <- bit k of plane j
        const uint32_t s0 = c0 >> k;
        const uint32_t s1 = k >= 1 ? c1 >> (k - 1) : c1 << 1;
        const uint32_t s2 = k >= 2 ? c2 >> (k - 2) : c2 << (2 - k);
        const uint32_t t = sel(0x33333333u, sel(0x11111111u, s0, s1), s2);
        w[k] = t & 0x07070707u;
        w[k + 4] = (t >> 4) & 0x07070707u;
    }
}

// heavier(u, l): u is not STONE, l yields (rank <= 4: GAS, EMPTY, OIL, WATER, HONEY) and rank(u) > rank(l)
FS3D_HD uint32_t heavier3(P3 u, P3 l) {
    const uint32_t g0 = u.p0 & ~l.p0;
    const uint32_t g1 = (u.p1 & ~l.p1) | (~(u.p1 ^ l.p1) & g0);
    const uint32_t gt = (u.p2 & ~l.p2) | (~(u.p2 ^ l.p2) & g1);
    const uint32_t yields = ~(l.p2 & (l.p1 | l.p0));
    const uint32_t movable = ~(u.p2 & u.p1 & u.p0);
    return gt & yields & movable;
}
FS3D_HD void cswap3(uint32_t m, P3 &x, P3 &y) {
    const uint32_t x0 = sel(m, y.p0, x.p0), x1 = sel(m, y.p1, x.p1), x2 = sel(m, y.p2, x.p2);
    y.p0 = sel(m, x.p0, y.p0); y.p1 = sel(m, x.p1, y.p1); y.p2 = sel(m, x.p2, y.p2);
    x.p0 = x0; x.p1 = x1; x.p2 = x2;
}
// second coin word of a block's hash word (SCHEDULE.md §7)
FS3D_HD uint32_t coin2_word(uint32_t rw) { uint32_t v = rw * 0x9E3779B1u; return v ^ (v >> 15); }

// The block rule F, D, L of schedule version 2 on 32 blocks at once.  a, b upper row; c, d lower row; r, r2 = first and
// second coin bits.  Returns the mask of enabled blocks (coins ignored).
FS3D_HD uint32_t block_rule3(P3 &a, P3 &b, P3 &c, P3 &d, uint32_t r, uint32_t r2) {
    const uint32_t fa = heavier3(a, c); cswap3(fa, a, c);
    const uint32_t fb = heavier3(b, d); cswap3(fb, b, d);
    // D: the cell beside the mover is not STONE, and GRAVEL (rank 6 = 110) never slides: mover rank bits (p2 p1) != 11
    const uint32_t da = heavier3(a, d) & ~(b.p0 & b.p1 & b.p2) & ~(a.p2 & a.p1);
    const uint32_t db = heavier3(b, c) & ~(a.p0 & a.p1 & a.p2) & ~(b.p2 & b.p1) & ~da;
    cswap3(da, a, d); cswap3(db, b, c);
    // L: two different yielding cells; HONEY (rank 4: among yielding ranks exactly p2 = 1) also needs the second coin
    const uint32_t ya = ~(a.p2 & (a.p1 | a.p0)), yb = ~(b.p2 & (b.p1 | b.p0));
    const uint32_t l = ((a.p0 ^ b.p0) | (a.p1 ^ b.p1) | (a.p2 ^ b.p2)) & ya & yb;
    cswap3(l & r & (~(a.p2 | b.p2) | r2), a, b);
    return fa | fb | da | db | l;
}

// ---- XY sub-step on both rows of a z-pair, every block evaluated once (see bitslice.cuh for the geometry) ----------
// Edge bits of the pair form, byte-packed like bitslice.cuh's E-packing with two more bits per row for the third plane:
//   planes 0, 1: bit 8·(2p + row) + j;   plane 2: bit 8·row + 2 + j;   j = 0 for the block's upper plane (U), 1 for L.
// Every other bit is "don't care" for the consumers (the warp-pair mailbox keeps its tag in bits 28-31).
constexpr uint32_t NB_STONE3 = 0x03030F0Fu;

FS3D_HD uint32_t xy3_pair_substep0(P3 &U0, P3 &L0, P3 &U1, P3 &L1, uint32_t rw0, uint32_t rw1) {
    P3 a{prmt(U0.p0, U1.p0, 0x6240), prmt(U0.p1, U1.p1, 0x6240), prmt(U0.p2, U1.p2, 0x6240)};
    P3 b{prmt(U0.p0, U1.p0, 0x7351), prmt(U0.p1, U1.p1, 0x7351), prmt(U0.p2, U1.p2, 0x7351)};
    P3 c{prmt(L0.p0, L1.p0, 0x6240), prmt(L0.p1, L1.p1, 0x6240), prmt(L0.p2, L1.p2, 0x6240)};
    P3 d{prmt(L0.p0, L1.p0, 0x7351), prmt(L0.p1, L1.p1, 0x7351), prmt(L0.p2, L1.p2, 0x7351)};
    // both coins live at the block's left cell
    const uint32_t en = block_rule3(a, b, c, d, prmt(rw0, rw1, 0x6240), prmt(coin2_word(rw0), coin2_word(rw1), 0x6240));
    U0.p0 = prmt(a.p0, b.p0, 0x6240); U1.p0 = prmt(a.p0, b.p0, 0x7351);
    U0.p1 = prmt(a.p1, b.p1, 0x6240); U1.p1 = prmt(a.p1, b.p1, 0x7351);
    U0.p2 = prmt(a.p2, b.p2, 0x6240); U1.p2 = prmt(a.p2, b.p2, 0x7351);
    L0.p0 = prmt(c.p0, d.p0, 0x6240); L1.p0 = prmt(c.p0, d.p0, 0x7351);
    L0.p1 = prmt(c.p1, d.p1, 0x6240); L1.p1 = prmt(c.p1, d.p1, 0x7351);
    L0.p2 = prmt(c.p2, d.p2, 0x6240); L1.p2 = prmt(c.p2, d.p2, 0x7351);
    return en;
}
FS3D_HD uint32_t xy3_first_bits(P3 U0, P3 L0, P3 U1, P3 L1) {
    // byte 0 of the twelve plane words, gathered: [U0.p0, U1.p0, U0.p1, U1.p1], the same of L, [U0.p2, U1.p2, L0.p2, L1.p2]
    const uint32_t fu = prmt(xy_gather1(U0.p0, U1.p0), xy_gather1(U0.p1, U1.p1), 0x7632);
    const uint32_t fl = prmt(xy_gather1(L0.p0, L1.p0), xy_gather1(L0.p1, L1.p1), 0x7632);
    const uint32_t f2 = prmt(xy_gather1(U0.p2, U1.p2), xy_gather1(L0.p2, L1.p2), 0x7632);
    const uint32_t t01 = (fu & 0x01010101u) | (shl_add<1>(fl, 0u) & ~0x01010101u);
    const uint32_t t2 = (shl_add<2>(f2, 0u) & 0x0404u) | (shr<13>(f2) & ~0x0404u);       // U: bits 0, 8 -> 2, 10; L: bits 16, 24 -> 3, 11
    return (t01 & 0x03030303u) | (t2 & ~0x03030303u);
}
FS3D_HD uint32_t xy3_pair_substep1(P3 &U0, P3 &L0, P3 &U1, P3 &L1, uint32_t rw0, uint32_t rw1, uint32_t nb, uint32_t &carry) {
    P3 a{prmt(U0.p0, U1.p0, 0x7351), prmt(U0.p1, U1.p1, 0x7351), prmt(U0.p2, U1.p2, 0x7351)};
    P3 c{prmt(L0.p0, L1.p0, 0x7351), prmt(L0.p1, L1.p1, 0x7351), prmt(L0.p2, L1.p2, 0x7351)};
    const uint32_t nl = shr<1>(nb);
    P3 b{xy_right1<0>(U0.p0, U1.p0, nb), xy_right1<1>(U0.p1, U1.p1, nb), xy_right1<0>(U0.p2, U1.p2, shr<2>(nb))};
    P3 d{xy_right1<0>(L0.p0, L1.p0, nl), xy_right1<1>(L0.p1, L1.p1, nl), xy_right1<0>(L0.p2, L1.p2, shr<3>(nb))};
    const uint32_t en = block_rule3(a, b, c, d, prmt(rw0, rw1, 0x7351), prmt(coin2_word(rw0), coin2_word(rw1), 0x7351));
    // new values of the next word's voxel 0: bit 7 of the right words' bytes 2 (row 0) and 3 (row 1)
    const uint32_t c2 = prmt(b.p2, d.p2, 0x7632);        // bit 7, 15: U.p2 rows 0, 1; bit 23, 31: L.p2 rows 0, 1
    carry = (shr<7>(prmt(b.p0, b.p1, 0x7632)) & 0x01010101u) | (shr<6>(prmt(d.p0, d.p1, 0x7632)) & 0x02020202u) |
            (shr<5>(c2) & 0x0404u) | (shr<20>(c2) & 0x0808u);
    // the words are finished by xy3_pair_post1 (see xy_pair_substep1 in bitslice.cuh)
    U0.p0 = prmt(a.p0, b.p0, 0x2406); U1.p0 = prmt(a.p0, b.p0, 0x3517);
    U0.p1 = prmt(a.p1, b.p1, 0x2406); U1.p1 = prmt(a.p1, b.p1, 0x3517);
    U0.p2 = prmt(a.p2, b.p2, 0x2406); U1.p2 = prmt(a.p2, b.p2, 0x3517);
    L0.p0 = prmt(c.p0, d.p0, 0x2406); L1.p0 = prmt(c.p0, d.p0, 0x3517);
    L0.p1 = prmt(c.p1, d.p1, 0x2406); L1.p1 = prmt(c.p1, d.p1, 0x3517);
    L0.p2 = prmt(c.p2, d.p2, 0x2406); L1.p2 = prmt(c.p2, d.p2, 0x3517);
    return en;
}
FS3D_HD void xy3_pair_post1(P3 &U0, P3 &L0, P3 &U1, P3 &L1, uint32_t pb) {
    const uint32_t cu = pb & 0x01010101u, cl = shr<1>(pb) & 0x01010101u, cu2 = shr<2>(pb) & 0x0101u, cl2 = shr<3>(pb) & 0x0101u;
    U0.p0 = xy_finish1(U0.p0, cu);  U1.p0 = xy_finish1(U1.p0, shr<8>(cu));  U0.p1 = xy_finish1(U0.p1, shr<16>(cu)); U1.p1 = xy_finish1(U1.p1, shr<24>(cu));
    L0.p0 = xy_finish1(L0.p0, cl);  L1.p0 = xy_finish1(L1.p0, shr<8>(cl));  L0.p1 = xy_finish1(L0.p1, shr<16>(cl)); L1.p1 = xy_finish1(L1.p1, shr<24>(cl));
    U0.p2 = xy_finish1(U0.p2, cu2); U1.p2 = xy_finish1(U1.p2, shr<8>(cu2));
    L0.p2 = xy_finish1(L0.p2, cl2); L1.p2 = xy_finish1(L1.p2, shr<8>(cl2));
}
// voxel 0 of a row's first word (global x = 0) under odd x-offset: the block's left column is the wall, so only F
// applies (the wall is STONE: D needs a non-STONE neighbour beside the mover, L two yielding cells).  `first` = the
// cells' bits before the sub-step (xy3_first_bits); returns their new values; en |= rows whose block is enabled.
FS3D_HD uint32_t xy3_wall_first(uint32_t first, uint32_t &en) {
    // both rows at once, aligned to bits 0 (row 0) and 8 (row 1); the other bits are garbage and masked out of h
    const P3 u{first, shr<16>(first), shr<2>(first)}, l{shr<1>(first), shr<17>(first), shr<3>(first)};
    const uint32_t h = heavier3(u, l) & 0x0101u;
    en |= h;
    // the U and L bit of a plane flip together where the cells swap and differ in that plane
    const uint32_t t0 = (u.p0 ^ l.p0) & h, t1 = (u.p1 ^ l.p1) & h, t2 = (u.p2 ^ l.p2) & h;
    return first ^ ((t0 + shl_add<2>(t2, 0u) + shl_add<16>(t1, 0u)) * 3u);
}

}  // namespace fs3d
