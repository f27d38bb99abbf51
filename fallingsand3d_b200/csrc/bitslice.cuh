// bitslice.cuh — the bit-sliced block rule of SCHEDULE.md on 32 voxels per 32-bit word.
//
// Pure functions only (no memory access, no intrinsics that need a GPU) so the same source
// compiles under nvcc for the kernels and under g++ for tests/host/bitslice_host_test.cpp.
// No reference counterpart.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FS3D_HD __host__ __device__ __forceinline__
#define FS3D_CX __host__ __device__ constexpr
#else
#define FS3D_HD inline
#define FS3D_CX constexpr
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

// x >> K and (x << K) + y behind one name each, so that the pipe they issue on can be chosen in one place.  The step
// kernels are bound by the integer ALU pipe (LOP3 / SHF / PRMT) while the FMA pipe idles, which makes IMAD.HI (x *
// 2^(32-K), upper half, multiplier read from constant memory so the compiler cannot turn it back into a shift) look
// attractive for constant right shifts — measured, it is not: IMAD.HI issues at half rate and stalls the dispatch port,
// 0.893 against 0.859 ms/step at 2048^3 (profiles/r02p_experiments_alu.txt).  Plain shifts are the default;
// -DFS3D_IMAD_SHR=1 selects the IMAD.HI form everywhere, -DFS3D_IMAD_SHR=2 only in hash_word.
#ifndef FS3D_IMAD_SHR
#define FS3D_IMAD_SHR 0
#endif
#if defined(__CUDACC__) && FS3D_IMAD_SHR
static __constant__ uint32_t c_pow2[32] = {
    1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,  1u << 8,  1u << 9,  1u << 10,
    1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21,
    1u << 22, 1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};
#endif
#if defined(__CUDA_ARCH__) && FS3D_IMAD_SHR
template <int K> __device__ __forceinline__ uint32_t shr_imad(uint32_t x) {
    static_assert(K >= 1 && K <= 31, "shift count");
    return __umulhi(x, c_pow2[32 - K]);
}
#else
template <int K> FS3D_HD uint32_t shr_imad(uint32_t x) { return x >> K; }
#endif
#if defined(__CUDA_ARCH__) && FS3D_IMAD_SHR == 1
template <int K> __device__ __forceinline__ uint32_t shr(uint32_t x) { return shr_imad<K>(x); }
template <int K> __device__ __forceinline__ uint32_t shl_add(uint32_t x, uint32_t y) { return x * c_pow2[K] + y; }
#else
template <int K> FS3D_HD uint32_t shr(uint32_t x) { return x >> K; }
template <int K> FS3D_HD uint32_t shl_add(uint32_t x, uint32_t y) { return (x << K) + y; }
#endif
#if FS3D_IMAD_SHR == 2
#define FS3D_SHR_HASH shr_imad
#else
#define FS3D_SHR_HASH shr
#endif
constexpr uint32_t HC1 = 0x9E3779B1u, HC2 = 0x85EBCA77u, HC3 = 0xC2B2AE3Du;

struct P2 { uint32_t p0, p1; };   // two bit-planes of 32 voxels

// 32 bytes (codes 0..3) -> two bit-planes, voxel x = 4k + i (word k, byte i) at bit 8i + k.
FS3D_HD P2 pack(const uint32_t (&w)[8]) {
    uint32_t te = w[0] + (w[2] << 2) + (w[4] << 4) + (w[6] << 6);   // even k: 2-bit codes at 2·(k/2)
    uint32_t to = w[1] + (w[3] << 2) + (w[5] << 4) + (w[7] << 6);   // odd k
    P2 c;
    c.p0 = (te & 0x55555555u) | ((to << 1) & 0xAAAAAAAAu);
    c.p1 = (shr<1>(te) & 0x55555555u) | (to & 0xAAAAAAAAu);
    return c;
}
FS3D_HD void unpack(P2 c, uint32_t (&w)[8]) {
    uint32_t te = (c.p0 & 0x55555555u) | ((c.p1 << 1) & 0xAAAAAAAAu);
    uint32_t to = (shr<1>(c.p0) & 0x55555555u) | (c.p1 & 0xAAAAAAAAu);
    w[0] = te & 0x03030303u;         w[1] = to & 0x03030303u;
    w[2] = shr<2>(te) & 0x03030303u; w[3] = shr<2>(to) & 0x03030303u;
    w[4] = shr<4>(te) & 0x03030303u; w[5] = shr<4>(to) & 0x03030303u;
    w[6] = shr<6>(te) & 0x03030303u; w[7] = shr<6>(to) & 0x03030303u;
}

FS3D_HD uint32_t hash_word(uint32_t v) {   // SCHEDULE.md §3 H, after the linear part
    v ^= FS3D_SHR_HASH<16>(v); v *= 0x7FEB352Du;
    v ^= FS3D_SHR_HASH<15>(v); v *= 0x846CA68Bu;
    v ^= FS3D_SHR_HASH<16>(v);
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

// The block rule F, D, L in its plain mask-and-swap form: the specification block_rule() below is tested against
// (tests/host/bitslice_host_test.cpp); the kernels do not call it.
FS3D_HD uint32_t block_rule_plain(P2 &a, P2 &b, P2 &c, P2 &d, uint32_t r) {
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

// One three-input logic operation = one LOP3.  The block rule below is written as an explicit network of them (26 for
// the whole rule; the compiler's own synthesis from the mask-and-swap formulation needs 39) and the step kernels are
// bound by exactly this pipe.  FS3D_LOP3(name, expr) defines name(a, b, c) = expr through its truth table (expr
// evaluated at compile time on the three selector constants): on the device one lop3.b32 with that table, on the host
// the same table applied minterm by minterm — so tests/host/ checks the very tables the kernels run.
template <uint32_t LUT>
FS3D_HD uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
#else
    uint32_t d = 0;
    for (uint32_t i = 0; i < 8; ++i)
        if ((LUT >> i) & 1u) d |= ((i & 4u) ? a : ~a) & ((i & 2u) ? b : ~b) & ((i & 1u) ? c : ~c);
    return d;
#endif
}
#define FS3D_LOP3(NAME, EXPR)                                                                         \
    FS3D_CX uint32_t NAME##_expr(uint32_t a, uint32_t b, uint32_t c) { return (EXPR); }                \
    FS3D_HD uint32_t NAME(uint32_t a, uint32_t b, uint32_t c) { return lop3<NAME##_expr(0xF0u, 0xCCu, 0xAAu) & 0xFFu>(a, b, c); }
// codes: EMPTY 00, SAND p0, WATER p1, STONE both
FS3D_LOP3(br_keep0,   a & (b | c))              // upper.p0 after a fall: SAND (a = p0, b = p1) stays only over c = "cannot fall"
FS3D_LOP3(br_take0,   a | (b & ~c))             // lower.p0 after F: own (a), or SAND above (b = p0, c = p1)
FS3D_LOP3(br_p1pair,  (a ^ b) & ~(a ^ c))       // upper movable (a = p0, b = p1) and p1 differs from the lower's (c): SAND/WATER, WATER/EMPTY
FS3D_LOP3(br_flip_nc, a ^ (b & ~c))             // flip a where b, unless c
FS3D_LOP3(br_flip_c,  a ^ (b & c))              // flip a where b and c
FS3D_LOP3(br_open,    ~a & ~(b & c))            // D: target not p0 (EMPTY or WATER) and the cell beside the mover not STONE
FS3D_LOP3(br_keep0d,  a & (b | ~c))             // upper.p0 after D: SAND stays unless the way is open (c)
FS3D_LOP3(br_take0d,  a | (b ^ c))              // lower.p0 after D: own, or what the upper cell gave up
FS3D_LOP3(br_lat,     (a ^ b) & ~c)             // L, first half: p1 differs, first cell's p0 clear
FS3D_LOP3(br_lat2,    a & ~b & c)               // L, second half: second cell's p0 clear, coin set
FS3D_LOP3(br_lat2e,   a & ~b)                   // the same without the coin (c unused)
FS3D_LOP3(br_chg,     a | (b ^ c))

// The block rule F, D, L on 32 blocks at once. a,b upper row; c,d lower row; r = coin bits.
// Returns the mask of enabled blocks (coin ignored; dead code where the caller does not track activity).
FS3D_HD uint32_t block_rule(P2 &a, P2 &b, P2 &c, P2 &d, uint32_t r) {
    const P2 a_in = a, b_in = b;
    // F: both columns.  p0 moves only with SAND; p1 is exchanged for SAND over WATER and WATER over EMPTY.
    {
        const uint32_t ta = br_p1pair(a.p0, a.p1, c.p1), tb = br_p1pair(b.p0, b.p1, d.p1);
        const uint32_t a0 = br_keep0(a.p0, a.p1, c.p0), c0 = br_take0(c.p0, a.p0, a.p1);
        const uint32_t b0 = br_keep0(b.p0, b.p1, d.p0), d0 = br_take0(d.p0, b.p0, b.p1);
        a.p1 = br_flip_nc(a.p1, ta, c.p0); c.p1 = br_flip_nc(c.p1, ta, c.p0);
        b.p1 = br_flip_nc(b.p1, tb, d.p0); d.p1 = br_flip_nc(d.p1, tb, d.p0);
        a.p0 = a0; c.p0 = c0; b.p0 = b0; d.p0 = d0;
    }
    // D: a -> d unless b is STONE, b -> c unless a is STONE (never both: densities are ordered)
    {
        const uint32_t ya = br_open(d.p0, b.p0, b.p1), yb = br_open(c.p0, a.p0, a.p1);
        const uint32_t ta = br_p1pair(a.p0, a.p1, d.p1), tb = br_p1pair(b.p0, b.p1, c.p1);
        const uint32_t a0 = br_keep0d(a.p0, a.p1, ya), b0 = br_keep0d(b.p0, b.p1, yb);
        d.p0 = br_take0d(d.p0, a.p0, a0); c.p0 = br_take0d(c.p0, b.p0, b0);
        a.p1 = br_flip_c(a.p1, ta, ya); d.p1 = br_flip_c(d.p1, ta, ya);
        b.p1 = br_flip_c(b.p1, tb, yb); c.p1 = br_flip_c(c.p1, tb, yb);
        a.p0 = a0; b.p0 = b0;
    }
    // what F and D moved shows in the upper cells (a moved block always changes its upper cell)
    const uint32_t moved = br_chg(br_chg(a_in.p0 ^ a.p0, a_in.p1, a.p1), b_in.p0, b.p0) | (b_in.p1 ^ b.p1);
    // L: WATER (p1 only) beside EMPTY — the two cells differ in p1 alone, so the swap is a flip of both p1 bits
    const uint32_t x = br_lat(a.p1, b.p1, a.p0);
    const uint32_t m = br_lat2(x, b.p0, r);
    const uint32_t l = br_lat2e(x, b.p0, 0u);
    a.p1 ^= m; b.p1 ^= m;
    return moved | l;
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

// Edge packing ("E") of the eight voxel-0 bits a lane exchanges with its x-neighbours under OX = 1: bit 8·(2p + row) + j
// with p = bit-plane, row = row of the z-pair, j = 0 for the block's upper plane (U), 1 for its lower plane (L):
//   U0.p0 -> 0, U1.p0 -> 8, U0.p1 -> 16, U1.p1 -> 24, L0.p0 -> 1, L1.p0 -> 9, L0.p1 -> 17, L1.p1 -> 25.
// Every other bit is "don't care" for all consumers below, so producers skip the masks and the warp-pair mailbox of the
// step kernels keeps its tag in bits 28-31.  STONE beyond the wall = NB_STONE2.
constexpr uint32_t NB_STONE2 = 0x03030303u;
// [w0.b2, w1.b2, w0.b0, w1.b0] of one plane of the two rows: the right cells of blocks (4k+1, 4k+2) and, in bytes 2-3,
// the words' own voxels 0, 4, .. 28 (right cells of the straddling blocks of the PREVIOUS positions)
FS3D_HD uint32_t xy_gather1(uint32_t w0, uint32_t w1) { return prmt(w0, w1, 0x4062); }
// voxel-0 bits of the eight plane words, E-packed (two PRMTs on top of the gathers xy_pair_substep1 needs anyway)
FS3D_HD uint32_t xy_first_bits(P2 U0, P2 L0, P2 U1, P2 L1) {
    const uint32_t fu = prmt(xy_gather1(U0.p0, U1.p0), xy_gather1(U0.p1, U1.p1), 0x7632);   // byte 2p + row = byte 0 of that word
    const uint32_t fl = prmt(xy_gather1(L0.p0, L1.p0), xy_gather1(L0.p1, L1.p1), 0x7632);
    return (fu & 0x01010101u) | (shl_add<1>(fl, 0u) & ~0x01010101u);
}
// right-cell word of one plane for OX = 1: [r0.b2, r1.b2, r0.b0 >> 1 | nb_r0 << 7, r1.b0 >> 1 | nb_r1 << 7].
// n = a word whose bytes 2p and 2p + 1 carry the next word's voxel-0 bit of rows 0 / 1 in their bit 0 (E-packed first
// bits for U planes, the same shifted right by one for L planes); the bit rides into bit 7 on the shift.
template <int PL>
FS3D_HD uint32_t xy_right1(uint32_t w0, uint32_t w1, uint32_t n) {
    const uint32_t g = xy_gather1(w0, w1);
    const uint32_t x = prmt(g, n, PL == 0 ? 0x5342u : 0x7362u);     // [g.b2, n.b(2p), g.b3, n.b(2p+1)]
    return prmt(g, shr<1>(x), 0x6410);                               // [g.b0, g.b1, (x >> 1).b0, (x >> 1).b2]
}
// The sub-step proper.  Leaves in U0 .. L1 the words [R.byte(2+row), L.byte(row), R.byte(row), L.byte(2+row)], whose
// byte 0 still has to move up by one bit and take voxel 0 from the previous word: xy_pair_post1 finishes them.
// carry = the new values of the NEXT word's voxel 0, E-packed.
FS3D_HD uint32_t xy_pair_substep1(P2 &U0, P2 &L0, P2 &U1, P2 &L1, uint32_t rw0, uint32_t rw1, uint32_t nb, uint32_t &carry) {
    P2 a{prmt(U0.p0, U1.p0, 0x7351), prmt(U0.p1, U1.p1, 0x7351)}, c{prmt(L0.p0, L1.p0, 0x7351), prmt(L0.p1, L1.p1, 0x7351)};
    const uint32_t nl = shr<1>(nb);
    P2 b{xy_right1<0>(U0.p0, U1.p0, nb), xy_right1<1>(U0.p1, U1.p1, nb)}, d{xy_right1<0>(L0.p0, L1.p0, nl), xy_right1<1>(L0.p1, L1.p1, nl)};
    const uint32_t en = block_rule(a, b, c, d, prmt(rw0, rw1, 0x7351));
    // new values of the next word's voxel 0: bit 7 of the right words' bytes 2 (row 0) and 3 (row 1)
    carry = (shr<7>(prmt(b.p0, b.p1, 0x7632)) & 0x01010101u) | (shr<6>(prmt(d.p0, d.p1, 0x7632)) & 0x02020202u);
    U0.p0 = prmt(a.p0, b.p0, 0x2406); U1.p0 = prmt(a.p0, b.p0, 0x3517);
    U0.p1 = prmt(a.p1, b.p1, 0x2406); U1.p1 = prmt(a.p1, b.p1, 0x3517);
    L0.p0 = prmt(c.p0, d.p0, 0x2406); L1.p0 = prmt(c.p0, d.p0, 0x3517);
    L0.p1 = prmt(c.p1, d.p1, 0x2406); L1.p1 = prmt(c.p1, d.p1, 0x3517);
    return en;
}
// byte 0 <- (byte 0 << 1) | voxel 0, where `add` has the voxel-0 bit in bit 0 and nothing else in its byte 0 (what its
// upper bytes add to s << 1 never reaches byte 0 and is not taken)
FS3D_HD uint32_t xy_finish1(uint32_t s, uint32_t add) { return prmt(s, shl_add<1>(s, add), 0x3214); }
// finishes the words of xy_pair_substep1 with voxel 0 from the previous word's carry (E-packed)
FS3D_HD void xy_pair_post1(P2 &U0, P2 &L0, P2 &U1, P2 &L1, uint32_t pb) {
    const uint32_t cu = pb & 0x01010101u, cl = shr<1>(pb) & 0x01010101u;
    U0.p0 = xy_finish1(U0.p0, cu); U1.p0 = xy_finish1(U1.p0, shr<8>(cu)); U0.p1 = xy_finish1(U0.p1, shr<16>(cu)); U1.p1 = xy_finish1(U1.p1, shr<24>(cu));
    L0.p0 = xy_finish1(L0.p0, cl); L1.p0 = xy_finish1(L1.p0, shr<8>(cl)); L0.p1 = xy_finish1(L0.p1, shr<16>(cl)); L1.p1 = xy_finish1(L1.p1, shr<24>(cl));
}
// voxel 0 of the first word of a row (global x = 0) under OX = 1: its block's left column is the wall, so only
// F applies (STONE never moves, blocks D and L need a movable left cell).  `first` = the cells' bits BEFORE the
// sub-step (xy_first_bits); returns their new values in the same packing; en |= blocks enabled.
FS3D_HD uint32_t xy_wall_first(uint32_t first, uint32_t &en) {
    // both rows at once, aligned to bits 0 (row 0) and 8 (row 1)
    const uint32_t u0 = first, u1 = shr<16>(first), l0 = shr<1>(first), l1 = shr<17>(first);
    const uint32_t h = (u0 ^ u1) & (u0 | ~l1) & ~l0 & 0x0101u;      // heavier(U, L)
    const uint32_t t0 = (u0 ^ l0) & h, t1 = (u1 ^ l1) & h;           // plane bits that change when the cells swap
    en |= h;
    return first ^ (shl_add<16>(t1, t0) * 3u);                       // U and L bit of a plane flip together
}

// ---- the plain forms of the two helpers above, with the neighbour bits as a small integer (bit 0: row 0, bit 1: row 1);
// bitslice3.cuh (schedule version 2) builds its OX = 1 sub-step from these
FS3D_HD uint32_t xy_right1(uint32_t w0, uint32_t w1, uint32_t nb2) {
    const uint32_t g = prmt(w0, w1, 0x4062);
    return (g & 0x0000FFFFu) | (shr<1>(g) & 0x7F7F0000u) | (((nb2 & 3u) * 0x40800000u) & 0x80800000u);
}
// one row's plane word back from the left/right words (OX = 1); bit 0 (voxel 0) is left clear for the post step
FS3D_HD uint32_t xy_merge1(uint32_t l, uint32_t r, int row) {
    const uint32_t s = prmt(l, r, row == 0 ? 0x2406u : 0x3517u);   // [R.byte(2+row), L.byte(row), R.byte(row), L.byte(2+row)]
    return prmt(s, s << 1, 0x3214);                                 // byte 0 <- (byte 0 << 1) & 0xFF
}

}  // namespace fs3d
