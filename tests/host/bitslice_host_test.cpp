// bitslice_host_test.cpp — CPU unit test of fallingsand3d_b200/csrc/bitslice.cuh (g++, no GPU).
//
// 1. pack/unpack, xpartner<0|1>, block_rule against their scalar definitions (SCHEDULE.md §2).
// 2. A word-level emulation of one full step built from the SAME helper functions the kernel
//    uses (xy_substep + block_rule + edge words), written to stdout as a grid so the Python test
//    can compare it with the oracle.  This pins the bit layout, the coin indexing and the edge
//    exchange before any GPU time is spent.  Test infrastructure only.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>
#include "../../fallingsand3d_b200/csrc/bitslice.cuh"

using namespace fs3d;

static int bitpos(int x) { return 8 * (x & 3) + ((x >> 2) & 7); }
enum { E = 0, S = 1, W = 2, X = 3 };
static int dens(int m) { return m == S ? 2 : (m == W ? 1 : 0); }
static int heav(int u, int l) { if (u != S && u != W) return 0; if (l != E && l != W) return 0; return dens(u) > dens(l); }
// returns whether the block is enabled: something fell or slid, or a lateral move is possible whatever the coin says
static int rule(int &a, int &b, int &c, int &d, int coin) {
    int en = 0;
    if (heav(a, c)) { std::swap(a, c); en = 1; }
    if (heav(b, d)) { std::swap(b, d); en = 1; }
    if (heav(a, d) && b != X) { std::swap(a, d); en = 1; } else if (heav(b, c) && a != X) { std::swap(b, c); en = 1; }
    if ((a == W && b == E) || (b == W && a == E)) { en = 1; if (coin) std::swap(a, b); }
    return en;
}

static int unit_tests() {
    srand(1);
    int bad = 0;
    for (int trial = 0; trial < 20000; ++trial) {
        uint8_t by[32]; uint32_t w[8];
        for (int i = 0; i < 32; ++i) by[i] = rand() & 3;
        memcpy(w, by, 32);
        P2 p = pack(w);
        for (int x = 0; x < 32; ++x) {
            int m = ((p.p0 >> bitpos(x)) & 1) | (((p.p1 >> bitpos(x)) & 1) << 1);
            if (m != by[x]) bad++;
        }
        uint32_t o[8]; unpack(p, o);
        if (memcmp(o, w, 32)) bad++;
        uint32_t P = (uint32_t)rand() * 65536u ^ (uint32_t)rand(), pb = rand() & 1, nb = rand() & 1;
        uint32_t q0 = xpartner<0>(P, pb, nb), q1 = xpartner<1>(P, pb, nb);
        for (int x = 0; x < 32; ++x) {
            if (((q0 >> bitpos(x)) & 1) != ((P >> bitpos(x ^ 1)) & 1)) bad++;
            int px1 = (x & 1) ? x + 1 : x - 1;
            uint32_t exp = px1 < 0 ? pb : (px1 > 31 ? nb : ((P >> bitpos(px1)) & 1));
            if (((q1 >> bitpos(x)) & 1) != exp) bad++;
        }
        int a[32], b[32], c[32], d[32], r[32];
        P2 A{0, 0}, B{0, 0}, C{0, 0}, D{0, 0}; uint32_t R = 0;
        for (int i = 0; i < 32; ++i) {
            a[i] = rand() & 3; b[i] = rand() & 3; c[i] = rand() & 3; d[i] = rand() & 3; r[i] = rand() & 1;
            A.p0 |= (a[i] & 1u) << i; A.p1 |= ((a[i] >> 1) & 1u) << i; B.p0 |= (b[i] & 1u) << i; B.p1 |= ((b[i] >> 1) & 1u) << i;
            C.p0 |= (c[i] & 1u) << i; C.p1 |= ((c[i] >> 1) & 1u) << i; D.p0 |= (d[i] & 1u) << i; D.p1 |= ((d[i] >> 1) & 1u) << i;
            R |= (uint32_t)r[i] << i;
        }
        {   // the LOP3 network against the plain mask-and-swap formulation, on all 32 bit positions at once
            P2 A2 = A, B2 = B, C2 = C, D2 = D, A3 = A, B3 = B, C3 = C, D3 = D;
            const uint32_t e2 = block_rule(A2, B2, C2, D2, R), e3 = block_rule_plain(A3, B3, C3, D3, R);
            if (e2 != e3 || A2.p0 != A3.p0 || A2.p1 != A3.p1 || B2.p0 != B3.p0 || B2.p1 != B3.p1 || C2.p0 != C3.p0 || C2.p1 != C3.p1 ||
                D2.p0 != D3.p0 || D2.p1 != D3.p1) bad++;
        }
        const uint32_t EN = block_rule(A, B, C, D, R);
        for (int i = 0; i < 32; ++i) {
            if (rule(a[i], b[i], c[i], d[i], r[i]) != (int)((EN >> i) & 1u)) bad++;
            int ga = ((A.p0 >> i) & 1) | (((A.p1 >> i) & 1) << 1), gb = ((B.p0 >> i) & 1) | (((B.p1 >> i) & 1) << 1);
            int gc = ((C.p0 >> i) & 1) | (((C.p1 >> i) & 1) << 1), gd = ((D.p0 >> i) & 1) | (((D.p1 >> i) & 1) << 1);
            if (ga != a[i] || gb != b[i] || gc != c[i] || gd != d[i]) bad++;
        }
    }
    return bad;
}

// ---- word-level emulation of one step --------------------------------------------------------------
struct Grid {
    int nx, ny, nz, wpr;
    std::vector<P2> w;   // [z][y][xw]
    P2 &at(int z, int y, int xw) { return w[((size_t)z * ny + y) * wpr + xw]; }
};
static const P2 STONE2 = {ONES, ONES};
static P2 rdw(Grid &g, int z, int y, int xw) { if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) return STONE2; return g.at(z, y, xw); }
static void wrw(Grid &g, int z, int y, int xw, P2 v) { if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) return; g.at(z, y, xw) = v; }

template <int OX>
static void emu_xy(Grid &g, uint32_t key, int oy) {
    for (int z = 0; z < g.nz; ++z)
        for (int y0 = oy ? -1 : 0; y0 < g.ny; y0 += 2) {
            std::vector<P2> U(g.wpr), L(g.wpr); std::vector<uint32_t> rw(g.wpr), e(g.wpr);
            for (int xw = 0; xw < g.wpr; ++xw) {
                U[xw] = rdw(g, z, y0 + 1, xw); L[xw] = rdw(g, z, y0, xw);
                rw[xw] = hash_word(key + (uint32_t)xw * HC1 + (uint32_t)(y0 + 1) * HC2 + (uint32_t)z * HC3);
                e[xw] = edge_pack(U[xw], L[xw], rw[xw]);
            }
            for (int xw = 0; xw < g.wpr; ++xw) {
                uint32_t ep = xw > 0 ? e[xw - 1] : EDGE_STONE, en = xw + 1 < g.wpr ? e[xw + 1] : EDGE_STONE;
                P2 u = U[xw], l = L[xw];
                xy_substep<OX>(u, l, rw[xw], ep, en);
                wrw(g, z, y0 + 1, xw, u); wrw(g, z, y0, xw, l);
            }
        }
}
// the same sub-step through the pair functions (every block evaluated once, rows taken two at a time)
// the exchanged edge words are E-packed (bitslice.cuh): every bit outside 0x03030303 is "don't care" — the kernels
// put the mailbox tag there — so the emulation sets them all
static const uint32_t GARBAGE = ~NB_STONE2;
template <int OX>
static void emu_xy_pair(Grid &g, uint32_t key, int oy) {
    for (int z = 0; z < g.nz; z += 2)
        for (int y0 = oy ? -1 : 0; y0 < g.ny; y0 += 2) {
            const int W = g.wpr;
            std::vector<P2> U0(W), L0(W), U1(W), L1(W); std::vector<uint32_t> r0(W), r1(W), first(W), carry(W);
            for (int xw = 0; xw < W; ++xw) {
                U0[xw] = rdw(g, z, y0 + 1, xw); L0[xw] = rdw(g, z, y0, xw); U1[xw] = rdw(g, z + 1, y0 + 1, xw); L1[xw] = rdw(g, z + 1, y0, xw);
                r0[xw] = hash_word(key + (uint32_t)xw * HC1 + (uint32_t)(y0 + 1) * HC2 + (uint32_t)z * HC3);
                r1[xw] = hash_word(key + (uint32_t)xw * HC1 + (uint32_t)(y0 + 1) * HC2 + (uint32_t)(z + 1) * HC3);
                first[xw] = xy_first_bits(U0[xw], L0[xw], U1[xw], L1[xw]);
            }
            for (int xw = 0; xw < W; ++xw) {
                if (OX == 0) xy_pair_substep0(U0[xw], L0[xw], U1[xw], L1[xw], r0[xw], r1[xw]);
                else xy_pair_substep1(U0[xw], L0[xw], U1[xw], L1[xw], r0[xw], r1[xw], xw + 1 < W ? (first[xw + 1] | GARBAGE) : NB_STONE2, carry[xw]);
            }
            if (OX == 1)
                for (int xw = 0; xw < W; ++xw) {
                    uint32_t en = 0;
                    xy_pair_post1(U0[xw], L0[xw], U1[xw], L1[xw], xw > 0 ? (carry[xw - 1] | GARBAGE) : xy_wall_first(first[0] | GARBAGE, en));
                }
            for (int xw = 0; xw < W; ++xw) {
                wrw(g, z, y0 + 1, xw, U0[xw]); wrw(g, z, y0, xw, L0[xw]); wrw(g, z + 1, y0 + 1, xw, U1[xw]); wrw(g, z + 1, y0, xw, L1[xw]);
            }
        }
}
static void emu_zy(Grid &g, uint32_t key, int oz, int oy) {
    for (int z0 = oz ? -1 : 0; z0 < g.nz; z0 += 2)
        for (int y0 = oy ? -1 : 0; y0 < g.ny; y0 += 2)
            for (int xw = 0; xw < g.wpr; ++xw) {
                P2 a = rdw(g, z0, y0 + 1, xw), b = rdw(g, z0 + 1, y0 + 1, xw), c = rdw(g, z0, y0, xw), d = rdw(g, z0 + 1, y0, xw);
                uint32_t rw = hash_word(key + (uint32_t)xw * HC1 + (uint32_t)(y0 + 1) * HC2 + (uint32_t)z0 * HC3);
                block_rule(a, b, c, d, rw);
                wrw(g, z0, y0 + 1, xw, a); wrw(g, z0 + 1, y0 + 1, xw, b); wrw(g, z0, y0, xw, c); wrw(g, z0 + 1, y0, xw, d);
            }
}

// usage: prog                          -> unit tests
//        prog nx ny nz kxy kzy t [p]   -> reads nx*ny*nz bytes on stdin, writes one emulated step to stdout
//                                         (p = 1: XY through the pair functions the kernel uses)
int main(int argc, char **argv) {
    if (argc < 7) {
        int bad = unit_tests();
        printf("bad=%d\n", bad);
        return bad != 0;
    }
    Grid g; g.nx = atoi(argv[1]); g.ny = atoi(argv[2]); g.nz = atoi(argv[3]); g.wpr = g.nx / 32;
    uint32_t kxy = (uint32_t)strtoul(argv[4], 0, 10), kzy = (uint32_t)strtoul(argv[5], 0, 10);
    unsigned long t = strtoul(argv[6], 0, 10);
    std::vector<uint8_t> bytes((size_t)g.nx * g.ny * g.nz);
    if (fread(bytes.data(), 1, bytes.size(), stdin) != bytes.size()) return 2;
    g.w.resize((size_t)g.nz * g.ny * g.wpr);
    for (size_t i = 0; i < g.w.size(); ++i) { uint32_t w[8]; memcpy(w, &bytes[i * 32], 32); g.w[i] = pack(w); }
    int hoff = (t >> 1) & 1;
    const bool pairfn = argc > 7 && atoi(argv[7]) != 0;   // 1: xy_pair_substep* (what the kernel runs), 0: xy_substep
    auto xy = [&](int oy) {
        if (pairfn) { if (hoff) emu_xy_pair<1>(g, kxy, oy); else emu_xy_pair<0>(g, kxy, oy); }
        else        { if (hoff) emu_xy<1>(g, kxy, oy); else emu_xy<0>(g, kxy, oy); }
    };
    if ((t & 1) == 0) { xy(0); emu_zy(g, kzy, hoff, 1); }
    else              { emu_zy(g, kzy, hoff, 0); xy(1); }
    for (size_t i = 0; i < g.w.size(); ++i) { uint32_t w[8]; unpack(g.w[i], w); memcpy(&bytes[i * 32], w, 32); }
    fwrite(bytes.data(), 1, bytes.size(), stdout);
    return 0;
}
