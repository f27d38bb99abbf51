// bitslice3_host_test.cpp — CPU unit test of fallingsand3d_b200/csrc/bitslice3.cuh (schedule version 2, g++, no GPU).
//
// 1. pack3 / unpack3 (codes <-> rank planes), heavier3, block_rule3 against scalar definitions of SCHEDULE.md §7.
// 2. A word-level emulation of one full version-2 step built from the SAME helper functions the kernels run
//    (xy3_pair_substep0/1 + first bits / carry / wall + block_rule3 + coin2_word), written to stdout as a grid so the
//    Python test can compare it with the version-2 oracle.  Test infrastructure only.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>
#include "../../fallingsand3d_b200/csrc/bitslice3.cuh"

using namespace fs3d;

static int bitpos(int x) { return 8 * (x & 3) + ((x >> 2) & 7); }
enum { E = 0, S = 1, W = 2, X = 3, G = 4, O = 5, H = 6, V = 7 };   // V = GRAVEL
static const int RANK[8] = {1, 5, 3, 7, 0, 2, 4, 6};
static const int CODE_OF_RANK[8] = {G, E, O, W, H, S, V, X};
static int yields(int m) { return m == G || m == E || m == O || m == W || m == H; }
static int heav(int u, int l) { return u != X && yields(l) && RANK[u] > RANK[l]; }
static void rule(int &a, int &b, int &c, int &d, int coin, int coin2) {
    if (heav(a, c)) std::swap(a, c);
    if (heav(b, d)) std::swap(b, d);
    if (heav(a, d) && b != X && a != V) std::swap(a, d); else if (heav(b, c) && a != X && b != V) std::swap(b, c);
    if (a != b && yields(a) && yields(b)) { int go = coin; if (go && (a == H || b == H)) go = coin2; if (go) std::swap(a, b); }
}
static P3 from_codes(const int *m) {        // 32 lanes, lane i = bit i (no voxel layout): rank planes
    P3 p{0, 0, 0};
    for (int i = 0; i < 32; ++i) { int r = RANK[m[i]]; p.p0 |= (uint32_t)(r & 1) << i; p.p1 |= (uint32_t)((r >> 1) & 1) << i; p.p2 |= (uint32_t)((r >> 2) & 1) << i; }
    return p;
}
static int code_at(P3 p, int i) { return CODE_OF_RANK[((p.p0 >> i) & 1) | (((p.p1 >> i) & 1) << 1) | (((p.p2 >> i) & 1) << 2)]; }

static int unit_tests() {
    srand(2);
    int bad = 0;
    for (int trial = 0; trial < 20000; ++trial) {
        uint8_t by[32]; uint32_t w[8];
        for (int i = 0; i < 32; ++i) by[i] = rand() & 7;
        memcpy(w, by, 32);
        P3 p = pack3(w);
        for (int x = 0; x < 32; ++x) if (code_at(p, bitpos(x)) != by[x]) bad++;
        uint32_t o[8]; unpack3(p, o);
        if (memcmp(o, w, 32)) bad++;
        int a[32], b[32], c[32], d[32]; uint32_t R = (uint32_t)rand() * 65536u ^ (uint32_t)rand(), R2 = (uint32_t)rand() * 65536u ^ (uint32_t)rand();
        for (int i = 0; i < 32; ++i) { a[i] = rand() & 7; b[i] = rand() & 7; c[i] = rand() & 7; d[i] = rand() & 7; }
        P3 A = from_codes(a), B = from_codes(b), C = from_codes(c), D = from_codes(d);
        uint32_t hv = heavier3(A, C);
        for (int i = 0; i < 32; ++i) if ((int)((hv >> i) & 1) != heav(a[i], c[i])) bad++;
        block_rule3(A, B, C, D, R, R2);
        for (int i = 0; i < 32; ++i) {
            rule(a[i], b[i], c[i], d[i], (R >> i) & 1, (R2 >> i) & 1);
            if (code_at(A, i) != a[i] || code_at(B, i) != b[i] || code_at(C, i) != c[i] || code_at(D, i) != d[i]) bad++;
        }
    }
    return bad;
}

// ---- word-level emulation of one step --------------------------------------------------------------
struct Grid {
    int nx, ny, nz, wpr;
    std::vector<P3> w;   // [z][y][xw]
    P3 &at(int z, int y, int xw) { return w[((size_t)z * ny + y) * wpr + xw]; }
};
static const P3 STONE3 = {ONES, ONES, ONES};
static P3 rdw(Grid &g, int z, int y, int xw) { if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) return STONE3; return g.at(z, y, xw); }
static void wrw(Grid &g, int z, int y, int xw, P3 v) { if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) return; g.at(z, y, xw) = v; }

// the exchanged edge words are byte-packed (bitslice3.cuh): every bit outside NB_STONE3 is "don't care" — the kernels put
// the mailbox tag there — so the emulation sets them all
static const uint32_t GARBAGE3 = ~NB_STONE3;
template <int OX>
static void emu_xy_pair(Grid &g, uint32_t key, int oy) {
    for (int z = 0; z < g.nz; z += 2)
        for (int y0 = oy ? -1 : 0; y0 < g.ny; y0 += 2) {
            const int W = g.wpr;
            std::vector<P3> U0(W), L0(W), U1(W), L1(W); std::vector<uint32_t> r0(W), r1(W), first(W), carry(W);
            for (int xw = 0; xw < W; ++xw) {
                U0[xw] = rdw(g, z, y0 + 1, xw); L0[xw] = rdw(g, z, y0, xw); U1[xw] = rdw(g, z + 1, y0 + 1, xw); L1[xw] = rdw(g, z + 1, y0, xw);
                r0[xw] = hash_word(key + (uint32_t)xw * HC1 + (uint32_t)(y0 + 1) * HC2 + (uint32_t)z * HC3);
                r1[xw] = hash_word(key + (uint32_t)xw * HC1 + (uint32_t)(y0 + 1) * HC2 + (uint32_t)(z + 1) * HC3);
                first[xw] = xy3_first_bits(U0[xw], L0[xw], U1[xw], L1[xw]);
            }
            for (int xw = 0; xw < W; ++xw) {
                if (OX == 0) xy3_pair_substep0(U0[xw], L0[xw], U1[xw], L1[xw], r0[xw], r1[xw]);
                else xy3_pair_substep1(U0[xw], L0[xw], U1[xw], L1[xw], r0[xw], r1[xw], xw + 1 < W ? (first[xw + 1] | GARBAGE3) : NB_STONE3, carry[xw]);
            }
            if (OX == 1)
                for (int xw = 0; xw < W; ++xw) {
                    uint32_t en = 0;
                    xy3_pair_post1(U0[xw], L0[xw], U1[xw], L1[xw], xw > 0 ? (carry[xw - 1] | GARBAGE3) : xy3_wall_first(first[0] | GARBAGE3, en));
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
                P3 a = rdw(g, z0, y0 + 1, xw), b = rdw(g, z0 + 1, y0 + 1, xw), c = rdw(g, z0, y0, xw), d = rdw(g, z0 + 1, y0, xw);
                uint32_t rw = hash_word(key + (uint32_t)xw * HC1 + (uint32_t)(y0 + 1) * HC2 + (uint32_t)z0 * HC3);
                block_rule3(a, b, c, d, rw, coin2_word(rw));
                wrw(g, z0, y0 + 1, xw, a); wrw(g, z0 + 1, y0 + 1, xw, b); wrw(g, z0, y0, xw, c); wrw(g, z0 + 1, y0, xw, d);
            }
}

// usage: prog                      -> unit tests
//        prog nx ny nz kxy kzy t   -> reads nx*ny*nz bytes on stdin, writes one emulated version-2 step to stdout
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
    for (size_t i = 0; i < g.w.size(); ++i) { uint32_t w[8]; memcpy(w, &bytes[i * 32], 32); g.w[i] = pack3(w); }
    int hoff = (t >> 1) & 1;
    auto xy = [&](int oy) { if (hoff) emu_xy_pair<1>(g, kxy, oy); else emu_xy_pair<0>(g, kxy, oy); };
    if ((t & 1) == 0) { xy(0); emu_zy(g, kzy, hoff, 1); }
    else              { emu_zy(g, kzy, hoff, 0); xy(1); }
    for (size_t i = 0; i < g.w.size(); ++i) { uint32_t w[8]; unpack3(g.w[i], w); memcpy(&bytes[i * 32], w, 32); }
    fwrite(bytes.data(), 1, bytes.size(), stdout);
    return 0;
}
