/*
 * fs3d_sweep.c — STAND-IN SEQUENTIAL SWEEP (builder-written).  TEST INFRASTRUCTURE ONLY.
 *
 * NOT the reference: the reference snapshot has no update sweep at all (SURVEY.md §0;
 * /root/reference/src/engine/engine.cpp:59-70 only handles events and draws).  This is the
 * classic in-place, bottom-up, one-cell-at-a-time falling-sand sweep, using the same local move
 * predicates as SCHEDULE.md (fall into a lighter cell; slide diagonally in ±x / ±z when the cell
 * below does not yield, the target is lighter and the cell beside is not STONE).  It exists to
 * document order dependence (SCHEDULE.md §6) and for two order-independent checks: material
 * counts, and "a configuration settled under the partitioned schedule is a fixed point of this
 * sweep's fall/slide moves".  Liquid lateral spread is deliberately left out of the fixed-point
 * check's move set via `with_lateral = 0` (a partial water layer never settles under either rule).
 */
#include <stdint.h>
#include <stddef.h>

enum { EMPTY = 0, SAND = 1, WATER = 2, STONE = 3 };

static int density(uint8_t m) { return m == SAND ? 2 : (m == WATER ? 1 : 0); }
static int heavier(uint8_t u, uint8_t l) {
    if (u != SAND && u != WATER) return 0;
    if (l != EMPTY && l != WATER) return 0;
    return density(u) > density(l);
}

/* first direction tried by cell (x, y, z) in sweep `parity`: a small integer hash, so that a lone liquid cell on a full
 * layer performs a random walk (and finds the last hole) instead of orbiting on a fixed cycle */
static int first_dir(int64_t x, int64_t y, int64_t z, uint64_t parity) {
    uint32_t v = (uint32_t)x * 0x9E3779B1u + (uint32_t)y * 0x85EBCA77u + (uint32_t)z * 0xC2B2AE3Du + (uint32_t)parity * 0x27D4EB2Fu;
    v ^= v >> 15; v *= 0x2C1B3C6Du; v ^= v >> 12;
    return (int)(v & 3u);
}

/* One in-place sweep, y ascending (bottom-up), then z, then x.  Returns the number of moves. */
int64_t fs3d_sweep_step(uint8_t *g, int64_t nx, int64_t ny, int64_t nz, int with_lateral, uint64_t parity) {
#define AT(x, y, z) g[(x) + nx * ((y) + ny * (z))]
#define INB(x, y, z) ((x) >= 0 && (x) < nx && (y) >= 0 && (y) < ny && (z) >= 0 && (z) < nz)
    static const int DX[4] = { 1, -1, 0, 0 }, DZ[4] = { 0, 0, 1, -1 };
    int64_t moves = 0;
    for (int64_t y = 0; y < ny; ++y)
        for (int64_t z = 0; z < nz; ++z)
            for (int64_t x = 0; x < nx; ++x) {
                uint8_t m = AT(x, y, z);
                if (m != SAND && m != WATER) continue;
                if (y > 0 && heavier(m, AT(x, y - 1, z))) {
                    uint8_t t = AT(x, y - 1, z); AT(x, y - 1, z) = m; AT(x, y, z) = t; ++moves; continue;
                }
                int moved = 0;
                const int d0 = first_dir(x, y, z, parity);
                for (int k = 0; k < 4 && !moved; ++k) {
                    int dir = (k + d0) & 3;
                    int64_t xs = x + DX[dir], zs = z + DZ[dir];
                    if (y == 0 || !INB(xs, y, zs)) continue;
                    if (AT(xs, y, zs) == STONE) continue;
                    if (heavier(m, AT(xs, y - 1, zs))) {
                        uint8_t t = AT(xs, y - 1, zs); AT(xs, y - 1, zs) = m; AT(x, y, z) = t; ++moves; moved = 1;
                    }
                }
                if (moved || !with_lateral || m != WATER) continue;
                for (int k = 0; k < 4 && !moved; ++k) {
                    int dir = (k + d0) & 3;
                    int64_t xs = x + DX[dir], zs = z + DZ[dir];
                    if (!INB(xs, y, zs)) continue;
                    if (AT(xs, y, zs) == EMPTY) { AT(xs, y, zs) = m; AT(x, y, z) = EMPTY; ++moves; moved = 1; }
                }
            }
    return moves;
#undef AT
#undef INB
}
