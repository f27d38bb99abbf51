/*
 * fs3d_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Scalar, per-cell restatement of SCHEDULE.md (schedule version 1).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it;
 * the product (fallingsand3d_b200/, libfs3d.so) never links, imports or falls back to it.
 *
 * PARITY UNPINNED: the reference snapshot has no voxel simulation, no tests and no golden
 * vectors for this path (grep of /root/reference/src and shaders finds no
 * grid/step code; the frame loop is handleEvents(); draw(); only —
 * /root/reference/src/engine/engine.cpp:59-70).  There is therefore no reference output to pin
 * this oracle against.  What pins it instead: hand-derived micro-scenes (tests/golden/), exact
 * conservation of material counts, and a second independent restatement in numpy
 * (oracle/oracle_np.py) that must agree with it bit-for-bit.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

enum { EMPTY = 0, SAND = 1, WATER = 2, STONE = 3, GAS = 4, OIL = 5, HONEY = 6, GRAVEL = 7 };
enum { AXIS_XY = 0, AXIS_ZY = 1 };

/* ---- SCHEDULE.md §3: the coin ------------------------------------------------------------ */

static uint64_t mix64(uint64_t v) {
    v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ull;
    v ^= v >> 27; v *= 0x94D049BB133111EBull;
    v ^= v >> 31;
    return v;
}

uint32_t fs3d_oracle_key(uint64_t seed, uint64_t t, uint32_t axis) {
    uint64_t v = seed ^ (t * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(axis + 1) * 0xD1B54A32D192ED03ull);
    v = mix64(v);
    return (uint32_t)(v ^ (v >> 32));
}

uint32_t fs3d_oracle_hash(uint32_t key, uint32_t xw, uint32_t y, uint32_t z) {
    uint32_t v = key + xw * 0x9E3779B1u + y * 0x85EBCA77u + z * 0xC2B2AE3Du;
    v ^= v >> 16; v *= 0x7FEB352Du;
    v ^= v >> 15; v *= 0x846CA68Bu;
    v ^= v >> 16;
    return v;
}

static int coin_at(uint32_t key, int64_t X, int64_t Y, int64_t Z) {
    uint32_t x = (uint32_t)X;
    uint32_t bit = 8u * (x & 3u) + ((x >> 2) & 7u);
    return (int)((fs3d_oracle_hash(key, x >> 5, (uint32_t)Y, (uint32_t)Z) >> bit) & 1u);
}

int fs3d_oracle_coin(uint64_t seed, uint64_t t, uint32_t axis, uint32_t X, uint32_t Y, uint32_t Z) {
    return coin_at(fs3d_oracle_key(seed, t, axis), X, Y, Z);
}

/* SCHEDULE.md §7 (version 2): the second coin of a block, drawn from the same hash word through C2 */
static int coin2_at(uint32_t key, int64_t X, int64_t Y, int64_t Z) {
    uint32_t x = (uint32_t)X;
    uint32_t bit = 8u * (x & 3u) + ((x >> 2) & 7u);
    uint32_t v = fs3d_oracle_hash(key, x >> 5, (uint32_t)Y, (uint32_t)Z) * 0x9E3779B1u;
    v ^= v >> 15;
    return (int)((v >> bit) & 1u);
}

int fs3d_oracle_coin2(uint64_t seed, uint64_t t, uint32_t axis, uint32_t X, uint32_t Y, uint32_t Z) {
    return coin2_at(fs3d_oracle_key(seed, t, axis), X, Y, Z);
}

/* ---- SCHEDULE.md §1-2: the block rule ------------------------------------------------------ */

static int density(uint8_t m) { return m == SAND ? 2 : (m == WATER ? 1 : 0); }

static int heavier(uint8_t u, uint8_t l) {
    if (u != SAND && u != WATER) return 0;
    if (l != EMPTY && l != WATER) return 0;
    return density(u) > density(l);
}

static void swap8(uint8_t *p, uint8_t *q) { uint8_t t = *p; *p = *q; *q = t; }

/* ---- SCHEDULE.md §7: schedule version 2, eight materials ----------------------------------------
 * rank (density order): GAS 0 < EMPTY 1 < OIL 2 < WATER 3 < HONEY 4 < SAND 5 < GRAVEL 6; STONE never moves.
 * A cell yields to a denser mover only if it is not granular / solid: GAS, EMPTY, OIL, WATER, HONEY. */
static const int RANK_V2[8] = { /*EMPTY*/ 1, /*SAND*/ 5, /*WATER*/ 3, /*STONE*/ 7, /*GAS*/ 0, /*OIL*/ 2, /*HONEY*/ 4, /*GRAVEL*/ 6 };
static int yields_v2(uint8_t m) { return m == GAS || m == EMPTY || m == OIL || m == WATER || m == HONEY; }
static int heavier_v2(uint8_t u, uint8_t l) {
    if (u == STONE || !yields_v2(l)) return 0;
    return RANK_V2[u & 7] > RANK_V2[l & 7];
}
static int block_rule_v2(uint8_t *a, uint8_t *b, uint8_t *c, uint8_t *d, uint32_t key, int64_t X, int64_t Y, int64_t Z) {
    int enabled = 0;
    /* F */
    if (heavier_v2(*a, *c)) { swap8(a, c); enabled = 1; }
    if (heavier_v2(*b, *d)) { swap8(b, d); enabled = 1; }
    /* D: GRAVEL never slides */
    if (heavier_v2(*a, *d) && *b != STONE && *a != GRAVEL) { swap8(a, d); enabled = 1; }
    else if (heavier_v2(*b, *c) && *a != STONE && *b != GRAVEL) { swap8(b, c); enabled = 1; }
    /* L: any two different yielding cells mix sideways under the coin; HONEY is viscous: it also needs the second coin */
    if (*a != *b && yields_v2(*a) && yields_v2(*b)) {
        enabled = 1;
        int go = coin_at(key, X, Y, Z);
        if (go && (*a == HONEY || *b == HONEY)) go = coin2_at(key, X, Y, Z);
        if (go) swap8(a, b);
    }
    return enabled;
}

/* Applies F, D, L to one block whose upper-left cell `a` is global cell (X, Y, Z).  Returns 1 if the
 * block is "enabled" (something would move with coin = 1), else 0.  The coin is only drawn when L is
 * enabled: then a and b are WATER/EMPTY, i.e. both inside the grid, so (X, Y, Z) is in range
 * (SCHEDULE.md §3 "the coin only matters when both upper cells are inside the grid"). */
static int block_rule_v1(uint8_t *a, uint8_t *b, uint8_t *c, uint8_t *d, uint32_t key, int64_t X, int64_t Y, int64_t Z) {
    int enabled = 0;
    /* F */
    if (heavier(*a, *c)) { swap8(a, c); enabled = 1; }
    if (heavier(*b, *d)) { swap8(b, d); enabled = 1; }
    /* D */
    if (heavier(*a, *d) && *b != STONE) { swap8(a, d); enabled = 1; }
    else if (heavier(*b, *c) && *a != STONE) { swap8(b, c); enabled = 1; }
    /* L */
    if ((*a == WATER && *b == EMPTY) || (*b == WATER && *a == EMPTY)) {
        enabled = 1;
        if (coin_at(key, X, Y, Z)) swap8(a, b);
    }
    return enabled;
}

#define block_rule(a, b, c, d, key, X, Y, Z) \
    (g->version == 2 ? block_rule_v2(a, b, c, d, key, X, Y, Z) : block_rule_v1(a, b, c, d, key, X, Y, Z))

/* ---- grid access: arr holds global planes [zbase, zbase + narr) --------------------------- */

typedef struct {
    uint8_t *arr;
    int64_t nx, ny, nzg;   /* global dims */
    int64_t zbase, narr;   /* planes held */
    int version;           /* schedule version: 1 (four materials) or 2 (eight, SCHEDULE.md §7) */
} grid_t;

static uint8_t rd(const grid_t *g, int64_t x, int64_t y, int64_t z) {
    if (x < 0 || x >= g->nx || y < 0 || y >= g->ny || z < 0 || z >= g->nzg) return STONE;
    if (z < g->zbase || z >= g->zbase + g->narr) return STONE; /* not held: caller never needs it */
    return g->arr[x + g->nx * (y + g->ny * (z - g->zbase))];
}

static void wr(const grid_t *g, int64_t x, int64_t y, int64_t z, uint8_t m) {
    if (x < 0 || x >= g->nx || y < 0 || y >= g->ny || z < 0 || z >= g->nzg) return;
    if (z < g->zbase || z >= g->zbase + g->narr) return;
    g->arr[x + g->nx * (y + g->ny * (z - g->zbase))] = m;
}

static int first_origin(int64_t lo, int off) {
    /* smallest h0 >= lo - 1 with h0 ≡ off (mod 2) — blocks {h0, h0+1} then cover lo */
    int64_t h0 = lo - 1;
    if (((h0 % 2) + 2) % 2 != off) h0 += 1;
    return (int)h0;
}

/* A block through the bounds-checked accessors (cells outside the grid read STONE and are not written). */
static int block_checked(const grid_t *g, uint32_t key, int64_t xa, int64_t ya, int64_t za, int64_t xb, int64_t zb) {
    /* a = (xa, ya, za), b = (xb, ya, zb) upper row; c, d the cells below them */
    uint8_t a = rd(g, xa, ya, za), b = rd(g, xb, ya, zb), c = rd(g, xa, ya - 1, za), d = rd(g, xb, ya - 1, zb);
    int en = block_rule(&a, &b, &c, &d, key, xa, ya, za);
    wr(g, xa, ya, za, a); wr(g, xb, ya, zb, b); wr(g, xa, ya - 1, za, c); wr(g, xb, ya - 1, zb, d);
    return en;
}

static int held(const grid_t *g, int64_t z) { return z >= 0 && z < g->nzg && z >= g->zbase && z < g->zbase + g->narr; }

/* One XY sub-step on planes z in [zlo, zhi). Returns number of enabled blocks.
 * Blocks are disjoint, so (plane, block-row) pairs are independent work items: the loop nest is
 * collapsed over both so that a thin slab sample still feeds every host thread.  Blocks that lie
 * wholly inside the held planes take the direct-pointer path; the rest go through rd()/wr(). */
static int64_t substep_xy(const grid_t *g, uint32_t key, int ox, int oy, int64_t zlo, int64_t zhi) {
    int64_t enabled = 0;
    const int64_t ystart = first_origin(0, oy), xstart = first_origin(0, ox);
    const int64_t nyb = (g->ny - ystart + 1) / 2;   /* y0 = ystart + 2 yb < ny */
    int64_t z, yb;
#pragma omp parallel for collapse(2) reduction(+ : enabled) schedule(static)
    for (z = zlo; z < zhi; ++z) {
        for (yb = 0; yb < nyb; ++yb) {
            const int64_t y0 = ystart + 2 * yb;
            int64_t x0 = xstart;
            if (y0 >= 0 && y0 + 1 < g->ny && held(g, z)) {
                uint8_t *lo = g->arr + g->nx * (y0 + g->ny * (z - g->zbase)), *up = lo + g->nx;
                if (x0 < 0) { enabled += block_checked(g, key, x0, y0 + 1, z, x0 + 1, z); x0 += 2; }
                for (; x0 + 1 < g->nx; x0 += 2)
                    enabled += block_rule(up + x0, up + x0 + 1, lo + x0, lo + x0 + 1, key, x0, y0 + 1, z);
            }
            for (; x0 < g->nx; x0 += 2) enabled += block_checked(g, key, x0, y0 + 1, z, x0 + 1, z);
        }
    }
    return enabled;
}

/* One ZY sub-step over the blocks that intersect global planes [zlo, zhi). */
static int64_t substep_zy(const grid_t *g, uint32_t key, int oz, int oy, int64_t zlo, int64_t zhi) {
    int64_t enabled = 0;
    const int64_t zstart = first_origin(zlo, oz), ystart = first_origin(0, oy);
    const int64_t nblk = (zhi - zstart + 1) / 2; /* z0 = zstart + 2k < zhi */
    const int64_t nyb = (g->ny - ystart + 1) / 2;
    int64_t k, yb;
#pragma omp parallel for collapse(2) reduction(+ : enabled) schedule(static)
    for (k = 0; k < nblk; ++k) {
        for (yb = 0; yb < nyb; ++yb) {
            const int64_t z0 = zstart + 2 * k, y0 = ystart + 2 * yb;
            if (y0 >= 0 && y0 + 1 < g->ny && held(g, z0) && held(g, z0 + 1)) {
                uint8_t *lo0 = g->arr + g->nx * (y0 + g->ny * (z0 - g->zbase)), *up0 = lo0 + g->nx;
                uint8_t *lo1 = lo0 + g->nx * g->ny, *up1 = lo1 + g->nx;
                for (int64_t x = 0; x < g->nx; ++x)
                    enabled += block_rule(up0 + x, up1 + x, lo0 + x, lo1 + x, key, x, y0 + 1, z0);
            } else {
                for (int64_t x = 0; x < g->nx; ++x) enabled += block_checked(g, key, x, y0 + 1, z0, x, z0 + 1);
            }
        }
    }
    return enabled;
}

/*
 * One full step (SCHEDULE.md §2) on an array holding global planes [zbase, zbase + narr) of an
 * nx × ny × nzg world, of which [own_lo, own_hi) are owned (the rest are ghost planes holding
 * the neighbour slabs' cells; they are read, and scribbled on, but never trusted afterwards).
 * A whole grid is zbase = 0, narr = nzg, own = [0, nzg).  Returns the number of enabled blocks
 * that intersect the owned planes' dependency region (0 ⇒ nothing could move this step).
 */
int64_t fs3d_oracle_step_range_v(uint8_t *arr, int64_t nx, int64_t ny, int64_t nzg,
                                 int64_t zbase, int64_t narr, int64_t own_lo, int64_t own_hi,
                                 uint64_t seed, uint64_t t, int version) {
    grid_t g = { arr, nx, ny, nzg, zbase, narr, version };
    int hoff = (int)((t >> 1) & 1);
    uint32_t kxy = fs3d_oracle_key(seed, t, AXIS_XY), kzy = fs3d_oracle_key(seed, t, AXIS_ZY);
    int64_t en = 0;
    if ((t & 1) == 0) {
        en += substep_xy(&g, kxy, hoff, 0, zbase, zbase + narr);
        en += substep_zy(&g, kzy, hoff, 1, own_lo, own_hi);
    } else {
        en += substep_zy(&g, kzy, hoff, 0, own_lo, own_hi);
        en += substep_xy(&g, kxy, hoff, 1, zbase, zbase + narr);
    }
    return en;
}

int64_t fs3d_oracle_step_range(uint8_t *arr, int64_t nx, int64_t ny, int64_t nzg,
                               int64_t zbase, int64_t narr, int64_t own_lo, int64_t own_hi,
                               uint64_t seed, uint64_t t) {
    return fs3d_oracle_step_range_v(arr, nx, ny, nzg, zbase, narr, own_lo, own_hi, seed, t, 1);
}

int64_t fs3d_oracle_step_v(uint8_t *grid, int64_t nx, int64_t ny, int64_t nz, uint64_t seed, uint64_t t, int version) {
    return fs3d_oracle_step_range_v(grid, nx, ny, nz, 0, nz, 0, nz, seed, t, version);
}

int64_t fs3d_oracle_step(uint8_t *grid, int64_t nx, int64_t ny, int64_t nz, uint64_t seed, uint64_t t) {
    return fs3d_oracle_step_v(grid, nx, ny, nz, seed, t, 1);
}

void fs3d_oracle_run_v(uint8_t *grid, int64_t nx, int64_t ny, int64_t nz, uint64_t seed,
                       uint64_t t0, uint64_t nsteps, int version) {
    for (uint64_t i = 0; i < nsteps; ++i) fs3d_oracle_step_v(grid, nx, ny, nz, seed, t0 + i, version);
}

void fs3d_oracle_run(uint8_t *grid, int64_t nx, int64_t ny, int64_t nz, uint64_t seed,
                     uint64_t t0, uint64_t nsteps) {
    fs3d_oracle_run_v(grid, nx, ny, nz, seed, t0, nsteps, 1);
}

/* ---- SCHEDULE.md §5: scenes ---------------------------------------------------------------- */

static int in_box(int64_t x, int64_t y, int64_t z, int64_t nx, int64_t ny, int64_t nz, const int *b) {
    /* b = {x0,x1,y0,y1,z0,z1} in 64ths of each dimension, half-open */
    return x >= b[0] * nx / 64 && x < b[1] * nx / 64 && y >= b[2] * ny / 64 && y < b[3] * ny / 64 &&
           z >= b[4] * nz / 64 && z < b[5] * nz / 64;
}

static const int STONE_BOXES[8][6] = {
    {  8, 28, 20, 22,  8, 28 }, { 36, 56, 20, 22, 36, 56 }, { 30, 34,  1, 30, 30, 34 },
    {  8, 28, 32, 34, 36, 56 }, { 36, 56, 32, 34,  8, 28 }, { 20, 22,  1, 12,  4, 60 },
    {  4, 60,  1, 10, 42, 44 }, { 44, 52,  1,  6, 12, 20 },
};
static const int SAND_BOX[6]  = { 10, 30, 44, 60, 10, 54 };
static const int WATER_BOX[6] = { 34, 54, 44, 60, 10, 54 };

static uint8_t random_cell(uint64_t seed, int64_t x, int64_t y, int64_t z) {
    uint32_t u = fs3d_oracle_hash(fs3d_oracle_key(seed, 0, 7), (uint32_t)x, (uint32_t)y, (uint32_t)z) & 3u;
    return u == 0 ? SAND : (u == 1 ? WATER : EMPTY);
}

/* SCHEDULE.md §7 scenes: RANDOM8 draws from all seven movable materials, half of the cells stay EMPTY */
static uint8_t random8_cell(uint64_t seed, int64_t x, int64_t y, int64_t z) {
    static const uint8_t PICK[16] = { SAND, SAND, WATER, WATER, OIL, GAS, HONEY, GRAVEL, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY };
    return PICK[(fs3d_oracle_hash(fs3d_oracle_key(seed, 0, 7), (uint32_t)x, (uint32_t)y, (uint32_t)z) >> 4) & 15u];
}
static const int GAS_BOX[6]    = { 10, 26, 12, 19, 10, 26 };   /* under the first stone ledge: rises and pools below it */
static const int OIL_BOX[6]    = { 34, 54, 36, 43, 10, 30 };
static const int HONEY_BOX[6]  = { 38, 50, 24, 31, 38, 54 };
static const int GRAVEL_BOX[6] = { 12, 24, 36, 43, 38, 54 };

static uint8_t mixed_cell(int64_t nx, int64_t ny, int64_t nz, int64_t x, int64_t y, int64_t z) {
    int64_t floor_h = ny / 64 > 1 ? ny / 64 : 1;
    if (y < floor_h) return STONE;
    for (int i = 0; i < 8; ++i) if (in_box(x, y, z, nx, ny, nz, STONE_BOXES[i])) return STONE;
    if (in_box(x, y, z, nx, ny, nz, SAND_BOX)) return SAND;
    if (in_box(x, y, z, nx, ny, nz, WATER_BOX)) return WATER;
    return EMPTY;
}

uint8_t fs3d_oracle_scene_cell(int scene, uint64_t seed, int64_t nx, int64_t ny, int64_t nz,
                               int64_t x, int64_t y, int64_t z) {
    switch (scene) {
    case 1:
        return (x >= 3 * nx / 8 && x < 5 * nx / 8 && z >= 3 * nz / 8 && z < 5 * nz / 8 &&
                y >= 5 * ny / 8 && y < 7 * ny / 8) ? SAND : EMPTY;
    case 2: return mixed_cell(nx, ny, nz, x, y, z);
    case 3: return random_cell(seed, x, y, z);
    case 4: {
        uint8_t m = mixed_cell(nx, ny, nz, x, y, z);
        if (m == EMPTY && y >= ny / 2) m = random_cell(seed, x, y, z);
        return m;
    }
    case 5: return random8_cell(seed, x, y, z);
    case 6: {   /* MIXED8: the MIXED layout, boxes of the four new materials, RANDOM8 noise in the empty upper half */
        uint8_t m = mixed_cell(nx, ny, nz, x, y, z);
        if (m != EMPTY) return m;
        if (in_box(x, y, z, nx, ny, nz, GAS_BOX)) return GAS;
        if (in_box(x, y, z, nx, ny, nz, OIL_BOX)) return OIL;
        if (in_box(x, y, z, nx, ny, nz, HONEY_BOX)) return HONEY;
        if (in_box(x, y, z, nx, ny, nz, GRAVEL_BOX)) return GRAVEL;
        if (y >= ny / 2) m = random8_cell(seed, x, y, z);
        return m;
    }
    default: return EMPTY;
    }
}

/* Fills planes [zlo, zhi) of the global scene into out (plane-major, zhi - zlo planes). */
void fs3d_oracle_generate(uint8_t *out, int64_t nx, int64_t ny, int64_t nz, int64_t zlo, int64_t zhi,
                          int scene, uint64_t seed) {
    int64_t z;
#pragma omp parallel for schedule(static)
    for (z = zlo; z < zhi; ++z)
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t x = 0; x < nx; ++x)
                out[x + nx * (y + ny * (z - zlo))] = fs3d_oracle_scene_cell(scene, seed, nx, ny, nz, x, y, z);
}

/* ---- SCHEDULE.md §4: histogram and digest -------------------------------------------------- */

void fs3d_oracle_histogram(const uint8_t *grid, int64_t ncells, uint64_t counts[256]) {
    memset(counts, 0, 256 * sizeof(uint64_t));
    for (int64_t i = 0; i < ncells; ++i) counts[grid[i]]++;
}

/* grid holds planes [zlo, zhi) of an nx × ny × * world; idx is the GLOBAL linear index. */
uint64_t fs3d_oracle_digest(const uint8_t *grid, int64_t nx, int64_t ny, int64_t zlo, int64_t zhi) {
    uint64_t sum = 0;
    int64_t plane = nx * ny, z;
#pragma omp parallel for reduction(+ : sum) schedule(static)
    for (z = zlo; z < zhi; ++z) {
        const uint8_t *p = grid + (z - zlo) * plane;
        uint64_t base = (uint64_t)z * (uint64_t)plane;
        for (int64_t i = 0; i < plane; ++i)
            if (p[i]) sum += mix64(8ull * (base + (uint64_t)i) + p[i]);
    }
    return sum;
}

int fs3d_oracle_schedule_version(void) { return 1; }

/* host threads the parallel loops above actually get (1 without OpenMP) */
#ifdef _OPENMP
#include <omp.h>
int fs3d_oracle_threads(void) {
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
}
#else
int fs3d_oracle_threads(void) { return 1; }
#endif
