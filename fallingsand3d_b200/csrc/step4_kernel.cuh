// step4_kernel.cuh — FOUR SCHEDULE.md steps per pass (0.5 B of DRAM traffic per voxel-update), schedule version 1.
//
// Steps t .. t+3 with t ≡ 0 (mod 4): steps t, t+1 use hoff = 0 ("stage A": x-offset 0, ZY pairs (0,1), (2,3), …), steps
// t+2, t+3 use hoff = 1 ("stage B": x-offset 1, ZY pairs (−1,0), (1,2), …).  Inside a stage a z-pair of rows is closed
// under both steps (that is the two-step pass of step_kernel.cuh); between the stages the pairing shifts by one row, so
// stage B's pair q = (plane 2q, 2q+1) takes the upper row of stage-A pair q−1 and the lower row of stage-A pair q.
//
// A warp stays autonomous (no barrier, no other warp's data): it owns a BAND of P (= 4) consecutive B-pairs and walks the
// grid in y-blocks of K (= 4) iterations; inside a y-block it visits the band's P+1 A-pairs one after the other —
//     load pair a  →  stage A (two steps)  →  its upper row's finished planes go into a K-deep shared-memory buffer,
//     stage B (two more steps) on (the previous pair's upper row out of that buffer, this pair's lower row)  →  store
// — and parks each pair's pipeline state (the carried planes of both stages: 28 words per lane) in shared memory until
// the next y-block.  The first A-pair of a band is recomputed by the band below ((P+1)/P of the loads and of stage A);
// everything else is done once.  DRAM sees every byte read (P+1)/P times and written once per FOUR steps — on big
// single slabs 25/24 times: the units of a CTA then take neighbouring bands, staggered so that the shared pair's second
// load hits L2 (`grp` below).
// Rows of 2048 / 4096 voxels use two / four warps per band (XW = 2, 4); neighbouring warps exchange the word-boundary
// cells of stage B's odd x-offset through the tagged mailboxes of step_kernel.cuh.
//
// Used by fs3d_step for worlds of schedule version 1 without skipping whose rows are 1024, 2048 or 4096 voxels wide, when four
// steps remain and the step index is a multiple of four: single slabs, and z-slabs with fused-push neighbours (NBR = 1:
// two ghost planes per side, delivered by halo4_kernel after the pass), and by fs3d_step_host(…, 4) on chunks of bands.
// Everything else keeps the two-step pass.  Results are identical by construction and tested against the oracle (every
// parity test with nx in {1024, 2048, 4096}, the full 2048^3 compare, bench digests).
// No reference counterpart; call site: /root/reference/src/engine/engine.cpp:59-70.
#pragma once
#include <type_traits>
#include "step_kernel.cuh"

namespace fs3d {

// Band size P, y-block K and CTA size were chosen by A/B on a B200 (profiles/r02m_experiments_step4.txt): P = 4, K = 4
// leaves 17.5 KB of parked state per warp, so TWELVE warps fit an SM (384-thread CTAs) instead of eight with P = 6,
// K = 8 — 7 % faster at 2048^3 although 5/4 instead of 7/6 of stage A is recomputed (the kernel is ALU-bound and latency
// hiding wins), the same at 1024^3.  Still the best shape after the ALU diet (profiles/r02p_experiments_alu.txt).
#ifndef FS3D_S4_P
#define FS3D_S4_P 4
#endif
#ifndef FS3D_S4_K
#define FS3D_S4_K 4
#endif
#ifndef FS3D_S4_ONE_ISSUE
#define FS3D_S4_ONE_ISSUE 1      // one load-issue site per iteration instead of three (1 % faster, a tenth less code)
#endif
constexpr int S4_P = FS3D_S4_P;    // B-pairs per band        (tuning hooks: profiles/r02m_experiments_step4.txt)
constexpr int S4_K = FS3D_S4_K;    // iterations per y-block
constexpr uint32_t S4_LEAD = 7;    // warm-up iterations that rebuild both stages' carried planes
constexpr uint32_t S4_WORDS_PER_LANE = (S4_P + 1) * 12 + S4_P * 16 + S4_K * 4;     // A states, B states, hand-off buffer

struct Step4Params {
    const uint8_t *src;
    uint8_t *dst;
    uint32_t nx, ny, wpr;
    uint32_t z0;                   // global z of local plane 1 (even)
    uint32_t nzl;                  // owned planes (local 1 .. nzl); local plane 0 / nzl+1 are the near ghosts
    uint32_t nA, nB;               // stage-A pairs a: local planes (1+2a, 2+2a); stage-B pairs q: local planes (2q, 2q+1)
    uint32_t nbands;               // bands this launch marches: ceil(nB / S4_P), or a chunk of them (fs3d_step_host)
    uint32_t band0;                // first band of this launch (single slabs only; 0 with neighbours)
    uint32_t nit;                  // march iterations: ny / 2 + 4
    uint32_t key_xy[4], key_zy[4]; // SCHEDULE.md §3 keys of steps t .. t+3
    // z-slabs (NBR = 1 instantiations): a neighbour holds the planes below / above.  Stage B's pair across a slab boundary
    // needs the neighbour's edge row AFTER stage A, so the neighbour's edge A-pair is recomputed here from TWO ghost
    // planes per side: the near ghosts (local planes 0 and nzl+1) and the far ghosts, which live behind the slab in the
    // same buffer — local plane nzl+2 = global z1+1 (so A-pair nA is contiguous) and local plane nzl+3 = global z0-2.
    // halo4_kernel delivers all four after every four-step pass and bumps the arrival counters of step_kernel.cuh.
    int has_lo, has_hi;
    const unsigned long long *my_flags;   // arrive[0] (from below), arrive[1] (from above)
    unsigned long long wait_target;
    unsigned long long *push_err;
    unsigned long long push_timeout_ns;
    int edge_late;                 // NBR: units take their share of the two edge bands AFTER their share of the interior bands
    int groups;                    // single slab: CTAs own spans of (group of neighbouring bands x iteration), units staggered
};

// Delivers this slab's two edge rows on each side into the z-neighbours' near and far ghost planes (peer memory) and
// adds gridDim.x to their arrival counters.  near = 0: only the far planes (refresh before a four-step pass that follows
// two-step passes, whose kernels push the near planes themselves).
struct Halo4Params {
    const uint8_t *src;            // my buffer (the one the neighbours will read ghosts of)
    uint8_t *lo_buf, *hi_buf;      // the same buffer of the neighbour below / above, or nullptr
    uint32_t nzl, lo_nzl, hi_nzl;
    uint64_t plane_bytes;
    unsigned long long *lo_flag, *hi_flag;   // neighbour below: its arrive[1]; above: its arrive[0]
    int near;
};
constexpr unsigned HALO4_BLOCKS = 32;
static __global__ void halo4_kernel(const Halo4Params h) {
    const uint64_t n16 = h.plane_bytes / 16;
    auto copy = [&](const uint8_t *s, uint8_t *d) {
        for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n16; v += (uint64_t)gridDim.x * blockDim.x)
            reinterpret_cast<uint4 *>(d)[v] = reinterpret_cast<const uint4 *>(s)[v];
    };
    if (h.lo_buf) {    // the neighbour below reads my planes 1 (z0) and 2 (z0+1) as its near ghost-high and far ghost-high
        if (h.near) copy(h.src + 1 * h.plane_bytes, h.lo_buf + (uint64_t)(h.lo_nzl + 1) * h.plane_bytes);
        copy(h.src + 2 * h.plane_bytes, h.lo_buf + (uint64_t)(h.lo_nzl + 2) * h.plane_bytes);
    }
    if (h.hi_buf) {    // the neighbour above reads my planes nzl (z1-1) and nzl-1 (z1-2) as its near ghost-low and far ghost-low
        if (h.near) copy(h.src + (uint64_t)h.nzl * h.plane_bytes, h.hi_buf);
        copy(h.src + (uint64_t)(h.nzl - 1) * h.plane_bytes, h.hi_buf + (uint64_t)(h.hi_nzl + 3) * h.plane_bytes);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (h.lo_flag) atomicAdd_system(h.lo_flag, 1ull);
        if (h.hi_flag) atomicAdd_system(h.hi_flag, 1ull);
    }
}

template <int THREADS>
constexpr uint32_t step4_smem_bytes() { return (THREADS / 32) * S4_WORDS_PER_LANE * 32u * 4u; }

template <int XW, int NBR, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) step4_kernel(const Step4Params p) {
    using R = Rules1;
    using Cell = P2;
    static_assert(XW == 1 || XW == 2 || XW == 4, "one warp per 1024 voxels of a row");
    static_assert((THREADS / 32) % XW == 0, "whole bands per CTA");
    constexpr bool XCH = XW > 1;
    constexpr uint32_t UNITS = THREADS / 32 / XW;           // bands marched side by side in a CTA
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wic = threadIdx.x >> 5;
    const uint32_t pic = wic / XW, half = wic % XW;
    const uint32_t gw = blockIdx.x * UNITS + pic, nw = gridDim.x * UNITS;
    // one mailbox (two slots) per boundary between neighbouring warps of a band
    constexpr uint32_t NBOX = XCH ? (XW - 1) * 2 : 1;
    __shared__ uint32_t xch_smem[XCH ? UNITS * NBOX : 1];
    volatile uint32_t *const xch = xch_smem + (XCH ? pic * NBOX : 0);
    uint32_t xseq = 0;
    if (XCH) {
        if (threadIdx.x < UNITS * NBOX) xch_smem[threadIdx.x] = 0u;
        __syncthreads();
    }
    extern __shared__ __align__(16) uint32_t s4_smem[];
    uint32_t *const sm = s4_smem + (size_t)wic * S4_WORDS_PER_LANE * 32u + lane;     // word w of this lane: sm[w * 32]
    auto sA = [&](int k, int w) -> uint32_t & { return sm[(k * 12 + w) * 32]; };
    auto sB = [&](int k, int w) -> uint32_t & { return sm[((S4_P + 1) * 12 + (k - 1) * 16 + w) * 32]; };
    auto sH = [&](int slot, int w) -> uint32_t & { return sm[((S4_P + 1) * 12 + S4_P * 16 + slot * 4 + w) * 32]; };

    const uint32_t xw = lane + half * 32u;                 // this lane's 32-voxel word of the row
    const bool hasp = xw > 0u, hasn = xw + 1u < p.wpr;
    const size_t row_bytes = p.nx, plane_bytes = (size_t)p.nx * p.ny;
    const uint32_t ylast = p.ny - 1u;
    // Work split.  The (band x iteration) space is cut into contiguous spans, one per unit.  With neighbours the two edge
    // bands (permuted to the end of the position order) must wait for the neighbours' deliveries of the previous pass: a
    // unit that owned a span of an edge band would wait at the very start of the kernel, putting every bit of inter-GPU
    // skew on the critical path.  So (edge_late) the edge bands are cut into chunks that some units take as a SECOND span
    // behind a correspondingly shorter interior span: edge work starts in the second half of the kernel and skew up
    // to that much costs nothing.
    const uint64_t total = (uint64_t)p.nbands * p.nit;
    uint64_t lo0 = total * gw / nw, hi0 = total * (gw + 1) / nw, lo1 = 0, hi1 = 0;     // first span, second span (edge chunk)
    // Single slab with long spans (p.groups, set by the host): the units of a CTA take NEIGHBOURING bands and march the
    // same iterations of them side by side, each one y-block (in y) behind the unit below it.  The A-pair two
    // neighbouring bands both load — the lower band as its last pair, the upper one as its first — is then read twice
    // within one pair-visit's time and the second read hits L2 (those loads say evict_last, all others evict_first):
    // DRAM reads fall from 5/4 to 25/24 of the grid.  The CTA, not the unit, owns a contiguous span of the (group of
    // UNITS bands x iteration) space; inner segment boundaries of unit `pic` move down by pic * K iterations, so a
    // band's segments still tile [0, nit).  Without the stagger the second read comes four visits BEFORE the first one
    // of the next y-block and misses (profiles/r02r_experiments_groups.txt).
    const bool grp = !NBR && p.groups != 0;
    if (grp) {
        const uint64_t totalg = (uint64_t)((p.nbands + UNITS - 1u) / UNITS) * p.nit;
        // equal spans, except that the last CTA's is (UNITS - 1) * K shorter: nothing follows the last group, so its
        // top unit would otherwise run that much longer than every other unit of the grid
        const uint64_t per = (totalg + (UNITS - 1u) * (uint64_t)S4_K + gridDim.x - 1u) / gridDim.x;
        lo0 = per * blockIdx.x < totalg ? per * blockIdx.x : totalg;
        hi0 = per * (blockIdx.x + 1u) < totalg ? per * (blockIdx.x + 1u) : totalg;
    }

    // Messages between neighbouring warps of a band (the tagged mailboxes of step_kernel.cuh, one per warp boundary).
    // xmail<0>: every warp hands `payload` of its lane 0 DOWN to the warp on its left and returns what the warp on its
    // right sent (undefined for the last warp); xmail<1>: lane 31's payload goes UP to the right, the left one's comes
    // back (undefined for the first warp).  On every boundary the two directions alternate strictly, so message k + 2
    // reuses the slot of message k only after its reader has answered message k + 1; four tag bits are plenty.
    auto xmail = [&](auto dir, uint32_t payload) -> uint32_t {
        constexpr uint32_t UP = decltype(dir)::value;
        ++xseq;
        const uint32_t tag = xseq << 28;
        const uint32_t s = xseq & 1u;
        const bool sends = UP ? half + 1u < (uint32_t)XW : half > 0u, receives = UP ? half > 0u : half + 1u < (uint32_t)XW;
        // boundary b lies between warps b and b + 1
        if (sends && lane == (UP ? 31u : 0u)) xch[(UP ? half : half - 1u) * 2u + s] = (payload & 0x0FFFFFFFu) | tag;
        uint32_t v = 0u;
        if (receives) {
            volatile uint32_t *slot = xch + (UP ? half - 1u : half) * 2u + s;
            do { v = *slot; } while ((v & 0xF0000000u) != tag);
        }
        return v;                                         // Rules1's edge bits tolerate the tag (bitslice.cuh, E-packing)
    };
    // XY sub-step on (upper, lower) of both rows; x-offset 0 (stage A) or 1 (stage B, cells of the word-straddling block
    // cross lanes by shuffle and the warp pair's boundary through the mailbox)
    auto xy0 = [&](Cell (&up)[2], Cell (&lw)[2], uint32_t yu, uint32_t key, const uint32_t (&hxy)[2]) {
        R::xy0(up[0], lw[0], up[1], lw[1], hash_word(key + hxy[0] + yu * HC2), hash_word(key + hxy[1] + yu * HC2));
    };
    auto xy1 = [&](Cell (&up)[2], Cell (&lw)[2], uint32_t yu, uint32_t key, const uint32_t (&hxy)[2]) {
        const uint32_t r0 = hash_word(key + hxy[0] + yu * HC2), r1 = hash_word(key + hxy[1] + yu * HC2);
        const uint32_t first = R::first_bits(up[0], lw[0], up[1], lw[1]);
        uint32_t xin = R::NB_STONE;
        if (XCH) xin = xmail(std::integral_constant<uint32_t, 0u>{}, first);
        uint32_t b = __shfl_down_sync(ONES, first, 1);
        if (XCH && half + 1u < (uint32_t)XW && lane == 31u) b = xin;
        uint32_t carry;
        R::xy1(up[0], lw[0], up[1], lw[1], r0, r1, hasn ? b : R::NB_STONE, carry);
        if (XCH) xin = xmail(std::integral_constant<uint32_t, 1u>{}, carry);
        uint32_t a = __shfl_up_sync(ONES, carry, 1);
        if (XCH && half > 0u && lane == 0u) a = xin;
        uint32_t enw = 0;
        const uint32_t wall = R::wall_first(first, enw);
        if (!hasp) a = wall;
        R::post1(up[0], lw[0], up[1], lw[1], a);
    };
    auto zy = [&](Cell (&up)[2], Cell (&lw)[2], uint32_t yu, uint32_t key, uint32_t hzy) {
        R::zy(up[0], up[1], lw[0], lw[1], hash_word(key + hzy + yu * HC2));
    };

    // A-pairs that exist: -1 and nA are the neighbours' edge pairs, read from the ghost planes
    const int a_first = (NBR && p.has_lo) ? -1 : 0, a_last = (int)p.nA - 1 + ((NBR && p.has_hi) ? 1 : 0);

    for (int sp = 0; sp < 2; ++sp) {
    uint64_t pos = sp == 0 ? lo0 : lo1;
    const uint64_t end = sp == 0 ? hi0 : hi1;
    while (pos < end) {
        uint32_t band = (uint32_t)(pos / p.nit);
        uint32_t it_a = (uint32_t)(pos - (uint64_t)band * p.nit);
        // with neighbours the two edge bands come LAST: they wait for the neighbours' deliveries of the previous pass
        if (NBR && p.nbands >= 3u) band = band < p.nbands - 2u ? band + 1u : (band == p.nbands - 2u ? 0u : p.nbands - 1u);
        const uint64_t left = end - pos;
        uint32_t it_b = (left < (uint64_t)(p.nit - it_a)) ? it_a + (uint32_t)left : p.nit;
        pos += it_b - it_a;
        if (grp) {
            band = band * UNITS + pic;                 // `band` was the group
            if (band >= p.nbands) continue;
            const uint32_t sh = pic * (uint32_t)S4_K;
            auto shifted = [&](uint32_t s) { return s == 0u ? 0u : (s >= p.nit ? p.nit : (s > sh ? s - sh : 0u)); };
            it_a = shifted(it_a); it_b = shifted(it_b);
            if (it_a >= it_b) continue;
        }

        band += p.band0;
        const int qa = (int)(band * S4_P);
        const int nq = min(S4_P, (int)p.nB - qa);                 // B-pairs of this band; its A-pairs are a = qa-1 .. qa-1+nq
        if (NBR) {
            // bands that read ghost planes wait (bounded) until the neighbour delivered them after its previous pass
            if (p.has_lo && qa == 0) wait_arrival(p.my_flags + 0, p.wait_target, p.push_err, p.push_timeout_ns, 0u);
            if (p.has_hi && qa - 1 + nq >= (int)p.nA) wait_arrival(p.my_flags + 1, p.wait_target, p.push_err, p.push_timeout_ns, 1u);
            __syncwarp();
        }
        const uint32_t warm = it_a < S4_LEAD ? it_a : S4_LEAD;
        const uint32_t it0 = it_a - warm;
        // fresh pipelines: STONE below the segment's first plane (exact at the floor, rebuilt by the warm-up elsewhere)
        for (int k = 0; k <= nq; ++k)
            for (int w = 0; w < 12; ++w) sA(k, w) = ONES;
        for (int k = 1; k <= nq; ++k)
            for (int w = 0; w < 16; ++w) sB(k, w) = ONES;

        // plane pair `itt` of A-pair kk of this band: four unconditional 256-bit loads (clamped addresses; what lies
        // outside the grid is replaced by STONE when the words are packed)
        Raw<1> raw;
        auto issue = [&](int kk, uint32_t itt) {
            const int a = qa - 1 + kk;
            const bool ok = a >= a_first && a <= a_last;
            // the pair the neighbouring band of this CTA loads too (grp)
            const bool again = (kk == 0 && pic != 0u) || (kk == nq && pic + 1u != UNITS);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                // local plane 1 + 2a + r; plane -1 (global z0 - 2) is the far ghost kept at local plane nzl + 3
                const int pl = 1 + 2 * a + r;
                const uint8_t *base = p.src + (ok ? (size_t)(pl < 0 ? (int)p.nzl + 3 : pl) * plane_bytes + (size_t)xw * 32u : (size_t)0);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t y = 2u * itt + h;
                    const uint8_t *ad = base + (size_t)(y < ylast ? y : ylast) * row_bytes;
                    if (NBR) ld256_coherent(ad, raw.w[0][r][h]);     // ghost planes are written by a peer GPU
                    else if (!grp) ld256(ad, raw.w[0][r][h]);
                    else if (again) ld256_keep(ad, raw.w[0][r][h]);
                    else ld256_once(ad, raw.w[0][r][h]);
                }
            }
        };
        issue(0, it0);

        for (uint32_t ib = it0; ib < it_b; ib += S4_K) {
            const uint32_t ie = ib + S4_K < it_b ? ib + S4_K : it_b;
            for (int k = 0; k <= nq; ++k) {
                const int a = qa - 1 + k, q = qa + k - 1;          // stage-A pair loaded now; stage-B pair finished now (k >= 1)
                const bool a_ok = a >= a_first && a <= a_last;
                // hash coordinates are GLOBAL: local plane pl is global z0 + pl - 1
                const uint32_t hxyA[2] = {xw * HC1 + (p.z0 + (uint32_t)(2 * a)) * HC3, xw * HC1 + (p.z0 + (uint32_t)(2 * a + 1)) * HC3};
                const uint32_t hzyA = hxyA[0];
                const uint32_t hxyB[2] = {xw * HC1 + (p.z0 + (uint32_t)(2 * q - 1)) * HC3, xw * HC1 + (p.z0 + (uint32_t)(2 * q)) * HC3};
                const uint32_t hzyB = hxyB[0];
                const bool own[2] = {k >= 1 && 2 * q >= 1 && 2 * q <= (int)p.nzl, k >= 1 && 2 * q + 1 <= (int)p.nzl};
                uint8_t *const drow0 = p.dst + (size_t)(k >= 1 ? 2 * q : 0) * plane_bytes + (size_t)xw * 32u;

                Cell prev1[2], c2[2], c3[2];                       // stage A's carried planes (rows r0, r1 of pair a)
                Cell prev1B[2], c2B[2], c3B[2], bprev[2];          // stage B's (rows: pair a-1's r1, pair a's r0) + the even plane in waiting
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    prev1[r] = {sA(k, 6 * r + 0), sA(k, 6 * r + 1)}; c2[r] = {sA(k, 6 * r + 2), sA(k, 6 * r + 3)}; c3[r] = {sA(k, 6 * r + 4), sA(k, 6 * r + 5)};
                }
                if (k >= 1) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        prev1B[r] = {sB(k, 8 * r + 0), sB(k, 8 * r + 1)}; c2B[r] = {sB(k, 8 * r + 2), sB(k, 8 * r + 3)};
                        c3B[r] = {sB(k, 8 * r + 4), sB(k, 8 * r + 5)}; bprev[r] = {sB(k, 8 * r + 6), sB(k, 8 * r + 7)};
                    }
                }

                for (uint32_t it = ib; it < ie; ++it) {
                    const uint32_t y1 = 2u * it;
                    Cell lo[2], hi[2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        lo[r] = R::pack(raw.w[0][r][0]);
                        hi[r] = R::pack(raw.w[0][r][1]);
                        if (!(a_ok && y1 < p.ny))      lo[r] = R::stone();
                        if (!(a_ok && y1 + 1u < p.ny)) hi[r] = R::stone();
                    }
                    // the next plane pair's loads are in flight while this one is evaluated: next iteration of this pair,
                    // else the next pair's first, else the next y-block's first pair
#if FS3D_S4_ONE_ISSUE
                    {
                        const bool more = it + 1u < ie, nextk = k < nq;
                        if (more || nextk || ie < it_b) issue(more ? k : (nextk ? k + 1 : 0), more ? it + 1u : (nextk ? ib : ie));
                    }
#else
                    if (it + 1u < ie) issue(k, it + 1u);
                    else if (k < nq) issue(k + 1, ib);
                    else if (ie < it_b) issue(0, ie);
#endif

                    // ---- stage A: steps t (XY then ZY) and t+1 (ZY then XY), x-offset 0
                    xy0(hi, lo, y1 + 1u, p.key_xy[0], hxyA);
                    zy(lo, prev1, y1, p.key_zy[0], hzyA);
                    zy(prev1, c2, y1 - 1u, p.key_zy[1], hzyA);
                    xy0(c2, c3, y1 - 2u, p.key_xy[1], hxyA);
                    // planes y1-3 (c3) and y1-2 (c2) are final after step t+1.  Hand this pair's upper row to the next pair
                    // through the buffer, after taking the previous pair's upper row out of the same slot.
                    const uint32_t slot = it - ib;
                    Cell L3 = {sH(slot, 0), sH(slot, 1)}, L2 = {sH(slot, 2), sH(slot, 3)};
                    sH(slot, 0) = c3[1].p0; sH(slot, 1) = c3[1].p1; sH(slot, 2) = c2[1].p0; sH(slot, 3) = c2[1].p1;

                    // a segment that starts in mid-grid warms up for 7 iterations; stage A's planes are not right before the
                    // fourth, so stage B only idles through the first three (its pipeline converges from any state in 3)
                    if (k >= 1 && warm == S4_LEAD && it < it0 + 3u) {
                        bprev[0] = L2; bprev[1] = c2[0];
                    } else if (k >= 1) {
                        // ---- stage B: steps t+2 (XY then ZY) and t+3 (ZY then XY), x-offset 1, on rows (pair a-1's r1, pair a's r0)
                        // its plane pair is (y1-4, y1-3): the even plane waited one iteration in `bprev`
                        Cell loB[2] = {bprev[0], bprev[1]}, hiB[2] = {L3, c3[0]};
                        const uint32_t yb1 = y1 - 4u;
                        xy1(hiB, loB, yb1 + 1u, p.key_xy[2], hxyB);
                        zy(loB, prev1B, yb1, p.key_zy[2], hzyB);
                        zy(prev1B, c2B, yb1 - 1u, p.key_zy[3], hzyB);
                        xy1(c2B, c3B, yb1 - 2u, p.key_xy[3], hxyB);
                        if (it >= it_a) {
                            // planes y1-7 (c3B) and y1-6 (c2B) are final after step t+3
                            const uint32_t ya = yb1 - 3u, yb = yb1 - 2u;          // wrap to huge when negative
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                if (!own[r]) continue;
                                uint32_t o[8];
                                uint8_t *d = drow0 + (size_t)r * plane_bytes;
                                if (ya < p.ny) { R::unpack(c3B[r], o); st256(d + (size_t)ya * row_bytes, o); }
                                if (yb < p.ny) { R::unpack(c2B[r], o); st256(d + (size_t)yb * row_bytes, o); }
                            }
                        }
#pragma unroll
                        for (int r = 0; r < 2; ++r) { c3B[r] = prev1B[r]; c2B[r] = loB[r]; prev1B[r] = hiB[r]; }
                        bprev[0] = L2; bprev[1] = c2[0];
                    }
#pragma unroll
                    for (int r = 0; r < 2; ++r) { c3[r] = prev1[r]; c2[r] = lo[r]; prev1[r] = hi[r]; }
                }

#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    sA(k, 6 * r + 0) = prev1[r].p0; sA(k, 6 * r + 1) = prev1[r].p1; sA(k, 6 * r + 2) = c2[r].p0; sA(k, 6 * r + 3) = c2[r].p1;
                    sA(k, 6 * r + 4) = c3[r].p0; sA(k, 6 * r + 5) = c3[r].p1;
                }
                if (k >= 1) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        sB(k, 8 * r + 0) = prev1B[r].p0; sB(k, 8 * r + 1) = prev1B[r].p1; sB(k, 8 * r + 2) = c2B[r].p0; sB(k, 8 * r + 3) = c2B[r].p1;
                        sB(k, 8 * r + 4) = c3B[r].p0; sB(k, 8 * r + 5) = c3B[r].p1; sB(k, 8 * r + 6) = bprev[r].p0; sB(k, 8 * r + 7) = bprev[r].p1;
                    }
                }
            }
        }
    }
    }
}

}  // namespace fs3d
