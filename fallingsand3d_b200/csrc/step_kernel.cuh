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
//  * XY blocks pair x with x±1: inside a word that is a byte permute (PRMT) — the two rows' left and
//    right cells are gathered into separate words so each block is evaluated once; with odd
//    x-offset the cells of the word-straddling block cross lanes by warp shuffle.
//  * Loads for the next plane pair are issued before the current pair is evaluated
//    (register double-buffering), so ~J·4 KiB per warp is always in flight.  The loads are
//    unconditional (clamped addresses): predicated ones made nvcc wait for the data right
//    behind the LDG, which switched the double-buffering off.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "bitslice.cuh"
#include "bitslice3.cuh"

namespace fs3d {

// ---- rule sets: what a kernel instantiation needs to know about the cells it marches -------------------------------
// Rules1 = schedule version 1 (four materials, two bit-planes, SCHEDULE.md §1-5); Rules3 = schedule version 2 (eight
// materials, three rank-encoded bit-planes, SCHEDULE.md §7).  Everything else in the step kernel — the march, the
// loads and stores, the halo push, the settled-tile bookkeeping, the word-boundary exchanges — is shared.
struct Rules1 {
    using Cell = P2;
    static constexpr bool COLUMN_FORM = true;          // the per-column XY form (xy_substep) exists for two planes only
    static constexpr uint32_t NB_STONE = NB_STONE2;     // packed voxel-0 bits of a word beyond the wall
    static constexpr bool EDGE_TOLERANT = true;         // consumers of the packed edge bits ignore every other bit
    static __device__ __forceinline__ Cell stone() { return {ONES, ONES}; }
    static __device__ __forceinline__ Cell pack(const uint32_t (&w)[8]) { return fs3d::pack(w); }
    static __device__ __forceinline__ void unpack(Cell c, uint32_t (&w)[8]) { fs3d::unpack(c, w); }
    static __device__ __forceinline__ uint32_t zy(Cell &a, Cell &b, Cell &c, Cell &d, uint32_t rw) { return block_rule(a, b, c, d, rw); }
    static __device__ __forceinline__ uint32_t xy0(Cell &U0, Cell &L0, Cell &U1, Cell &L1, uint32_t r0, uint32_t r1) { return xy_pair_substep0(U0, L0, U1, L1, r0, r1); }
    static __device__ __forceinline__ uint32_t first_bits(Cell U0, Cell L0, Cell U1, Cell L1) { return xy_first_bits(U0, L0, U1, L1); }
    static __device__ __forceinline__ uint32_t xy1(Cell &U0, Cell &L0, Cell &U1, Cell &L1, uint32_t r0, uint32_t r1, uint32_t nb, uint32_t &carry) {
        return xy_pair_substep1(U0, L0, U1, L1, r0, r1, nb, carry);
    }
    static __device__ __forceinline__ void post1(Cell &U0, Cell &L0, Cell &U1, Cell &L1, uint32_t pb) { xy_pair_post1(U0, L0, U1, L1, pb); }
    static __device__ __forceinline__ uint32_t wall_first(uint32_t first, uint32_t &en) { return xy_wall_first(first, en); }
};
struct Rules3 {
    using Cell = P3;
    static constexpr bool COLUMN_FORM = false;
    static constexpr uint32_t NB_STONE = NB_STONE3;
    static constexpr bool EDGE_TOLERANT = true;
    static __device__ __forceinline__ Cell stone() { return {ONES, ONES, ONES}; }
    static __device__ __forceinline__ Cell pack(const uint32_t (&w)[8]) { return pack3(w); }
    static __device__ __forceinline__ void unpack(Cell c, uint32_t (&w)[8]) { unpack3(c, w); }
    static __device__ __forceinline__ uint32_t zy(Cell &a, Cell &b, Cell &c, Cell &d, uint32_t rw) { return block_rule3(a, b, c, d, rw, coin2_word(rw)); }
    static __device__ __forceinline__ uint32_t xy0(Cell &U0, Cell &L0, Cell &U1, Cell &L1, uint32_t r0, uint32_t r1) { return xy3_pair_substep0(U0, L0, U1, L1, r0, r1); }
    static __device__ __forceinline__ uint32_t first_bits(Cell U0, Cell L0, Cell U1, Cell L1) { return xy3_first_bits(U0, L0, U1, L1); }
    static __device__ __forceinline__ uint32_t xy1(Cell &U0, Cell &L0, Cell &U1, Cell &L1, uint32_t r0, uint32_t r1, uint32_t nb, uint32_t &carry) {
        return xy3_pair_substep1(U0, L0, U1, L1, r0, r1, nb, carry);
    }
    static __device__ __forceinline__ void post1(Cell &U0, Cell &L0, Cell &U1, Cell &L1, uint32_t pb) { xy3_pair_post1(U0, L0, U1, L1, pb); }
    static __device__ __forceinline__ uint32_t wall_first(uint32_t first, uint32_t &en) { return xy3_wall_first(first, en); }
};

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
    uint32_t nit;            // march iterations per pair: ny / 2 + NS
    uint32_t key_xy, key_zy; // SCHEDULE.md §3 key(seed, t, axis)
    uint32_t key_xy2, key_zy2; // keys of step t + 1 (NS = 2 only)
    // settled-tile skipping (SKIP = 1 instantiations only)
    uint32_t *last_active;   // [ztiles][ytiles] (step + 1) of the last enabled block seen in the tile
    uint32_t ytile_log2;     // y-tile height = 1 << ytile_log2 planes (>= 2)
    uint32_t ztile_log2;     // z-tile depth = 1 << ztile_log2 owned planes
    uint32_t nytiles;
    uint32_t step_plus1;     // (uint32)(t + NS)
    const uint32_t *runs;    // [n][3] = (pair group, it_a, it_b): the live march segments of this launch, built by skip_plan_kernel
    const uint32_t *nruns;
    // fused halo push over peer memory (PUSH = 1 instantiations only): the warps that compute this
    // slab's first / last owned plane also store it into the z-neighbour's ghost plane (its dst
    // buffer, mapped through CUDA IPC / NVLink) and then add the iterations they finished to the
    // neighbour's arrival counter; warps that touch a ghost or edge plane first wait until their own
    // counters show that the neighbours delivered (and stopped reading) the previous pass.
    uint8_t *peer_lo_dst;                 // neighbour below: its ghost-HIGH plane (dst buffer) or nullptr
    uint8_t *peer_hi_dst;                 // neighbour above: its ghost-LOW plane (dst buffer) or nullptr
    unsigned long long *peer_lo_flag;     // neighbour below: its arrive[1]
    unsigned long long *peer_hi_flag;     // neighbour above: its arrive[0]
    const unsigned long long *my_flags;   // arrive[0] (from below), arrive[1] (from above)
    unsigned long long wait_target;       // cumulative iterations the neighbours must have delivered
    // watchdog of that wait: a neighbour that died, missed a pass or was stepped a different number of times would
    // otherwise hang this kernel (and every later fs3d_sync / fs3d_destroy) forever.  A warp that has waited
    // push_timeout_ns sets push_err (bit 0: the neighbour below, bit 1: above; bits 8..: the pass it waited for) and
    // goes on; every later wait of this world falls through at once, and the host turns the word into FS3D_ERR_CUDA.
    unsigned long long *push_err;
    unsigned long long push_timeout_ns;
};

// cache-policy qualifiers of the streaming accesses (tuning hooks; defaults measured best)
#ifndef FS3D_LD_POLICY
#define FS3D_LD_POLICY ".L1::no_allocate"
#endif
#ifndef FS3D_ST_POLICY
#define FS3D_ST_POLICY ".L1::no_allocate"
#endif
__device__ __forceinline__ void ld256(const uint8_t *p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc" FS3D_LD_POLICY ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
// same with an L2 eviction priority: ".L2::evict_last" for rows a second warp of the CTA reads again shortly (the pair two
// neighbouring bands of the four-step kernel share), ".L2::evict_first" for rows nobody reads again
#define FS3D_LD256_PRIO(NAME, PRIO)                                                                                  \
    __device__ __forceinline__ void NAME(const uint8_t *p, uint32_t (&r)[8]) {                                       \
        asm volatile("ld.global.nc.L1::no_allocate" PRIO ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                  \
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
                     : "l"(p));                                                                                      \
    }
FS3D_LD256_PRIO(ld256_keep, ".L2::evict_last")
FS3D_LD256_PRIO(ld256_once, ".L2::evict_first")
// same, through the coherent path: ghost planes are written by a peer GPU while this kernel runs.
// (asm volatile statements keep their program order, so these stay behind the acquire of the arrival flag.)
__device__ __forceinline__ void ld256_coherent(const uint8_t *p, uint32_t (&r)[8]) {
    asm volatile("ld.global" FS3D_LD_POLICY ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// programmatic dependent launch (PDL): used by the settled-tile path, where a pass is two short kernels
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait for a z-neighbour's arrival counter (PUSH kernels).  side: 0 = the neighbour below, 1 = above.
// err[0] = watchdog word; err[1..3] = statistics of the waits that really blocked (ns summed over warps, longest
// single wait, number of waits) — what inter-GPU skew costs, read back through fs3d_push_wait_stats.
__device__ __forceinline__ void wait_arrival(const unsigned long long *flag, unsigned long long target,
                                             unsigned long long *err, unsigned long long timeout_ns, unsigned side) {
    if (ld_acquire_sys(flag) >= target) return;
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    bool gave_up = false;
    while (ld_acquire_sys(flag) < target) {
        __nanosleep(64);
        if ((++spins & 63u) == 0u) {
            if (ld_relaxed_u64(err) != 0ull) { gave_up = true; break; }   // this world already failed: do not wait again
            if (global_ns() - t0 > timeout_ns) {
                atomicOr(err, (1ull << side) | (target << 8));
                gave_up = true;
                break;
            }
        }
    }
    if ((threadIdx.x & 31u) == 0u && !gave_up) {
        const unsigned long long dt = global_ns() - t0;
        atomicAdd(err + 1, dt);
        atomicMax(err + 2, dt);
        atomicAdd(err + 3, 1ull);
    }
}
__device__ __forceinline__ void st256(uint8_t *p, const uint32_t (&r)[8]) {
    asm volatile("st.global" FS3D_ST_POLICY ".v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// ---- experiment: bulk-async (TMA engine) staging of the row segments in a per-warp shared-memory ring ----
// -DFS3D_STAGE_LOADS=1 replaces the LDG register double-buffering of the J = 2, non-PUSH kernels by
// cp.async.bulk copies (2 KiB per row-plane) completing on an mbarrier, FS3D_NSTAGE plane pairs deep.
#ifndef FS3D_STAGE_LOADS
#define FS3D_STAGE_LOADS 0
#endif
#ifndef FS3D_NSTAGE
#define FS3D_NSTAGE 2
#endif
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void lds256(uint32_t addr, uint32_t (&r)[8]) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr + 16u));
}
template <int J, int THREADS>
constexpr uint32_t stage_smem_bytes() { return (THREADS / 32) * FS3D_NSTAGE * (4u * J * 1024u) + (THREADS / 32) * FS3D_NSTAGE * 8u; }

template <int J>
struct Raw { uint32_t w[J][2][2][8]; };   // [word][row l/r][plane lo/hi][8 x u32]

// ---------------------------------------------------------------------------------------------
// NS = number of SCHEDULE.md steps fused into this pass (1 or 2).  Two steps can be fused without
// any halo because hoff = (t >> 1) & 1 is the same for t = 2k and 2k + 1: both steps use the same
// z-pairing and the same x-offset, so a z-pair of rows is closed under both; only the pipeline
// along y gets two plane pairs deeper.  NS = 2 requires t even (TODD = 0 for the first step).
//
// March, iteration `it` (y1 = 2·it), planes lo = y1 and hi = y1 + 1 freshly loaded:
//   step t   : sub-step 1 on (hi, lo), sub-step 2 on (lo, prev1)            -> planes y1-1, y1 done
//   step t+1 : sub-step 3 (ZY) on (prev1, c2), sub-step 4 (XY) on (c2, c3)  -> planes y1-3, y1-2 done
// carried to the next iteration: prev1 <- hi, c2 <- lo, c3 <- prev1.  A segment that starts above
// the floor first runs LEAD = 2·NS − 1 iterations without storing to rebuild the carried planes.
#ifndef FS3D_MINB
#define FS3D_MINB 1
#endif
//
// XW = warps per z-pair (1 or 2).  Rows wider than 64 words (nx > 2048) do not fit one warp's registers
// at J = 4 (255 registers + spills, measured 40 % slower), so two adjacent warps of a CTA share the
// z-pair instead: warp half h owns words [64h, 64h + 64).  The only thing they exchange is what a
// single warp moves by shuffle at the word boundary in XY sub-steps with odd x-offset.  It goes through
// 32-bit shared-memory mailboxes, tagged with a sequence number and polled by the consumer — no
// barrier, so the two warps only ever wait for the one value they need.
template <class R, int J, int XW, int OX, int TODD, int SKIP, int NS, int PUSH, int THREADS>
__global__ void __launch_bounds__(THREADS, FS3D_MINB) step_kernel(const StepParams p) {
    using Cell = typename R::Cell;
    static_assert(NS == 1 || (NS == 2 && TODD == 0), "a fused pair of steps starts on an even step");
    static_assert(XW == 1 || (XW == 2 && J > 1 && THREADS % 64 == 0), "warp pairs live in one CTA");
    constexpr bool XCH = XW == 2 && OX == 1;    // edge words cross the warp-pair boundary
    constexpr uint32_t LEAD = 2 * NS - 1;      // warm-up iterations to rebuild the carried planes
    constexpr uint32_t LAG = 2 * NS - 2;       // iteration `it` stores planes 2·it − LAG − 1 and 2·it − LAG
    const uint32_t lane = threadIdx.x & 31u;
    // The two warps of a pair are adjacent warps of the CTA (2k, 2k + 1).  Putting them on the same warp
    // scheduler instead (k, k + THREADS/64) measured 12 % slower (-DFS3D_EXP_PAIR_SPLIT=1).
#ifndef FS3D_EXP_PAIR_SPLIT
#define FS3D_EXP_PAIR_SPLIT 0
#endif
    constexpr uint32_t PAIRS = XW == 2 ? THREADS / 64 : THREADS / 32;   // work units per CTA
    const uint32_t wic = threadIdx.x >> 5;        // warp in CTA
    const uint32_t pic = FS3D_EXP_PAIR_SPLIT ? wic % PAIRS : wic / XW;   // work unit in CTA
    const uint32_t half = FS3D_EXP_PAIR_SPLIT ? wic / PAIRS : wic % XW;  // which 32·J-word half of the row this warp owns
    const uint32_t gw = blockIdx.x * PAIRS + pic; // work unit (z-pair marcher): one warp, or a warp pair
    const uint32_t nw = gridDim.x * PAIRS;
    __shared__ uint32_t xch_smem[XCH ? PAIRS * 4 : 1];   // [pair in CTA][mailboxes]
    volatile uint32_t *const xch = xch_smem + (XCH ? pic * 4 : 0);
    uint32_t xseq = 0;                             // exchanges done; both warps of a pair count alike
    if (XCH) {
        if (threadIdx.x < PAIRS * 4) xch_smem[threadIdx.x] = 0u;   // tag 0 is never expected first
        __syncthreads();
    }

    // staging ring (experiment): [warp][stage][row][plane][J KiB] then one mbarrier per warp and stage
    constexpr bool STG = FS3D_STAGE_LOADS && !PUSH && J == 2;
    constexpr uint32_t ROWB = J * 1024u, STAGE_BYTES = 4u * ROWB;
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    const uint32_t ring = STG ? smem_u32(dyn_smem) + wic * (FS3D_NSTAGE * STAGE_BYTES) : 0u;
    const uint32_t bars = STG ? smem_u32(dyn_smem) + (THREADS / 32) * (FS3D_NSTAGE * STAGE_BYTES) + wic * (FS3D_NSTAGE * 8u) : 0u;
    uint32_t kc = 0, ki = 0;                       // stages consumed / issued so far (ring position and phase)
    if (STG) {
        if (lane == 0) for (int st = 0; st < FS3D_NSTAGE; ++st) mbar_init(bars + 8u * st, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }

    uint32_t g = 0, xw0 = lane + half * (32u * J);
    bool lane_ok = true;
    if (J == 1) { g = lane / p.lpr; xw0 = lane - g * p.lpr; lane_ok = g < p.groups; }

    const uint32_t npairs = p.pair_end - p.pair_begin;
    const uint32_t npg = (npairs + p.groups - 1) / p.groups;
    const uint64_t total = (uint64_t)npg * p.nit;
    // Work split.  Without skipping every iteration costs the same, so the (pair-group x iteration)
    // space is cut into equal contiguous ranges, one per resident warp (no tail, one lead-in each).
    // With skipping only the live segments exist as work: skip_plan_kernel compacts them into a list of
    // runs (a few y-blocks of one pair group each) that is dealt round-robin to the warps.
    if (SKIP) { pdl_launch_dependents(); pdl_wait(); }      // the run list comes from the plan kernel just in front (PDL)
    const uint32_t nruns = SKIP ? *p.nruns : 0xFFFFFFFFu;
    uint32_t run = gw;
    uint64_t pos = total * gw / nw;
    const uint64_t end = total * (gw + 1) / nw;

    const size_t row_bytes = p.nx;
    const size_t plane_rows = p.ny;

    for (;;) {
        uint32_t pg, it_a, it_b;
        if (SKIP) {
            if (run >= nruns) break;
            pg = p.runs[3u * run]; it_a = p.runs[3u * run + 1u]; it_b = p.runs[3u * run + 2u];
            run += nw;
        } else {
            if (pos >= end) break;
            pg = (uint32_t)(pos / p.nit);
            it_a = (uint32_t)(pos - (uint64_t)pg * p.nit);
            const uint64_t left = end - pos;
            it_b = (left < (uint64_t)(p.nit - it_a)) ? it_a + (uint32_t)left : p.nit;
            pos += it_b - it_a;
        }

        // PUSH kernels visit the two slab-edge pairs LAST: their warps must wait for the neighbours'
        // previous pass, and at the end of the kernel that wait has a whole pass of slack, whereas at
        // the start it would put every bit of inter-GPU skew on the critical path.
        uint32_t pair = p.pair_begin + pg * p.groups + g;
        const bool pair_ok = lane_ok && pair < p.pair_end;
        if (PUSH && !SKIP && npairs >= 3u && pair_ok) {
            const uint32_t pl = pair - p.pair_begin;
            pair = p.pair_begin + (pl < npairs - 2u ? pl + 1u : (pl == npairs - 2u ? 0u : npairs - 1u));
        }
        const uint32_t lzl = p.lz_first + 2u * pair;          // local plane of the left row
        const int32_t zgl = p.z0 + (int32_t)lzl - 1;          // its global z
        bool own[2];
        own[0] = pair_ok && lzl >= 1u && lzl <= p.nzl;
        own[1] = pair_ok && lzl + 1u <= p.nzl;
        // pairs are enumerated so that lzl <= nzl, hence lzl + 1 <= nzl + 1 = ghost-high.

        bool wok[J];                 // this lane's word j exists
        uint32_t xw[J];
        bool hasp[J], hasn[J];
        uint32_t hxy[J][2], hzy[J];  // linear hash parts (without the key)
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
                srow[j][r] = p.src + (wok[j] ? off : (size_t)0);   // lanes without a word load (and drop) plane 0
                drow[j][r] = p.dst + off;
                hxy[j][r] = xw[j] * HC1 + (uint32_t)(zgl + r) * HC3;
            }
            hzy[j] = xw[j] * HC1 + (uint32_t)zgl * HC3;
        }

        // fused halo push: which of my two rows (if any) is the slab's first / last owned plane
        int rlo = -1, rhi = -1;
        ptrdiff_t dlo = 0, dhi = 0;    // peer ghost plane address = my dst row address + delta
        if (PUSH) {
            if (pair_ok && p.peer_lo_dst != nullptr) rlo = lzl == 1u ? 0 : (lzl == 0u ? 1 : -1);
            if (pair_ok && p.peer_hi_dst != nullptr) rhi = lzl == p.nzl ? 0 : (lzl + 1u == p.nzl ? 1 : -1);
            dlo = p.peer_lo_dst - (p.dst + (size_t)1u * plane_rows * row_bytes);
            dhi = p.peer_hi_dst - (p.dst + (size_t)p.nzl * plane_rows * row_bytes);
            // pairs that read a ghost plane, or write a neighbour's, wait for that neighbour's previous pass
            const bool near_lo = pair_ok && lzl <= 1u && p.peer_lo_flag != nullptr;
            const bool near_hi = pair_ok && lzl + 1u >= p.nzl && p.peer_hi_flag != nullptr;
            if (__any_sync(ONES, near_lo)) wait_arrival(p.my_flags + 0, p.wait_target, p.push_err, p.push_timeout_ns, 0u);
            if (__any_sync(ONES, near_hi)) wait_arrival(p.my_flags + 1, p.wait_target, p.push_err, p.push_timeout_ns, 1u);
            __syncwarp();
        }

        // Loads are UNCONDITIONAL (addresses clamped into the buffer) and out-of-grid planes are replaced by
        // STONE when the words are packed: a predicated load makes the compiler merge its result into
        // the raw registers right behind the LDG, which waits for the data there and silently turns the
        // register double-buffering off (seen in SASS; 3.34 ms instead of 2.71 ms per pass).
        // PUSH kernels read everything through the coherent path (ghost planes are written by a peer GPU
        // while this kernel runs), one code path and no branch per load.
        Raw<J> raw;
        const uint32_t ylast = p.ny - 1u;
        auto load_pair = [&](uint32_t it) {     // planes y1 = 2·it (lo) and y1 + 1 (hi)
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t y = 2u * it + h;
                        const uint8_t *a = srow[j][r] + (size_t)(y < ylast ? y : ylast) * row_bytes;
                        if (PUSH) ld256_coherent(a, raw.w[j][r][h]);
                        else ld256(a, raw.w[j][r][h]);
                    }
        };

        // staged variant: lane 0 asks the copy engine for this warp's four row-plane segments of plane pair `it`
        const uint32_t seg_w = (pair_ok && p.wpr > half * (32u * J)) ? (p.wpr - half * (32u * J) < 32u * J ? p.wpr - half * (32u * J) : 32u * J) : 0u;
        const uint32_t seg_bytes = seg_w * 32u;
        const uint8_t *seg_src = p.src + (size_t)lzl * plane_rows * row_bytes + (size_t)half * (32u * J) * 32u;
        auto issue_pair = [&](uint32_t it) {
            if (lane == 0) {
                const uint32_t st = ki % FS3D_NSTAGE, bar = bars + 8u * st;
                mbar_expect_tx(bar, 4u * seg_bytes);
                if (seg_bytes) {
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t y = 2u * it + h;
                            bulk_g2s(ring + st * STAGE_BYTES + (uint32_t)(r * 2 + h) * ROWB,
                                     seg_src + (size_t)r * plane_rows * row_bytes + (size_t)(y < ylast ? y : ylast) * row_bytes,
                                     seg_bytes, bar);
                        }
                }
            }
            ++ki;
        };
        auto consume_pair = [&]() {
            const uint32_t st = kc % FS3D_NSTAGE;
            mbar_wait(bars + 8u * st, (kc / FS3D_NSTAGE) & 1u);
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        lds256(ring + st * STAGE_BYTES + (uint32_t)(r * 2 + h) * ROWB + (lane + 32u * j) * 32u, raw.w[j][r][h]);
            ++kc;
        };

        Cell prev1[J][2], c2[J][2], c3[J][2], lo[J][2], hi[J][2];
        auto reset_carry = [&]() {
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) { prev1[j][r] = R::stone(); c2[j][r] = R::stone(); c3[j][r] = R::stone(); }
        };

        // Odd x-offset has two implementations of the XY sub-step (even offset always uses xy_pair_substep0):
        //   pair  (xy_pair_substep1): every block evaluated once, ~30 % fewer ALU instructions and 15-35 fewer
        //         registers, but two dependent shuffle round trips per sub-step (next word's cells in, new voxel 0 back);
        //   column (xy_substep): every block evaluated from both columns, one round trip.
        // Measured on one box (profiles/r01d_experiments_xy_pair.txt): the pair form wins where issue slots or
        // registers are the limit (single-step kernels 1 %, warp-pair PUSH kernels 8 %, sparse SKIP passes 10 %), the
        // column form where two warps per scheduler must hide the extra round trip (fused XW = 1 kernels, 5 %).
        // -DFS3D_XY_PAIR_OX1=0/1 forces one of them everywhere.
#ifdef FS3D_XY_PAIR_OX1
        constexpr bool PAIR1 = !R::COLUMN_FORM || FS3D_XY_PAIR_OX1 != 0;
#else
        constexpr bool PAIR1 = !R::COLUMN_FORM || NS == 1 || XW == 2 || SKIP == 1;
#endif
        // One-way message between the two warps of a pair (XW = 2): a tagged 32-bit mailbox in shared memory (tag in
        // bits 28-31, 28 payload bits).  Both warps count messages alike; messages alternate direction (pre: half
        // 1 -> 0, post: half 0 -> 1), so message k + 2 reuses the slot of message k only after its reader has
        // answered message k + 1.
        auto xmail = [&](uint32_t from_half, uint32_t from_lane, uint32_t payload) -> uint32_t {
            ++xseq;
            const uint32_t tag = xseq << 28;              // strict alternation: four tag bits are plenty
            volatile uint32_t *slot = xch + (xseq & 1u);
            if (half == from_half) {
                if (lane == from_lane) *slot = (payload & 0x0FFFFFFFu) | tag;
                return 0u;
            }
            uint32_t v;
            do { v = *slot; } while ((v & 0xF0000000u) != tag);
            return R::EDGE_TOLERANT ? v : (v & 0x0FFFFFFFu);
        };
        // XY sub-step on (upper, lower) of both rows at once, every block evaluated once (bitslice.cuh,
        // xy_pair_substep*); the upper row is plane yu
        auto do_xy = [&](Cell (&up)[J][2], Cell (&lw)[J][2], uint32_t yu, uint32_t key) {
            uint32_t en = 0;
            uint32_t rw[J][2];
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) rw[j][r] = hash_word(key + hxy[j][r] + yu * HC2);
            if constexpr (OX == 0) {
#pragma unroll
                for (int j = 0; j < J; ++j) en |= R::xy0(up[j][0], lw[j][0], up[j][1], lw[j][1], rw[j][0], rw[j][1]);
            } else if constexpr (!PAIR1) {
                // per-column evaluation (xy_substep): one exchange of edge words, every block computed from both columns
                uint32_t e[J][2];
                uint32_t xe[2] = {EDGE_STONE, EDGE_STONE};
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int j = 0; j < J; ++j) e[j][r] = edge_pack(up[j][r], lw[j][r], rw[j][r]);
                if (XCH) {
                    ++xseq;
                    const uint32_t tag = (xseq & 0x3FFFu) << 18;
                    volatile uint32_t *slot = xch + (xseq & 1u) * 2u;
                    const uint32_t mine = half == 0u ? (e[J - 1][0] | (e[J - 1][1] << 9)) : (e[0][0] | (e[0][1] << 9));
                    if (lane == (half == 0u ? 31u : 0u)) slot[half] = tag | mine;
                    uint32_t v;
                    do { v = slot[half ^ 1u]; } while ((v & 0xFFFC0000u) != tag);
                    xe[0] = v & 0x1FFu;
                    xe[1] = (v >> 9) & 0x1FFu;
                }
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        uint32_t a = __shfl_up_sync(ONES, e[j][r], 1);
                        uint32_t b = __shfl_down_sync(ONES, e[j][r], 1);
                        if (J > 1) {
                            if (j > 0)     { uint32_t t = __shfl_sync(ONES, e[j - 1][r], 31); if (lane == 0)  a = t; }
                            if (j < J - 1) { uint32_t t = __shfl_sync(ONES, e[j + 1][r], 0);  if (lane == 31) b = t; }
                        }
                        if (XCH && j == 0     && half == 1u && lane == 0u)  a = xe[r];
                        if (XCH && j == J - 1 && half == 0u && lane == 31u) b = xe[r];
                        en |= xy_substep<1>(up[j][r], lw[j][r], rw[j][r], hasp[j] ? a : EDGE_STONE, hasn[j] ? b : EDGE_STONE);
                    }
            } else {
                // blocks (4k+3, 4k+4) straddle words: a lane evaluates the blocks whose LEFT cell it owns.
                // pre: the next word's voxel-0 bits (shuffle down); post: the new voxel 0 from the previous word's
                // evaluation (shuffle up) or, at the grid wall, the fall-only rule
                uint32_t first[J], carry[J];
#pragma unroll
                for (int j = 0; j < J; ++j) first[j] = R::first_bits(up[j][0], lw[j][0], up[j][1], lw[j][1]);
                uint32_t xin = R::NB_STONE;
                if (XCH) xin = xmail(1u, 0u, first[0]);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    uint32_t b = __shfl_down_sync(ONES, first[j], 1);
                    if (J > 1 && j < J - 1) { uint32_t t = __shfl_sync(ONES, first[j + 1], 0); if (lane == 31) b = t; }
                    if (XCH && j == J - 1 && half == 0u && lane == 31u) b = xin;
                    en |= R::xy1(up[j][0], lw[j][0], up[j][1], lw[j][1], rw[j][0], rw[j][1], hasn[j] ? b : R::NB_STONE, carry[j]);
                }
                if (XCH) xin = xmail(0u, 31u, carry[J - 1]);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    uint32_t a = __shfl_up_sync(ONES, carry[j], 1);
                    if (J > 1 && j > 0) { uint32_t t = __shfl_sync(ONES, carry[j - 1], 31); if (lane == 0) a = t; }
                    if (XCH && j == 0 && half == 1u && lane == 0u) a = xin;
                    if (j == 0) {           // only a row's first word can sit at the wall
                        uint32_t enw = 0;
                        const uint32_t wall = R::wall_first(first[0], enw);
                        if (!hasp[0]) { a = wall; en |= enw; }
                    }
                    R::post1(up[j][0], lw[j][0], up[j][1], lw[j][1], a);
                }
            }
            return en;
        };
        // ZY sub-step across the two rows, upper row is plane yu
        auto do_zy = [&](Cell (&up)[J][2], Cell (&lw)[J][2], uint32_t yu, uint32_t key) {
            uint32_t en = 0;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                uint32_t rw = hash_word(key + hzy[j] + yu * HC2);
                en |= R::zy(up[j][0], up[j][1], lw[j][0], lw[j][1], rw);
            }
            return en;
        };

        // settled-tile bookkeeping (SCHEDULE.md §4): y is cut into blocks of BLK iterations; block b
        // stores planes 2·BLK·b − LAG − 1 … 2·BLK·(b+1) − LAG − 2, i.e. y-tile b plus the top LAG + 1
        // planes of tile b − 1
        const uint32_t blk_log2 = SKIP ? p.ytile_log2 - 1u : 31u;
        const uint32_t blk_mask = (1u << blk_log2) - 1u;
        const int32_t ozl = (int32_t)lzl - 1, ozr = (int32_t)lzl;      // owned-plane indices of the two rows
        auto mark = [&](int32_t oz, uint32_t yt) {
            if (oz >= 0 && oz < (int32_t)p.nzl && yt < p.nytiles)
                p.last_active[(size_t)((uint32_t)oz >> p.ztile_log2) * p.nytiles + yt] = p.step_plus1;
        };
        uint32_t en_main = 0, en_strad = 0;
        auto flush_marks = [&](uint32_t bi) {
            if (en_main | en_strad) { mark(ozl, bi); mark(ozr, bi); }
            if (en_strad && bi > 0) { mark(ozl, bi - 1); mark(ozr, bi - 1); }
            en_main = 0; en_strad = 0;
        };

        // `it` runs over [it_a, it_b); the LEAD iterations below it_a are replayed first without storing
        // (`warm` counts them down).  SKIP runs start and end on y-block boundaries and hold live blocks only.
        bool loaded = false;
        uint32_t it = it_a;
        uint32_t warm = 0;             // > 0: this iteration only rebuilds carried planes
        bool need_restart = true;
        while (it < it_b) {
            if (need_restart) {
                reset_carry();                      // STONE below the floor; harmless garbage otherwise
                warm = it < LEAD ? it : LEAD;
                it -= warm;
                need_restart = false; loaded = false;
                if (STG) for (uint32_t a = it; a < it_b && a < it + FS3D_NSTAGE; ++a) issue_pair(a);
            }
            if (STG) consume_pair();
            else if (!loaded) load_pair(it);

            const uint32_t y1 = 2u * it;
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    lo[j][r] = R::pack(raw.w[j][r][0]);
                    hi[j][r] = R::pack(raw.w[j][r][1]);
                    if (!(wok[j] && y1 < p.ny))      lo[j][r] = R::stone();   // STONE outside the grid
                    if (!(wok[j] && y1 + 1u < p.ny)) hi[j][r] = R::stone();
                }

            // the next iteration's loads are in flight while this one is evaluated
            const uint32_t nxt = it + 1;
            if (STG) {
                // the words are in registers: hand the stage back to the copy engine for plane pair it + NSTAGE
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (it + FS3D_NSTAGE < it_b) issue_pair(it + FS3D_NSTAGE);
            } else {
                loaded = nxt < it_b;
                if (loaded) load_pair(nxt);
            }

            uint32_t e1, e2, e3 = 0, e4 = 0;
            if (TODD == 0) { e1 = do_xy(hi, lo, y1 + 1, p.key_xy); e2 = do_zy(lo, prev1, y1, p.key_zy); }
            else           { e1 = do_zy(hi, lo, y1 + 1, p.key_zy); e2 = do_xy(lo, prev1, y1, p.key_xy); }
            if (NS == 2) {   // step t + 1 (odd): ZY with oy = 0 on (y1-1, y1-2), then XY with oy = 1 on (y1-2, y1-3)
                e3 = do_zy(prev1, c2, y1 - 1u, p.key_zy2);
                e4 = do_xy(c2, c3, y1 - 2u, p.key_xy2);
            }
            if (SKIP && warm == 0) {
                // enabled blocks seen while storing block b: those of the first LAG/2 + 1 iterations
                // may lie in tile b − 1
                if ((it & blk_mask) <= (LAG >> 1)) en_strad |= e1 | e2 | e3 | e4; else en_main |= e1 | e2 | e3 | e4;
            }

            if (warm == 0) {
                // the two oldest planes are final: NS = 1: (prev1, lo) = y1-1, y1; NS = 2: (c3, c2) = y1-3, y1-2
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (wok[j] && own[r]) {
                            uint32_t o[8];
                            const uint32_t ya = y1 - LAG - 1u, yb = y1 - LAG;    // wrap to huge when negative
                            if (ya < p.ny) {
                                R::unpack(NS == 2 ? c3[j][r] : prev1[j][r], o);
                                st256(drow[j][r] + (size_t)ya * row_bytes, o);
                                if (PUSH && r == rlo) st256(drow[j][r] + (size_t)ya * row_bytes + dlo, o);
                                if (PUSH && r == rhi) st256(drow[j][r] + (size_t)ya * row_bytes + dhi, o);
                            }
                            if (yb < p.ny) {
                                R::unpack(NS == 2 ? c2[j][r] : lo[j][r], o);
                                st256(drow[j][r] + (size_t)yb * row_bytes, o);
                                if (PUSH && r == rlo) st256(drow[j][r] + (size_t)yb * row_bytes + dlo, o);
                                if (PUSH && r == rhi) st256(drow[j][r] + (size_t)yb * row_bytes + dhi, o);
                            }
                        }
                    }
            }
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (NS == 2) { c3[j][r] = prev1[j][r]; c2[j][r] = lo[j][r]; }
                    prev1[j][r] = hi[j][r];
                }

            if (warm > 0) {
                --warm;
            } else if (SKIP && ((nxt & blk_mask) == 0u || nxt >= it_b)) {
                flush_marks(it >> blk_log2);
            }
            it = nxt;
        }

        if (PUSH) {
            // publish: my stores into the neighbour's ghost plane happen-before the counter update
            const bool push_lo = rlo >= 0, push_hi = rhi >= 0;
            if (push_lo || push_hi) __threadfence_system();
            __syncwarp();
            const bool leader = pair_ok && xw0 == half * (32u * J);   // one lane per z-pair and warp (XW adds per pair)
            if (leader && push_lo) atomicAdd_system(p.peer_lo_flag, (unsigned long long)(it_b - it_a));
            if (leader && push_hi) atomicAdd_system(p.peer_hi_flag, (unsigned long long)(it_b - it_a));
        }
    }
}

// ---- settled-tile plan: ONE small kernel per launch of a SKIP step kernel --------------------------------------
// A tile is *quiet* at step t_now iff it and its 8 neighbours in (z-tile, y-tile) space saw no enabled block during
// the last four steps (all four offset phases), so nothing in or around it can change now (SCHEDULE.md §4).  Edge
// z-tiles next to another slab are never quiet (the neighbour's activity is not visible here).
//
// The plan kernel turns the tiles' last_active stamps straight into the list of live march segments the SKIP step
// kernel deals out to its warps (round 1 used a skip-map kernel, a run-list kernel and two memsets per pass):
// one WARP per pair group looks at its y-tiles 32 at a time (live unless quiet for both rows) and lane 0 emits the
// runs.  y-block b (BLK iterations) stores tile b's planes except its top LAG + 1, which the first NS iterations of
// block b + 1 store; so a maximal range of live tiles [Ta, Tb) needs iterations [BLK·Ta, BLK·Tb + NS) — whole blocks
// plus a short tail.  Ranges are chopped into pieces of `chop` blocks so that every marching warp gets several runs.
// Order in the list is arbitrary: runs write disjoint planes (what a run stores into neighbouring static tiles is
// their unchanged content).
//
// No memset, no second kernel: the counters rotate.  stats[3][4] = (tiles live, tiles total, live ranges, live
// iterations) of the pass before (read: piece length), of this pass (accumulated) and of the next (zeroed here);
// nruns[2] likewise per launch.
struct PlanParams {
    const uint32_t *last_active;
    uint32_t nztiles, nytiles, ztile_log2, blk_log2;
    uint32_t nzl, lz_first, pair_begin, pair_end, groups, nit, ns;
    uint32_t nw;                          // marching work units of the step kernel that follows
    uint32_t t_now;
    int has_lo_neighbour, has_hi_neighbour;
    int force_live;                       // the grid just came from the host: nothing is known to be static
    const unsigned long long *stats_prev;
    unsigned long long *stats_cur, *stats_next;
    uint32_t *runs, *nruns, *nruns_next;
};

static __global__ void skip_plan_kernel(const PlanParams q) {   // static: step_kernel.cuh is included by two translation units
    // Programmatic dependent launch (both no-ops in an ordinary launch): let the step kernel behind this one be
    // scheduled already, and do not read what the step kernel in front of this one wrote before it has completed.
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t npairs = q.pair_end - q.pair_begin;
    const uint32_t npg = (npairs + q.groups - 1) / q.groups;
    const uint32_t blk = 1u << q.blk_log2;
    const uint32_t nblk = (q.nit + blk - 1u) >> q.blk_log2;
    const uint32_t lane = threadIdx.x & 31u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        q.stats_next[0] = 0ull; q.stats_next[2] = 0ull; q.stats_next[3] = 0ull;
        *q.nruns_next = 0u;
        q.stats_cur[1] = (unsigned long long)q.nztiles * q.nytiles;
    }
    // Piece length.  A maximal range of live tiles of one pair group is cut into equal pieces, one run each.  Every
    // run re-reads LEAD plane pairs, so long pieces waste least — but in a sparse phase a pass is latency-bound per
    // warp, not bandwidth-bound, and what counts is that the marching warps get equal shares in as few rounds as
    // possible.  From the previous pass's ranges and live iterations: take the number of pieces per (mean) range with
    // the smallest estimated makespan  ceil(runs / warps) x (piece + lead-in).
    const unsigned long long ranges_prev = q.stats_prev[2], iters_prev = q.stats_prev[3];
    uint32_t chop_it = 64u;
    if (ranges_prev) {
        const unsigned long long mean = (iters_prev + ranges_prev - 1) / ranges_prev;
        unsigned long long best = ~0ull;
        for (uint32_t pcs = 1u; pcs <= 16u; ++pcs) {
            const unsigned long long piece = (mean + pcs - 1) / pcs;
            if (piece < 8u && pcs > 1u) break;                       // lead-in would dominate
            const unsigned long long nr = ranges_prev * pcs;
            const unsigned long long span = ((nr + q.nw - 1) / (q.nw ? q.nw : 1u)) * (piece + 3u);
            if (best == ~0ull || span * 20ull < best * 19ull) { best = span; chop_it = (uint32_t)piece; }   // shorter pieces must win by > 5 %
        }
        if (chop_it < 8u) chop_it = 8u;
    }
    auto tile_quiet = [&](uint32_t zt, uint32_t yt) -> bool {
        if (q.force_live) return false;
        if ((zt == 0 && q.has_lo_neighbour) || (zt == q.nztiles - 1 && q.has_hi_neighbour)) return false;
        // newest stamp of the 3 x 3 neighbourhood: nine independent loads (indices clamped into the map — a clamped
        // index repeats a tile of the neighbourhood, which cannot change the maximum), not a chain of nine
        uint32_t newest = 0;
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                int z = (int)zt + dz, y = (int)yt + dy;
                z = z < 0 ? 0 : (z >= (int)q.nztiles ? (int)q.nztiles - 1 : z);
                y = y < 0 ? 0 : (y >= (int)q.nytiles ? (int)q.nytiles - 1 : y);
                const uint32_t la = q.last_active[(size_t)z * q.nytiles + y];
                newest = la > newest ? la : newest;
            }
        // quiet for steps t_now-4 .. t_now-1  <=>  la (= last active step + 1) + 4 <= t_now
        return (uint64_t)newest + 4ull <= (uint64_t)q.t_now;
    };
    const uint32_t zmask = (1u << q.ztile_log2) - 1u;
    const uint32_t wpb = blockDim.x >> 5;
    uint32_t live_tiles = 0;              // this lane's count of live tiles (each tile is counted by the pair holding its first plane)
    uint32_t my_ranges = 0;               // lane 0: live ranges and live iterations of this warp's pair groups
    unsigned long long my_iters = 0;
    for (uint32_t pg = blockIdx.x * wpb + (threadIdx.x >> 5); pg < npg; pg += gridDim.x * wpb) {
        // live mask of y-tiles [base, base + 32) for the rows of this pair group
        auto live_mask = [&](uint32_t base, bool count) -> uint32_t {
            const uint32_t yt = base + lane;
            bool act = false;
            if (yt < nblk) {
                for (uint32_t g = 0; g < q.groups; ++g) {
                    const uint32_t pair = q.pair_begin + pg * q.groups + g;
                    if (pair >= q.pair_end) break;
                    const int32_t ozl = (int32_t)(q.lz_first + 2u * pair) - 1, ozr = ozl + 1;
                    const bool inl = ozl >= 0 && ozl < (int32_t)q.nzl && yt < q.nytiles;
                    const bool inr = ozr >= 0 && ozr < (int32_t)q.nzl && yt < q.nytiles;
                    const uint32_t ztl = inl ? (uint32_t)ozl >> q.ztile_log2 : 0u, ztr = inr ? (uint32_t)ozr >> q.ztile_log2 : 0u;
                    const bool ql = inl ? tile_quiet(ztl, yt) : true;
                    const bool qr = inr ? ((inl && ztr == ztl) ? ql : tile_quiet(ztr, yt)) : true;
                    act = act || !(ql & qr);
                    if (count) {
                        if (inl && ((uint32_t)ozl & zmask) == 0u) live_tiles += ql ? 0u : 1u;
                        if (inr && ((uint32_t)ozr & zmask) == 0u) live_tiles += qr ? 0u : 1u;
                    }
                }
            }
            return __ballot_sync(0xFFFFFFFFu, act);
        };
        // One sweep over the y-tiles, 32 at a time.  Lane 0 walks the live mask range by range (find-first-set jumps, not
        // bit by bit); a range that is still open at the end of a word carries over.  Each finished range reserves its
        // pieces in the run list with one atomic (the order of runs in the list is arbitrary).
        int32_t start = -1;
        for (uint32_t base = 0; base <= nblk; base += 32u) {
            const uint32_t mask = live_mask(base, true);
            if (lane != 0) continue;
            const uint32_t nvalid = nblk + 1u - base < 32u ? nblk + 1u - base : 32u;     // tile nblk is never live: it closes a range
            uint32_t pos = 0;
            while (pos < nvalid) {
                if (start < 0) {
                    const uint32_t m = mask >> pos;
                    if (!m) break;
                    pos += (uint32_t)__ffs((int)m) - 1u;
                    start = (int32_t)(base + pos);
                } else {
                    const uint32_t m = ~mask >> pos;                 // first tile that is not live
                    const uint32_t skip = m ? (uint32_t)__ffs((int)m) - 1u : 32u;
                    if (pos + skip >= 32u && base + 32u <= nblk) break;               // the range runs on into the next word
                    pos += skip;
                    const uint32_t bend = base + pos;                // live tiles [start, bend): iterations [BLK·start, BLK·bend + NS)
                    const uint32_t it_a = (uint32_t)start << q.blk_log2;
                    uint32_t it_e = (bend << q.blk_log2) + q.ns;     // the tail stores the last tile's top planes
                    if (it_e > q.nit) it_e = q.nit;
                    const uint32_t len = it_e - it_a;
                    const uint32_t pcs = (len + chop_it - 1u) / chop_it;
                    const uint32_t plen = (len + pcs - 1u) / pcs;
                    uint32_t at = atomicAdd(q.nruns, pcs);
                    for (uint32_t i = 0; i < pcs; ++i) {
                        const uint32_t ra = it_a + i * plen, rb = ra + plen < it_e ? ra + plen : it_e;
                        q.runs[3u * at] = pg; q.runs[3u * at + 1u] = ra < rb ? ra : rb; q.runs[3u * at + 2u] = rb;
                        ++at;
                    }
                    ++my_ranges; my_iters += len;
                    start = -1;
                }
            }
        }
    }
    if (lane == 0 && my_ranges) { atomicAdd(&q.stats_cur[2], (unsigned long long)my_ranges); atomicAdd(&q.stats_cur[3], my_iters); }
    for (int o = 16; o > 0; o >>= 1) live_tiles += __shfl_xor_sync(0xFFFFFFFFu, live_tiles, o);
    if (lane == 0 && live_tiles) atomicAdd(&q.stats_cur[0], (unsigned long long)live_tiles);
}

}  // namespace fs3d
