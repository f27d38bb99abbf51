// fs3d.cu — libfs3d: world/slab management and the C ABI declared in include/fs3d.h.
//
// The reference has no voxel world; this library is what its frame loop
// (/root/reference/src/engine/engine.cpp:59-70) would call once per frame, and what its material
// binding builder (/root/reference/src/engine/rendering/materials.cpp:388-418) would be handed a
// volume by.  No CPU fallback exists anywhere in this file: without a CUDA device every
// computing entry point fails with FS3D_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda.h>      // driver-API TYPES only: the functions come from cudaGetDriverEntryPoint, so libfs3d.so has no
                       // link-time dependency on libcuda (it must load on a box without a driver and fail in fs3d_create)
#include <unistd.h>

#include "common.cuh"
#include "aux_kernels.cuh"
#include "step_dispatch.cuh"
#include "step4_kernel.cuh"
#include "raymarch.cuh"

namespace fs3d {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) { g_err = msg; return code; }

cudaError_t step4_launch(int xw, int nbr, const Step4Params &p, unsigned grid, cudaStream_t stream);   // fs3d_s4.cu
cudaError_t halo4_launch(const Halo4Params &h, cudaStream_t stream);
uint32_t step4_units_per_cta(int xw);

constexpr uint64_t SMALL_GRID_VOXELS = 1ull << 27;      // per slab; 512^3 measured the same either way

struct Slab {
    int device = 0;
    uint32_t z0 = 0, nzl = 0;         // global planes [z0, z0 + nzl)
    uint8_t *buf[2] = {nullptr, nullptr};   // (nzl + 2) planes each; plane 0 / nzl+1 are ghosts
    size_t bytes = 0;
    // FS3D_FLAG_EXPORTABLE: the two buffers are driver VMM allocations that can be exported as POSIX file descriptors
    bool vmm = false;
    unsigned long long vmm_handle[2] = {0, 0};   // CUmemGenericAllocationHandle
    size_t vmm_size = 0;                          // mapped size (bytes rounded up to the allocation granularity)
    cudaStream_t s_main = nullptr, s_comm = nullptr;
    cudaEvent_t ev_edges = nullptr, ev_done = nullptr;
    cudaEvent_t ev_out_lo = nullptr, ev_out_hi = nullptr;   // my edge plane has landed in the lower / upper neighbour's ghost
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // fs3d_step_host: copies overlap the kernels
    std::vector<cudaEvent_t> ev_chunk;               // 2 per in-flight chunk (uploaded, computed)
    // fs3d_step_host_packed: two packed chunks in flight each way, and the events that hand them back
    uint8_t *d_stage[4] = {nullptr, nullptr, nullptr, nullptr};   // [0,1] upload staging, [2,3] download staging
    size_t stage_bytes = 0;
    cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};   // [0,1] staging unpacked (free for the next upload), [2,3] staging downloaded
    unsigned long long *d_scratch = nullptr;   // 256 + 1 u64 + 1 u32 flag
    uint8_t *d_img = nullptr; size_t img_bytes = 0;
    float *d_palette = nullptr;
    float *d_thr = nullptr;          // 256 sRGB thresholds (ray-march of this slab into a frame slot)
    bool palette_current = false;    // d_palette holds the world's palette
    // ray-march empty-space skipping: occupancy bits of this slab's 8^3 bricks, valid for content epoch bricks_epoch
    uint32_t *d_bricks = nullptr; size_t bricks_bytes = 0; uint64_t bricks_epoch = ~0ull;
    unsigned long long *d_rm_steps = nullptr;   // loop iterations of the last frame's rays on this slab
    bool rm_bricks_on = false;                  // last decision (hysteresis)
    uint64_t rm_pixels = 0;                     // rays of the last frame
    int num_sms = 0;
    int blocks_per_sm[2][2][2][2][2] = {};   // [PUSH][NS-1][SKIP][OX][TODD] for this world's J
    // fused halo push (one process per GPU, CUDA IPC): my arrival counters and the neighbours' memory
    unsigned long long *d_flags = nullptr;     // [0] iterations delivered from below, [1] from above, [2] watchdog error word, [3..5] wait statistics
    struct Peer {
        uint8_t *buf[2] = {nullptr, nullptr};  // neighbour's two slab buffers (IPC-mapped)
        unsigned long long *flags = nullptr;   // neighbour's d_flags
        uint32_t nzl = 0;
        bool valid = false;
        bool ipc = false;                      // mapped through CUDA IPC (another process) rather than plain peer access
    } peer_lo, peer_hi;
    // fused multi-rank ray-march: the compositor's frame (one slot of W x H 64-bit words per rank)
    struct Frame {
        unsigned long long *base = nullptr;   // slot 0; mine if `owner`, else the compositor's memory mapped through CUDA IPC
        bool owner = false;
        bool ipc = false;                     // mapped through CUDA IPC (closed on destroy); neither: borrowed from a world of this process
        uint32_t width = 0, height = 0, nslots = 0, slot = 0;
        uint32_t *d_rgba = nullptr;           // compositor only: resolved image
    } frame;
    // settled-tile skipping
    uint32_t *d_last_active = nullptr;         // non-null <=> FS3D_FLAG_SKIP_SETTLED
    unsigned long long *d_stats = nullptr;     // [3][4] (tiles live, tiles total, live ranges, live iterations), rotating per pass (skip_plan_kernel)
    uint32_t *d_runs = nullptr, *d_nruns = nullptr;   // live march segments of the launch in flight; nruns[2] rotates per launch
    uint64_t plan_pass = 0, plan_launch = 0;   // passes / SKIP launches planned so far (indices of the rotating counters)
    uint32_t nztiles = 0, nytiles = 0;
};

}  // namespace fs3d

struct fs3d_world {
    fs3d_desc desc{};
    std::vector<fs3d::Slab> slabs;
    std::vector<int32_t> devices;
    int cur = 0;                 // index of the front buffer
    uint64_t step = 0;
    bool external = false;       // created by fs3d_create_slab: caller exchanges halos
    bool halo_pending = false;   // in-process: ghost planes of `cur` are being filled on s_comm
    uint64_t launches = 0;       // kernels launched since creation
    int jidx = 0;                // kernel shape, see step_threads()
    uint32_t lpr = 32, groups = 1;
    float palette[256 * 4];
    int edges_phase = 0;         // external stepping protocol state
    int pass_ns = 1;             // steps fused in the current external pass
    bool p2p = false;            // slab world with IPC-attached neighbours: fused halo push, fs3d_step allowed
    unsigned long long wait_target = 0;   // iterations each neighbour has delivered before the next pass
    bool ghosts_stale = false;   // slab world after fs3d_slab_step_host: fs3d_slab_push_halos must run before fs3d_step
    unsigned long long push_timeout_ns = 20000ull * 1000000ull;   // watchdog of the fused halo push (FS3D_PUSH_TIMEOUT_MS)
    int version = 1;             // schedule version: 1 (four materials) or 2 (FS3D_FLAG_MATERIALS8: eight, SCHEDULE.md §7)
    uint8_t max_material = FS3D_STONE;
    uint64_t content_epoch = 0;  // bumps whenever cells may have changed (steps, edits): the ray-marcher's brick maps follow it
    bool far_valid = false;      // multi-slab p2p worlds: the far ghost planes of the front buffer are current (four-step passes)
    bool fuse4_allowed = false;  // slab worlds of several processes: every rank agreed that its slab supports four-step passes
    bool force_live = false;     // fs3d_step_host in flight: settled-tile plans treat every tile as live
    bool failed = false;         // the watchdog fired: cells are undefined, stepping is refused
};

namespace fs3d {

// ---- kernel dispatch ----------------------------------------------------------------------------
StepFn step_fn_v1(int jidx, int ox, int todd, int skip, int ns, int push) { return step_fn_of<Rules1>(jidx, ox, todd, skip, ns, push); }
static StepFn step_fn(int version, int jidx, int ox, int todd, int skip, int ns, int push) {
    return version == 2 ? step_fn_v2(jidx, ox, todd, skip, ns, push) : step_fn_v1(jidx, ox, todd, skip, ns, push);
}
#ifndef FS3D_HOST_CHUNK_MIB
#define FS3D_HOST_CHUNK_MIB 256ull   // fs3d_step_host streams the grid in chunks of about this size (64 MiB measured 3 % slower)
#endif
// dynamic shared memory of a step kernel: only the staged-load experiment (-DFS3D_STAGE_LOADS=1) uses any
static size_t step_smem(int jidx, int push) {
    return (FS3D_STAGE_LOADS && !push && (jidx == 1 || jidx == 2)) ? stage_smem_bytes<2, STEP_THREADS>() : 0;
}
constexpr uint32_t YTILE_LOG2 = 5, ZTILE_LOG2 = 3;   // activity tile = nx x 32 x 8 voxels

static size_t plane_bytes(const fs3d_world *w) { return (size_t)w->desc.nx * w->desc.ny; }
static uint8_t *owned_ptr(const fs3d_world *w, const Slab &s, int b) { return s.buf[b] + plane_bytes(w); }

static int check_dims(const fs3d_desc *d) {
    if (!d) return fail(FS3D_ERR_INVALID_ARG, "desc is NULL");
    if (d->nx == 0 || d->ny == 0 || d->nz == 0) return fail(FS3D_ERR_BAD_DIMS, "grid dimensions must be non-zero");
    if (d->nx % 32 != 0) return fail(FS3D_ERR_BAD_DIMS, "nx must be a multiple of 32");
    if (d->nx > 4096) return fail(FS3D_ERR_BAD_DIMS, "nx must be <= 4096");
    return FS3D_OK;
}

// ---- exportable volume memory: driver virtual-memory-management allocations ----------------
struct DriverApi {
    CUresult (*memGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*memCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*memAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*memExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    bool ok = false;
};
static const DriverApi &driver_api() {
    static DriverApi api = [] {
        DriverApi a;
        bool ok = true;
        auto get = [&](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult st;
            if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*fn) ok = false;
        };
        get("cuMemGetAllocationGranularity", (void **)&a.memGetAllocationGranularity);
        get("cuMemCreate", (void **)&a.memCreate);
        get("cuMemAddressReserve", (void **)&a.memAddressReserve);
        get("cuMemMap", (void **)&a.memMap);
        get("cuMemSetAccess", (void **)&a.memSetAccess);
        get("cuMemUnmap", (void **)&a.memUnmap);
        get("cuMemAddressFree", (void **)&a.memAddressFree);
        get("cuMemRelease", (void **)&a.memRelease);
        get("cuMemExportToShareableHandle", (void **)&a.memExportToShareableHandle);
        cudaGetLastError();
        a.ok = ok;
        return a;
    }();
    return api;
}
#define FS3D_CU(expr)                                                                                   \
    do {                                                                                                \
        CUresult _r = (expr);                                                                           \
        if (_r != CUDA_SUCCESS)                                                                         \
            return ::fs3d::fail(_r == CUDA_ERROR_OUT_OF_MEMORY ? FS3D_ERR_OOM : FS3D_ERR_CUDA,          \
                                std::string(#expr) + ": CUresult " + std::to_string((int)_r));          \
    } while (0)

static int vmm_alloc(fs3d_world *w, Slab &s, int b) {
    const DriverApi &d = driver_api();
    if (!d.ok) return fail(FS3D_ERR_UNSUPPORTED, "FS3D_FLAG_EXPORTABLE: the driver's virtual memory management API is not available");
    FS3D_CUDA(cudaFree(nullptr));                       // make sure this device's primary context exists and is current
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = s.device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    FS3D_CU(d.memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    s.vmm_size = (s.bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h = 0;
    FS3D_CU(d.memCreate(&h, s.vmm_size, &prop, 0));
    CUdeviceptr ptr = 0;
    CUresult r = d.memAddressReserve(&ptr, s.vmm_size, 0, 0, 0);
    if (r == CUDA_SUCCESS) r = d.memMap(ptr, s.vmm_size, 0, h, 0);
    if (r == CUDA_SUCCESS) {
        // readable and writable from every device of this world (peer pushes, the compositor's ray-march)
        std::vector<CUmemAccessDesc> acc;
        for (int32_t dev : w->devices) {
            bool seen = false;
            for (auto &a : acc) seen = seen || a.location.id == dev;
            if (seen) continue;
            CUmemAccessDesc a{};
            a.location.type = CU_MEM_LOCATION_TYPE_DEVICE; a.location.id = dev; a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            acc.push_back(a);
        }
        r = d.memSetAccess(ptr, s.vmm_size, acc.data(), acc.size());
    }
    if (r != CUDA_SUCCESS) {
        if (ptr) { d.memUnmap(ptr, s.vmm_size); d.memAddressFree(ptr, s.vmm_size); }
        d.memRelease(h);
        return fail(r == CUDA_ERROR_OUT_OF_MEMORY ? FS3D_ERR_OOM : FS3D_ERR_CUDA, "mapping an exportable slab buffer failed: CUresult " + std::to_string((int)r));
    }
    s.vmm = true;
    s.vmm_handle[b] = h;
    s.buf[b] = reinterpret_cast<uint8_t *>(ptr);
    return FS3D_OK;
}
static void vmm_free(Slab &s, int b) {
    const DriverApi &d = driver_api();
    if (!s.buf[b] || !d.ok) return;
    d.memUnmap((CUdeviceptr)s.buf[b], s.vmm_size);
    d.memAddressFree((CUdeviceptr)s.buf[b], s.vmm_size);
    d.memRelease(s.vmm_handle[b]);
    s.buf[b] = nullptr; s.vmm_handle[b] = 0;
}

static int init_slab(fs3d_world *w, Slab &s) {
    FS3D_CUDA(cudaSetDevice(s.device));
    const size_t pb = plane_bytes(w);
    // planes: 0 near ghost-low, 1 .. nzl owned, nzl+1 near ghost-high, then the FAR ghosts of the four-step pass
    // (step4_kernel.cuh): nzl+2 = global plane z1+1, nzl+3 = global plane z0-2
    s.bytes = pb * ((size_t)s.nzl + 4);
    for (int b = 0; b < 2; ++b) {
        if (w->desc.flags & FS3D_FLAG_EXPORTABLE) { int rc = vmm_alloc(w, s, b); if (rc) return rc; }
        else FS3D_CUDA(cudaMalloc(&s.buf[b], s.bytes));
    }
    FS3D_CUDA(cudaStreamCreateWithFlags(&s.s_main, cudaStreamNonBlocking));
    FS3D_CUDA(cudaStreamCreateWithFlags(&s.s_comm, cudaStreamNonBlocking));
    cudaEvent_t *evs[] = {&s.ev_edges, &s.ev_done, &s.ev_out_lo, &s.ev_out_hi};
    for (auto e : evs) FS3D_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    FS3D_CUDA(cudaEventCreate(&s.ev_t0));
    FS3D_CUDA(cudaEventCreate(&s.ev_t1));
    FS3D_CUDA(cudaMalloc(&s.d_scratch, 260 * sizeof(unsigned long long)));
    FS3D_CUDA(cudaMalloc(&s.d_palette, 256 * 4 * sizeof(float)));
    FS3D_CUDA(cudaMalloc(&s.d_thr, 256 * sizeof(float)));
    FS3D_CUDA(cudaDeviceGetAttribute(&s.num_sms, cudaDevAttrMultiProcessorCount, s.device));
    for (int pu = 0; pu < 2; ++pu)
        for (int ns = 1; ns <= 2; ++ns)
            for (int sk = 0; sk < 2; ++sk)
                for (int ox = 0; ox < 2; ++ox)
                    for (int td = 0; td < (ns == 2 ? 1 : 2); ++td) {
                        int nb = 0;
                        const size_t smem = step_smem(w->jidx, pu);
                        if (smem) FS3D_CUDA(cudaFuncSetAttribute(step_fn(w->version, w->jidx, ox, td, sk, ns, pu), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        FS3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, step_fn(w->version, w->jidx, ox, td, sk, ns, pu), step_threads(w->jidx), smem));
                        s.blocks_per_sm[pu][ns - 1][sk][ox][td] = std::max(nb, 1);
                    }
    FS3D_CUDA(cudaMalloc(&s.d_flags, 8 * sizeof(unsigned long long)));
    FS3D_CUDA(cudaMemsetAsync(s.d_flags, 0, 8 * sizeof(unsigned long long), s.s_main));
    if (w->desc.flags & FS3D_FLAG_SKIP_SETTLED) {
        s.nztiles = (s.nzl + (1u << ZTILE_LOG2) - 1) >> ZTILE_LOG2;
        s.nytiles = (w->desc.ny + (1u << YTILE_LOG2) - 1) >> YTILE_LOG2;
        const size_t nt = (size_t)s.nztiles * s.nytiles;
        FS3D_CUDA(cudaMalloc(&s.d_last_active, nt * sizeof(uint32_t)));
        FS3D_CUDA(cudaMalloc(&s.d_stats, 12 * sizeof(unsigned long long)));
        // runs are pieces of >= 4 iterations (skip_plan_kernel: piece length >= 8 before the equal split), at most one
        // range per two y-blocks
        const size_t its = (size_t)w->desc.ny / 2 + 2;
        const size_t max_runs = ((size_t)s.nzl / 2 + 2) * (its / 4 + its / (1u << YTILE_LOG2) + 4);
        FS3D_CUDA(cudaMalloc(&s.d_runs, max_runs * 3 * sizeof(uint32_t)));
        FS3D_CUDA(cudaMalloc(&s.d_nruns, 2 * sizeof(uint32_t)));
        FS3D_CUDA(cudaMemsetAsync(s.d_last_active, 0, nt * sizeof(uint32_t), s.s_main));
        FS3D_CUDA(cudaMemsetAsync(s.d_stats, 0, 12 * sizeof(unsigned long long), s.s_main));
        FS3D_CUDA(cudaMemsetAsync(s.d_nruns, 0, 2 * sizeof(uint32_t), s.s_main));
    }
    // both buffers start as EMPTY with STONE ghost planes (closed box / not-yet-exchanged halo)
    for (int b = 0; b < 2; ++b) {
        FS3D_CUDA(cudaMemsetAsync(s.buf[b], 0, s.bytes, s.s_main));
        FS3D_CUDA(cudaMemsetAsync(s.buf[b], FS3D_STONE, pb, s.s_main));
        FS3D_CUDA(cudaMemsetAsync(s.buf[b] + pb * ((size_t)s.nzl + 1), FS3D_STONE, 3 * pb, s.s_main));
    }
    FS3D_CUDA(cudaStreamSynchronize(s.s_main));
    return FS3D_OK;
}

static void free_slab(Slab &s) {
    cudaSetDevice(s.device);
    if (s.s_main) cudaStreamSynchronize(s.s_main);
    if (s.s_comm) cudaStreamSynchronize(s.s_comm);
    for (int b = 0; b < 2; ++b) if (s.buf[b]) { if (s.vmm) vmm_free(s, b); else cudaFree(s.buf[b]); }
    for (Slab::Peer *pr : {&s.peer_lo, &s.peer_hi}) {
        if (!pr->valid || !pr->ipc) continue;
        for (int b = 0; b < 2; ++b) if (pr->buf[b]) cudaIpcCloseMemHandle(pr->buf[b]);
        if (pr->flags) cudaIpcCloseMemHandle(pr->flags);
    }
    if (s.frame.base) { if (s.frame.owner) cudaFree(s.frame.base); else if (s.frame.ipc) cudaIpcCloseMemHandle(s.frame.base); }
    if (s.frame.d_rgba) cudaFree(s.frame.d_rgba);
    if (s.d_flags) cudaFree(s.d_flags);
    if (s.d_scratch) cudaFree(s.d_scratch);
    if (s.d_img) cudaFree(s.d_img);
    if (s.d_palette) cudaFree(s.d_palette);
    if (s.d_thr) cudaFree(s.d_thr);
    if (s.d_bricks) cudaFree(s.d_bricks);
    if (s.d_rm_steps) cudaFree(s.d_rm_steps);
    if (s.d_last_active) cudaFree(s.d_last_active);
    if (s.d_stats) cudaFree(s.d_stats);
    if (s.d_runs) cudaFree(s.d_runs);
    if (s.d_nruns) cudaFree(s.d_nruns);
    cudaEvent_t evs[] = {s.ev_edges, s.ev_done, s.ev_out_lo, s.ev_out_hi, s.ev_t0, s.ev_t1};
    for (auto e : evs) if (e) cudaEventDestroy(e);
    for (auto e : s.ev_chunk) if (e) cudaEventDestroy(e);
    for (auto e : s.ev_stage) if (e) cudaEventDestroy(e);
    for (auto p : s.d_stage) if (p) cudaFree(p);
    if (s.s_h2d) cudaStreamDestroy(s.s_h2d);
    if (s.s_d2h) cudaStreamDestroy(s.s_d2h);
    if (s.s_main) cudaStreamDestroy(s.s_main);
    if (s.s_comm) cudaStreamDestroy(s.s_comm);
    s = Slab();
}

static void default_palette(float *p) {
    // EMPTY transparent black, SAND, WATER, STONE; the rest a grey ramp.  A host that owns the
    // reference's colors[256] (renderer.cpp:136-393) passes it to fs3d_set_palette instead.
    for (int i = 0; i < 256; ++i) { float g = i / 255.0f; p[4 * i] = g; p[4 * i + 1] = g; p[4 * i + 2] = g; p[4 * i + 3] = 1.0f; }
    const float base[8][4] = {{0, 0, 0, 0}, {0.86f, 0.72f, 0.40f, 1}, {0.15f, 0.40f, 0.85f, 1}, {0.45f, 0.45f, 0.48f, 1},
                              {0.80f, 0.90f, 0.75f, 1} /* GAS */, {0.25f, 0.20f, 0.10f, 1} /* OIL */,
                              {0.95f, 0.65f, 0.10f, 1} /* HONEY */, {0.55f, 0.50f, 0.45f, 1} /* GRAVEL */};
    for (int i = 0; i < 8; ++i) for (int c = 0; c < 4; ++c) p[4 * i + c] = base[i][c];
}

static int finish_create(fs3d_world *w) {
    if (const char *ms = std::getenv("FS3D_PUSH_TIMEOUT_MS")) {
        const long long v = std::atoll(ms);
        if (v > 0) w->push_timeout_ns = (unsigned long long)v * 1000000ull;
    }
    if (w->desc.flags & FS3D_FLAG_MATERIALS8) { w->version = 2; w->max_material = FS3D_GRAVEL; }
    const uint32_t wpr = w->desc.nx / 32;
    w->jidx = wpr <= 32 ? 0 : (wpr <= 64 ? 1 : 2);
    w->lpr = wpr < 32 ? wpr : 32;
    w->groups = w->jidx == 0 ? 32 / w->lpr : 1;
    uint64_t slab_voxels = 0;
    for (auto &s : w->slabs) slab_voxels = std::max<uint64_t>(slab_voxels, (uint64_t)w->desc.nx * w->desc.ny * s.nzl);
    if (w->jidx == 0 && slab_voxels < SMALL_GRID_VOXELS) w->jidx = 3;
    default_palette(w->palette);
    for (auto &s : w->slabs) { int rc = init_slab(w, s); if (rc) return rc; }
    return FS3D_OK;
}

// ---- one step of one slab: pairs [pb, pe) --------------------------------------------------------
struct PairLayout { uint32_t lz_first, npairs; };
static PairLayout pair_layout(const Slab &s, uint32_t oz) {
    PairLayout L;
    L.lz_first = ((s.z0 & 1u) == oz) ? 1u : 0u;
    L.npairs = (s.nzl - L.lz_first) / 2 + 1;
    return L;
}

// launch with programmatic stream serialization: the kernel may be scheduled while its predecessor in the stream is
// still running and waits for it in griddepcontrol.wait (step_kernel.cuh, pdl_wait)
template <class Param>
static cudaError_t launch_pdl(void (*kernel)(const Param), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, const Param &arg) {
    static const bool off = std::getenv("FS3D_NO_PDL") != nullptr;      // A/B switch
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, arg);
}

static int launch_pairs(fs3d_world *w, Slab &s, uint32_t pb, uint32_t pe, int ns, int push = 0) {
    if (pe <= pb) return FS3D_OK;
    const uint64_t t = w->step;
    const uint32_t hoff = (uint32_t)((t >> 1) & 1), todd = (uint32_t)(t & 1);
    PairLayout L = pair_layout(s, hoff);
    StepParams p{};
    p.src = s.buf[w->cur];
    p.dst = s.buf[w->cur ^ 1];
    p.nx = w->desc.nx; p.ny = w->desc.ny;
    p.wpr = w->desc.nx / 32; p.lpr = w->lpr; p.groups = w->groups;
    p.z0 = (int32_t)s.z0; p.nzl = s.nzl;
    p.lz_first = L.lz_first;
    p.pair_begin = pb; p.pair_end = pe;
    p.nit = w->desc.ny / 2 + (uint32_t)ns;
    p.key_xy = step_key(w->desc.seed, t, 0);
    p.key_zy = step_key(w->desc.seed, t, 1);
    p.key_xy2 = step_key(w->desc.seed, t + 1, 0);
    p.key_zy2 = step_key(w->desc.seed, t + 1, 1);
    const int sk = s.d_last_active ? 1 : 0;
    p.last_active = s.d_last_active;
    p.ytile_log2 = YTILE_LOG2; p.ztile_log2 = ZTILE_LOG2; p.nytiles = s.nytiles;
    p.step_plus1 = (uint32_t)(t + (uint64_t)ns);
    if (push) {
        const size_t pbytes = plane_bytes(w);
        const int back = w->cur ^ 1;
        if (s.peer_lo.valid) {
            p.peer_lo_dst = s.peer_lo.buf[back] + pbytes * ((size_t)s.peer_lo.nzl + 1);
            p.peer_lo_flag = s.peer_lo.flags + 1;
        }
        if (s.peer_hi.valid) {
            p.peer_hi_dst = s.peer_hi.buf[back];
            p.peer_hi_flag = s.peer_hi.flags + 0;
        }
        p.my_flags = s.d_flags;
        p.wait_target = w->wait_target;
        p.push_err = s.d_flags + 2;
        p.push_timeout_ns = w->push_timeout_ns;
    }

    const uint64_t npg = ((uint64_t)(pe - pb) + w->groups - 1) / w->groups;
    const uint64_t total = npg * p.nit;
    // enough warps to fill the machine, but never fewer than ~8 iterations per warp
    const int bps = s.blocks_per_sm[push][ns - 1][sk][hoff][todd];
    uint64_t max_blocks = (uint64_t)s.num_sms * bps;
    const int threads = step_threads(w->jidx);
    const uint64_t warps_per_block = threads / 32;
    uint64_t want_warps = std::max<uint64_t>(1, total / 8) * warps_per_pair(w->jidx);
    uint64_t blocks = std::min<uint64_t>(max_blocks, (want_warps + warps_per_block - 1) / warps_per_block);
    blocks = std::max<uint64_t>(blocks, 1);
    if (sk) {
        // plan: compact the live (pair group, y-block) segments of this launch into the run list the warps share out
        PlanParams q{};
        q.last_active = s.d_last_active;
        q.nztiles = s.nztiles; q.nytiles = s.nytiles; q.ztile_log2 = ZTILE_LOG2; q.blk_log2 = YTILE_LOG2 - 1;
        q.nzl = s.nzl; q.lz_first = L.lz_first; q.pair_begin = pb; q.pair_end = pe; q.groups = w->groups;
        q.nit = p.nit; q.ns = (uint32_t)ns;
        q.nw = (uint32_t)(blocks * warps_per_block / warps_per_pair(w->jidx));
        q.t_now = (uint32_t)t;
        q.has_lo_neighbour = s.z0 > 0; q.has_hi_neighbour = s.z0 + s.nzl < w->desc.nz;
        q.force_live = w->force_live ? 1 : 0;
        const uint64_t k = s.plan_pass;           // begin_skip_pass advanced it for this pass
        q.stats_prev = s.d_stats + 4 * ((k + 2) % 3); q.stats_cur = s.d_stats + 4 * (k % 3); q.stats_next = s.d_stats + 4 * ((k + 1) % 3);
        q.runs = s.d_runs; q.nruns = s.d_nruns + (s.plan_launch & 1); q.nruns_next = s.d_nruns + ((s.plan_launch + 1) & 1);
        s.plan_launch++;
        FS3D_CUDA(launch_pdl(skip_plan_kernel, (unsigned)((npg + 3) / 4), 128, 0, s.s_main, q));
        w->launches++;
        p.runs = q.runs; p.nruns = q.nruns;
        // the plan and the march are short when most tiles sleep: programmatic dependent launch hides the launch latency
        FS3D_CUDA(launch_pdl(step_fn(w->version, w->jidx, (int)hoff, (int)todd, sk, ns, push), (unsigned)blocks, (unsigned)threads,
                             step_smem(w->jidx, push), s.s_main, p));
        w->launches++;
        return FS3D_OK;
    }
    step_fn(w->version, w->jidx, (int)hoff, (int)todd, sk, ns, push)<<<(unsigned)blocks, threads, step_smem(w->jidx, push), s.s_main>>>(p);
    FS3D_CUDA(cudaGetLastError());
    w->launches++;
    return FS3D_OK;
}

// settled-tile skipping: a new pass begins on this slab (the statistics counters rotate; the plan itself is made by
// skip_plan_kernel in front of every SKIP launch, straight from the tiles' last_active stamps)
static int launch_skip_map(fs3d_world *, Slab &s) {
    if (s.d_last_active) s.plan_pass++;
    return FS3D_OK;
}

// after the front buffer was edited from outside the step (upload / generate / set_cell / fill_box):
// every tile counts as active "just now", so the next four steps run everywhere
static int touch_all_tiles(fs3d_world *w) {
    w->content_epoch++;
    if (!w->external) w->far_valid = false;      // in-process worlds: an edit may have touched a far-ghost source plane
                                                 // (ranks of a multi-process world reset it together in fs3d_slab_push_halos)
    for (auto &s : w->slabs) {
        if (!s.d_last_active) continue;
        FS3D_CUDA(cudaSetDevice(s.device));
        const uint64_t nt = (uint64_t)s.nztiles * s.nytiles;
        // last_active holds (step + 1); "active at step - 1" = step; a fresh world (step 0) holds 0
        std::vector<uint32_t> v(nt, (uint32_t)w->step);
        FS3D_CUDA(cudaMemcpyAsync(s.d_last_active, v.data(), nt * sizeof(uint32_t), cudaMemcpyHostToDevice, s.s_main));
        FS3D_CUDA(cudaStreamSynchronize(s.s_main));
    }
    return FS3D_OK;
}

// in-process multi-slab halo exchange of the BACK buffer after the edge pairs are written
static int exchange_halos(fs3d_world *w) {
    const int n = (int)w->slabs.size();
    const size_t pb = plane_bytes(w);
    const int back = w->cur ^ 1;
    for (int i = 0; i < n; ++i) {
        Slab &s = w->slabs[i];
        FS3D_CUDA(cudaSetDevice(s.device));
        FS3D_CUDA(cudaStreamWaitEvent(s.s_comm, s.ev_edges, 0));
        if (i > 0) {   // my first owned plane -> lower neighbour's ghost-high
            Slab &d = w->slabs[i - 1];
            FS3D_CUDA(cudaStreamWaitEvent(s.s_comm, d.ev_done, 0));   // d finished reading that ghost last step
            FS3D_CUDA(cudaMemcpyPeerAsync(d.buf[back] + pb * ((size_t)d.nzl + 1), d.device,
                                          s.buf[back] + pb, s.device, pb, s.s_comm));
            FS3D_CUDA(cudaEventRecord(s.ev_out_lo, s.s_comm));   // events live on the recording stream's device
        }
        if (i + 1 < n) {   // my last owned plane -> upper neighbour's ghost-low
            Slab &d = w->slabs[i + 1];
            FS3D_CUDA(cudaStreamWaitEvent(s.s_comm, d.ev_done, 0));
            FS3D_CUDA(cudaMemcpyPeerAsync(d.buf[back], d.device, s.buf[back] + pb * (size_t)s.nzl, s.device, pb, s.s_comm));
            FS3D_CUDA(cudaEventRecord(s.ev_out_hi, s.s_comm));
        }
    }
    return FS3D_OK;
}

// Four steps per pass (step4_kernel.cuh): worlds of schedule version 1 without skipping, rows of exactly 1024 or 2048
// voxels; the step index must be a multiple of four.  Single slabs, or z-slabs wired for the fused halo push whose
// internal boundaries lie on even planes (stage A then needs no halo at all) with at least four planes each.
static bool slab_fuse4_capable(const fs3d_world *w, const Slab &s) {
    if (w->version != 1 || (w->desc.flags & (FS3D_FLAG_NO_FUSE | FS3D_FLAG_NO_FUSE4 | FS3D_FLAG_SKIP_SETTLED))) return false;
    if (w->desc.nx != 1024 && w->desc.nx != 2048 && w->desc.nx != 4096) return false;
    if (s.nzl < 2 || w->desc.ny < 2) return false;
    const bool has_lo = s.z0 > 0, has_hi = s.z0 + s.nzl < w->desc.nz;
    if ((has_lo || has_hi) && ((s.z0 & 1u) || s.nzl < 4)) return false;
    if (has_hi && (s.nzl & 1u)) return false;
    return true;
}
static bool fuse4_ok(const fs3d_world *w) {
    static const bool off = std::getenv("FS3D_NO_FUSE4") != nullptr;      // A/B switch
    if (off) return false;
    for (auto &s : w->slabs) if (!slab_fuse4_capable(w, s)) return false;
    if (w->slabs.size() == 1 && w->slabs[0].nzl == w->desc.nz) return !w->p2p;    // the whole grid in one slab
    if (!w->p2p) return false;                          // copy / NCCL halo paths keep the two-step pass
    return w->external ? w->fuse4_allowed : true;       // ranks decide together (fs3d_slab_allow_fuse4)
}

// deliver slab s's edge rows of buffer `b` into the neighbours' ghost planes of their buffer `b` (near = 0: far planes only)
static int launch_halo4(fs3d_world *w, Slab &s, int b, int near) {
    if (!s.peer_lo.valid && !s.peer_hi.valid) return FS3D_OK;
    Halo4Params h{};
    h.src = s.buf[b];
    h.nzl = s.nzl;
    h.plane_bytes = plane_bytes(w);
    h.near = near;
    if (s.peer_lo.valid) { h.lo_buf = s.peer_lo.buf[b]; h.lo_nzl = s.peer_lo.nzl; h.lo_flag = s.peer_lo.flags + 1; }
    if (s.peer_hi.valid) { h.hi_buf = s.peer_hi.buf[b]; h.hi_nzl = s.peer_hi.nzl; h.hi_flag = s.peer_hi.flags + 0; }
    FS3D_CUDA(halo4_launch(h, s.s_main));
    w->launches++;
    return FS3D_OK;
}

// band_lo / band_hi: the bands of this launch (fs3d_step_host streams the grid through in chunks of bands); default = all
static int launch_fused4(fs3d_world *w, Slab &s, uint32_t band_lo = 0, uint32_t band_hi = ~0u) {
    Step4Params p{};
    p.src = s.buf[w->cur];
    p.dst = s.buf[w->cur ^ 1];
    p.nx = w->desc.nx; p.ny = w->desc.ny; p.wpr = w->desc.nx / 32;
    p.z0 = s.z0; p.nzl = s.nzl;
    p.nA = (s.nzl - 1) / 2 + 1;                 // pair_layout with lz_first = 1 (z0 is even)
    p.nB = s.nzl / 2 + 1;                       // pair_layout with lz_first = 0
    p.nbands = (p.nB + S4_P - 1) / S4_P;
    band_hi = std::min(band_hi, p.nbands);
    if (band_lo >= band_hi) return FS3D_OK;
    p.band0 = band_lo; p.nbands = band_hi - band_lo;
    p.nit = w->desc.ny / 2 + 4;
    for (int i = 0; i < 4; ++i) {
        p.key_xy[i] = step_key(w->desc.seed, w->step + (uint64_t)i, 0);
        p.key_zy[i] = step_key(w->desc.seed, w->step + (uint64_t)i, 1);
    }
    p.has_lo = s.peer_lo.valid ? 1 : 0;
    p.has_hi = s.peer_hi.valid ? 1 : 0;
    p.my_flags = s.d_flags; p.wait_target = w->wait_target;
    p.push_err = s.d_flags + 2; p.push_timeout_ns = w->push_timeout_ns;
    // Edge bands as a late second span: removes every blocking wait (profiles/r02l_bench_n{4,8}.json: 0 waits) but pays a
    // second warm-up on ~100 units — measured 1-2 % slower at N = 4 and 8 than letting the edge units wait at the start
    // (r02l_bench_n{4,8}_edge_early.json), so it is opt-in: useful when ranks are badly skewed.
    static const bool edge_late = std::getenv("FS3D_S4_EDGE_LATE") != nullptr;
    p.edge_late = edge_late ? 1 : 0;
    const int xw = (int)(p.wpr / 32);
    // one CTA per SM, but never fewer than ~16 iterations per unit
    const uint64_t total = (uint64_t)p.nbands * p.nit;
    const uint64_t units = std::max<uint64_t>(1, total / 16);
    const uint64_t upc = step4_units_per_cta(xw);
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)s.num_sms, (units + upc - 1) / upc));
    // Group split with staggered units (step4_kernel.cuh): only where a CTA's span is long against the stagger, i.e. the
    // big grids that are DRAM-heavy anyway.  FS3D_S4_GROUP_SPAN overrides the threshold (0 = never; tests use 1).
    if (!p.has_lo && !p.has_hi) {
        uint64_t min_span = 8ull * upc * S4_K;
        if (const char *e = std::getenv("FS3D_S4_GROUP_SPAN")) { min_span = std::strtoull(e, nullptr, 10); if (min_span == 0) min_span = ~0ull; }
        const uint64_t span = (uint64_t)((p.nbands + upc - 1) / upc) * p.nit / grid;
        p.groups = span >= min_span ? 1 : 0;
    }
    FS3D_CUDA(step4_launch(xw, (p.has_lo || p.has_hi) ? 1 : 0, p, grid, s.s_main));
    w->launches++;
    return FS3D_OK;
}

// a four-step pass of a p2p world: (far-ghost refresh if two-step passes ran since the last delivery) -> step4 on every
// slab -> delivery of the new edge rows into the neighbours' ghost planes of the new front buffer
static int fused4_pass_p2p(fs3d_world *w) {
    if (!w->far_valid) {
        for (auto &s : w->slabs) { FS3D_CUDA(cudaSetDevice(s.device)); int rc = launch_halo4(w, s, w->cur, 0); if (rc) return rc; }
        w->wait_target += HALO4_BLOCKS;
    }
    for (auto &s : w->slabs) { FS3D_CUDA(cudaSetDevice(s.device)); int rc = launch_fused4(w, s); if (rc) return rc; }
    for (auto &s : w->slabs) { FS3D_CUDA(cudaSetDevice(s.device)); int rc = launch_halo4(w, s, w->cur ^ 1, 1); if (rc) return rc; }
    w->wait_target += HALO4_BLOCKS;
    w->far_valid = true;
    return FS3D_OK;
}

// one pass over the grid = ns fused SCHEDULE.md steps (ns = 2 needs an even step index)
static int step_pass(fs3d_world *w, int ns) {
    const int n = (int)w->slabs.size();
    const uint32_t hoff = (uint32_t)((w->step >> 1) & 1);
    if (n == 1) {
        Slab &s = w->slabs[0];
        FS3D_CUDA(cudaSetDevice(s.device));
        PairLayout L = pair_layout(s, hoff);
        int rc = launch_skip_map(w, s);
        static const bool force_push = std::getenv("FS3D_DEBUG_FORCE_PUSH") != nullptr;   // timing experiments only
        if (!rc && ns == 4) rc = w->p2p ? fused4_pass_p2p(w) : launch_fused4(w, s);
        else if (!rc) rc = launch_pairs(w, s, 0, L.npairs, ns, (w->p2p || force_push) ? 1 : 0);
        if (rc) return rc;
        // every warp of an edge pair adds the iterations it finished: nit of this pass x warps per pair
        if (w->p2p && ns != 4) { w->wait_target += (unsigned long long)(w->desc.ny / 2 + (uint32_t)ns) * warps_per_pair(w->jidx); w->far_valid = false; }
    } else if (w->p2p) {
        // one process, one slab per GPU, peer access both ways: the same fused halo push as between ranks — one
        // kernel per slab per pass, edge planes stored straight into the neighbour's ghost plane over NVLink
        if (ns == 4) {
            int rc = fused4_pass_p2p(w);
            if (rc) return rc;
        } else {
            for (auto &s : w->slabs) {
                FS3D_CUDA(cudaSetDevice(s.device));
                PairLayout L = pair_layout(s, hoff);
                int rc = launch_skip_map(w, s);
                if (!rc) rc = launch_pairs(w, s, 0, L.npairs, ns, 1);
                if (rc) return rc;
            }
            w->wait_target += (unsigned long long)(w->desc.ny / 2 + (uint32_t)ns) * warps_per_pair(w->jidx);
            w->far_valid = false;
        }
    } else {
        // 1. edge pairs of every slab, 2. halo copies on the comm streams, 3. interiors
        for (auto &s : w->slabs) {
            FS3D_CUDA(cudaSetDevice(s.device));
            // ghosts of the front buffer must have arrived (written during the previous step)
            if (w->halo_pending) {
                const int i = (int)(&s - &w->slabs[0]);
                if (i > 0) FS3D_CUDA(cudaStreamWaitEvent(s.s_main, w->slabs[i - 1].ev_out_hi, 0));
                if (i + 1 < n) FS3D_CUDA(cudaStreamWaitEvent(s.s_main, w->slabs[i + 1].ev_out_lo, 0));
            }
            PairLayout L = pair_layout(s, hoff);
            int rc = launch_skip_map(w, s);
            if (!rc) rc = launch_pairs(w, s, 0, 1, ns);
            if (!rc && L.npairs > 1) rc = launch_pairs(w, s, L.npairs - 1, L.npairs, ns);
            if (rc) return rc;
            FS3D_CUDA(cudaEventRecord(s.ev_edges, s.s_main));
        }
        int rc = exchange_halos(w);
        if (rc) return rc;
        for (auto &s : w->slabs) {
            FS3D_CUDA(cudaSetDevice(s.device));
            PairLayout L = pair_layout(s, hoff);
            if (L.npairs > 2) { rc = launch_pairs(w, s, 1, L.npairs - 1, ns); if (rc) return rc; }
            FS3D_CUDA(cudaEventRecord(s.ev_done, s.s_main));
        }
        w->halo_pending = true;
    }
    w->cur ^= 1;
    w->step += (uint64_t)ns;
    w->content_epoch++;
    return FS3D_OK;
}

// After the front buffer was rewritten by upload/generate/set_cell: refresh ghost planes between
// in-process slabs (and leave STONE at the global boundary).
static int refresh_ghosts(fs3d_world *w) {
    const int n = (int)w->slabs.size();
    { int rc = touch_all_tiles(w); if (rc) return rc; }
    if (n <= 1 || w->external) return FS3D_OK;
    const size_t pb = plane_bytes(w);
    for (auto &s : w->slabs) { FS3D_CUDA(cudaSetDevice(s.device)); FS3D_CUDA(cudaStreamSynchronize(s.s_main)); FS3D_CUDA(cudaStreamSynchronize(s.s_comm)); }
    for (int i = 0; i < n; ++i) {
        Slab &s = w->slabs[i];
        FS3D_CUDA(cudaSetDevice(s.device));
        if (i > 0) {
            Slab &d = w->slabs[i - 1];
            FS3D_CUDA(cudaMemcpyPeerAsync(d.buf[w->cur] + pb * ((size_t)d.nzl + 1), d.device, s.buf[w->cur] + pb, s.device, pb, s.s_main));
        }
        if (i + 1 < n) {
            Slab &d = w->slabs[i + 1];
            FS3D_CUDA(cudaMemcpyPeerAsync(d.buf[w->cur], d.device, s.buf[w->cur] + pb * (size_t)s.nzl, s.device, pb, s.s_main));
        }
    }
    for (auto &s : w->slabs) { FS3D_CUDA(cudaSetDevice(s.device)); FS3D_CUDA(cudaStreamSynchronize(s.s_main)); }
    w->halo_pending = false;
    return FS3D_OK;
}

// fused halo push: did a kernel give up waiting for a neighbour (step_kernel.cuh, wait_arrival)?
static int check_push_watchdog(fs3d_world *w) {
    if (!w->p2p || w->failed) return FS3D_OK;      // reported once; afterwards only stepping is refused
    for (auto &s : w->slabs) {
        unsigned long long e = 0;
        FS3D_CUDA(cudaSetDevice(s.device));
        FS3D_CUDA(cudaMemcpy(&e, s.d_flags + 2, sizeof(e), cudaMemcpyDeviceToHost));
        if (e == 0) continue;
        w->failed = true;
        char msg[400];
        std::snprintf(msg, sizeof(msg),
                      "fused halo push: slab z [%u, %u) waited more than %llu ms for its %s%s%s neighbour to deliver (arrival "
                      "target %llu): the neighbour rank died or the ranks issued different fs3d_step sequences; this world's "
                      "cells are undefined and it will not step again (destroy and re-create the slab worlds)",
                      s.z0, s.z0 + s.nzl, w->push_timeout_ns / 1000000ull, (e & 1) ? "lower" : "", (e & 3) == 3 ? " and " : "",
                      (e & 2) ? "upper" : "", e >> 8);
        return fail(FS3D_ERR_CUDA, msg);
    }
    return FS3D_OK;
}

static int sync_all(fs3d_world *w) {
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        FS3D_CUDA(cudaStreamSynchronize(s.s_main));
        FS3D_CUDA(cudaStreamSynchronize(s.s_comm));
    }
    return check_push_watchdog(w);
}

static Slab *slab_of_z(fs3d_world *w, uint32_t z) {
    for (auto &s : w->slabs) if (z >= s.z0 && z < s.z0 + s.nzl) return &s;
    return nullptr;
}

static unsigned grid_for(uint64_t n, const Slab &s) {
    uint64_t b = (n + 255) / 256;
    uint64_t cap = (uint64_t)s.num_sms * 16;
    return (unsigned)std::max<uint64_t>(1, std::min(b, cap));
}

// ---- raymarch host side ----------------------------------------------------------------------------
static void rm_camera(const fs3d_world *w, const fs3d_camera *cam, uint32_t width, uint32_t height, uint32_t mode, RMParams &p) {
    p.ox = cam->pos[0]; p.oy = cam->pos[1]; p.oz = cam->pos[2];
    const double yaw = (double)cam->yaw_deg * 3.14159265358979323846 / 180.0;
    p.cs = cam->yaw_deg == 0.0f ? 1.0f : (float)std::cos(yaw);
    p.sn = cam->yaw_deg == 0.0f ? 0.0f : (float)std::sin(yaw);
    p.aspect = cam->aspect;
    p.W = width; p.H = height; p.mode = mode;
    p.nx = w->desc.nx; p.ny = w->desc.ny; p.nz = w->desc.nz;
    const uint32_t nmax = std::max(p.nx, std::max(p.ny, p.nz));
    p.h = 1.0f / (float)nmax;
    p.ex = (float)p.nx * p.h * 0.5f; p.ey = (float)p.ny * p.h * 0.5f; p.ez = (float)p.nz * p.h * 0.5f;
}

// (re)allocates the frame a world owns: n_slots x (width x height) 64-bit words + the resolved image
static int ensure_frame(Slab &s, uint32_t width, uint32_t height, uint32_t n_slots) {
    if (s.frame.base && s.frame.owner && s.frame.width == width && s.frame.height == height && s.frame.nslots == n_slots)
        return FS3D_OK;
    if (s.frame.base) { if (s.frame.owner) cudaFree(s.frame.base); else if (s.frame.ipc) cudaIpcCloseMemHandle(s.frame.base); }
    if (s.frame.d_rgba) cudaFree(s.frame.d_rgba);
    s.frame = Slab::Frame();
    const size_t npix = (size_t)width * height;
    FS3D_CUDA(cudaMalloc(&s.frame.base, npix * n_slots * sizeof(unsigned long long)));
    s.frame.owner = true;
    FS3D_CUDA(cudaMalloc(&s.frame.d_rgba, npix * (sizeof(uint32_t) + sizeof(float))));   // image, then depth
    FS3D_CUDA(cudaMemset(s.frame.base, 0xFF, npix * n_slots * sizeof(unsigned long long)));   // every slot: all misses
    s.frame.width = width; s.frame.height = height; s.frame.nslots = n_slots; s.frame.slot = 0;
    return FS3D_OK;
}

// Empty-space skipping of one slab for the frame about to be marched.  Whether bricks pay is decided from the previous
// frame on this slab: without bricks, more than 48 loop iterations per ray means rays cross a lot of air; with bricks,
// fewer than 8 means they hit at once anyway (the map costs one read of the slab).  Returns the device pointer or
// nullptr, and resets the step counter.  The slab's stream must be idle or ordered (it is: frames are synchronous
// per slab).
static int prepare_bricks(fs3d_world *w, Slab &s, uint32_t mode, uint64_t pixels, const uint32_t **out) {
    *out = nullptr;
    FS3D_CUDA(cudaSetDevice(s.device));
    if (!s.d_rm_steps) {
        FS3D_CUDA(cudaMalloc(&s.d_rm_steps, sizeof(unsigned long long)));
        FS3D_CUDA(cudaMemsetAsync(s.d_rm_steps, 0, sizeof(unsigned long long), s.s_main));
    }
    unsigned long long steps = 0;
    FS3D_CUDA(cudaMemcpyAsync(&steps, s.d_rm_steps, sizeof(steps), cudaMemcpyDeviceToHost, s.s_main));
    FS3D_CUDA(cudaStreamSynchronize(s.s_main));
    FS3D_CUDA(cudaMemsetAsync(s.d_rm_steps, 0, sizeof(unsigned long long), s.s_main));
    const double per_ray = s.rm_pixels ? (double)steps / (double)s.rm_pixels : 0.0;
    s.rm_pixels = pixels;
    bool on = s.rm_bricks_on ? per_ray > 8.0 : per_ray > 48.0;
    if (mode & FS3D_RM_BRICKS) on = true;
    if (mode & FS3D_RM_NO_BRICKS) on = false;
    s.rm_bricks_on = on;
    if (!on) return FS3D_OK;
    const size_t bytes = brick_words(w->desc.nx, w->desc.ny, s.z0, s.z0 + s.nzl) * sizeof(uint32_t);
    if (s.bricks_bytes < bytes) {
        if (s.d_bricks) cudaFree(s.d_bricks);
        s.d_bricks = nullptr; s.bricks_bytes = 0;
        FS3D_CUDA(cudaMalloc(&s.d_bricks, bytes));
        s.bricks_bytes = bytes;
        s.bricks_epoch = ~0ull;
    }
    if (s.bricks_epoch != w->content_epoch) {
        const uint64_t warps = bytes / 16;       // one warp per 128 bricks
        brick_build_kernel<<<grid_for(warps * 32, s), 256, 0, s.s_main>>>(owned_ptr(w, s, w->cur), w->desc.nx, w->desc.ny, s.z0, s.z0 + s.nzl, s.d_bricks);
        FS3D_CUDA(cudaGetLastError());
        w->launches++;
        s.bricks_epoch = w->content_epoch;
    }
    *out = s.d_bricks;
    return FS3D_OK;
}

// march one slab of the world on its own device into `frame_slot` (asynchronous on the slab's stream)
static int raymarch_slab_to_frame(fs3d_world *w, Slab &s, const fs3d_camera *cam, uint32_t width, uint32_t height,
                                  uint32_t mode, unsigned long long *frame_slot) {
    FS3D_CUDA(cudaSetDevice(s.device));
    if (!s.palette_current) {
        FS3D_CUDA(cudaMemcpyAsync(s.d_palette, w->palette, sizeof(w->palette), cudaMemcpyHostToDevice, s.s_main));
        s.palette_current = true;
    }
    RMParams p{};
    rm_camera(w, cam, width, height, mode, p);
    p.nslabs = 1;
    p.slab_ptr[0] = owned_ptr(w, s, w->cur);
    p.slab_z0[0] = s.z0; p.slab_z1[0] = s.z0 + s.nzl;
    p.zheld0 = s.z0; p.zheld1 = s.z0 + s.nzl;
    if ((mode & 15u) == FS3D_RM_VOXELS) { int rc = prepare_bricks(w, s, mode, (uint64_t)width * height, &p.bricks[0]); if (rc) return rc; }
    p.steps_out = s.d_rm_steps;
    p.palette = s.d_palette;
    if (mode & FS3D_RM_SRGB) {
        float thr[256];
        srgb_thresholds(thr);
        thr[255] = INFINITY;
        FS3D_CUDA(cudaMemcpyAsync(s.d_thr, thr, sizeof(thr), cudaMemcpyHostToDevice, s.s_main));
        p.srgb_thr = s.d_thr;
    }
    p.frame = frame_slot;
    dim3 blk(16, 16), grd((width + 15) / 16, (height + 15) / 16);
    raymarch_kernel<<<grd, blk, 0, s.s_main>>>(p);
    FS3D_CUDA(cudaGetLastError());
    w->launches++;
    return FS3D_OK;
}

int raymarch_world(fs3d_world *w, const fs3d_camera *cam, uint32_t width, uint32_t height, uint32_t mode,
                   uint8_t *host_rgba8, float *host_depth, unsigned long long *frame_slot) {
    if ((mode & 15u) > FS3D_RM_VOXELS) return fail(FS3D_ERR_INVALID_ARG, "unknown raymarch mode");
    Slab &s0 = w->slabs[0];
    if (frame_slot) return raymarch_slab_to_frame(w, s0, cam, width, height, mode, frame_slot);   // a rank's slab -> the compositor

    const size_t npix = (size_t)width * height;
    if (w->slabs.size() > 1 && w->p2p && !w->external) {
        // One slab per device with peer access: every device marches its OWN slab and stores (t, rgba) straight
        // into its slot of a frame on the first device; that device keeps the nearest hit per pixel.  Nothing
        // but the finished pixels crosses NVLink.
        bool ok = true;
        for (size_t i = 1; i < w->slabs.size() && ok; ++i) {
            int can = 0;
            FS3D_CUDA(cudaDeviceCanAccessPeer(&can, w->slabs[i].device, s0.device));
            if (!can) { ok = false; break; }
            FS3D_CUDA(cudaSetDevice(w->slabs[i].device));
            cudaError_t e = cudaDeviceEnablePeerAccess(s0.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
            cudaGetLastError();
        }
        if (ok) {
            FS3D_CUDA(cudaSetDevice(s0.device));
            int rc = ensure_frame(s0, width, height, (uint32_t)w->slabs.size());
            if (rc) return rc;
            for (size_t i = 0; i < w->slabs.size(); ++i) {
                rc = raymarch_slab_to_frame(w, w->slabs[i], cam, width, height, mode, s0.frame.base + npix * i);
                if (rc) return rc;
            }
            rc = sync_all(w);
            if (rc) return rc;
            FS3D_CUDA(cudaSetDevice(s0.device));
            float *d_depth = reinterpret_cast<float *>(s0.frame.d_rgba + npix);
            frame_resolve_kernel<<<grid_for(npix, s0), 256, 0, s0.s_main>>>(s0.frame.base, s0.frame.nslots, npix, s0.frame.d_rgba,
                                                                            host_depth ? d_depth : nullptr);
            FS3D_CUDA(cudaGetLastError());
            w->launches++;
            FS3D_CUDA(cudaMemcpyAsync(host_rgba8, s0.frame.d_rgba, npix * 4, cudaMemcpyDeviceToHost, s0.s_main));
            if (host_depth) FS3D_CUDA(cudaMemcpyAsync(host_depth, d_depth, npix * sizeof(float), cudaMemcpyDeviceToHost, s0.s_main));
            FS3D_CUDA(cudaStreamSynchronize(s0.s_main));
            return FS3D_OK;
        }
    }

    // one device marches every slab (slabs on the same device, or reached through peer loads)
    if ((int)w->slabs.size() > RM_MAX_SLABS) return fail(FS3D_ERR_UNSUPPORTED, "too many slabs for raymarch");
    FS3D_CUDA(cudaSetDevice(s0.device));
    for (size_t i = 1; i < w->slabs.size(); ++i) {
        if (w->slabs[i].device == s0.device) continue;
        int can = 0;
        FS3D_CUDA(cudaDeviceCanAccessPeer(&can, s0.device, w->slabs[i].device));
        if (!can) return fail(FS3D_ERR_UNSUPPORTED, "raymarch needs peer access from the first slab's device");
        cudaError_t e = cudaDeviceEnablePeerAccess(w->slabs[i].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FS3D_CUDA(e);
        cudaGetLastError();
    }
    const size_t need = npix * 4 + npix * sizeof(float);
    if (s0.img_bytes < need) {
        if (s0.d_img) cudaFree(s0.d_img);
        s0.d_img = nullptr; s0.img_bytes = 0;
        FS3D_CUDA(cudaMalloc(&s0.d_img, need));
        s0.img_bytes = need;
    }
    float *d_depth = reinterpret_cast<float *>(s0.d_img + npix * 4);
    if (!s0.palette_current) {
        FS3D_CUDA(cudaMemcpyAsync(s0.d_palette, w->palette, sizeof(w->palette), cudaMemcpyHostToDevice, s0.s_main));
        s0.palette_current = true;
    }
    RMParams p{};
    rm_camera(w, cam, width, height, mode, p);
    p.nslabs = (int)w->slabs.size();
    for (int i = 0; i < p.nslabs; ++i) {
        p.slab_ptr[i] = owned_ptr(w, w->slabs[i], w->cur);
        p.slab_z0[i] = w->slabs[i].z0;
        p.slab_z1[i] = w->slabs[i].z0 + w->slabs[i].nzl;
    }
    p.zheld0 = w->slabs.front().z0; p.zheld1 = w->slabs.back().z0 + w->slabs.back().nzl;
    // bricks only where the marching device owns the slab (peer-read slabs march voxel by voxel); the decision is s0's
    if ((mode & 15u) == FS3D_RM_VOXELS) {
        int rc = prepare_bricks(w, s0, mode, (uint64_t)npix, &p.bricks[0]);
        if (rc) return rc;
        for (int i = 1; i < p.nslabs; ++i) {
            p.bricks[i] = nullptr;
            if (w->slabs[i].device != s0.device || p.bricks[0] == nullptr) continue;
            const uint32_t *b = nullptr;
            rc = prepare_bricks(w, w->slabs[i], mode | FS3D_RM_BRICKS, (uint64_t)npix, &b);
            if (rc) return rc;
            FS3D_CUDA(cudaStreamSynchronize(w->slabs[i].s_main));      // built on that slab's stream, read on s0's
            p.bricks[i] = b;
        }
        FS3D_CUDA(cudaSetDevice(s0.device));
    }
    p.steps_out = s0.d_rm_steps;
    p.palette = s0.d_palette;
    p.srgb_thr = nullptr;
    if (mode & FS3D_RM_SRGB) {
        float thr[256];
        srgb_thresholds(thr);
        thr[255] = INFINITY;
        FS3D_CUDA(cudaMemcpyAsync(s0.d_thr, thr, sizeof(thr), cudaMemcpyHostToDevice, s0.s_main));
        p.srgb_thr = s0.d_thr;
    }
    p.img = s0.d_img;
    p.depth = d_depth;
    p.frame = nullptr;
    dim3 blk(16, 16), grd((width + 15) / 16, (height + 15) / 16);
    raymarch_kernel<<<grd, blk, 0, s0.s_main>>>(p);
    FS3D_CUDA(cudaGetLastError());
    w->launches++;
    FS3D_CUDA(cudaMemcpyAsync(host_rgba8, s0.d_img, npix * 4, cudaMemcpyDeviceToHost, s0.s_main));
    if (host_depth) FS3D_CUDA(cudaMemcpyAsync(host_depth, d_depth, npix * sizeof(float), cudaMemcpyDeviceToHost, s0.s_main));
    FS3D_CUDA(cudaStreamSynchronize(s0.s_main));
    return FS3D_OK;
}

}  // namespace fs3d

using namespace fs3d;

// =================================================================================================
extern "C" {

const char *fs3d_last_error(void) { return g_err.c_str(); }
int fs3d_schedule_version(void) { return FS3D_SCHEDULE_VERSION; }
int fs3d_world_schedule_version(fs3d_world *w) { return w ? w->version : 0; }

int fs3d_create(const fs3d_desc *desc, fs3d_world **out) {
    if (!out) return fail(FS3D_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int rc = check_dims(desc);
    if (rc) return rc;
    int ndev = 0;
    FS3D_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(FS3D_ERR_CUDA, "no CUDA device (libfs3d has no CPU fallback)");
    int n = desc->n_gpus <= 0 ? 1 : desc->n_gpus;
    if ((uint32_t)n > desc->nz) return fail(FS3D_ERR_BAD_DIMS, "more slabs than z-planes");
    fs3d_world *w = new (std::nothrow) fs3d_world();
    if (!w) return fail(FS3D_ERR_OOM, "host allocation failed");
    w->desc = *desc;
    w->desc.n_gpus = n;
    w->devices.resize(n);
    int curdev = 0;
    cudaGetDevice(&curdev);
    for (int i = 0; i < n; ++i) {
        w->devices[i] = desc->devices ? desc->devices[i] : (desc->n_gpus <= 1 ? curdev : i);
        if (w->devices[i] < 0 || w->devices[i] >= ndev) { delete w; return fail(FS3D_ERR_INVALID_ARG, "device ordinal out of range"); }
    }
    w->desc.devices = w->devices.data();
    w->slabs.resize(n);
    // contiguous slabs with even boundaries where possible (keeps ZY pairs inside slabs on oz = 0 steps)
    uint32_t zb = 0;
    for (int i = 0; i < n; ++i) {
        uint64_t ze = (uint64_t)desc->nz * (i + 1) / n;
        if (i + 1 < n && (ze & 1)) ze += 1;                                  // prefer even boundaries
        ze = std::min<uint64_t>(ze, (uint64_t)desc->nz - (uint64_t)(n - 1 - i)); // leave a plane for every later slab
        ze = std::max<uint64_t>(ze, (uint64_t)zb + 1);
        if (i + 1 == n) ze = desc->nz;
        w->slabs[i].device = w->devices[i];
        w->slabs[i].z0 = zb;
        w->slabs[i].nzl = (uint32_t)ze - zb;
        zb = (uint32_t)ze;
    }
    if (n > 1) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                if (i == j || std::abs(i - j) != 1) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, w->devices[i], w->devices[j]);
                if (can) { cudaSetDevice(w->devices[i]); cudaError_t e = cudaDeviceEnablePeerAccess(w->devices[j], 0); if (e != cudaSuccess) cudaGetLastError(); }
            }
    }
    rc = finish_create(w);
    if (rc) { std::string keep = g_err; fs3d_destroy(w); g_err = keep; return rc; }
    // Slabs on distinct devices with peer access both ways use the fused halo push (kernels store into the
    // neighbour's memory directly).  Slabs sharing a device keep the copy-based exchange: two persistent
    // kernels on one device could wait on each other for SMs.
    bool push_ok = n > 1 && !(desc->flags & FS3D_FLAG_NO_PEER_PUSH);
    for (int i = 0; i < n && push_ok; ++i)
        for (int j = 0; j < n && push_ok; ++j) {
            if (i == j) continue;
            if (w->devices[i] == w->devices[j]) {
                // FS3D_FLAG_PEER_PUSH_SHARED_DEVICE: run the PUSH kernels even so (what a one-GPU box needs to test them).
                // It cannot deadlock: only the warps of a slab's two edge pairs ever wait, every other CTA of a kernel
                // retires, so the neighbour's previous pass always finds SMs.
                if (!(desc->flags & FS3D_FLAG_PEER_PUSH_SHARED_DEVICE)) push_ok = false;
                continue;
            }
            if (std::abs(i - j) == 1) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, w->devices[i], w->devices[j]);
                if (!can) push_ok = false;
            }
        }
    if (push_ok) {
        for (int i = 0; i < n; ++i) {
            Slab &s = w->slabs[i];
            if (i > 0) { Slab &d = w->slabs[i - 1]; s.peer_lo.buf[0] = d.buf[0]; s.peer_lo.buf[1] = d.buf[1]; s.peer_lo.flags = d.d_flags; s.peer_lo.nzl = d.nzl; s.peer_lo.valid = true; }
            if (i + 1 < n) { Slab &d = w->slabs[i + 1]; s.peer_hi.buf[0] = d.buf[0]; s.peer_hi.buf[1] = d.buf[1]; s.peer_hi.flags = d.d_flags; s.peer_hi.nzl = d.nzl; s.peer_hi.valid = true; }
        }
        w->p2p = true;
    }
    if (n > 1) {   // ghost planes between slabs start as the neighbour's (EMPTY) edge plane, not as STONE
        rc = refresh_ghosts(w);
        if (rc) { std::string keep = g_err; fs3d_destroy(w); g_err = keep; return rc; }
    }
    *out = w;
    return FS3D_OK;
}

int fs3d_create_slab(const fs3d_desc *desc, uint32_t z_begin, uint32_t z_end, fs3d_world **out) {
    if (!out) return fail(FS3D_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int rc = check_dims(desc);
    if (rc) return rc;
    if (z_begin >= z_end || z_end > desc->nz) return fail(FS3D_ERR_OUT_OF_RANGE, "slab [z_begin, z_end) outside the grid");
    int ndev = 0;
    FS3D_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(FS3D_ERR_CUDA, "no CUDA device (libfs3d has no CPU fallback)");
    fs3d_world *w = new (std::nothrow) fs3d_world();
    if (!w) return fail(FS3D_ERR_OOM, "host allocation failed");
    w->desc = *desc;
    w->desc.n_gpus = 1;
    int curdev = 0;
    cudaGetDevice(&curdev);
    w->devices.assign(1, curdev);
    w->desc.devices = w->devices.data();
    w->external = true;
    w->slabs.resize(1);
    w->slabs[0].device = curdev;
    w->slabs[0].z0 = z_begin;
    w->slabs[0].nzl = z_end - z_begin;
    rc = finish_create(w);
    if (rc) { std::string keep = g_err; fs3d_destroy(w); g_err = keep; return rc; }
    *out = w;
    return FS3D_OK;
}

void fs3d_destroy(fs3d_world *w) {
    if (!w) return;
    for (auto &s : w->slabs) free_slab(s);
    delete w;
}

int fs3d_sync(fs3d_world *w) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    return sync_all(w);
}

int fs3d_kernel_launches(fs3d_world *w, uint64_t *out) {
    if (!w || !out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    *out = w->launches;
    return FS3D_OK;
}

int fs3d_step_index(fs3d_world *w, uint64_t *out) {
    if (!w || !out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    *out = w->step;
    return FS3D_OK;
}

int fs3d_step(fs3d_world *w, uint32_t n_steps) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (w->external && w->desc.nz != w->slabs[0].nzl && !w->p2p)
        return fail(FS3D_ERR_UNSUPPORTED, "slab worlds step through fs3d_slab_step_* with a caller-driven halo exchange, "
                                          "or through fs3d_step after fs3d_slab_ipc_attach");
    if (w->ghosts_stale)
        return fail(FS3D_ERR_UNSUPPORTED, "after fs3d_slab_step_host call fs3d_slab_push_halos on every rank (and barrier) before fs3d_step");
    if (w->failed)
        return fail(FS3D_ERR_CUDA, "the halo-push watchdog fired earlier: cells are undefined; destroy and re-create the slab worlds");
    uint32_t left = n_steps;
    while (left > 0) {
        // steps 2k and 2k + 1 share the z-pairing and x-offset, so they fuse into one pass
        // and four steps starting on a multiple of four share one pass where step4_kernel.cuh applies
        const int ns = (left >= 4 && (w->step & 3) == 0 && fuse4_ok(w)) ? 4
                     : (left >= 2 && (w->step & 1) == 0 && !(w->desc.flags & FS3D_FLAG_NO_FUSE)) ? 2 : 1;
        int rc = step_pass(w, ns);
        if (rc) return rc;
        left -= (uint32_t)ns;
    }
    return FS3D_OK;
}

int fs3d_step_timed(fs3d_world *w, uint32_t n_steps, float *ms, uint64_t *kernel_launches) {
    if (!w || !ms) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    int rc = sync_all(w);
    if (rc) return rc;
    const uint64_t l0 = w->launches;
    for (auto &s : w->slabs) { FS3D_CUDA(cudaSetDevice(s.device)); FS3D_CUDA(cudaEventRecord(s.ev_t0, s.s_main)); }
    rc = fs3d_step(w, n_steps);
    if (rc) return rc;
    for (auto &s : w->slabs) { FS3D_CUDA(cudaSetDevice(s.device)); FS3D_CUDA(cudaEventRecord(s.ev_t1, s.s_main)); }
    rc = sync_all(w);
    if (rc) return rc;
    float best = 0.f;
    for (auto &s : w->slabs) {
        float t = 0.f;
        FS3D_CUDA(cudaSetDevice(s.device));
        FS3D_CUDA(cudaEventElapsedTime(&t, s.ev_t0, s.ev_t1));
        best = std::max(best, t);
    }
    *ms = best;
    if (kernel_launches) *kernel_launches = w->launches - l0;
    return FS3D_OK;
}

// ---- cell access -----------------------------------------------------------------------------------
static const char *bad_material_msg(const fs3d_world *w) {
    return w->version == 2 ? "material codes 8-255 are reserved" : "material codes 4-255 are reserved (codes 4-7 need FS3D_FLAG_MATERIALS8)";
}
static uint32_t bad_bits(const fs3d_world *w) { return w->version == 2 ? 0xF8F8F8F8u : 0xFCFCFCFCu; }   // bits no valid code has

static int check_cell(fs3d_world *w, uint32_t x, uint32_t y, uint32_t z) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (x >= w->desc.nx || y >= w->desc.ny || z >= w->desc.nz) return fail(FS3D_ERR_OUT_OF_RANGE, "cell outside the grid");
    return FS3D_OK;
}

int fs3d_set_cell(fs3d_world *w, uint32_t x, uint32_t y, uint32_t z, uint8_t m) {
    int rc = check_cell(w, x, y, z);
    if (rc) return rc;
    if (m > w->max_material) return fail(FS3D_ERR_BAD_MATERIAL, bad_material_msg(w));
    Slab *s = slab_of_z(w, z);
    if (!s) return FS3D_OK;   // not in this rank's slab: nothing to do
    rc = sync_all(w);
    if (rc) return rc;
    FS3D_CUDA(cudaSetDevice(s->device));
    uint8_t *p = owned_ptr(w, *s, w->cur) + x + (size_t)w->desc.nx * (y + (size_t)w->desc.ny * (z - s->z0));
    FS3D_CUDA(cudaMemcpy(p, &m, 1, cudaMemcpyHostToDevice));
    if (z == s->z0 || z == s->z0 + s->nzl - 1) return refresh_ghosts(w);
    return touch_all_tiles(w);
}

int fs3d_get_cell(fs3d_world *w, uint32_t x, uint32_t y, uint32_t z, uint8_t *m) {
    int rc = check_cell(w, x, y, z);
    if (rc) return rc;
    if (!m) return fail(FS3D_ERR_INVALID_ARG, "m is NULL");
    Slab *s = slab_of_z(w, z);
    if (!s) return fail(FS3D_ERR_OUT_OF_RANGE, "cell is not in this rank's slab");
    rc = sync_all(w);
    if (rc) return rc;
    FS3D_CUDA(cudaSetDevice(s->device));
    const uint8_t *p = owned_ptr(w, *s, w->cur) + x + (size_t)w->desc.nx * (y + (size_t)w->desc.ny * (z - s->z0));
    FS3D_CUDA(cudaMemcpy(m, p, 1, cudaMemcpyDeviceToHost));
    return FS3D_OK;
}

int fs3d_fill_box(fs3d_world *w, const uint32_t lo[3], const uint32_t hi[3], uint8_t m) {
    if (!w || !lo || !hi) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (m > w->max_material) return fail(FS3D_ERR_BAD_MATERIAL, bad_material_msg(w));
    if (hi[0] > w->desc.nx || hi[1] > w->desc.ny || hi[2] > w->desc.nz || lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2])
        return fail(FS3D_ERR_OUT_OF_RANGE, "box outside the grid");
    if (lo[0] == hi[0] || lo[1] == hi[1] || lo[2] == hi[2]) return FS3D_OK;
    int rc = sync_all(w);
    if (rc) return rc;
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        uint64_t n = (uint64_t)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
        fill_box_kernel<<<grid_for(n, s), 256, 0, s.s_main>>>(owned_ptr(w, s, w->cur), w->desc.nx, w->desc.ny, s.z0, s.nzl,
                                                               lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], m);
        FS3D_CUDA(cudaGetLastError());
    }
    rc = sync_all(w);
    if (rc) return rc;
    return refresh_ghosts(w);
}

int fs3d_paint_sphere(fs3d_world *w, int32_t cx, int32_t cy, int32_t cz, uint32_t radius, uint8_t m, int only_empty) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (m > w->max_material) return fail(FS3D_ERR_BAD_MATERIAL, bad_material_msg(w));
    if (radius > (1u << 20)) return fail(FS3D_ERR_INVALID_ARG, "brush radius too large");
    int rc = sync_all(w);
    if (rc) return rc;
    const uint64_t side = 2ull * radius + 1;
    const uint64_t n = std::min<uint64_t>(side, w->desc.nx) * std::min<uint64_t>(side, w->desc.ny) * std::min<uint64_t>(side, w->desc.nz);
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        paint_sphere_kernel<<<grid_for(n, s), 256, 0, s.s_main>>>(owned_ptr(w, s, w->cur), w->desc.nx, w->desc.ny, w->desc.nz,
                                                                   s.z0, s.nzl, cx, cy, cz, (int64_t)radius, m, only_empty);
        FS3D_CUDA(cudaGetLastError());
    }
    rc = sync_all(w);
    if (rc) return rc;
    return refresh_ghosts(w);
}

int fs3d_generate(fs3d_world *w, int scene_id, uint64_t seed) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (scene_id < 0 || scene_id > FS3D_SCENE_MIXED8) return fail(FS3D_ERR_INVALID_ARG, "unknown scene id");
    if (scene_id > FS3D_SCENE_MIXED_NOISE && w->version != 2)
        return fail(FS3D_ERR_BAD_MATERIAL, "scenes RANDOM8 / MIXED8 hold materials 4-7: create the world with FS3D_FLAG_MATERIALS8");
    int rc = sync_all(w);
    if (rc) return rc;
    const uint32_t key = step_key(seed, 0, 7);
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        uint64_t nvec = (uint64_t)w->desc.nx / 16 * w->desc.ny * s.nzl;
        generate_kernel<<<grid_for(nvec, s), 256, 0, s.s_main>>>(owned_ptr(w, s, w->cur), w->desc.nx, w->desc.ny, w->desc.nz,
                                                                  s.z0, s.nzl, scene_id, key);
        FS3D_CUDA(cudaGetLastError());
    }
    rc = sync_all(w);
    if (rc) return rc;
    return refresh_ghosts(w);
}

int fs3d_upload(fs3d_world *w, const uint8_t *host) {
    if (!w || !host) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    int rc = sync_all(w);
    if (rc) return rc;
    const size_t pb = plane_bytes(w);
    const uint32_t zbase = w->slabs[0].z0;
    // stage into the BACK buffer, validate there, and only then flip: a rejected upload leaves
    // the world untouched
    const int back = w->cur ^ 1;
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        FS3D_CUDA(cudaMemcpyAsync(owned_ptr(w, s, back), host + pb * (size_t)(s.z0 - zbase), pb * s.nzl, cudaMemcpyHostToDevice, s.s_main));
        uint32_t *flag = reinterpret_cast<uint32_t *>(s.d_scratch + 258);
        FS3D_CUDA(cudaMemsetAsync(flag, 0, sizeof(uint32_t), s.s_main));
        uint64_t n16 = pb * s.nzl / 16;
        validate_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(owned_ptr(w, s, back), n16, bad_bits(w), flag);
        FS3D_CUDA(cudaGetLastError());
    }
    bool bad = false;
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        uint32_t f = 0;
        FS3D_CUDA(cudaMemcpyAsync(&f, reinterpret_cast<uint32_t *>(s.d_scratch + 258), sizeof(uint32_t), cudaMemcpyDeviceToHost, s.s_main));
        FS3D_CUDA(cudaStreamSynchronize(s.s_main));
        bad = bad || f != 0;
    }
    if (bad) return fail(FS3D_ERR_BAD_MATERIAL, std::string("upload: ") + bad_material_msg(w));
    w->cur = back;
    return refresh_ghosts(w);
}

int fs3d_download(fs3d_world *w, uint8_t *host) {
    if (!w || !host) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    int rc = sync_all(w);
    if (rc) return rc;
    const size_t pb = plane_bytes(w);
    const uint32_t zbase = w->slabs[0].z0;
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        FS3D_CUDA(cudaMemcpyAsync(host + pb * (size_t)(s.z0 - zbase), owned_ptr(w, s, w->cur), pb * s.nzl, cudaMemcpyDeviceToHost, s.s_main));
    }
    return sync_all(w);
}

// packed transfers (the checkpoint encoding without its header): companions of fs3d_step_host_packed
static int transfer_packed(fs3d_world *w, uint8_t *host, bool up) {
    int rc = sync_all(w);
    if (rc) return rc;
    const size_t pb = plane_bytes(w);
    const uint32_t div = w->version == 2 ? 2u : 4u;
    const uint32_t zbase = w->slabs[0].z0;
    const int buf = up ? (w->cur ^ 1) : w->cur;          // uploads are staged in the back buffer and validated there
    const uint32_t planes_per_chunk = (uint32_t)std::max<uint64_t>(1, (64ull << 20) / (pb / div));
    bool bad = false;
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        uint32_t *packed = nullptr;
        const uint32_t cp = std::min(planes_per_chunk, s.nzl);
        FS3D_CUDA(cudaMalloc(&packed, (size_t)cp * (pb / div)));
        uint32_t *flag = reinterpret_cast<uint32_t *>(s.d_scratch + 258);
        FS3D_CUDA(cudaMemsetAsync(flag, 0, sizeof(uint32_t), s.s_main));
        cudaError_t err = cudaSuccess;
        for (uint32_t z = 0; z < s.nzl && err == cudaSuccess; z += cp) {
            const uint32_t n = std::min(cp, s.nzl - z);
            const uint64_t n16 = pb * n / 16, nbytes = n16 * (16 / div);
            uint8_t *h = host + pb / div * (size_t)(s.z0 - zbase + z);
            uint8_t *cells = owned_ptr(w, s, buf) + pb * z;
            if (up) {
                err = cudaMemcpyAsync(packed, h, nbytes, cudaMemcpyHostToDevice, s.s_main);
                if (w->version == 2) unpack4_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(reinterpret_cast<const uint2 *>(packed), n16, cells);
                else unpack2_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(packed, n16, cells);
                if (w->version == 2) validate_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(cells, n16, bad_bits(w), flag);
            } else {
                if (w->version == 2) pack4_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(cells, n16, reinterpret_cast<uint2 *>(packed));
                else pack2_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(cells, n16, packed);
                err = cudaMemcpyAsync(h, packed, nbytes, cudaMemcpyDeviceToHost, s.s_main);
            }
            if (err == cudaSuccess) err = cudaStreamSynchronize(s.s_main);      // `packed` is reused by the next chunk
            w->launches++;
        }
        uint32_t f = 0;
        if (err == cudaSuccess) err = cudaMemcpy(&f, flag, sizeof(f), cudaMemcpyDeviceToHost);
        cudaFree(packed);
        FS3D_CUDA(err);
        bad = bad || f != 0;
    }
    if (!up) return FS3D_OK;
    if (bad) return fail(FS3D_ERR_BAD_MATERIAL, std::string("packed upload: ") + bad_material_msg(w));
    w->cur ^= 1;
    return refresh_ghosts(w);
}

int fs3d_upload_packed(fs3d_world *w, const uint8_t *packed) {
    if (!w || !packed) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    return transfer_packed(w, const_cast<uint8_t *>(packed), true);
}

int fs3d_download_packed(fs3d_world *w, uint8_t *packed) {
    if (!w || !packed) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    return transfer_packed(w, packed, false);
}

// ---- reductions --------------------------------------------------------------------------------------
int fs3d_histogram(fs3d_world *w, uint64_t counts[256]) {
    if (!w || !counts) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    int rc = sync_all(w);
    if (rc) return rc;
    std::memset(counts, 0, 256 * sizeof(uint64_t));
    const size_t pb = plane_bytes(w);
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        FS3D_CUDA(cudaMemsetAsync(s.d_scratch, 0, 256 * sizeof(unsigned long long), s.s_main));
        uint64_t n16 = pb * s.nzl / 16;
        histogram_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(owned_ptr(w, s, w->cur), n16, s.d_scratch);
        FS3D_CUDA(cudaGetLastError());
        unsigned long long h[256];
        FS3D_CUDA(cudaMemcpyAsync(h, s.d_scratch, sizeof(h), cudaMemcpyDeviceToHost, s.s_main));
        FS3D_CUDA(cudaStreamSynchronize(s.s_main));
        for (int i = 0; i < 256; ++i) counts[i] += h[i];
    }
    return FS3D_OK;
}

int fs3d_digest(fs3d_world *w, uint64_t *out) {
    if (!w || !out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    int rc = sync_all(w);
    if (rc) return rc;
    const size_t pb = plane_bytes(w);
    uint64_t sum = 0;
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        unsigned long long *d = s.d_scratch + 256;
        FS3D_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), s.s_main));
        uint64_t n16 = pb * s.nzl / 16;
        digest_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(owned_ptr(w, s, w->cur), n16, (uint64_t)s.z0 * pb, d);
        FS3D_CUDA(cudaGetLastError());
        unsigned long long h = 0;
        FS3D_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s.s_main));
        FS3D_CUDA(cudaStreamSynchronize(s.s_main));
        sum += h;
    }
    *out = sum;
    return FS3D_OK;
}

int fs3d_activity(fs3d_world *w, uint64_t *tiles_run, uint64_t *tiles_total) {
    if (!w || !tiles_run || !tiles_total) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    *tiles_run = 0; *tiles_total = 0;
    int rc = sync_all(w);
    if (rc) return rc;
    for (auto &s : w->slabs) {
        if (!s.d_last_active) { *tiles_run += 1; *tiles_total += 1; continue; }   // skipping off: everything runs
        FS3D_CUDA(cudaSetDevice(s.device));
        unsigned long long st[2] = {0, 0};
        FS3D_CUDA(cudaMemcpy(st, s.d_stats + 4 * (s.plan_pass % 3), sizeof(st), cudaMemcpyDeviceToHost));
        *tiles_run += st[0];
        *tiles_total += st[1] ? st[1] : (unsigned long long)s.nztiles * s.nytiles;
    }
    return FS3D_OK;
}

// ---- renderer hand-off -----------------------------------------------------------------------------
int fs3d_num_slabs(fs3d_world *w, int32_t *out) {
    if (!w || !out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    *out = (int32_t)w->slabs.size();
    return FS3D_OK;
}

int fs3d_volume_view(fs3d_world *w, int32_t slab, fs3d_view *out) {
    if (!w || !out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (slab < 0 || slab >= (int32_t)w->slabs.size()) return fail(FS3D_ERR_OUT_OF_RANGE, "slab index out of range");
    int rc = sync_all(w);
    if (rc) return rc;
    Slab &s = w->slabs[slab];
    out->dev_ptr = owned_ptr(w, s, w->cur);
    out->device = s.device;
    out->nx = w->desc.nx; out->ny = w->desc.ny;
    out->z0 = s.z0; out->z1 = s.z0 + s.nzl;
    out->pitch_y = w->desc.nx;
    out->pitch_z = plane_bytes(w);
    out->step = w->step;
    return FS3D_OK;
}

int fs3d_volume_export_fd(fs3d_world *w, int32_t slab, fs3d_export *out) {
    if (!w || !out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (slab < 0 || slab >= (int32_t)w->slabs.size()) return fail(FS3D_ERR_OUT_OF_RANGE, "slab index out of range");
    Slab &s = w->slabs[slab];
    if (!s.vmm) return fail(FS3D_ERR_UNSUPPORTED, "create the world with FS3D_FLAG_EXPORTABLE to export its buffers");
    int rc = sync_all(w);
    if (rc) return rc;
    FS3D_CUDA(cudaSetDevice(s.device));
    const DriverApi &d = driver_api();
    int fds[2] = {-1, -1};
    for (int b = 0; b < 2; ++b) {
        CUresult r = d.memExportToShareableHandle(&fds[b], s.vmm_handle[b], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
        if (r != CUDA_SUCCESS) {
            if (fds[0] >= 0) close(fds[0]);
            return fail(FS3D_ERR_CUDA, "cuMemExportToShareableHandle: CUresult " + std::to_string((int)r));
        }
    }
    out->fd[0] = fds[0]; out->fd[1] = fds[1];
    out->alloc_bytes = s.vmm_size;
    out->first_cell_offset = plane_bytes(w);         // local plane 0 is the ghost plane below the slab
    out->front = (uint32_t)w->cur;
    out->device = s.device;
    out->nx = w->desc.nx; out->ny = w->desc.ny;
    out->z0 = s.z0; out->z1 = s.z0 + s.nzl;
    out->pitch_y = w->desc.nx;
    out->pitch_z = plane_bytes(w);
    out->step = w->step;
    return FS3D_OK;
}

int fs3d_set_palette(fs3d_world *w, const float *rgba256x4) {
    if (!w || !rgba256x4) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    std::memcpy(w->palette, rgba256x4, sizeof(w->palette));
    for (auto &s : w->slabs) s.palette_current = false;
    return FS3D_OK;
}

int fs3d_raymarch_bricks_in_use(fs3d_world *w, int32_t slab) {
    if (!w || slab < 0 || slab >= (int32_t)w->slabs.size()) return 0;
    return w->slabs[slab].rm_bricks_on ? 1 : 0;
}

int fs3d_raymarch(fs3d_world *w, const fs3d_camera *cam, uint32_t width, uint32_t height, uint32_t mode, uint8_t *host_rgba8) {
    if (!w || !cam || !host_rgba8) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (width == 0 || height == 0 || width > 16384 || height > 16384) return fail(FS3D_ERR_INVALID_ARG, "bad image size");
    int rc = sync_all(w);
    if (rc) return rc;
    return raymarch_world(w, cam, width, height, mode, host_rgba8, nullptr);
}

int fs3d_raymarch_depth(fs3d_world *w, const fs3d_camera *cam, uint32_t width, uint32_t height, uint32_t mode,
                        uint8_t *host_rgba8, float *host_depth) {
    if (!w || !cam || !host_rgba8 || !host_depth) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (width == 0 || height == 0 || width > 16384 || height > 16384) return fail(FS3D_ERR_INVALID_ARG, "bad image size");
    int rc = sync_all(w);
    if (rc) return rc;
    return raymarch_world(w, cam, width, height, mode, host_rgba8, host_depth);
}

// ---- one-process-per-GPU slab protocol -----------------------------------------------------------
int fs3d_slab_halo(fs3d_world *w, int back, fs3d_halo *out) {
    if (!w || !out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (w->slabs.size() != 1) return fail(FS3D_ERR_UNSUPPORTED, "fs3d_slab_halo needs a world made by fs3d_create_slab");
    Slab &s = w->slabs[0];
    const size_t pb = plane_bytes(w);
    uint8_t *b = s.buf[back ? (w->cur ^ 1) : w->cur];
    out->recv_lo = b;
    out->send_lo = b + pb;
    out->send_hi = b + pb * (size_t)s.nzl;
    out->recv_hi = b + pb * ((size_t)s.nzl + 1);
    out->plane_bytes = pb;
    out->stream = (void *)s.s_main;
    return FS3D_OK;
}

int fs3d_slab_step_edges(fs3d_world *w) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (w->slabs.size() != 1) return fail(FS3D_ERR_UNSUPPORTED, "needs a single-slab world");
    if (w->edges_phase != 0) return fail(FS3D_ERR_INVALID_ARG, "fs3d_slab_step_edges called twice without finish");
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    PairLayout L = pair_layout(s, (uint32_t)((w->step >> 1) & 1));
    if (w->pass_ns == 2 && (w->step & 1)) return fail(FS3D_ERR_INVALID_ARG, "a fused 2-step pass must start on an even step");
    int rc = launch_skip_map(w, s);
    if (!rc) rc = launch_pairs(w, s, 0, 1, w->pass_ns);
    if (!rc && L.npairs > 1) rc = launch_pairs(w, s, L.npairs - 1, L.npairs, w->pass_ns);
    if (rc) return rc;
    w->edges_phase = 1;
    return FS3D_OK;
}

int fs3d_slab_step_interior(fs3d_world *w) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (w->slabs.size() != 1) return fail(FS3D_ERR_UNSUPPORTED, "needs a single-slab world");
    if (w->edges_phase != 1) return fail(FS3D_ERR_INVALID_ARG, "call fs3d_slab_step_edges first");
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    PairLayout L = pair_layout(s, (uint32_t)((w->step >> 1) & 1));
    if (L.npairs > 2) { int rc = launch_pairs(w, s, 1, L.npairs - 1, w->pass_ns); if (rc) return rc; }
    w->edges_phase = 2;
    return FS3D_OK;
}

int fs3d_slab_step_finish(fs3d_world *w) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (w->edges_phase != 2) return fail(FS3D_ERR_INVALID_ARG, "call fs3d_slab_step_interior first");
    w->cur ^= 1;
    w->step += (uint64_t)w->pass_ns;
    w->content_epoch++;
    w->edges_phase = 0;
    return FS3D_OK;
}

int fs3d_slab_pass_steps(fs3d_world *w, uint32_t n_steps) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (n_steps != 1 && n_steps != 2) return fail(FS3D_ERR_INVALID_ARG, "a pass fuses 1 or 2 steps");
    if (w->edges_phase != 0) return fail(FS3D_ERR_INVALID_ARG, "cannot change the pass size inside a pass");
    w->pass_ns = (int)n_steps;
    return FS3D_OK;
}

// ---- out-of-core / end-to-end step: host grid in, host grid out, copies overlapped with compute ----
// Streams the planes a single-slab world holds through the GPU; the ghost planes of the front buffer must
// already be right (STONE at the global boundary, the neighbours' edge planes for a rank's slab).
// packed = the host keeps the grid in the checkpoint encoding (2 bits per voxel, 4 for schedule version 2): only those
// bytes cross PCIe; every chunk is unpacked after its upload and packed before its download on the step stream.
static int step_host_stream(fs3d_world *w, const uint8_t *host_in, uint8_t *host_out, uint32_t n_steps, bool packed = false) {
    int rc = sync_all(w);
    if (rc) return rc;
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    if (!s.s_h2d) {
        FS3D_CUDA(cudaStreamCreateWithFlags(&s.s_h2d, cudaStreamNonBlocking));
        FS3D_CUDA(cudaStreamCreateWithFlags(&s.s_d2h, cudaStreamNonBlocking));
    }
    // A z-pair of planes is closed under a pass, so the grid streams through in chunks
    // of whole pairs: upload chunk k+1 | step chunk k | download chunk k-1 all overlap.
    // n_steps = 4 (fuse4_ok worlds, step index a multiple of four): chunks of whole BANDS of the four-step kernel
    // (step4_kernel.cuh).  Band b finishes local planes [8b, 8b + 8) and reads planes [8b - 1, 8b + 8], so a chunk's
    // upload runs one plane ahead of its download: planes (8 b0, 8 b1] go up, planes [8 b0, 8 b1) come down.
    const int ns = (int)n_steps;
    const bool four = ns == 4;
    const uint32_t hoff = (uint32_t)((w->step >> 1) & 1);
    const PairLayout L = pair_layout(s, hoff);
    const size_t pb = plane_bytes(w);
    const uint32_t unit_planes = four ? 2u * S4_P : 2u;                        // planes per band / per z-pair
    const uint32_t nunits = four ? (s.nzl / 2 + 1 + S4_P - 1) / S4_P : L.npairs;
    uint64_t chunk_bytes = FS3D_HOST_CHUNK_MIB << 20;
    if (const char *e = std::getenv("FS3D_HOST_CHUNK_BYTES")) chunk_bytes = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));   // tests: many small chunks
    const uint32_t pairs_per_chunk = (uint32_t)std::min<uint64_t>(nunits ? nunits : 1, std::max<uint64_t>(1, chunk_bytes / (unit_planes * pb)));
    const uint32_t nchunks = (nunits + pairs_per_chunk - 1) / pairs_per_chunk;
    while (s.ev_chunk.size() < 2 * (size_t)nchunks) {
        cudaEvent_t e;
        FS3D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s.ev_chunk.push_back(e);
    }
    const uint32_t div = w->version == 2 ? 2u : 4u;         // voxels per packed byte
    if (packed) {
        const size_t need = (size_t)pairs_per_chunk * unit_planes * pb / div;
        if (s.stage_bytes < need) {
            for (auto &p : s.d_stage) { if (p) cudaFree(p); p = nullptr; }
            s.stage_bytes = 0;
            for (auto &p : s.d_stage) FS3D_CUDA(cudaMalloc(&p, need));
            s.stage_bytes = need;
        }
        for (auto &e : s.ev_stage) if (!e) FS3D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    uint32_t *flag = reinterpret_cast<uint32_t *>(s.d_scratch + 258);
    FS3D_CUDA(cudaMemsetAsync(flag, 0, sizeof(uint32_t), s.s_main));
    launch_skip_map(w, s);
    w->force_live = true;            // the grid comes from the host: nothing is known to be static
    FS3D_CUDA(cudaEventRecord(s.ev_t0, s.s_main));
    FS3D_CUDA(cudaStreamWaitEvent(s.s_h2d, s.ev_t0, 0));
    uint8_t *src = s.buf[w->cur], *dst = s.buf[w->cur ^ 1];
    for (uint32_t c = 0; c < nchunks; ++c) {
        const uint32_t p0 = c * pairs_per_chunk, p1 = std::min(nunits, p0 + pairs_per_chunk);
        // owned local planes that go up before this chunk is stepped: lz in [lo, hi); that come down after it: [dlo, dhi)
        uint32_t lo = four ? unit_planes * p0 + 1 : L.lz_first + 2 * p0, hi = four ? unit_planes * p1 + 1 : L.lz_first + 2 * p1;
        lo = std::max(lo, 1u); hi = std::min(hi, s.nzl + 1);
        uint32_t dlo = lo, dhi = hi;
        if (four) { dlo = std::max(unit_planes * p0, 1u); dhi = std::min(unit_planes * p1, s.nzl + 1); }
        const size_t off = pb * lo, bytes = pb * (size_t)(hi - lo);
        const size_t doff = pb * dlo, dbytes = pb * (size_t)(dhi - dlo);
        const uint64_t n16 = bytes / 16, dn16 = dbytes / 16;
        if (packed) {
            // upload the packed chunk into staging slot c & 1 (free once chunk c - 2 was unpacked), unpack it into place
            if (c >= 2) FS3D_CUDA(cudaStreamWaitEvent(s.s_h2d, s.ev_stage[c & 1], 0));
            FS3D_CUDA(cudaMemcpyAsync(s.d_stage[c & 1], host_in + pb / div * (size_t)(lo - 1), bytes / div, cudaMemcpyHostToDevice, s.s_h2d));
            FS3D_CUDA(cudaEventRecord(s.ev_chunk[2 * c], s.s_h2d));
            FS3D_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_chunk[2 * c], 0));
            if (!n16) {}       // the last chunk of a four-step call may have nothing left to upload
            else if (w->version == 2) unpack4_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(reinterpret_cast<const uint2 *>(s.d_stage[c & 1]), n16, src + off);
            else unpack2_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(reinterpret_cast<const uint32_t *>(s.d_stage[c & 1]), n16, src + off);
            FS3D_CUDA(cudaGetLastError());
            FS3D_CUDA(cudaEventRecord(s.ev_stage[c & 1], s.s_main));
            w->launches++;
        } else {
            FS3D_CUDA(cudaMemcpyAsync(src + off, host_in + pb * (size_t)(lo - 1), bytes, cudaMemcpyHostToDevice, s.s_h2d));
            FS3D_CUDA(cudaEventRecord(s.ev_chunk[2 * c], s.s_h2d));
            FS3D_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_chunk[2 * c], 0));
        }
        if ((!packed || w->version == 2) && n16) {       // 2-bit codes cannot be out of range
            validate_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(src + off, n16, bad_bits(w), flag);
            FS3D_CUDA(cudaGetLastError());
            w->launches++;
        }
        rc = four ? launch_fused4(w, s, p0, p1) : launch_pairs(w, s, p0, p1, ns);
        if (rc) { w->force_live = false; return rc; }
        if (packed) {
            // pack the stepped chunk into staging slot 2 + (c & 1) (free once chunk c - 2 was downloaded)
            if (c >= 2) FS3D_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_stage[2 + (c & 1)], 0));
            if (!dn16) {}
            else if (w->version == 2) pack4_kernel<<<grid_for(dn16, s), 256, 0, s.s_main>>>(dst + doff, dn16, reinterpret_cast<uint2 *>(s.d_stage[2 + (c & 1)]));
            else pack2_kernel<<<grid_for(dn16, s), 256, 0, s.s_main>>>(dst + doff, dn16, reinterpret_cast<uint32_t *>(s.d_stage[2 + (c & 1)]));
            FS3D_CUDA(cudaGetLastError());
            w->launches++;
        }
        FS3D_CUDA(cudaEventRecord(s.ev_chunk[2 * c + 1], s.s_main));
        FS3D_CUDA(cudaStreamWaitEvent(s.s_d2h, s.ev_chunk[2 * c + 1], 0));
        if (packed) {
            if (dbytes) FS3D_CUDA(cudaMemcpyAsync(host_out + pb / div * (size_t)(dlo - 1), s.d_stage[2 + (c & 1)], dbytes / div, cudaMemcpyDeviceToHost, s.s_d2h));
            FS3D_CUDA(cudaEventRecord(s.ev_stage[2 + (c & 1)], s.s_d2h));
        } else if (dbytes) {
            FS3D_CUDA(cudaMemcpyAsync(host_out + pb * (size_t)(dlo - 1), dst + doff, dbytes, cudaMemcpyDeviceToHost, s.s_d2h));
        }
    }
    w->force_live = false;
    uint32_t bad = 0;
    FS3D_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, s.s_main));
    FS3D_CUDA(cudaStreamSynchronize(s.s_main));
    FS3D_CUDA(cudaStreamSynchronize(s.s_d2h));
    w->cur ^= 1;
    w->step += (uint64_t)ns;
    rc = touch_all_tiles(w);
    if (rc) return rc;
    if (bad) return fail(FS3D_ERR_BAD_MATERIAL, std::string("host grid: ") + bad_material_msg(w) + "; the world now holds "
                                                "undefined cells - upload or generate before stepping again");
    return FS3D_OK;
}

int fs3d_step_host(fs3d_world *w, const uint8_t *host_in, uint8_t *host_out, uint32_t n_steps) {
    if (!w || !host_in || !host_out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (n_steps != 1 && n_steps != 2 && n_steps != 4) return fail(FS3D_ERR_INVALID_ARG, "fs3d_step_host advances 1, 2 or 4 steps per call");
    if (n_steps == 2 && (w->step & 1)) return fail(FS3D_ERR_INVALID_ARG, "a fused 2-step pass must start on an even step");
    if (n_steps == 4 && (w->step & 3)) return fail(FS3D_ERR_INVALID_ARG, "a 4-step pass must start on a step index that is a multiple of four");
    if (w->slabs.size() != 1 || (w->external && w->desc.nz != w->slabs[0].nzl))
        return fail(FS3D_ERR_UNSUPPORTED, "fs3d_step_host needs a single-slab world that holds the whole grid "
                                          "(ranks use fs3d_slab_step_host_begin + fs3d_slab_step_host)");
    if (n_steps == 4 && !fuse4_ok(w)) {        // no four-step kernel for this world: two streamed two-step passes
        int rc = step_host_stream(w, host_in, host_out, 2);
        return rc ? rc : step_host_stream(w, host_out, host_out, 2);
    }
    return step_host_stream(w, host_in, host_out, n_steps);
}

int fs3d_step_host_packed(fs3d_world *w, const uint8_t *packed_in, uint8_t *packed_out, uint32_t n_steps) {
    if (!w || !packed_in || !packed_out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (n_steps != 1 && n_steps != 2 && n_steps != 4) return fail(FS3D_ERR_INVALID_ARG, "fs3d_step_host_packed advances 1, 2 or 4 steps per call");
    if (n_steps == 2 && (w->step & 1)) return fail(FS3D_ERR_INVALID_ARG, "a fused 2-step pass must start on an even step");
    if (n_steps == 4 && (w->step & 3)) return fail(FS3D_ERR_INVALID_ARG, "a 4-step pass must start on a step index that is a multiple of four");
    if (w->slabs.size() != 1 || (w->external && w->desc.nz != w->slabs[0].nzl))
        return fail(FS3D_ERR_UNSUPPORTED, "fs3d_step_host_packed needs a single-slab world that holds the whole grid");
    if (n_steps == 4 && !fuse4_ok(w)) {
        int rc = step_host_stream(w, packed_in, packed_out, 2, true);
        return rc ? rc : step_host_stream(w, packed_out, packed_out, 2, true);
    }
    return step_host_stream(w, packed_in, packed_out, n_steps, true);
}

// One rank's share of the same end-to-end step.  _begin uploads the slab's two edge planes and stores them into the
// neighbours' ghost planes over peer memory; after a barrier, fs3d_slab_step_host streams the slab through.
int fs3d_slab_step_host_begin(fs3d_world *w, const uint8_t *host_in) {
    if (!w || !host_in) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (!w->external || !w->p2p) return fail(FS3D_ERR_UNSUPPORTED, "needs a slab world with attached neighbours (fs3d_slab_ipc_attach)");
    int rc = sync_all(w);
    if (rc) return rc;
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    const size_t pb = plane_bytes(w);
    uint8_t *src = s.buf[w->cur];
    FS3D_CUDA(cudaMemcpyAsync(src + pb, host_in, pb, cudaMemcpyHostToDevice, s.s_main));
    if (s.nzl > 1)
        FS3D_CUDA(cudaMemcpyAsync(src + pb * (size_t)s.nzl, host_in + pb * (size_t)(s.nzl - 1), pb, cudaMemcpyHostToDevice, s.s_main));
    if (s.peer_lo.valid)
        FS3D_CUDA(cudaMemcpyAsync(s.peer_lo.buf[w->cur] + pb * ((size_t)s.peer_lo.nzl + 1), src + pb, pb, cudaMemcpyDefault, s.s_main));
    if (s.peer_hi.valid)
        FS3D_CUDA(cudaMemcpyAsync(s.peer_hi.buf[w->cur], src + pb * (size_t)s.nzl, pb, cudaMemcpyDefault, s.s_main));
    FS3D_CUDA(cudaStreamSynchronize(s.s_main));
    return FS3D_OK;
}

int fs3d_slab_step_host(fs3d_world *w, const uint8_t *host_in, uint8_t *host_out, uint32_t n_steps) {
    if (!w || !host_in || !host_out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (n_steps != 1 && n_steps != 2) return fail(FS3D_ERR_INVALID_ARG, "fs3d_slab_step_host advances 1 or 2 steps per call");
    if (n_steps == 2 && (w->step & 1)) return fail(FS3D_ERR_INVALID_ARG, "a fused 2-step pass must start on an even step");
    if (!w->external || !w->p2p) return fail(FS3D_ERR_UNSUPPORTED, "needs a slab world with attached neighbours (fs3d_slab_ipc_attach)");
    int rc = step_host_stream(w, host_in, host_out, n_steps);
    w->ghosts_stale = true;     // the new front buffer's ghost planes were not exchanged
    return rc;
}

// ---- fused halo push between ranks: CUDA IPC plumbing ---------------------------------------------
struct IpcBlob {
    uint32_t magic, nzl, z0, cur;
    cudaIpcMemHandle_t buf[2];
    cudaIpcMemHandle_t flags;
};

int fs3d_slab_ipc_export(fs3d_world *w, void *blob, uint64_t blob_bytes) {
    if (!w || !blob) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (w->slabs.size() != 1 || !w->external) return fail(FS3D_ERR_UNSUPPORTED, "needs a world made by fs3d_create_slab");
    if (blob_bytes < sizeof(IpcBlob)) return fail(FS3D_ERR_INVALID_ARG, "blob too small (need FS3D_IPC_BLOB_BYTES)");
    Slab &s = w->slabs[0];
    if (s.vmm) return fail(FS3D_ERR_UNSUPPORTED, "FS3D_FLAG_EXPORTABLE buffers are exported with fs3d_volume_export_fd; CUDA IPC handles "
                                                 "(the fused halo push between processes) need ordinary allocations");
    FS3D_CUDA(cudaSetDevice(s.device));
    IpcBlob b{};
    b.magic = 0xF53D1BCu; b.nzl = s.nzl; b.z0 = s.z0; b.cur = (uint32_t)w->cur;
    for (int i = 0; i < 2; ++i) FS3D_CUDA(cudaIpcGetMemHandle(&b.buf[i], s.buf[i]));
    FS3D_CUDA(cudaIpcGetMemHandle(&b.flags, s.d_flags));
    std::memset(blob, 0, (size_t)blob_bytes);
    std::memcpy(blob, &b, sizeof(b));
    return FS3D_OK;
}

static int open_peer(Slab::Peer &pr, const void *blob, uint32_t expect_z, bool expect_end, int my_cur) {
    IpcBlob b;
    std::memcpy(&b, blob, sizeof(b));
    if (b.magic != 0xF53D1BCu) return fail(FS3D_ERR_INVALID_ARG, "not an fs3d IPC blob");
    if ((expect_end ? b.z0 + b.nzl : b.z0) != expect_z) return fail(FS3D_ERR_INVALID_ARG, "IPC blob is not the adjacent slab");
    if ((int)b.cur != my_cur) return fail(FS3D_ERR_INVALID_ARG, "neighbour's buffer parity differs (edit worlds collectively)");
    for (int i = 0; i < 2; ++i)
        FS3D_CUDA(cudaIpcOpenMemHandle((void **)&pr.buf[i], b.buf[i], cudaIpcMemLazyEnablePeerAccess));
    FS3D_CUDA(cudaIpcOpenMemHandle((void **)&pr.flags, b.flags, cudaIpcMemLazyEnablePeerAccess));
    pr.nzl = b.nzl;
    pr.valid = true;
    pr.ipc = true;
    return FS3D_OK;
}

int fs3d_slab_ipc_attach(fs3d_world *w, const void *lower_blob, const void *upper_blob) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (w->slabs.size() != 1 || !w->external) return fail(FS3D_ERR_UNSUPPORTED, "needs a world made by fs3d_create_slab");
    if (w->p2p) return fail(FS3D_ERR_INVALID_ARG, "neighbours already attached");
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    if ((lower_blob != nullptr) != (s.z0 > 0)) return fail(FS3D_ERR_INVALID_ARG, "lower neighbour blob must be given iff z_begin > 0");
    if ((upper_blob != nullptr) != (s.z0 + s.nzl < w->desc.nz)) return fail(FS3D_ERR_INVALID_ARG, "upper neighbour blob must be given iff z_end < nz");
    if (lower_blob) { int rc = open_peer(s.peer_lo, lower_blob, s.z0, true, w->cur); if (rc) return rc; }
    if (upper_blob) { int rc = open_peer(s.peer_hi, upper_blob, s.z0 + s.nzl, false, w->cur); if (rc) return rc; }
    FS3D_CUDA(cudaMemset(s.d_flags, 0, 2 * sizeof(unsigned long long)));
    w->wait_target = 0;
    w->p2p = true;
    return FS3D_OK;
}

// The same wiring for slab worlds that live in ONE process (one host thread per GPU, or several slabs on one GPU in
// the tests): plain pointers instead of CUDA IPC handles.
static int attach_local_peer(Slab &s, Slab::Peer &pr, fs3d_world *nb, uint32_t expect_z, bool expect_end, int my_cur) {
    if (!nb->external || nb->slabs.size() != 1) return fail(FS3D_ERR_UNSUPPORTED, "neighbour must be a world made by fs3d_create_slab");
    Slab &d = nb->slabs[0];
    if ((expect_end ? d.z0 + d.nzl : d.z0) != expect_z) return fail(FS3D_ERR_INVALID_ARG, "neighbour world is not the adjacent slab");
    if (nb->cur != my_cur) return fail(FS3D_ERR_INVALID_ARG, "neighbour's buffer parity differs (edit worlds collectively)");
    if (d.device != s.device) {
        int can = 0;
        FS3D_CUDA(cudaDeviceCanAccessPeer(&can, s.device, d.device));
        if (!can) return fail(FS3D_ERR_UNSUPPORTED, "no peer access to the neighbour's device");
        cudaError_t e = cudaDeviceEnablePeerAccess(d.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FS3D_CUDA(e);
        cudaGetLastError();
    }
    pr.buf[0] = d.buf[0]; pr.buf[1] = d.buf[1]; pr.flags = d.d_flags; pr.nzl = d.nzl;
    pr.valid = true; pr.ipc = false;
    return FS3D_OK;
}

int fs3d_slab_attach_local(fs3d_world *w, fs3d_world *lower, fs3d_world *upper) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (w->slabs.size() != 1 || !w->external) return fail(FS3D_ERR_UNSUPPORTED, "needs a world made by fs3d_create_slab");
    if (w->p2p) return fail(FS3D_ERR_INVALID_ARG, "neighbours already attached");
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    if ((lower != nullptr) != (s.z0 > 0)) return fail(FS3D_ERR_INVALID_ARG, "lower neighbour must be given iff z_begin > 0");
    if ((upper != nullptr) != (s.z0 + s.nzl < w->desc.nz)) return fail(FS3D_ERR_INVALID_ARG, "upper neighbour must be given iff z_end < nz");
    if (lower) { int rc = attach_local_peer(s, s.peer_lo, lower, s.z0, true, w->cur); if (rc) return rc; }
    if (upper) { int rc = attach_local_peer(s, s.peer_hi, upper, s.z0 + s.nzl, false, w->cur); if (rc) { s.peer_lo = Slab::Peer(); return rc; } }
    FS3D_CUDA(cudaMemset(s.d_flags, 0, 8 * sizeof(unsigned long long)));
    w->wait_target = 0;
    w->p2p = true;
    return FS3D_OK;
}

/* Four-step passes on a multi-process slab world need every rank's slab to support them (even internal boundaries,
 * at least four planes, rows of 1024 or 2048 voxels, schedule version 1, no skipping): ranks ask fs3d_slab_can_fuse4,
 * combine the answers (all-reduce AND) and call fs3d_slab_allow_fuse4 with the result before the first fs3d_step. */
int fs3d_slab_can_fuse4(fs3d_world *w) {
    if (!w || w->slabs.size() != 1) return 0;
    return slab_fuse4_capable(w, w->slabs[0]) ? 1 : 0;
}

int fs3d_slab_allow_fuse4(fs3d_world *w, int allow) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (allow && !fs3d_slab_can_fuse4(w)) return fail(FS3D_ERR_UNSUPPORTED, "this slab cannot run four-step passes (fs3d_slab_can_fuse4)");
    w->fuse4_allowed = allow != 0;
    return FS3D_OK;
}

/* What the halo waits cost so far (and resets the counters): out[0] = ns spent waiting for a neighbour's arrival counter,
 * summed over the warps that really blocked; out[1] = the longest single wait; out[2] = number of blocking waits. */
int fs3d_push_wait_stats(fs3d_world *w, uint64_t out[3]) {
    if (!w || !out) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    out[0] = out[1] = out[2] = 0;
    if (!w->p2p) return FS3D_OK;
    int rc = sync_all(w);
    if (rc) return rc;
    for (auto &s : w->slabs) {
        unsigned long long v[3] = {0, 0, 0};
        FS3D_CUDA(cudaSetDevice(s.device));
        FS3D_CUDA(cudaMemcpy(v, s.d_flags + 3, sizeof(v), cudaMemcpyDeviceToHost));
        FS3D_CUDA(cudaMemset(s.d_flags + 3, 0, sizeof(v)));
        out[0] += v[0]; out[1] = std::max<uint64_t>(out[1], v[1]); out[2] += v[2];
    }
    return FS3D_OK;
}

/* Copies this slab's two edge planes of the FRONT buffer into the neighbours' ghost planes (after
 * generate / upload / edits).  Every rank calls it, then all ranks barrier before the next step. */
int fs3d_slab_push_halos(fs3d_world *w) {
    if (!w) return fail(FS3D_ERR_INVALID_ARG, "world is NULL");
    if (!w->p2p) return fail(FS3D_ERR_UNSUPPORTED, "call fs3d_slab_ipc_attach first");
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    const size_t pb = plane_bytes(w);
    FS3D_CUDA(cudaStreamSynchronize(s.s_main));
    if (s.peer_lo.valid)
        FS3D_CUDA(cudaMemcpyAsync(s.peer_lo.buf[w->cur] + pb * ((size_t)s.peer_lo.nzl + 1), s.buf[w->cur] + pb, pb, cudaMemcpyDefault, s.s_main));
    if (s.peer_hi.valid)
        FS3D_CUDA(cudaMemcpyAsync(s.peer_hi.buf[w->cur], s.buf[w->cur] + pb * (size_t)s.nzl, pb, cudaMemcpyDefault, s.s_main));
    FS3D_CUDA(cudaStreamSynchronize(s.s_main));
    w->ghosts_stale = false;
    w->far_valid = false;            // only the near planes were copied: the next four-step pass refreshes the far ones
    return FS3D_OK;
}

// ---- fused multi-rank ray-march: every rank's kernel stores into the compositor's frame over NVLink ----
struct FrameBlob {
    uint32_t magic, width, height, nslots;
    cudaIpcMemHandle_t mem;
};

int fs3d_frame_export(fs3d_world *w, uint32_t width, uint32_t height, uint32_t n_slots, void *blob, uint64_t blob_bytes) {
    if (!w || !blob) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    if (width == 0 || height == 0 || width > 16384 || height > 16384 || n_slots == 0 || n_slots > 64)
        return fail(FS3D_ERR_INVALID_ARG, "bad frame size");
    if (blob_bytes < sizeof(FrameBlob)) return fail(FS3D_ERR_INVALID_ARG, "blob too small (need FS3D_IPC_BLOB_BYTES)");
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    int rc = sync_all(w);
    if (rc) return rc;
    s.frame.width = 0;                      // force a fresh (cleared) frame
    rc = ensure_frame(s, width, height, n_slots);
    if (rc) return rc;
    FrameBlob b{};
    b.magic = 0xF53DF4A3u; b.width = width; b.height = height; b.nslots = n_slots;
    FS3D_CUDA(cudaIpcGetMemHandle(&b.mem, s.frame.base));
    std::memset(blob, 0, (size_t)blob_bytes);
    std::memcpy(blob, &b, sizeof(b));
    return FS3D_OK;
}

int fs3d_frame_attach(fs3d_world *w, const void *blob, uint32_t slot) {
    if (!w || !blob) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    FrameBlob b;
    std::memcpy(&b, blob, sizeof(b));
    if (b.magic != 0xF53DF4A3u) return fail(FS3D_ERR_INVALID_ARG, "not an fs3d frame blob");
    if (slot >= b.nslots) return fail(FS3D_ERR_OUT_OF_RANGE, "slot outside the frame");
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    if (s.frame.owner) {   // the compositor attaches to its own frame: only the slot changes
        if (b.width != s.frame.width || b.height != s.frame.height || b.nslots != s.frame.nslots)
            return fail(FS3D_ERR_INVALID_ARG, "blob does not describe this world's frame");
        s.frame.slot = slot;
        return FS3D_OK;
    }
    if (s.frame.base) { if (s.frame.ipc) cudaIpcCloseMemHandle(s.frame.base); s.frame = Slab::Frame(); }
    FS3D_CUDA(cudaIpcOpenMemHandle((void **)&s.frame.base, b.mem, cudaIpcMemLazyEnablePeerAccess));
    s.frame.ipc = true;
    s.frame.width = b.width; s.frame.height = b.height; s.frame.nslots = b.nslots; s.frame.slot = slot;
    return FS3D_OK;
}

int fs3d_frame_attach_local(fs3d_world *w, fs3d_world *owner, uint32_t slot) {
    if (!w || !owner) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    Slab &o = owner->slabs[0];
    if (!o.frame.base || !o.frame.owner) return fail(FS3D_ERR_UNSUPPORTED, "owner has no frame (call fs3d_frame_export on it first)");
    if (slot >= o.frame.nslots) return fail(FS3D_ERR_OUT_OF_RANGE, "slot outside the frame");
    Slab &s = w->slabs[0];
    FS3D_CUDA(cudaSetDevice(s.device));
    if (w == owner) { s.frame.slot = slot; return FS3D_OK; }
    if (s.device != o.device) {
        int can = 0;
        FS3D_CUDA(cudaDeviceCanAccessPeer(&can, s.device, o.device));
        if (!can) return fail(FS3D_ERR_UNSUPPORTED, "no peer access to the frame owner's device");
        cudaError_t e = cudaDeviceEnablePeerAccess(o.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FS3D_CUDA(e);
        cudaGetLastError();
    }
    if (s.frame.base) { if (s.frame.owner) cudaFree(s.frame.base); else if (s.frame.ipc) cudaIpcCloseMemHandle(s.frame.base); }
    if (s.frame.d_rgba) cudaFree(s.frame.d_rgba);
    s.frame = Slab::Frame();
    s.frame.base = o.frame.base;            // borrowed: the owner frees it
    s.frame.width = o.frame.width; s.frame.height = o.frame.height; s.frame.nslots = o.frame.nslots; s.frame.slot = slot;
    return FS3D_OK;
}

int fs3d_raymarch_to_frame(fs3d_world *w, const fs3d_camera *cam, uint32_t mode) {
    if (!w || !cam) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    Slab &s = w->slabs[0];
    if (!s.frame.base) return fail(FS3D_ERR_UNSUPPORTED, "call fs3d_frame_export / fs3d_frame_attach first");
    const size_t npix = (size_t)s.frame.width * s.frame.height;
    return raymarch_world(w, cam, s.frame.width, s.frame.height, mode, nullptr, nullptr, s.frame.base + npix * s.frame.slot);
}

int fs3d_frame_resolve(fs3d_world *w, uint8_t *host_rgba8) {
    if (!w || !host_rgba8) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    Slab &s = w->slabs[0];
    if (!s.frame.base || !s.frame.owner) return fail(FS3D_ERR_UNSUPPORTED, "only the rank that called fs3d_frame_export resolves");
    FS3D_CUDA(cudaSetDevice(s.device));
    const size_t npix = (size_t)s.frame.width * s.frame.height;
    frame_resolve_kernel<<<grid_for(npix, s), 256, 0, s.s_main>>>(s.frame.base, s.frame.nslots, npix, s.frame.d_rgba, nullptr);
    FS3D_CUDA(cudaGetLastError());
    w->launches++;
    FS3D_CUDA(cudaMemcpyAsync(host_rgba8, s.frame.d_rgba, npix * 4, cudaMemcpyDeviceToHost, s.s_main));
    FS3D_CUDA(cudaStreamSynchronize(s.s_main));
    return FS3D_OK;
}

// ---- checkpoint: the planes this world holds + step index + seed, 2 bits per voxel ----------------
struct CkptHeader {
    char     magic[8];            // "FS3DCKPT"
    uint32_t format_version;      // FS3D_CKPT_VERSION
    uint32_t schedule_version;    // FS3D_SCHEDULE_VERSION the state was produced under
    uint32_t nx, ny, nz;          // global grid
    uint32_t z_begin, z_end;      // planes in this file
    uint32_t encoding;            // 1 = 2 bits per voxel (schedule version 1), 2 = 4 bits per voxel (version 2), x fastest
    uint64_t step, seed;
    uint64_t digest;              // fs3d_digest of these planes (global indices): checked on load
    uint64_t payload_bytes;
    uint64_t reserved;
};
static_assert(sizeof(CkptHeader) == FS3D_CKPT_HEADER_BYTES, "checkpoint header layout is part of the file format");

namespace {
struct File {          // closes on every return path
    FILE *f = nullptr;
    ~File() { if (f) std::fclose(f); }
    int close() { const int r = f ? std::fclose(f) : 0; f = nullptr; return r; }
};
struct DevBuf {        // frees on every return path (on whatever device is current: set it before the scope ends)
    uint32_t *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

int fs3d_save(fs3d_world *w, const char *path) {
    if (!w || !path) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    uint64_t dg = 0;
    int rc = fs3d_digest(w, &dg);          // also synchronises
    if (rc) return rc;
    const size_t pb = plane_bytes(w);
    const uint32_t zb = w->slabs.front().z0, ze = w->slabs.back().z0 + w->slabs.back().nzl;
    CkptHeader h{};
    std::memcpy(h.magic, "FS3DCKPT", 8);
    // schedule version 1: 2 bits per voxel (encoding 1, div = 4); version 2: 4 bits per voxel (encoding 2, div = 2)
    const uint32_t div = w->version == 2 ? 2u : 4u;
    h.format_version = FS3D_CKPT_VERSION; h.schedule_version = (uint32_t)w->version;
    h.nx = w->desc.nx; h.ny = w->desc.ny; h.nz = w->desc.nz; h.z_begin = zb; h.z_end = ze;
    h.encoding = w->version == 2 ? 2u : 1u; h.step = w->step; h.seed = w->desc.seed; h.digest = dg;
    h.payload_bytes = pb / div * (uint64_t)(ze - zb);    // nx % 32 == 0, so a plane packs to whole bytes
    File out;
    out.f = std::fopen(path, "wb");
    if (!out.f) return fail(FS3D_ERR_IO, std::string("cannot open ") + path + " for writing");
    bool io_ok = std::fwrite(&h, sizeof(h), 1, out.f) == 1;
    // pack on the device, stream to the file in chunks of whole planes (<= ~64 MiB packed)
    const uint32_t planes_per_chunk = (uint32_t)std::max<uint64_t>(1, (64ull << 20) / (pb / div));
    std::vector<uint8_t> host((size_t)std::min<uint64_t>(planes_per_chunk, ze - zb) * (pb / div));
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        DevBuf packed;
        const uint32_t cp = std::min(planes_per_chunk, s.nzl);
        FS3D_CUDA(cudaMalloc(&packed.p, (size_t)cp * (pb / div)));
        for (uint32_t z = 0; z < s.nzl && io_ok; z += cp) {
            const uint32_t n = std::min(cp, s.nzl - z);
            const uint64_t n16 = pb * n / 16, nbytes = n16 * (16 / div);
            if (w->version == 2) pack4_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(owned_ptr(w, s, w->cur) + pb * z, n16, reinterpret_cast<uint2 *>(packed.p));
            else pack2_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(owned_ptr(w, s, w->cur) + pb * z, n16, packed.p);
            FS3D_CUDA(cudaGetLastError());
            w->launches++;
            FS3D_CUDA(cudaMemcpyAsync(host.data(), packed.p, nbytes, cudaMemcpyDeviceToHost, s.s_main));
            FS3D_CUDA(cudaStreamSynchronize(s.s_main));
            io_ok = std::fwrite(host.data(), 1, nbytes, out.f) == nbytes;
        }
    }
    io_ok = (out.close() == 0) && io_ok;
    if (!io_ok) return fail(FS3D_ERR_IO, std::string("short write to ") + path);
    return FS3D_OK;
}

int fs3d_load(fs3d_world *w, const char *path) {
    if (!w || !path) return fail(FS3D_ERR_INVALID_ARG, "NULL argument");
    File in;
    in.f = std::fopen(path, "rb");
    if (!in.f) return fail(FS3D_ERR_IO, std::string("cannot open ") + path);
    CkptHeader h{};
    if (std::fread(&h, sizeof(h), 1, in.f) != 1 || std::memcmp(h.magic, "FS3DCKPT", 8) != 0)
        return fail(FS3D_ERR_IO, std::string(path) + " is not an fs3d checkpoint");
    const size_t pb = plane_bytes(w);
    const uint32_t zb = w->slabs.front().z0, ze = w->slabs.back().z0 + w->slabs.back().nzl;
    const uint32_t div = w->version == 2 ? 2u : 4u;
    if (h.format_version != FS3D_CKPT_VERSION || (h.encoding != 1 && h.encoding != 2))
        return fail(FS3D_ERR_INVALID_ARG, "unknown checkpoint format version / encoding");
    if (h.schedule_version != (uint32_t)w->version || h.encoding != (w->version == 2 ? 2u : 1u))
        return fail(FS3D_ERR_INVALID_ARG, "checkpoint was written under another schedule version");
    if (h.nx != w->desc.nx || h.ny != w->desc.ny || h.nz != w->desc.nz)
        return fail(FS3D_ERR_INVALID_ARG, "checkpoint grid differs from the world's");
    if (h.z_begin != zb || h.z_end != ze) return fail(FS3D_ERR_INVALID_ARG, "checkpoint holds other z-planes than this world");
    if (h.payload_bytes != pb / div * (uint64_t)(ze - zb)) return fail(FS3D_ERR_INVALID_ARG, "checkpoint payload size is inconsistent");
    int rc = sync_all(w);
    if (rc) return rc;
    // unpack into the BACK buffer, verify the digest there, then flip: a bad file leaves the world untouched
    const int back = w->cur ^ 1;
    const uint32_t planes_per_chunk = (uint32_t)std::max<uint64_t>(1, (64ull << 20) / (pb / div));
    std::vector<uint8_t> host((size_t)std::min<uint64_t>(planes_per_chunk, ze - zb) * (pb / div));
    uint64_t sum = 0;
    for (auto &s : w->slabs) {
        FS3D_CUDA(cudaSetDevice(s.device));
        DevBuf packed;
        const uint32_t cp = std::min(planes_per_chunk, s.nzl);
        FS3D_CUDA(cudaMalloc(&packed.p, (size_t)cp * (pb / div)));
        for (uint32_t z = 0; z < s.nzl; z += cp) {
            const uint32_t n = std::min(cp, s.nzl - z);
            const uint64_t n16 = pb * n / 16, nbytes = n16 * (16 / div);
            if (std::fread(host.data(), 1, nbytes, in.f) != nbytes) return fail(FS3D_ERR_IO, std::string(path) + " is truncated");
            FS3D_CUDA(cudaMemcpyAsync(packed.p, host.data(), nbytes, cudaMemcpyHostToDevice, s.s_main));
            if (w->version == 2) unpack4_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(reinterpret_cast<const uint2 *>(packed.p), n16, owned_ptr(w, s, back) + pb * z);
            else unpack2_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(packed.p, n16, owned_ptr(w, s, back) + pb * z);
            FS3D_CUDA(cudaGetLastError());
            w->launches++;
            FS3D_CUDA(cudaStreamSynchronize(s.s_main));      // `host` is reused by the next chunk
        }
        unsigned long long *d = s.d_scratch + 256, hsum = 0;
        FS3D_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), s.s_main));
        const uint64_t n16 = pb * s.nzl / 16;
        digest_kernel<<<grid_for(n16, s), 256, 0, s.s_main>>>(owned_ptr(w, s, back), n16, (uint64_t)s.z0 * pb, d);
        FS3D_CUDA(cudaGetLastError());
        FS3D_CUDA(cudaMemcpyAsync(&hsum, d, sizeof(hsum), cudaMemcpyDeviceToHost, s.s_main));
        FS3D_CUDA(cudaStreamSynchronize(s.s_main));
        sum += hsum;
    }
    if (sum != h.digest) return fail(FS3D_ERR_IO, std::string(path) + ": digest mismatch (corrupt checkpoint); world unchanged");
    w->cur = back;
    w->step = h.step;
    w->desc.seed = h.seed;
    return refresh_ghosts(w);   // also marks every activity tile live; ranks of a p2p world call fs3d_slab_push_halos next
}

}  // extern "C"
