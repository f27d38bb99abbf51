/*
 * fs3d.h — C ABI of libfs3d: the B200-native voxel-world step for FallingSand3D.
 *
 * The reference engine (/root/reference) has no plugin/FFI boundary and, in the surveyed
 * snapshot, no voxel world at all.  Each entry point below therefore cites
 * the reference *seam* it plugs into rather than a function it replaces:
 *
 *   - frame loop, where step() is called once per frame between handleEvents() and draw():
 *       src/engine/engine.cpp:59-70  (VulkanEngine::run)
 *   - material binding builder, the only place a volume could be handed to a ray-march shader:
 *       src/engine/rendering/materials.cpp:26-59 (addDataBinding/addImageBinding),
 *       src/engine/rendering/materials.cpp:388-418 (writeBuffer/writeImage/finalize)
 *   - full-screen ray-march camera + shading model:
 *       shaders/fs_raymarch.vert:30-37, shaders/fs_raymarch.frag:38-81,
 *       quad UVs src/engine/rendering/renderer.cpp:1253-1267,
 *       camera defaults src/engine/rendering/renderer.h:148-149
 *   - error convention (log "ERROR: ..." then throw std::runtime_error), mirrored by the C++
 *     wrapper include/fs3d.hpp:  src/util/debug.cpp:23-27
 *   - 256-entry colour palette indexed by uint8 material:
 *       src/engine/rendering/renderer.cpp:136-393 (colors[], unused by the reference)
 *
 * Conventions: plain C types only; every function returns 0 (FS3D_OK) or a negative error code
 * and never throws; fs3d_last_error() returns a thread-local message.  A world handle is not
 * thread-safe (the reference is single-threaded): use one host thread per handle.  The library
 * owns all device memory; host pointers passed in are borrowed for the duration of the call.
 * Kernels run asynchronously on library-owned streams; fs3d_sync() blocks.
 * There is NO CPU fallback: every call that computes fails with FS3D_ERR_CUDA without a GPU.
 *
 * The update rule is specified in SCHEDULE.md (schedule version FS3D_SCHEDULE_VERSION).
 */
#ifndef FS3D_H
#define FS3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS3D_SCHEDULE_VERSION 1      /* the default rule set; worlds made with FS3D_FLAG_MATERIALS8 run version 2 */

/* materials (SCHEDULE.md §1) */
#define FS3D_EMPTY 0
#define FS3D_SAND  1
#define FS3D_WATER 2
#define FS3D_STONE 3
/* schedule version 2 only (SCHEDULE.md §7; FS3D_FLAG_MATERIALS8) */
#define FS3D_GAS    4   /* lighter than EMPTY: rises, spreads sideways */
#define FS3D_OIL    5   /* liquid, floats on WATER */
#define FS3D_HONEY  6   /* viscous liquid between WATER and SAND: spreads sideways a quarter as often */
#define FS3D_GRAVEL 7   /* granular, denser than SAND, never slides diagonally */

/* error codes */
#define FS3D_OK                 0
#define FS3D_ERR_INVALID_ARG   -1
#define FS3D_ERR_BAD_DIMS      -2
#define FS3D_ERR_BAD_MATERIAL  -3
#define FS3D_ERR_OUT_OF_RANGE  -4
#define FS3D_ERR_CUDA          -5
#define FS3D_ERR_OOM           -6
#define FS3D_ERR_UNSUPPORTED   -7
#define FS3D_ERR_IO            -8

/* fs3d_desc.flags */
#define FS3D_FLAG_SKIP_SETTLED  1u   /* settled-tile skipping (bit-exact; SCHEDULE.md §4) */
#define FS3D_FLAG_NO_FUSE       2u   /* one kernel pass per step (default: steps 2k, 2k+1 fuse into one pass) */
#define FS3D_FLAG_NO_PEER_PUSH  4u   /* n_gpus > 1: exchange halos with peer copies on a side stream instead of
                                        storing them from inside the step kernels (the default with peer access) */
#define FS3D_FLAG_PEER_PUSH_SHARED_DEVICE 8u /* n_gpus > 1 with several slabs on ONE device: use the fused halo push
                                        anyway (default there: peer copies).  Lets a one-GPU box run the PUSH kernels. */

#define FS3D_FLAG_EXPORTABLE    16u  /* allocate the slab buffers through the driver's virtual-memory-management API so that
                                        fs3d_volume_export_fd can hand them to another API or process as file descriptors */

#define FS3D_FLAG_MATERIALS8    32u  /* schedule version 2: eight materials (codes 0-7) on three bit-planes — a separate set
                                        of kernels, so worlds that hold only codes 0-3 pay nothing for it.  A version-2
                                        world that holds only codes 0-3 evolves exactly like a version-1 world. */

#define FS3D_FLAG_NO_FUSE4      64u  /* never fuse FOUR steps into one pass (default: single-GPU worlds with rows of 1024 or
                                        2048 voxels do, when four steps remain and the step index is a multiple of four) */

/* scene ids for fs3d_generate (SCHEDULE.md §5) */
#define FS3D_SCENE_EMPTY        0
#define FS3D_SCENE_SAND_BLOCK   1
#define FS3D_SCENE_MIXED        2
#define FS3D_SCENE_RANDOM       3
#define FS3D_SCENE_MIXED_NOISE  4
#define FS3D_SCENE_RANDOM8      5   /* FS3D_FLAG_MATERIALS8 worlds only (SCHEDULE.md §7) */
#define FS3D_SCENE_MIXED8       6

typedef struct fs3d_world fs3d_world;

typedef struct {
    uint32_t nx, ny, nz;      /* global grid; nx % 32 == 0, nx <= 4096 */
    uint64_t seed;            /* coin seed (SCHEDULE.md §3) */
    int32_t  n_gpus;          /* z-slabs, one per device, driven by this process; 0 or 1 = current device */
    const int32_t *devices;   /* n_gpus CUDA device ordinals, or NULL for 0..n_gpus-1 */
    uint32_t flags;           /* FS3D_FLAG_* */
} fs3d_desc;

/* A borrowed view of one slab of the current (front) buffer — the volume hand-off to a renderer.
 * Valid until the next fs3d_step / upload / generate / destroy on the world. */
typedef struct {
    const uint8_t *dev_ptr;   /* device pointer to cell (0, 0, z0) */
    int32_t  device;          /* CUDA ordinal that owns dev_ptr */
    uint32_t nx, ny;
    uint32_t z0, z1;          /* global planes [z0, z1) held by this slab */
    uint64_t pitch_y, pitch_z;/* bytes between rows / planes */
    uint64_t step;            /* step index the view corresponds to */
} fs3d_view;

/* The two buffers of one slab as POSIX file descriptors (hand-off of the volume to a real Vulkan
 * consumer).  The world is double-buffered: `front` says which of the two holds the current step; after every pass
 * the other one does (fs3d_step(w, n) makes ceil(n / 2) passes when it starts on an even step).  Each descriptor
 * refers to an allocation of alloc_bytes; cell (0, 0, z0) sits first_cell_offset bytes into it, then pitch_y / pitch_z
 * as in fs3d_view.  Import with
 *   - Vulkan:  VkImportMemoryFdInfoKHR{handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT, fd} chained into
 *              VkMemoryAllocateInfo{allocationSize = alloc_bytes}, bound to a VkBuffer created with
 *              VkExternalMemoryBufferCreateInfo (INTEGRATION.md has the engine-side patch), or
 *   - CUDA in another process:  cuMemImportFromShareableHandle + cuMemAddressReserve / cuMemMap / cuMemSetAccess
 *              (tests/vmm_import_child.py does exactly this and reproduces the digest).
 * The caller owns the descriptors (close() them; an import takes its own reference).  Synchronise (fs3d_sync) before
 * the consumer reads; the library does not know about foreign readers. */
typedef struct {
    int32_t  fd[2];
    uint64_t alloc_bytes;
    uint64_t first_cell_offset;
    uint32_t front;           /* index into fd[] of the buffer holding step `step` */
    int32_t  device;
    uint32_t nx, ny;
    uint32_t z0, z1;
    uint64_t pitch_y, pitch_z;
    uint64_t step;
} fs3d_export;

/* Camera of shaders/fs_raymarch.{vert,frag}: origin + aspect; yaw is the engine's camRot.y
 * (renderer.cpp:460-467), which fs_raymarch itself ignores (0 reproduces the shader). */
typedef struct {
    float pos[3];
    float yaw_deg;
    float aspect;             /* reference: 1700/900 (materials.cpp:540) */
} fs3d_camera;

/* fs3d_raymarch modes */
#define FS3D_RM_SDF_SPHERE 0   /* fs_raymarch.frag as shipped: analytic sphere r=0.5 at origin */
#define FS3D_RM_VOXELS     1   /* same camera/light, voxel DDA through the grid + palette */
#define FS3D_RM_SRGB       16  /* OR-able: sRGB-encode like the reference's B8G8R8A8_SRGB swapchain */
#define FS3D_RM_BRICKS     32  /* OR-able: always (re)build the 8^3-brick occupancy map and skip empty bricks */
#define FS3D_RM_NO_BRICKS  64  /* OR-able: never; default: decided from the previous frame's steps per ray.  The image
                                  is identical either way (the jumps are exact). */

/* Device pointers for an external (one-process-per-GPU) halo exchange; see fs3d_create_slab. */
typedef struct {
    uint8_t *send_lo, *send_hi;  /* first / last owned plane of the buffer being exchanged */
    uint8_t *recv_lo, *recv_hi;  /* ghost planes below / above the slab in the same buffer */
    uint64_t plane_bytes;        /* nx * ny */
    void    *stream;             /* cudaStream_t the step kernels run on */
} fs3d_halo;

/* ---- lifetime ---- */
int  fs3d_create(const fs3d_desc *desc, fs3d_world **out);
/* One rank's slab [z_begin, z_end) of a desc->nz-plane world on the CURRENT device; the caller
 * exchanges halos (fs3d_slab_* below).  desc->n_gpus/devices are ignored. */
int  fs3d_create_slab(const fs3d_desc *desc, uint32_t z_begin, uint32_t z_end, fs3d_world **out);
void fs3d_destroy(fs3d_world *w);

/* ---- cell access (global coordinates) ---- */
int  fs3d_set_cell(fs3d_world *w, uint32_t x, uint32_t y, uint32_t z, uint8_t m);
int  fs3d_get_cell(fs3d_world *w, uint32_t x, uint32_t y, uint32_t z, uint8_t *m);
int  fs3d_fill_box(fs3d_world *w, const uint32_t lo[3], const uint32_t hi[3], uint8_t m); /* half-open */
/* Brush for paint / erase input (the engine's key and mouse events, src/engine/window.cpp:34-107): every cell
 * whose centre lies within `radius` cells of (cx, cy, cz) becomes m (FS3D_EMPTY erases); the centre may lie
 * outside the grid; with only_empty != 0 cells that already hold a material are kept. */
int  fs3d_paint_sphere(fs3d_world *w, int32_t cx, int32_t cy, int32_t cz, uint32_t radius, uint8_t m, int only_empty);
int  fs3d_generate(fs3d_world *w, int scene_id, uint64_t seed);
int  fs3d_upload(fs3d_world *w, const uint8_t *host);    /* host: the planes this world holds, x fastest */
int  fs3d_download(fs3d_world *w, uint8_t *host);

/* ---- stepping ---- */
int  fs3d_step(fs3d_world *w, uint32_t n_steps);          /* asynchronous */
int  fs3d_sync(fs3d_world *w);
int  fs3d_step_index(fs3d_world *w, uint64_t *out);
int  fs3d_kernel_launches(fs3d_world *w, uint64_t *out);   /* kernels this world has launched since it was created */
/* Runs n_steps and returns the device time between CUDA events recorded on the step stream
 * (max over slabs).  kernel_launches, if non-NULL, receives the number of kernels launched. */
int  fs3d_step_timed(fs3d_world *w, uint32_t n_steps, float *ms, uint64_t *kernel_launches);

/* End-to-end step for a HOST-resident grid: equivalent to fs3d_upload(host_in); fs3d_step(n);
 * fs3d_download(host_out) with n = 1, 2 (needs an even step index) or 4 (a step index that is a
 * multiple of four), but the grid streams through the GPU in chunks so the H2D copy, the kernels
 * and the D2H copy overlap (use pinned host memory): chunks of whole z-pairs for n = 1, 2, of whole
 * bands of the four-step kernel for n = 4 (worlds without one run two streamed 2-step passes).
 * The call is bound by PCIe, so its cost barely depends on n: ask for as many steps as you need.
 * host_in may equal host_out.  Single-slab worlds only. */
int  fs3d_step_host(fs3d_world *w, const uint8_t *host_in, uint8_t *host_out, uint32_t n_steps);

/* The same for a grid the host keeps PACKED in the checkpoint encoding (the payload of fs3d_save without its header:
 * 2 bits per voxel, 4 for FS3D_FLAG_MATERIALS8 worlds): a quarter (half) of the bytes cross PCIe — which is what bounds
 * fs3d_step_host — and a quarter of the host memory holds the grid.  Chunks are unpacked after their upload and packed
 * before their download on the step stream.  packed_in may equal packed_out. */
int  fs3d_step_host_packed(fs3d_world *w, const uint8_t *packed_in, uint8_t *packed_out, uint32_t n_steps);
int  fs3d_upload_packed(fs3d_world *w, const uint8_t *packed);      /* the planes this world holds, in that encoding */
int  fs3d_download_packed(fs3d_world *w, uint8_t *packed);

/* The same for one rank's slab of a fused-halo-push world (after fs3d_slab_ipc_attach):
 *   fs3d_slab_step_host_begin(w, host_in)   uploads the slab's two edge planes and stores them into the
 *                                           neighbours' ghost planes over peer memory
 *   -- barrier across ranks --
 *   fs3d_slab_step_host(w, host_in, host_out, n)   streams the slab through like fs3d_step_host
 *   -- barrier across ranks before the next _begin --
 * Afterwards the device holds the stepped slab but not its neighbours' new edge planes: call
 * fs3d_slab_push_halos on every rank (and barrier) before going back to fs3d_step. */
int  fs3d_slab_step_host_begin(fs3d_world *w, const uint8_t *host_in);
int  fs3d_slab_step_host(fs3d_world *w, const uint8_t *host_in, uint8_t *host_out, uint32_t n_steps);

/* ---- reductions (over the planes this world holds; sum across ranks yourself) ---- */
int  fs3d_histogram(fs3d_world *w, uint64_t counts[256]);
int  fs3d_digest(fs3d_world *w, uint64_t *out);
/* Settled-tile statistics of the last step: tiles processed / tiles total (skipping on). */
int  fs3d_activity(fs3d_world *w, uint64_t *tiles_run, uint64_t *tiles_total);

/* ---- renderer hand-off ---- */
int  fs3d_num_slabs(fs3d_world *w, int32_t *out);
int  fs3d_volume_view(fs3d_world *w, int32_t slab, fs3d_view *out);
int  fs3d_volume_export_fd(fs3d_world *w, int32_t slab, fs3d_export *out);   /* needs FS3D_FLAG_EXPORTABLE */
int  fs3d_set_palette(fs3d_world *w, const float *rgba256x4);
int  fs3d_raymarch(fs3d_world *w, const fs3d_camera *cam, uint32_t width, uint32_t height,
                   uint32_t mode, uint8_t *host_rgba8);
/* 1 if the last voxel-mode frame marched on `slab` used the brick occupancy map (adaptive unless forced by the mode) */
int  fs3d_raymarch_bricks_in_use(fs3d_world *w, int32_t slab);
/* Same, plus the hit parameter t per pixel (+inf on a miss): ranks that each hold one slab
 * composite their images by taking, per pixel, the colour with the smallest t. */
int  fs3d_raymarch_depth(fs3d_world *w, const fs3d_camera *cam, uint32_t width, uint32_t height,
                         uint32_t mode, uint8_t *host_rgba8, float *host_depth);

/* ---- one-process-per-GPU slab stepping with caller-driven halo exchange ----
 * Per step:  fs3d_slab_step_edges (writes the slab's two edge planes of the back buffer)
 *            → caller sends send_lo/send_hi of fs3d_slab_halo(.., back=1) to its z-neighbours and
 *              receives into recv_lo/recv_hi, ordered after the halo.stream work so far
 *            → fs3d_slab_step_interior (overlaps with the exchange)
 *            → fs3d_slab_step_finish (flips buffers, ++step).
 * At a global boundary the ghost plane is STONE and must not be overwritten. */
int  fs3d_slab_halo(fs3d_world *w, int back, fs3d_halo *out);
/* Steps fused per pass (1, or 2 when the step index is even): a 2-step pass needs ONE halo exchange. */
int  fs3d_slab_pass_steps(fs3d_world *w, uint32_t n_steps);
int  fs3d_slab_step_edges(fs3d_world *w);
int  fs3d_slab_step_interior(fs3d_world *w);
int  fs3d_slab_step_finish(fs3d_world *w);

/* ---- fused halo push between ranks over peer memory (CUDA IPC + NVLink) ----
 * Each rank exports a blob, sends it to its z-neighbours by any means (bench.py: torch.distributed
 * all_gather_object), and attaches the blobs of the rank below / above (NULL at the global
 * boundary).  From then on fs3d_step() works on the slab world: ONE kernel per pass computes the
 * slab, stores its edge planes straight into the neighbours' ghost planes over NVLink and
 * signals an arrival counter; warps that need a ghost plane wait on their own counter.  No host
 * synchronisation, no NCCL and no separate edge launches are involved per step.  All ranks must
 * issue the same sequence of fs3d_step / upload / generate calls; after editing the front buffer
 * call fs3d_slab_push_halos on every rank and barrier. */
#define FS3D_IPC_BLOB_BYTES 256
int  fs3d_slab_ipc_export(fs3d_world *w, void *blob, uint64_t blob_bytes);
int  fs3d_slab_ipc_attach(fs3d_world *w, const void *lower_blob, const void *upper_blob);
/* The same wiring between slab worlds of ONE process (one host thread per GPU — or several slabs on one GPU):
 * plain pointers instead of IPC handles.  lower / upper = the worlds holding the adjacent slabs, NULL at the global
 * boundary.  Call it on every slab world, then fs3d_slab_push_halos on every one, before the first fs3d_step. */
int  fs3d_slab_attach_local(fs3d_world *w, fs3d_world *lower, fs3d_world *upper);
int  fs3d_slab_push_halos(fs3d_world *w);
/* Four-step passes (FS3D_FLAG_NO_FUSE4 off) on a multi-process slab world: every rank's slab must support them (even
 * internal boundaries, >= 4 planes, rows of 1024 / 2048 voxels, schedule version 1, no skipping) and all ranks must
 * decide alike: combine fs3d_slab_can_fuse4 over the ranks (AND) and pass the result to fs3d_slab_allow_fuse4 before
 * the first fs3d_step.  Default for slab worlds: not allowed (two-step passes).  In-process worlds decide by themselves. */
int  fs3d_slab_can_fuse4(fs3d_world *w);
int  fs3d_slab_allow_fuse4(fs3d_world *w, int allow);
/* Cost of inter-GPU skew: ns the PUSH kernels spent blocked on a neighbour's arrival counter since the last call
 * (out[0] summed over warps, out[1] longest single wait, out[2] number of blocking waits); resets the counters. */
int  fs3d_push_wait_stats(fs3d_world *w, uint64_t out[3]);
/* Watchdog: a PUSH kernel that has waited FS3D_PUSH_TIMEOUT_MS (environment, default 20000) for a neighbour's halo
 * gives up; the next fs3d_sync (or any call that synchronises) returns FS3D_ERR_CUDA naming the stalled side, and the
 * world refuses to step again (its cells are undefined).  Nothing hangs. */

/* ---- fused multi-rank ray-march (one process per GPU): march + composite over peer memory ----
 * The compositor rank allocates a frame of n_slots x (width x height) 64-bit words and exports it;
 * every rank (the compositor too) attaches with its own slot.  fs3d_raymarch_to_frame marches this
 * rank's slab and its kernel stores (hit parameter bits << 32 | rgba8) for every pixel straight into
 * that slot — NVLink peer stores for the other ranks, no host copy, no collective.  After every rank
 * has synchronised (fs3d_sync + a barrier of the caller's choice) the compositor calls
 * fs3d_frame_resolve: per pixel the slot with the smallest hit parameter wins (same image as one
 * world marching all slabs), copied to host_rgba8.  Barrier again before the next frame is marched.
 * Camera and shading are those of shaders/fs_raymarch.{vert,frag} as for fs3d_raymarch. */
int  fs3d_frame_export(fs3d_world *w, uint32_t width, uint32_t height, uint32_t n_slots, void *blob, uint64_t blob_bytes);
int  fs3d_frame_attach(fs3d_world *w, const void *blob, uint32_t slot);
int  fs3d_frame_attach_local(fs3d_world *w, fs3d_world *owner, uint32_t slot);   /* same, compositor world in this process */
int  fs3d_raymarch_to_frame(fs3d_world *w, const fs3d_camera *cam, uint32_t mode);   /* asynchronous */
int  fs3d_frame_resolve(fs3d_world *w, uint8_t *host_rgba8);

/* ---- checkpoint (the reference has no on-disk state, src/engine holds none) ----
 * One file per world handle (per rank for slab worlds).  Little-endian:
 *   offset  0  char[8]  "FS3DCKPT"
 *           8  u32 format version (FS3D_CKPT_VERSION)   12  u32 schedule version
 *          16  u32 nx, ny, nz                            28  u32 z_begin, z_end (planes in the file)
 *          36  u32 encoding (1 | 2)                          40  u64 step index   48  u64 seed
 *          56  u64 digest of these planes                64  u64 payload bytes   72  u64 reserved (0)
 *   offset 80  payload: cells in x-fastest order; encoding 1 (schedule version 1): 2 bits each, cell i in bits 2(i & 3) of
 *              byte i >> 2; encoding 2 (schedule version 2, FS3D_FLAG_MATERIALS8): 4 bits each, cell i in bits 4(i & 1) of byte i >> 1
 * fs3d_load restores cells, step index and seed (so the run continues bit-identically), verifies
 * dimensions, plane range, schedule version and the digest, and leaves the world untouched on any
 * error.  Ranks of a fused-halo-push world call fs3d_slab_push_halos + barrier afterwards. */
#define FS3D_CKPT_VERSION       1
#define FS3D_CKPT_HEADER_BYTES  80
int  fs3d_save(fs3d_world *w, const char *path);
int  fs3d_load(fs3d_world *w, const char *path);

const char *fs3d_last_error(void);
int  fs3d_schedule_version(void);
int  fs3d_world_schedule_version(fs3d_world *w);   /* 1, or 2 for FS3D_FLAG_MATERIALS8 worlds */

#ifdef __cplusplus
}
#endif
#endif /* FS3D_H */
