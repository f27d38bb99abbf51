"""VoxelWorld — host-side mirror of the voxel-world interface (create, set/get cell, step,
hand-off to the ray-marcher) over the C ABI in include/fs3d.h.

The reference has no such interface to copy names from; this class is what
its frame loop would hold (/root/reference/src/engine/engine.cpp:59-70) and it raises
RuntimeError("ERROR: ...") on failure the way util::displayError logs then throws
(/root/reference/src/util/debug.cpp:23-27).
"""
import ctypes as C
import sys

import numpy as np

from . import _lib

EMPTY, SAND, WATER, STONE = 0, 1, 2, 3
GAS, OIL, HONEY, GRAVEL = 4, 5, 6, 7            # schedule version 2 (FLAG_MATERIALS8), SCHEDULE.md §7
SCENE_EMPTY, SCENE_SAND_BLOCK, SCENE_MIXED, SCENE_RANDOM, SCENE_MIXED_NOISE = 0, 1, 2, 3, 4
SCENE_RANDOM8, SCENE_MIXED8 = 5, 6
FLAG_SKIP_SETTLED = 1
FLAG_NO_FUSE = 2
FLAG_NO_PEER_PUSH = 4
FLAG_PEER_PUSH_SHARED_DEVICE = 8
FLAG_EXPORTABLE = 16
FLAG_MATERIALS8 = 32
FLAG_NO_FUSE4 = 64
RM_SDF_SPHERE, RM_VOXELS, RM_SRGB = 0, 1, 16
RM_BRICKS, RM_NO_BRICKS = 32, 64      # force / forbid the 8^3-brick empty-space skipping (default: adaptive; same image)

ERROR_NAMES = {-1: "INVALID_ARG", -2: "BAD_DIMS", -3: "BAD_MATERIAL", -4: "OUT_OF_RANGE", -5: "CUDA",
               -6: "OOM", -7: "UNSUPPORTED", -8: "IO"}


class Fs3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ERROR: fs3d {ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        msg = _lib.load().fs3d_last_error().decode("utf-8", "replace")
        print(f"ERROR: {msg}", file=sys.stderr)   # debug.cpp:16 prints before debug.cpp:26 throws
        raise Fs3dError(rc, msg)


class VoxelWorld:
    """A double-buffered uint8 voxel grid on one or more B200s.

    VoxelWorld(nx, ny, nz, seed=1, n_gpus=1, devices=None, flags=0) owns the whole grid in this
    process (n_gpus z-slabs with in-library P2P halo exchange).
    VoxelWorld(..., slab=(z_begin, z_end)) owns one rank's slab on the current CUDA device; the
    caller drives the halo exchange (see slab.SlabWorld).
    """

    IPC_BLOB_BYTES = 256

    def __init__(self, nx, ny, nz, seed=1, n_gpus=1, devices=None, flags=0, slab=None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.nx, self.ny, self.nz, self.seed = int(nx), int(ny), int(nz), int(seed)
        dev_arr = None
        if devices is not None:
            dev_arr = (C.c_int32 * len(devices))(*devices)
            n_gpus = len(devices)
        d = _lib.Desc(self.nx, self.ny, self.nz, self.seed, int(n_gpus),
                      C.cast(dev_arr, C.POINTER(C.c_int32)) if dev_arr is not None else None, int(flags))
        if slab is None:
            self.z_begin, self.z_end = 0, self.nz
            _check(self._lib.fs3d_create(C.byref(d), C.byref(self._h)))
        else:
            self.z_begin, self.z_end = int(slab[0]), int(slab[1])
            _check(self._lib.fs3d_create_slab(C.byref(d), self.z_begin, self.z_end, C.byref(self._h)))

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.fs3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def shape(self):
        """(nz_held, ny, nx) — numpy order of upload()/download() arrays."""
        return (self.z_end - self.z_begin, self.ny, self.nx)

    # ---- cells ----
    def set_cell(self, x, y, z, m):
        _check(self._lib.fs3d_set_cell(self._h, x, y, z, m))

    def get_cell(self, x, y, z):
        out = C.c_uint8()
        _check(self._lib.fs3d_get_cell(self._h, x, y, z, C.byref(out)))
        return out.value

    def fill_box(self, lo, hi, m):
        a = (C.c_uint32 * 3)(*lo)
        b = (C.c_uint32 * 3)(*hi)
        _check(self._lib.fs3d_fill_box(self._h, a, b, m))

    def paint_sphere(self, centre, radius, m, only_empty=False):
        """Brush: cells within `radius` of `centre` (x, y, z; may lie outside the grid) become m."""
        _check(self._lib.fs3d_paint_sphere(self._h, int(centre[0]), int(centre[1]), int(centre[2]), int(radius), int(m),
                                           1 if only_empty else 0))

    def generate(self, scene_id, seed=None):
        _check(self._lib.fs3d_generate(self._h, int(scene_id), self.seed if seed is None else int(seed)))

    def upload(self, grid):
        g = np.ascontiguousarray(grid, dtype=np.uint8)
        if g.shape != self.shape:
            raise ValueError(f"expected array of shape {self.shape} (z, y, x), got {g.shape}")
        _check(self._lib.fs3d_upload(self._h, g.ctypes.data_as(C.c_void_p)))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == self.shape
        _check(self._lib.fs3d_download(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- stepping ----
    def step(self, n=1):
        _check(self._lib.fs3d_step(self._h, int(n)))

    def sync(self):
        _check(self._lib.fs3d_sync(self._h))

    @property
    def schedule_version(self):
        return int(self._lib.fs3d_world_schedule_version(self._h))

    @property
    def step_index(self):
        out = C.c_uint64()
        _check(self._lib.fs3d_step_index(self._h, C.byref(out)))
        return out.value

    @property
    def kernel_launches(self):
        out = C.c_uint64()
        _check(self._lib.fs3d_kernel_launches(self._h, C.byref(out)))
        return out.value

    def step_timed(self, n=1):
        """Runs n steps; returns (device milliseconds from CUDA events on the step stream, kernels launched)."""
        ms = C.c_float()
        nl = C.c_uint64()
        _check(self._lib.fs3d_step_timed(self._h, int(n), C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def step_host(self, grid_in, grid_out=None, n=1):
        """Steps a host-resident grid n (1, 2 or 4) steps: upload, kernels and download overlap chunk by chunk."""
        if grid_out is None:
            grid_out = grid_in
        for g in (grid_in, grid_out):
            assert g.dtype == np.uint8 and g.flags.c_contiguous and g.shape == self.shape
        _check(self._lib.fs3d_step_host(self._h, grid_in.ctypes.data_as(C.c_void_p),
                                        grid_out.ctypes.data_as(C.c_void_p), int(n)))
        return grid_out

    def _packed_size(self):
        return self.nx * self.ny * (self.z_end - self.z_begin) // (2 if self.schedule_version == 2 else 4)

    def upload_packed(self, packed):
        assert packed.dtype == np.uint8 and packed.flags.c_contiguous and packed.size == self._packed_size()
        _check(self._lib.fs3d_upload_packed(self._h, packed.ctypes.data_as(C.c_void_p)))

    def download_packed(self, out=None):
        if out is None:
            out = np.empty(self._packed_size(), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size == self._packed_size()
        _check(self._lib.fs3d_download_packed(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def step_host_packed(self, packed_in, packed_out=None, n=1):
        """Same for a grid the host keeps packed in the checkpoint encoding (checkpoint.pack2 / pack4): uint8 arrays of
        nx * ny * nz / 4 bytes (/ 2 for schedule version 2)."""
        if packed_out is None:
            packed_out = packed_in
        per = 2 if self.schedule_version == 2 else 4
        for g in (packed_in, packed_out):
            assert g.dtype == np.uint8 and g.flags.c_contiguous and g.size * per == self.nx * self.ny * (self.z_end - self.z_begin)
        _check(self._lib.fs3d_step_host_packed(self._h, packed_in.ctypes.data_as(C.c_void_p),
                                               packed_out.ctypes.data_as(C.c_void_p), int(n)))
        return packed_out

    # ---- checkpoint (format in include/fs3d.h; numpy reader/writer in checkpoint.py) ----
    def save(self, path):
        _check(self._lib.fs3d_save(self._h, str(path).encode()))

    def load(self, path):
        _check(self._lib.fs3d_load(self._h, str(path).encode()))
        from . import checkpoint
        self.seed = int(checkpoint.read_header(path)["seed"])   # the world now runs on the checkpoint's seed

    # ---- reductions ----
    def histogram(self):
        h = (C.c_uint64 * 256)()
        _check(self._lib.fs3d_histogram(self._h, h))
        return np.frombuffer(h, dtype=np.uint64).copy()

    def digest(self):
        out = C.c_uint64()
        _check(self._lib.fs3d_digest(self._h, C.byref(out)))
        return out.value

    def activity(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self._lib.fs3d_activity(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- renderer hand-off ----
    @property
    def num_slabs(self):
        out = C.c_int32()
        _check(self._lib.fs3d_num_slabs(self._h, C.byref(out)))
        return out.value

    def volume_view(self, slab=0):
        v = _lib.View()
        _check(self._lib.fs3d_volume_view(self._h, slab, C.byref(v)))
        return {k: getattr(v, k) for k, _ in _lib.View._fields_}

    def volume_export_fd(self, slab=0):
        """The slab's two buffers as POSIX file descriptors (needs FLAG_EXPORTABLE); the caller closes them."""
        e = _lib.Export()
        _check(self._lib.fs3d_volume_export_fd(self._h, slab, C.byref(e)))
        d = {k: getattr(e, k) for k, _ in _lib.Export._fields_ if k != "fd"}
        d["fd"] = (int(e.fd[0]), int(e.fd[1]))
        return d

    def set_palette(self, rgba):
        p = np.ascontiguousarray(rgba, dtype=np.float32)
        assert p.shape == (256, 4)
        _check(self._lib.fs3d_set_palette(self._h, p.ctypes.data_as(C.POINTER(C.c_float))))

    def raymarch(self, pos=(0.0, 0.0, -5.0), yaw_deg=0.0, aspect=1700.0 / 900.0, width=850, height=450,
                 mode=RM_VOXELS, with_depth=False):
        """Offscreen image (H, W, 4) uint8; defaults are the reference camera
        (renderer.h:148-149, window.h:41, materials.cpp:540)."""
        cam = _lib.Camera((C.c_float * 3)(*pos), yaw_deg, aspect)
        img = np.empty((height, width, 4), dtype=np.uint8)
        if with_depth:
            depth = np.empty((height, width), dtype=np.float32)
            _check(self._lib.fs3d_raymarch_depth(self._h, C.byref(cam), width, height, mode,
                                                 img.ctypes.data_as(C.c_void_p), depth.ctypes.data_as(C.c_void_p)))
            return img, depth
        _check(self._lib.fs3d_raymarch(self._h, C.byref(cam), width, height, mode, img.ctypes.data_as(C.c_void_p)))
        return img

    def raymarch_bricks_in_use(self, slab=0):
        return bool(self._lib.fs3d_raymarch_bricks_in_use(self._h, int(slab)))

    # ---- fused multi-rank ray-march (see slab.SlabWorld.raymarch) ----
    def frame_export(self, width, height, n_slots):
        buf = C.create_string_buffer(self.IPC_BLOB_BYTES)
        _check(self._lib.fs3d_frame_export(self._h, int(width), int(height), int(n_slots), buf, self.IPC_BLOB_BYTES))
        return buf.raw

    def frame_attach(self, blob, slot):
        _check(self._lib.fs3d_frame_attach(self._h, C.create_string_buffer(blob, self.IPC_BLOB_BYTES), int(slot)))

    def raymarch_to_frame(self, pos=(0.0, 0.0, -5.0), yaw_deg=0.0, aspect=1700.0 / 900.0, mode=RM_VOXELS):
        cam = _lib.Camera((C.c_float * 3)(*pos), yaw_deg, aspect)
        _check(self._lib.fs3d_raymarch_to_frame(self._h, C.byref(cam), mode))

    def frame_resolve(self, width, height, out=None):
        img = np.empty((height, width, 4), dtype=np.uint8) if out is None else out
        _check(self._lib.fs3d_frame_resolve(self._h, img.ctypes.data_as(C.c_void_p)))
        return img

    # ---- one-process-per-GPU slab protocol (see slab.SlabWorld) ----
    def slab_halo(self, back):
        h = _lib.Halo()
        _check(self._lib.fs3d_slab_halo(self._h, int(back), C.byref(h)))
        return h

    def slab_ipc_export(self):
        buf = C.create_string_buffer(self.IPC_BLOB_BYTES)
        _check(self._lib.fs3d_slab_ipc_export(self._h, buf, self.IPC_BLOB_BYTES))
        return buf.raw

    def slab_ipc_attach(self, lower_blob, upper_blob):
        lo = C.create_string_buffer(lower_blob, self.IPC_BLOB_BYTES) if lower_blob is not None else None
        hi = C.create_string_buffer(upper_blob, self.IPC_BLOB_BYTES) if upper_blob is not None else None
        _check(self._lib.fs3d_slab_ipc_attach(self._h, lo, hi))

    def slab_can_fuse4(self):
        return bool(self._lib.fs3d_slab_can_fuse4(self._h))

    def slab_allow_fuse4(self, allow):
        _check(self._lib.fs3d_slab_allow_fuse4(self._h, 1 if allow else 0))

    def push_wait_stats(self):
        """(ns blocked on neighbours' arrival counters summed over warps, longest single wait ns, blocking waits); resets."""
        v = (C.c_uint64 * 3)()
        _check(self._lib.fs3d_push_wait_stats(self._h, v))
        return int(v[0]), int(v[1]), int(v[2])

    def slab_attach_local(self, lower, upper):
        """Wire this slab world to the worlds holding the adjacent slabs in the same process (None at the boundary)."""
        _check(self._lib.fs3d_slab_attach_local(self._h, lower._h if lower is not None else None,
                                                upper._h if upper is not None else None))

    def frame_attach_local(self, owner, slot):
        _check(self._lib.fs3d_frame_attach_local(self._h, owner._h, int(slot)))

    def slab_step_host_begin(self, grid_in):
        assert grid_in.dtype == np.uint8 and grid_in.flags.c_contiguous and grid_in.shape == self.shape
        _check(self._lib.fs3d_slab_step_host_begin(self._h, grid_in.ctypes.data_as(C.c_void_p)))

    def slab_step_host(self, grid_in, grid_out=None, n=1):
        if grid_out is None:
            grid_out = grid_in
        for g in (grid_in, grid_out):
            assert g.dtype == np.uint8 and g.flags.c_contiguous and g.shape == self.shape
        _check(self._lib.fs3d_slab_step_host(self._h, grid_in.ctypes.data_as(C.c_void_p),
                                             grid_out.ctypes.data_as(C.c_void_p), int(n)))
        return grid_out

    def slab_push_halos(self):
        _check(self._lib.fs3d_slab_push_halos(self._h))

    def slab_pass_steps(self, n):
        _check(self._lib.fs3d_slab_pass_steps(self._h, int(n)))

    def slab_step_edges(self):
        _check(self._lib.fs3d_slab_step_edges(self._h))

    def slab_step_interior(self):
        _check(self._lib.fs3d_slab_step_interior(self._h))

    def slab_step_finish(self):
        _check(self._lib.fs3d_slab_step_finish(self._h))
