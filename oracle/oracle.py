"""ctypes binding of the C oracle (oracle/fs3d_oracle.c, fs3d_sweep.c, fs3d_raymarch_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by fallingsand3d_b200/.  PARITY UNPINNED: the
reference has no implementation, tests or golden vectors for the simulation.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libfs3d_oracle.so")
_SOURCES = ["fs3d_oracle.c", "fs3d_sweep.c", "fs3d_raymarch_oracle.c", "Makefile"]

REF_LIB_PATH = os.path.join(HERE, "_ref", "libfs_raymarch_ref.so")
REFERENCE_ROOT = "/root/reference"

_lib = None
_ref = None


def build_ref(force=False):
    """Compiles the reference's own shaders/fs_raymarch.frag (with its vendored glm) into oracle/_ref/ — only
    possible where /root/reference is mounted.  Returns the library path, or None when it cannot be built and no
    prebuilt copy travelled with the repo."""
    if os.path.isdir(REFERENCE_ROOT) and (force or not os.path.exists(REF_LIB_PATH)):
        res = subprocess.run(["make", "-C", HERE, "ref"] + (["-B"] if force else []), capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + res.stdout + res.stderr)
    return REF_LIB_PATH if os.path.exists(REF_LIB_PATH) else None


def ref_frame(pos=(0.0, 0.0, -5.0), aspect=1700.0 / 900.0, width=850, height=450):
    """Linear RGBA float32 frame (H, W, 4) of the REFERENCE fragment shader itself, or None without oracle/_ref."""
    global _ref
    if _ref is None:
        path = build_ref()
        if path is None:
            return None
        _ref = C.CDLL(path)
        _ref.fs_raymarch_ref_frame.restype = None
        _ref.fs_raymarch_ref_frame.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
    out = np.empty((height, width, 4), dtype=np.float32)
    _ref.fs_raymarch_ref_frame((C.c_float * 3)(*pos), aspect, width, height, out.ctypes.data_as(C.c_void_p))
    return out


def build(force=False):
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(os.path.join(HERE, s)) > os.path.getmtime(LIB_PATH) for s in _SOURCES)
    if stale:
        res = subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building the oracle failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        i64, u64, u32, u8p = C.c_int64, C.c_uint64, C.c_uint32, C.c_void_p
        L.fs3d_oracle_key.restype = u32
        L.fs3d_oracle_key.argtypes = [u64, u64, u32]
        L.fs3d_oracle_hash.restype = u32
        L.fs3d_oracle_hash.argtypes = [u32, u32, u32, u32]
        L.fs3d_oracle_coin.restype = C.c_int
        L.fs3d_oracle_coin.argtypes = [u64, u64, u32, u32, u32, u32]
        L.fs3d_oracle_step.restype = i64
        L.fs3d_oracle_step.argtypes = [u8p, i64, i64, i64, u64, u64]
        L.fs3d_oracle_step_range.restype = i64
        L.fs3d_oracle_step_range.argtypes = [u8p, i64, i64, i64, i64, i64, i64, i64, u64, u64]
        L.fs3d_oracle_run.restype = None
        L.fs3d_oracle_run.argtypes = [u8p, i64, i64, i64, u64, u64, u64]
        L.fs3d_oracle_step_v.restype = i64
        L.fs3d_oracle_step_v.argtypes = [u8p, i64, i64, i64, u64, u64, C.c_int]
        L.fs3d_oracle_run_v.restype = None
        L.fs3d_oracle_run_v.argtypes = [u8p, i64, i64, i64, u64, u64, u64, C.c_int]
        L.fs3d_oracle_step_range_v.restype = i64
        L.fs3d_oracle_step_range_v.argtypes = [u8p, i64, i64, i64, i64, i64, i64, i64, u64, u64, C.c_int]
        L.fs3d_oracle_coin2.restype = C.c_int
        L.fs3d_oracle_coin2.argtypes = [u64, u64, u32, u32, u32, u32]
        L.fs3d_oracle_scene_cell.restype = C.c_uint8
        L.fs3d_oracle_scene_cell.argtypes = [C.c_int, u64, i64, i64, i64, i64, i64, i64]
        L.fs3d_oracle_generate.restype = None
        L.fs3d_oracle_generate.argtypes = [u8p, i64, i64, i64, i64, i64, C.c_int, u64]
        L.fs3d_oracle_histogram.restype = None
        L.fs3d_oracle_histogram.argtypes = [u8p, i64, C.POINTER(u64)]
        L.fs3d_oracle_digest.restype = u64
        L.fs3d_oracle_digest.argtypes = [u8p, i64, i64, i64, i64]
        L.fs3d_oracle_schedule_version.restype = C.c_int
        L.fs3d_oracle_threads.restype = C.c_int
        L.fs3d_sweep_step.restype = i64
        L.fs3d_sweep_step.argtypes = [u8p, i64, i64, i64, C.c_int, u64]
        L.fs3d_oracle_raymarch.restype = None
        L.fs3d_oracle_raymarch.argtypes = [u8p, u32, u32, u32, u32, u32, C.POINTER(C.c_float), C.c_float, C.c_float,
                                           u32, u32, u32, C.POINTER(C.c_float), u8p, u8p]
        L.fs3d_oracle_raymarch_pixel.restype = None
        L.fs3d_oracle_raymarch_pixel.argtypes = [C.POINTER(C.c_float), C.c_float, u32, u32, u32, u32,
                                                 C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _chk(grid):
    assert grid.dtype == np.uint8 and grid.flags.c_contiguous and grid.ndim == 3
    return grid.shape  # (nz, ny, nx)


def step(grid, seed, t, version=1):
    """One in-place step of a whole (nz, ny, nx) uint8 grid under schedule `version`. Returns the enabled-block count."""
    nz, ny, nx = _chk(grid)
    return lib().fs3d_oracle_step_v(_ptr(grid), nx, ny, nz, seed, t, version)


def run(grid, seed, t0, nsteps, version=1):
    nz, ny, nx = _chk(grid)
    lib().fs3d_oracle_run_v(_ptr(grid), nx, ny, nz, seed, t0, nsteps, version)
    return grid


def step_range(arr, nzg, zbase, own_lo, own_hi, seed, t, version=1):
    """One step of a slab array holding global planes [zbase, zbase + arr.shape[0])."""
    narr, ny, nx = _chk(arr)
    return lib().fs3d_oracle_step_range_v(_ptr(arr), nx, ny, nzg, zbase, narr, own_lo, own_hi, seed, t, version)


def generate(nx, ny, nz, scene, seed, zlo=0, zhi=None):
    zhi = nz if zhi is None else zhi
    out = np.empty((zhi - zlo, ny, nx), dtype=np.uint8)
    lib().fs3d_oracle_generate(_ptr(out), nx, ny, nz, zlo, zhi, scene, seed)
    return out


def histogram(grid):
    g = np.ascontiguousarray(grid)
    counts = (C.c_uint64 * 256)()
    lib().fs3d_oracle_histogram(_ptr(g), g.size, counts)
    return np.frombuffer(counts, dtype=np.uint64).copy()


def digest(grid, zlo=0):
    nz, ny, nx = _chk(grid)
    return lib().fs3d_oracle_digest(_ptr(grid), nx, ny, zlo, zlo + nz)


def sweep_step(grid, with_lateral=0, parity=0):
    nz, ny, nx = _chk(grid)
    return lib().fs3d_sweep_step(_ptr(grid), nx, ny, nz, with_lateral, parity)


def threads():
    """Host threads the oracle's OpenMP loops actually run on (honours OMP_NUM_THREADS and the affinity mask)."""
    return int(lib().fs3d_oracle_threads())


def key(seed, t, axis):
    return lib().fs3d_oracle_key(seed, t, axis)


def hash_word(k, xw, y, z):
    return lib().fs3d_oracle_hash(k, xw & 0xFFFFFFFF, y & 0xFFFFFFFF, z & 0xFFFFFFFF)


def coin(seed, t, axis, x, y, z):
    return lib().fs3d_oracle_coin(seed, t, axis, x, y, z)


def coin2(seed, t, axis, x, y, z):
    return lib().fs3d_oracle_coin2(seed, t, axis, x, y, z)


def default_palette():
    p = np.zeros((256, 4), dtype=np.float32)
    g = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    p[:, 0] = g; p[:, 1] = g; p[:, 2] = g; p[:, 3] = 1.0
    p[0] = (0, 0, 0, 0)
    p[1] = (0.86, 0.72, 0.40, 1)
    p[2] = (0.15, 0.40, 0.85, 1)
    p[3] = (0.45, 0.45, 0.48, 1)
    p[4] = (0.80, 0.90, 0.75, 1)      # GAS
    p[5] = (0.25, 0.20, 0.10, 1)      # OIL
    p[6] = (0.95, 0.65, 0.10, 1)      # HONEY
    p[7] = (0.55, 0.50, 0.45, 1)      # GRAVEL
    return p


def raymarch(grid, nz_global=None, zlo=0, pos=(0.0, 0.0, -5.0), yaw_deg=0.0, aspect=1700.0 / 900.0,
             width=850, height=450, mode=1, palette=None, with_depth=False):
    if grid is None:
        grid = np.zeros((1, 1, 32), dtype=np.uint8)
    nzh, ny, nx = _chk(grid)
    nzg = nzh if nz_global is None else nz_global
    pal = np.ascontiguousarray(default_palette() if palette is None else palette, dtype=np.float32)
    img = np.empty((height, width, 4), dtype=np.uint8)
    depth = np.empty((height, width), dtype=np.float32) if with_depth else None
    p = (C.c_float * 3)(*pos)
    lib().fs3d_oracle_raymarch(_ptr(grid), nx, ny, nzg, zlo, zlo + nzh, p, yaw_deg, aspect, width, height, mode,
                               pal.ctypes.data_as(C.POINTER(C.c_float)), _ptr(img),
                               _ptr(depth) if with_depth else None)
    return (img, depth) if with_depth else img


def raymarch_pixel(px, py, pos=(0.0, 0.0, -5.0), aspect=1700.0 / 900.0, width=850, height=450):
    p = (C.c_float * 3)(*pos)
    d = (C.c_float * 3)()
    red = C.c_float()
    it = C.c_int()
    lib().fs3d_oracle_raymarch_pixel(p, aspect, width, height, px, py, d, C.byref(red), C.byref(it))
    return (d[0], d[1], d[2]), red.value, it.value
