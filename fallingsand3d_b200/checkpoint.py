"""Numpy reader / writer of the fs3d checkpoint format (include/fs3d.h, "checkpoint").

Host-side file-format logic only: it lets tools and tests inspect what fs3d_save wrote and prepare
states for fs3d_load without a GPU.  It does not step anything.  The reference has no on-disk
state to be compatible with (SURVEY.md §5 "Checkpoint / resume: No", §8f.3).
"""
import struct

import numpy as np

MAGIC = b"FS3DCKPT"
FORMAT_VERSION = 1
SCHEDULE_VERSION = 1
HEADER = struct.Struct("<8sIIIIIIIIQQQQQ")      # 80 bytes
assert HEADER.size == 80
FIELDS = ("magic", "format_version", "schedule_version", "nx", "ny", "nz", "z_begin", "z_end", "encoding",
          "step", "seed", "digest", "payload_bytes", "reserved")


def _mix64(v):
    v = v.copy()
    v ^= v >> np.uint64(30); v *= np.uint64(0xBF58476D1CE4E5B9)
    v ^= v >> np.uint64(27); v *= np.uint64(0x94D049BB133111EB)
    v ^= v >> np.uint64(31)
    return v


def digest(planes, nx, ny, z_begin=0):
    """SCHEDULE.md §4 digest of planes [z_begin, z_begin + len) of a grid with rows of nx and ny rows per plane."""
    flat = np.ascontiguousarray(planes, dtype=np.uint8).reshape(-1)
    total = np.uint64(0)
    base = np.uint64(z_begin) * np.uint64(nx) * np.uint64(ny)
    step = 1 << 22
    with np.errstate(over="ignore"):
        for i in range(0, flat.size, step):
            m = flat[i:i + step]
            nz_ = np.nonzero(m)[0]
            if nz_.size:
                idx = base + np.uint64(i) + nz_.astype(np.uint64)
                total += _mix64(np.uint64(8) * idx + m[nz_].astype(np.uint64)).sum(dtype=np.uint64)
    return int(total)


def pack2(cells):
    c = np.ascontiguousarray(cells, dtype=np.uint8).reshape(-1, 4)
    if (c > 3).any():
        raise ValueError("material codes 4-255 are reserved")
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def unpack2(packed, n):
    p = np.frombuffer(packed, dtype=np.uint8) if not isinstance(packed, np.ndarray) else packed
    out = np.empty((p.size, 4), dtype=np.uint8)
    for k in range(4):
        out[:, k] = (p >> (2 * k)) & 3
    return out.reshape(-1)[:n]


def pack4(cells):
    """encoding 2 (schedule version 2, codes 0..7): two voxels per byte, voxel i in bits 4(i & 1) of byte i >> 1"""
    c = np.ascontiguousarray(cells, dtype=np.uint8).reshape(-1, 2)
    if (c > 7).any():
        raise ValueError("material codes 8-255 are reserved")
    return (c[:, 0] | (c[:, 1] << 4)).astype(np.uint8)


def unpack4(packed, n):
    p = np.frombuffer(packed, dtype=np.uint8) if not isinstance(packed, np.ndarray) else packed
    out = np.empty((p.size, 2), dtype=np.uint8)
    out[:, 0] = p & 15
    out[:, 1] = p >> 4
    return out.reshape(-1)[:n]


def read_header(path):
    with open(path, "rb") as f:
        raw = f.read(HEADER.size)
    if len(raw) != HEADER.size:
        raise ValueError(f"{path}: too short for an fs3d checkpoint")
    h = dict(zip(FIELDS, HEADER.unpack(raw)))
    if h["magic"] != MAGIC:
        raise ValueError(f"{path}: not an fs3d checkpoint")
    return h


def read(path, verify=True):
    """-> (header dict, uint8 array of shape (z_end - z_begin, ny, nx))."""
    h = read_header(path)
    if h["format_version"] != FORMAT_VERSION or h["encoding"] not in (1, 2) or h["encoding"] != h["schedule_version"]:
        raise ValueError(f"{path}: unknown checkpoint format version / encoding")
    per_byte = 4 if h["encoding"] == 1 else 2
    nzh = h["z_end"] - h["z_begin"]
    n = h["nx"] * h["ny"] * nzh
    with open(path, "rb") as f:
        f.seek(HEADER.size)
        payload = f.read()
    if len(payload) != h["payload_bytes"] or h["payload_bytes"] * per_byte != n:
        raise ValueError(f"{path}: payload size is inconsistent with the header")
    grid = (unpack2 if h["encoding"] == 1 else unpack4)(payload, n).reshape(nzh, h["ny"], h["nx"])
    if verify and digest(grid, h["nx"], h["ny"], h["z_begin"]) != h["digest"]:
        raise ValueError(f"{path}: digest mismatch (corrupt checkpoint)")
    return h, grid


def write(path, planes, nz=None, z_begin=0, step=0, seed=1, schedule_version=SCHEDULE_VERSION):
    """Writes planes (shape (nzh, ny, nx), uint8 codes 0..3; 0..7 with schedule_version=2) as a checkpoint fs3d_load accepts."""
    g = np.ascontiguousarray(planes, dtype=np.uint8)
    nzh, ny, nx = g.shape
    if nx % 32:
        raise ValueError("nx must be a multiple of 32")
    payload = pack2(g) if schedule_version == 1 else pack4(g)
    hdr = HEADER.pack(MAGIC, FORMAT_VERSION, schedule_version, nx, ny, nzh if nz is None else nz, z_begin,
                      z_begin + nzh, 1 if schedule_version == 1 else 2, step, seed, digest(g, nx, ny, z_begin), payload.size, 0)
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(payload.tobytes())
