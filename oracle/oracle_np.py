"""Second, independent restatement of SCHEDULE.md in vectorised numpy.  TEST INFRASTRUCTURE ONLY.

Exists because parity is unpinned (no reference implementation or golden vectors — SURVEY.md
§0, §8c): the scalar C oracle (fs3d_oracle.c), this mask-algebra version and the bit-sliced CUDA
kernels are three differently-shaped implementations of one spec and must agree bit-for-bit.
"""
import numpy as np

EMPTY, SAND, WATER, STONE, GAS, OIL, HONEY, GRAVEL = 0, 1, 2, 3, 4, 5, 6, 7
M64 = (1 << 64) - 1


def mix64(v):
    v &= M64
    v ^= v >> 30; v = (v * 0xBF58476D1CE4E5B9) & M64
    v ^= v >> 27; v = (v * 0x94D049BB133111EB) & M64
    v ^= v >> 31
    return v


def key(seed, t, axis):
    v = (seed ^ ((t * 0x9E3779B97F4A7C15) & M64) ^ (((axis + 1) * 0xD1B54A32D192ED03) & M64)) & M64
    v = mix64(v)
    return (v ^ (v >> 32)) & 0xFFFFFFFF


def hash_words(k, xw, y, z):
    """Vectorised H(key, xw, y, z) on uint32 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        v = (np.uint32(k) + xw.astype(np.uint32) * np.uint32(0x9E3779B1) + y.astype(np.uint32) * np.uint32(0x85EBCA77)
             + z.astype(np.uint32) * np.uint32(0xC2B2AE3D)).astype(np.uint32)
        v ^= v >> np.uint32(16); v *= np.uint32(0x7FEB352D)
        v ^= v >> np.uint32(15); v *= np.uint32(0x846CA68B)
        v ^= v >> np.uint32(16)
    return v


def coins(k, X, Y, Z):
    X = X.astype(np.int64)
    xu = (X & 0xFFFFFFFF).astype(np.uint32)
    bit = (np.uint32(8) * (xu & np.uint32(3)) + ((xu >> np.uint32(2)) & np.uint32(7))).astype(np.uint32)
    h = hash_words(k, xu >> np.uint32(5), (Y.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32),
                   (Z.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32))
    return ((h >> bit) & np.uint32(1)).astype(bool)


def coins2(k, X, Y, Z):
    """SCHEDULE.md §7: the second coin, the same hash word passed through C2(v) = (v * 0x9E3779B1) ^ ((v * 0x9E3779B1) >> 15)."""
    X = X.astype(np.int64)
    xu = (X & 0xFFFFFFFF).astype(np.uint32)
    bit = (np.uint32(8) * (xu & np.uint32(3)) + ((xu >> np.uint32(2)) & np.uint32(7))).astype(np.uint32)
    h = hash_words(k, xu >> np.uint32(5), (Y.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32),
                   (Z.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32))
    with np.errstate(over="ignore"):
        v = (h * np.uint32(0x9E3779B1)).astype(np.uint32)
    v ^= v >> np.uint32(15)
    return ((v >> bit) & np.uint32(1)).astype(bool)


# ---- schedule version 2 (SCHEDULE.md §7): eight materials, as lookup tables over the codes ----
_RANK2 = np.array([1, 5, 3, 7, 0, 2, 4, 6], dtype=np.int64)                     # by material code
_YIELDS2 = np.array([1, 0, 1, 0, 1, 1, 1, 0], dtype=bool)                       # GAS, EMPTY, OIL, WATER, HONEY


def _heavier2(u, l):
    return (u != STONE) & _YIELDS2[l] & (_RANK2[u] > _RANK2[l])


def _rule2(a, b, c, d, coin, coin2):
    fa = _heavier2(a, c); a, c = _swap(fa, a, c)
    fb = _heavier2(b, d); b, d = _swap(fb, b, d)
    da = _heavier2(a, d) & (b != STONE) & (a != GRAVEL)
    db = _heavier2(b, c) & (a != STONE) & (b != GRAVEL) & ~da
    a, d = _swap(da, a, d)
    b, c = _swap(db, b, c)
    l = (a != b) & _YIELDS2[a] & _YIELDS2[b]
    viscous = (a == HONEY) | (b == HONEY)
    a, b = _swap(l & coin & (~viscous | coin2), a, b)
    return a, b, c, d


def _dens(m):
    return np.where(m == SAND, 2, np.where(m == WATER, 1, 0))


def _heavier(u, l):
    return ((u == SAND) | (u == WATER)) & ((l == EMPTY) | (l == WATER)) & (_dens(u) > _dens(l))


def _swap(mask, p, q):
    pn = np.where(mask, q, p)
    qn = np.where(mask, p, q)
    return pn, qn


def _rule(a, b, c, d, coin):
    fa = _heavier(a, c); a, c = _swap(fa, a, c)
    fb = _heavier(b, d); b, d = _swap(fb, b, d)
    da = _heavier(a, d) & (b != STONE)
    db = _heavier(b, c) & (a != STONE) & ~da
    a, d = _swap(da, a, d)
    b, c = _swap(db, b, c)
    l = ((a == WATER) & (b == EMPTY)) | ((b == WATER) & (a == EMPTY))
    a, b = _swap(l & coin, a, b)
    return a, b, c, d


def _substep(grid, k, axis, oh, oy, version=1):
    """axis 0: XY blocks (horizontal = x, numpy axis 2); axis 1: ZY blocks (horizontal = z, numpy axis 0)."""
    nz, ny, nx = grid.shape
    g = np.transpose(grid, (2, 1, 0)) if axis == 1 else grid     # -> (free, y, h)
    nf, _, nh = g.shape
    P = np.full((nf, ny + 2, nh + 2), STONE, dtype=np.uint8)
    P[:, 1:-1, 1:-1] = g
    sh, sy = (oh + 1) % 2, (oy + 1) % 2
    nbh, nby = (nh + 2 - sh) // 2, (ny + 2 - sy) // 2
    hs = slice(sh, sh + 2 * nbh, 2); hs1 = slice(sh + 1, sh + 1 + 2 * nbh, 2)
    ys = slice(sy, sy + 2 * nby, 2); ys1 = slice(sy + 1, sy + 1 + 2 * nby, 2)
    a, b, c, d = P[:, ys1, hs], P[:, ys1, hs1], P[:, ys, hs], P[:, ys, hs1]
    # coin at the upper-left cell a: unpadded coords
    H0 = (np.arange(nbh) * 2 + sh - 1)[None, None, :]
    Yu = (np.arange(nby) * 2 + sy + 1 - 1)[None, :, None]
    F = np.arange(nf)[:, None, None]
    H0b, Yub, Fb = np.broadcast_arrays(H0, Yu, F)
    XYZ = (H0b, Yub, Fb) if axis == 0 else (Fb, Yub, H0b)     # XY: (h0, y0+1, z); ZY: (x, y0+1, z0)
    coin = coins(k, *XYZ)
    if version == 2:
        # outside the grid the coordinates are meaningless, but there a or b reads STONE and L cannot fire
        a2, b2, c2, d2 = _rule2(a, b, c, d, coin, coins2(k, *XYZ))
    else:
        a2, b2, c2, d2 = _rule(a, b, c, d, coin)
    P[:, ys1, hs], P[:, ys1, hs1], P[:, ys, hs], P[:, ys, hs1] = a2, b2, c2, d2
    out = P[:, 1:-1, 1:-1]
    if axis == 1:
        out = np.transpose(out, (2, 1, 0))
    grid[...] = out


def step(grid, seed, t, version=1):
    """One in-place step of a whole (nz, ny, nx) uint8 grid under schedule `version` (1: SCHEDULE.md §1-5, 2: §7)."""
    hoff = (t >> 1) & 1
    kxy, kzy = key(seed, t, 0), key(seed, t, 1)
    if t % 2 == 0:
        _substep(grid, kxy, 0, hoff, 0, version)
        _substep(grid, kzy, 1, hoff, 1, version)
    else:
        _substep(grid, kzy, 1, hoff, 0, version)
        _substep(grid, kxy, 0, hoff, 1, version)
    return grid


def digest(grid, zlo=0):
    nz, ny, nx = grid.shape
    idx = np.nonzero(grid.ravel())[0].astype(np.uint64) + np.uint64(zlo * ny * nx)
    vals = grid.ravel()[np.nonzero(grid.ravel())[0]].astype(np.uint64)
    total = 0
    for i, m in zip(idx.tolist(), vals.tolist()):
        total = (total + mix64(8 * i + m)) & M64
    return total
