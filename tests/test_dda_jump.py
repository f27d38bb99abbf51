"""The exactness argument behind the ray-marcher's jumps (csrc/raymarch.cuh), checked in float32 on the CPU.

The voxel walk recomputes every crossing time from the integer boundary index (boundary_t), so its state depends only
on position.  jump() leaves an empty axis-aligned box in one move; this test replays both in numpy float32 — the walk
voxel by voxel until it leaves the box, and the jump — for random rays, origins and boxes (including rays along lattice
planes and exact ties), and requires the same voxel, the same crossing times, the same t and the same exit axis."""
import numpy as np

F = np.float32


def boundary_t(k, h, e, o, rcp):
    v = F(F(k) * h)
    v = F(v - e)
    v = F(v - o)
    return F(v * rcp)


def setup(o, d, idx, h, e):
    stp = [1 if d[a] > 0 else (-1 if d[a] < 0 else 0) for a in range(3)]
    rcp = [F(F(1.0) / d[a]) if stp[a] else F(0) for a in range(3)]
    tnext = [boundary_t(idx[a] + (1 if stp[a] > 0 else 0), h, e[a], o[a], rcp[a]) if stp[a] else F(np.inf) for a in range(3)]
    return stp, rcp, tnext


def walk_out(o, d, idx, h, e, lo, hi):
    """voxel by voxel until the ray is outside [lo, hi): state after the move that left the box"""
    idx = list(idx)
    stp, rcp, tnext = setup(o, d, idx, h, e)
    while True:
        a = 0
        if tnext[1] < tnext[a]:
            a = 1
        if tnext[2] < tnext[a]:
            a = 2
        t = tnext[a]
        idx[a] += stp[a]
        tnext[a] = boundary_t(idx[a] + (1 if stp[a] > 0 else 0), h, e[a], o[a], rcp[a])
        if not (lo[a] <= idx[a] < hi[a]):
            return idx, tnext, t, a


def jump(o, d, idx, h, e, lo, hi):
    idx = list(idx)
    stp, rcp, tnext = setup(o, d, idx, h, e)
    tex = [boundary_t(hi[a] if stp[a] > 0 else lo[a], h, e[a], o[a], rcp[a]) if stp[a] else F(np.inf) for a in range(3)]
    ax = 0
    if tex[1] < tex[ax]:
        ax = 1
    if tex[2] < tex[ax]:
        ax = 2
    T = tex[ax]
    for b in range(3):
        if b == ax:
            idx[b] = hi[b] if stp[b] > 0 else lo[b] - 1
            continue
        if stp[b] == 0:
            continue

        def crossed(k):
            tb = boundary_t(k, h, e[b], o[b], rcp[b])
            return tb < T or (tb == T and b < ax)

        with np.errstate(invalid="ignore", over="ignore"):
            g = int(np.floor(F(F(F(o[b] + F(T * d[b])) + e[b]) / h)))
        if stp[b] > 0:
            g = idx[b] if g < idx[b] else min(g, hi[b] - 1)
            while g > idx[b] and not crossed(g):
                g -= 1
            while g + 1 <= hi[b] - 1 and crossed(g + 1):
                g += 1
        else:
            g = idx[b] if g > idx[b] else max(g, lo[b])
            while g < idx[b] and not crossed(g + 1):
                g += 1
            while g - 1 >= lo[b] and crossed(g):
                g -= 1
        idx[b] = g
    tnext = [boundary_t(idx[a] + (1 if stp[a] > 0 else 0), h, e[a], o[a], rcp[a]) if stp[a] else F(np.inf) for a in range(3)]
    return idx, tnext, T, ax


def test_jump_lands_in_the_state_of_the_voxel_walk():
    rng = np.random.default_rng(5)
    n = (64, 48, 40)
    h = F(1.0 / 64)
    e = [F(n[a] * h / 2) for a in range(3)]
    cases = 0
    for trial in range(1500):
        kind = trial % 5
        o = [F(v) for v in rng.uniform(-2.0, 2.0, 3)]
        d = rng.normal(size=3)
        if kind == 1:
            d[rng.integers(3)] = 0.0                       # along a lattice plane
        if kind == 2:
            d = np.sign(d) * np.array([1.0, 1.0, 1.0])      # exact diagonal: ties on every crossing
            o = [F(np.round(v * 64) / 64) for v in o]
        if kind == 3:
            d[1] = d[0]                                     # ties between two axes
            o[1] = o[0]
        d = d / np.linalg.norm(d)
        d = [F(v) for v in d]
        # a box and a start voxel inside it
        lo = [int(rng.integers(0, n[a] - 1)) for a in range(3)]
        hi = [int(rng.integers(lo[a] + 1, min(n[a], lo[a] + 9) + 1)) for a in range(3)]
        if kind == 4:
            lo, hi = [0, 0, 0], [n[0], n[1], int(rng.integers(1, n[2]))]      # the planes another rank holds
        idx = [int(rng.integers(lo[a], hi[a])) for a in range(3)]
        if all(v == 0 for v in d):
            continue
        w = walk_out(o, d, idx, h, e, lo, hi)
        j = jump(o, d, idx, h, e, lo, hi)
        assert w[0] == j[0], (trial, kind, w, j)
        assert [v.tobytes() for v in w[1]] == [v.tobytes() for v in j[1]], (trial, kind, w, j)
        assert w[2].tobytes() == j[2].tobytes() and w[3] == j[3], (trial, kind, w, j)
        cases += 1
    assert cases > 1400
