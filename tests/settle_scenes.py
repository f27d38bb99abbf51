"""Deterministic (tie-free) scenes whose settled configuration is unique, shared by the CPU and GPU tests of
"same settled configurations as the stand-in sequential sweep" (north_star; SCHEDULE.md §6)."""
import numpy as np

E, S, W, X = 0, 1, 2, 3


def shaft_scene(nx=32, ny=40, nz=10, seed=3):
    """1 x 1 vertical shafts at (even x, even z) separated by full-height STONE walls, some with STONE plugs; every
    shaft segment holds a random stack of SAND / WATER / EMPTY.  No cell has a free lateral neighbour, so nothing is
    order-dependent: in any falling-sand rule each segment must end sorted by density (SAND, WATER, EMPTY upwards)."""
    rng = np.random.default_rng(seed)
    g = np.full((nz, ny, nx), X, np.uint8)
    g[0::2, :, 0::2] = rng.choice(np.array([E, S, W], np.uint8), size=(nz // 2, ny, nx // 2), p=[0.4, 0.3, 0.3])
    plugs = rng.random((nz // 2, ny, nx // 2)) < 0.04
    g[0::2, :, 0::2][plugs] = X
    return g


def settle_closed_form(g):
    out = g.copy()
    nz, ny, nx = g.shape
    for z in range(nz):
        for x in range(nx):
            col = out[z, :, x]
            y = 0
            while y < ny:
                if col[y] == X:
                    y += 1
                    continue
                y1 = y
                while y1 < ny and col[y1] != X:
                    y1 += 1
                seg = col[y:y1]
                ns, nw = int((seg == S).sum()), int((seg == W).sum())
                seg[:] = E
                seg[:ns] = S
                seg[ns:ns + nw] = W
                y = y1
    return out


def basin_scene(nx=32, ny=20, nz=8):
    """A stone basin (floor + the closed box as walls) with a ragged heap of WATER whose volume is exactly 3 full
    layers: whatever the order of moves, the only settled state is the flat 3-layer pool."""
    g = np.zeros((nz, ny, nx), np.uint8)
    g[:, 0, :] = X
    vol = 3 * nx * nz
    g[:, 1:1 + 12, 0:8][...] = W            # a 8-wide, 12-high slab of water against one wall = nz * 8 * 12 cells
    assert int((g == W).sum()) == vol
    return g


