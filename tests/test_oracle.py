"""CPU tests of the oracle itself (SCHEDULE.md): hand-derived micro-scenes, the independent numpy
restatement, conservation, and the committed golden digests.  Parity is unpinned against the
reference (it has no simulation: SURVEY.md §0), so these are what pin the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import oracle_np

E, S, W, X = 0, 1, 2, 3
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def cells(grid, m):
    z, y, x = np.nonzero(grid == m)
    return sorted(zip(x.tolist(), y.tolist(), z.tolist()))


def test_single_grain_trajectory(oracle):
    # hand-derived (see tests/golden/README.md): y = 6 -> 5 -> 3 -> 1 -> 0 -> 0, x and z fixed
    g = np.zeros((4, 8, 32), np.uint8)
    g[3, 6, 5] = S
    ys = []
    for t in range(5):
        oracle.step(g, seed=123, t=t)
        (c,) = cells(g, S)
        assert (c[0], c[2]) == (5, 3)
        ys.append(c[1])
    assert ys == [5, 3, 1, 0, 0]


def test_grain_slides_off_pillar(oracle):
    # STONE (5,0,3), SAND on top: step 0 sub-step 1 is XY, ox = 0, oy = 0; block x∈{4,5}, y∈{0,1}:
    # a=E b=SAND c=E d=STONE -> D: heavier(b,c) and a != STONE -> sand goes to (4,0,3) and rests
    g = np.zeros((8, 8, 32), np.uint8)
    g[3, 0, 5] = X
    g[3, 1, 5] = S
    oracle.step(g, seed=9, t=0)
    assert cells(g, S) == [(4, 0, 3)] and cells(g, X) == [(5, 0, 3)]
    for t in range(1, 9):
        oracle.step(g, seed=9, t=t)
    assert cells(g, S) == [(4, 0, 3)]


def test_no_slide_past_stone_corner(oracle):
    # like above, but a STONE beside the grain (4,1,3): D is blocked in XY; it may still go in z
    g = np.zeros((8, 8, 32), np.uint8)
    g[3, 0, 5] = X
    g[3, 1, 5] = S
    g[3, 1, 4] = X
    g[3, 0, 4] = E
    oracle.step_range(g, 8, 0, 0, 8, 9, 0)  # whole grid through the range entry point
    # XY sub-step could not move it (a = STONE); ZY sub-step 2 has oy = 1 so y=1 is a LOWER row: stays
    assert cells(g, S) == [(5, 1, 3)]


def test_water_lateral_uses_coin(oracle):
    # water on the floor at (6,0,2): step 0, sub-step 2 = ZY, oz = 0, oy = 1: block z∈{2,3}, y∈{-1,0};
    # upper row y=0: a=W b=E -> moves to z=3 iff coin(ZY, t=0, X=6, Y=0, Z=2)
    for seed in range(1, 9):
        g = np.zeros((4, 4, 32), np.uint8)
        g[2, 0, 6] = W
        oracle.step(g, seed=seed, t=0)
        k = oracle_np.key(seed, 0, 1)
        c = bool(oracle_np.coins(k, np.array([6]), np.array([0]), np.array([2]))[0])
        assert c == bool(oracle.coin(seed, 0, 1, 6, 0, 2))
        assert cells(g, W) == [(6, 0, 3 if c else 2)]


def test_sand_sinks_in_water_and_stone_is_static(oracle):
    g = np.zeros((4, 8, 32), np.uint8)
    g[:, 0, :] = X
    g[1, 1:4, 10] = W
    g[1, 4, 10] = S
    stone_before = g == X
    for t in range(40):
        oracle.step(g, 1, t)
        assert np.array_equal(g == X, stone_before)
    assert cells(g, S) == [(10, 1, 1)]          # the grain reached the floor through the water column
    assert (g == W).sum() == 3


@pytest.mark.parametrize("shape,scene", [((6, 8, 32), 3), ((5, 9, 64), 3), ((16, 16, 32), 4), ((3, 7, 96), 3),
                                         ((1, 5, 32), 3), ((7, 1, 32), 3)])
def test_c_oracle_equals_numpy_restatement(oracle, shape, scene):
    nz, ny, nx = shape
    g = oracle.generate(nx, ny, nz, scene, 5)
    h = g.copy()
    for t in range(12):
        oracle.step(g, 7, t)
        oracle_np.step(h, 7, t)
        assert np.array_equal(g, h), f"step {t}"
    assert oracle.digest(g) == oracle_np.digest(g)


def test_conservation_and_determinism(oracle):
    g = oracle.generate(64, 32, 24, 4, 3)
    h0 = oracle.histogram(g)
    a = g.copy()
    b = g.copy()
    for t in range(30):
        oracle.step(a, 11, t)
        assert np.array_equal(oracle.histogram(a), h0)
    oracle.run(b, 11, 0, 30)
    assert np.array_equal(a, b)
    c = g.copy()
    oracle.run(c, 12, 0, 30)
    assert not np.array_equal(a, c)   # the seed matters (water coins)


def test_slab_range_steps_equal_whole_grid(oracle):
    # two slabs with ghost planes, exchanged by hand, must reproduce the whole-grid result
    nx, ny, nz = 32, 12, 10
    whole = oracle.generate(nx, ny, nz, 3, 2)
    cuts = [(0, 4), (4, 10)]
    slabs = []
    for lo, hi in cuts:
        a = np.full((hi - lo + 2, ny, nx), X, np.uint8)
        a[1:-1] = whole[lo:hi]
        slabs.append(a)
    for t in range(8):
        # halo exchange of the current state
        slabs[0][-1] = slabs[1][1]
        slabs[1][0] = slabs[0][-2]
        for (lo, hi), a in zip(cuts, slabs):
            oracle.step_range(a, nz, lo - 1, lo, hi, 77, t)
        oracle.step(whole, 77, t)
        got = np.concatenate([slabs[0][1:-1], slabs[1][1:-1]])
        assert np.array_equal(got, whole), f"step {t}"


def test_scene_counts(oracle):
    g = oracle.generate(64, 64, 64, 1, 1)
    h = oracle.histogram(g)
    assert h[E] == 258048 and h[S] == 4096          # BASELINE config 1
    g = oracle.generate(64, 64, 64, 2, 1)
    assert set(np.unique(g).tolist()) == {E, S, W, X}
    g = oracle.generate(64, 64, 64, 3, 1)
    h = oracle.histogram(g)
    n = 64 ** 3
    assert abs(h[S] / n - 0.25) < 0.01 and abs(h[W] / n - 0.25) < 0.01 and h[X] == 0


def test_golden_digests(oracle):
    with open(os.path.join(GOLDEN, "digests.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        nx, ny, nz = case["dims"]
        g = oracle.generate(nx, ny, nz, case["scene"], case["scene_seed"])
        assert oracle.digest(g) == int(case["digest0"], 16)
        t = 0
        for upto, dg in case["digests"]:
            oracle.run(g, case["seed"], t, upto - t)
            t = upto
            assert oracle.digest(g) == int(dg, 16), f"{case['name']} step {upto}"
        assert [int(v) for v in oracle.histogram(g)[:4]] == case["histogram"]


def test_settled_state_is_fixed_point_of_standin_sweep(oracle):
    # SCHEDULE.md §6: order-independent facts shared with the builder-written sequential sweep
    g = oracle.generate(32, 24, 16, 1, 1)   # sand block only: settles completely
    t = 0
    quiet = 0
    while quiet < 4 and t < 400:
        quiet = quiet + 1 if oracle.step(g, 5, t) == 0 else 0
        t += 1
    assert quiet == 4, "sand block should settle"
    before = g.copy()
    assert oracle.sweep_step(g, with_lateral=0, parity=0) == 0
    assert np.array_equal(g, before)
    # and the sweep conserves counts from the same start
    s = oracle.generate(32, 24, 16, 1, 1)
    h0 = oracle.histogram(s)
    for i in range(50):
        oracle.sweep_step(s, with_lateral=1, parity=i)
    assert np.array_equal(oracle.histogram(s), h0)


# ---- settled configurations on deterministic (tie-free) scenes: schedule == stand-in sweep == closed form ----------
from tests.settle_scenes import basin_scene, settle_closed_form, shaft_scene  # noqa: E402


def run_schedule_until_quiet(oracle, g, seed, limit):
    t = quiet = 0
    while quiet < 4 and t < limit:
        quiet = quiet + 1 if oracle.step(g, seed, t) == 0 else 0
        t += 1
    assert quiet == 4, f"not settled after {limit} steps"
    return t


def run_sweep_until_quiet(oracle, g, with_lateral, limit):
    for i in range(limit):
        if oracle.sweep_step(g, with_lateral=with_lateral, parity=i) == 0:
            return i
    raise AssertionError(f"sweep not settled after {limit} sweeps")


def test_settled_shafts_equal_standin_sweep_and_closed_form(oracle):
    # north_star: "reach the same settled configurations as the [stand-in] sweep on deterministic scenes": MIXED
    # materials (stone + sand + water), tie-free by construction
    g0 = shaft_scene()
    want = settle_closed_form(g0)
    a, b = g0.copy(), g0.copy()
    run_schedule_until_quiet(oracle, a, seed=7, limit=600)
    run_sweep_until_quiet(oracle, b, with_lateral=1, limit=600)
    assert np.array_equal(a, want), "partitioned schedule: shafts not density-sorted"
    assert np.array_equal(b, want), "stand-in sweep: shafts not density-sorted"
    # and the seed (coins) cannot matter on a tie-free scene
    c = g0.copy()
    run_schedule_until_quiet(oracle, c, seed=12345, limit=600)
    assert np.array_equal(c, want)


def test_settled_basin_equals_standin_sweep(oracle):
    g0 = basin_scene()
    want = np.zeros_like(g0)
    want[:, 0, :] = X
    want[:, 1:4, :] = W
    a, b = g0.copy(), g0.copy()
    run_schedule_until_quiet(oracle, a, seed=3, limit=20000)
    run_sweep_until_quiet(oracle, b, with_lateral=1, limit=20000)
    assert np.array_equal(a, want), "partitioned schedule: pool not flat"
    assert np.array_equal(b, want), "stand-in sweep: pool not flat"
