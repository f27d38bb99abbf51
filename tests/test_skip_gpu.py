"""GPU tests of settled-tile skipping (FS3D_FLAG_SKIP_SETTLED): it must never change a result and
it must actually skip once a region is provably static (SCHEDULE.md §4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_skip_is_bit_exact_while_a_sand_block_settles(fs3d, oracle):
    nx, ny, nz = 64, 96, 48
    g = oracle.generate(nx, ny, nz, 1, 1)
    with fs3d.VoxelWorld(nx, ny, nz, seed=1, flags=fs3d.FLAG_SKIP_SETTLED) as w:
        w.generate(fs3d.SCENE_SAND_BLOCK, 1)
        fractions = []
        for t in range(400):
            w.step(1)
            oracle.step(g, 1, t)
            assert w.digest() == oracle.digest(g), f"step {t + 1}"
            run, total = w.activity()
            fractions.append(run / total)
        assert np.array_equal(w.download(), g)
        assert fractions[0] == 1.0            # nothing is known to be static at the start
        assert min(fractions[4:60]) < 1.0     # empty space far from the block is skipped early
        assert fractions[-1] == 0.0           # the settled pile costs nothing


@pytest.mark.parametrize("dims,scene,steps", [((128, 100, 40), 2, 300), ((2048, 70, 20), 4, 120), ((96, 33, 17), 3, 150),
                                              ((4096, 70, 12), 4, 60)])
def test_skip_on_equals_skip_off(fs3d, dims, scene, steps):
    nx, ny, nz = dims
    with fs3d.VoxelWorld(nx, ny, nz, seed=3) as a, fs3d.VoxelWorld(nx, ny, nz, seed=3, flags=fs3d.FLAG_SKIP_SETTLED) as b:
        a.generate(scene, 7)
        b.generate(scene, 7)
        saw_skip = False
        for t in range(0, steps, 10):
            a.step(10)
            b.step(10)
            assert a.digest() == b.digest(), f"step {t + 10}"
            run, total = b.activity()
            saw_skip = saw_skip or run < total
        assert np.array_equal(a.download(), b.download())
        if dims == (128, 100, 40):
            assert saw_skip                 # the empty upper air of the MIXED scene sleeps


def test_edits_wake_sleeping_tiles(fs3d, oracle):
    nx, ny, nz = 64, 128, 32
    g = np.zeros((nz, ny, nx), np.uint8)
    g[:, 0, :] = 3
    with fs3d.VoxelWorld(nx, ny, nz, seed=8, flags=fs3d.FLAG_SKIP_SETTLED) as w:
        w.upload(g)
        w.step(8)
        oracle.run(g, 8, 0, 8)
        assert w.activity()[0] == 0           # an empty box over a floor is asleep after 4 quiet steps
        w.set_cell(20, 120, 10, fs3d.SAND)    # drop a grain from the top
        w.fill_box((30, 100, 12), (34, 104, 16), fs3d.WATER)
        g[10, 120, 20] = 1
        g[12:16, 100:104, 30:34] = 2
        for t in range(8, 120):
            w.step(1)
            oracle.step(g, 8, t)
            assert w.digest() == oracle.digest(g), f"step {t + 1}"
        assert np.array_equal(w.download(), g)


def test_skip_with_slabs(fs3d, oracle):
    import torch
    k = torch.cuda.device_count()
    nx, ny, nz = 64, 96, 48
    g = oracle.generate(nx, ny, nz, 1, 1)
    with fs3d.VoxelWorld(nx, ny, nz, seed=2, flags=fs3d.FLAG_SKIP_SETTLED, devices=[i % k for i in range(3)]) as w:
        w.generate(fs3d.SCENE_SAND_BLOCK, 1)
        for t in range(0, 300, 5):
            w.step(5)
            oracle.run(g, 2, t, 5)
            assert w.digest() == oracle.digest(g), f"step {t + 5}"
        run, total = w.activity()
        assert run < total                    # interior tiles sleep; slab-edge tiles never do
