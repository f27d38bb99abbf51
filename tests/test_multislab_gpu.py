"""GPU tests of the in-process multi-slab world (fs3d_desc.n_gpus > 1: edge pairs first, P2P halo
copies on a side stream, interior overlapped).  With one GPU the slabs share the device
(devices=[0,0,...]) which exercises the same code; with >= 2 GPUs they are spread out."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch
    k = torch.cuda.device_count()
    return [i % k for i in range(n)]


# With >= nslabs GPUs the slabs sit on distinct devices and the kernels push their halos over peer memory
# (the default); FLAG_NO_PEER_PUSH (and any world whose slabs share a device) uses peer copies instead.
@pytest.mark.parametrize("flags", [0, 4])
@pytest.mark.parametrize("nslabs,dims", [(2, (64, 16, 12)), (3, (256, 10, 13)), (4, (2048, 8, 16)), (2, (32, 6, 2)),
                                         (2, (4096, 12, 10))])
def test_multislab_matches_oracle_and_single_slab(fs3d, oracle, nslabs, dims, flags):
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, 3, 5)
    with fs3d.VoxelWorld(nx, ny, nz, seed=9, devices=_devices(nslabs), flags=flags) as w:
        assert w.num_slabs == nslabs
        w.upload(g)
        for t in range(12):
            w.step(1)
            oracle.step(g, 9, t)
            assert w.digest() == oracle.digest(g), f"step {t + 1}"
        assert np.array_equal(w.download(), g)
        views = [w.volume_view(i) for i in range(nslabs)]
        assert views[0]["z0"] == 0 and views[-1]["z1"] == nz
        assert all(views[i]["z1"] == views[i + 1]["z0"] for i in range(nslabs - 1))


def test_multislab_async_pipelining_many_steps(fs3d, oracle):
    # many steps enqueued without host syncs: the event graph alone must keep the halos coherent
    nx, ny, nz = 128, 64, 40
    g = oracle.generate(nx, ny, nz, 4, 2)
    with fs3d.VoxelWorld(nx, ny, nz, seed=4, devices=_devices(4)) as w:
        w.generate(fs3d.SCENE_MIXED_NOISE, 2)
        w.step(60)
        oracle.run(g, 4, 0, 60)
        assert np.array_equal(w.download(), g)
        w.set_cell(5, 60, 9, fs3d.SAND)     # edit at a slab edge refreshes the neighbour's ghost
        w.set_cell(5, 60, 10, fs3d.WATER)
        g[9, 60, 5] = 1
        g[10, 60, 5] = 2
        w.step(30)
        oracle.run(g, 4, 60, 30)
        assert np.array_equal(w.download(), g)


def test_fresh_multislab_world_has_open_internal_boundaries(fs3d, oracle):
    # no generate/upload: a grain dropped next to a slab boundary must see EMPTY (not a STONE ghost) across it
    nx, ny, nz = 32, 12, 8
    g = np.zeros((nz, ny, nx), np.uint8)
    with fs3d.VoxelWorld(nx, ny, nz, seed=3, devices=_devices(2)) as w:
        for (x, y, z) in [(5, 10, 3), (6, 10, 3), (5, 11, 3), (7, 9, 4), (9, 10, 4)]:
            w.set_cell(x, y, z, fs3d.SAND)
            g[z, y, x] = 1
        w.step(30)
        oracle.run(g, 3, 0, 30)
        assert np.array_equal(w.download(), g)


def test_slab_protocol_single_rank(fs3d, oracle):
    # fs3d_create_slab over the whole z range + the edges/interior/finish protocol == fs3d_step
    nx, ny, nz = 64, 20, 14
    g = oracle.generate(nx, ny, nz, 3, 8)
    with fs3d.VoxelWorld(nx, ny, nz, seed=6, slab=(0, nz)) as w:
        w.upload(g)
        for t in range(10):
            w.slab_step_edges()
            w.slab_step_interior()
            w.slab_step_finish()
            oracle.step(g, 6, t)
        assert np.array_equal(w.download(), g)
        with pytest.raises(fs3d.Fs3dError):
            w.slab_step_interior()        # protocol misuse is an error, not silent
