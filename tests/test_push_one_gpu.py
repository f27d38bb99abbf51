"""The multi-GPU product kernels on ONE device: step_kernel<..., PUSH = 1> (halo rows stored straight into the
z-neighbour's ghost plane, arrival counters, bounded wait) and the peer-store frame composite, with real neighbours.

A driver box with a single B200 skips every test that needs >= 2 GPUs, so round 1's PUSH path had no oracle parity
there.  Here the neighbours are other slabs on the same device:
  * fs3d_create(n_gpus = k, devices = [0] * k, FS3D_FLAG_PEER_PUSH_SHARED_DEVICE): the in-process world, and
  * k fs3d_create_slab worlds wired with fs3d_slab_attach_local — exactly what ranks do through CUDA IPC, minus IPC.
Only the warps of a slab's two edge pairs ever wait, so several persistent kernels on one device cannot deadlock."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _passes(t0, n):
    """kernel passes fs3d_step(n) makes from step index t0: steps 2k, 2k + 1 fuse into one pass"""
    c, t, left = 0, t0, n
    while left:
        ns = 2 if (left >= 2 and t % 2 == 0) else 1
        t += ns
        left -= ns
        c += 1
    return c


def _split(nz, k):
    from fallingsand3d_b200.slab import slab_bounds
    return slab_bounds(nz, k)


# J = 1 grouped / J = 1 full warp / J = 2 / warp pairs (nx > 2048); odd and even plane counts; both x-offsets occur
@pytest.mark.parametrize("nslabs,dims,scene", [(2, (64, 16, 12), 3), (3, (256, 40, 13), 4), (4, (2048, 24, 16), 4),
                                               (3, (1024, 12, 9), 3), (2, (4096, 12, 10), 3), (3, (3104, 18, 11), 4)])
def test_inprocess_world_pushes_halos_on_one_device(fs3d, oracle, nslabs, dims, scene):
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, scene, 5)
    # FLAG_NO_FUSE4: this test is about the PUSH kernels of the one- and two-step passes (the four-step pass has its own below)
    with fs3d.VoxelWorld(nx, ny, nz, seed=9, devices=[0] * nslabs, flags=fs3d.FLAG_PEER_PUSH_SHARED_DEVICE | fs3d.FLAG_NO_FUSE4) as w:
        assert w.num_slabs == nslabs
        w.upload(g)
        t = 0
        for n in (1, 2, 3, 40, 2, 5, 80, 1):             # single passes, fused passes, long runs without a host sync
            ms, launches = w.step_timed(n)
            # one PUSH kernel per slab per pass and nothing else: the copy path launches 3 kernels per slab and pass
            assert launches == nslabs * _passes(t, n), (n, launches)
            oracle.run(g, 9, t, n)
            t += n
            assert w.digest() == oracle.digest(g), f"step {t}"
        assert np.array_equal(w.download(), g)


def test_inprocess_push_with_skipping_on_one_device(fs3d, oracle):
    nx, ny, nz = 64, 96, 48
    g = oracle.generate(nx, ny, nz, 1, 1)
    flags = fs3d.FLAG_PEER_PUSH_SHARED_DEVICE | fs3d.FLAG_SKIP_SETTLED
    with fs3d.VoxelWorld(nx, ny, nz, seed=2, devices=[0, 0, 0], flags=flags) as w:
        w.generate(fs3d.SCENE_SAND_BLOCK, 1)
        for t in range(0, 240, 6):
            w.step(6)
            oracle.run(g, 2, t, 6)
            assert w.digest() == oracle.digest(g), f"step {t + 6}"
        run, total = w.activity()
        assert run < total


class LocalRanks:
    """k slab worlds in this process, attached to each other like ranks (fs3d_slab_attach_local)."""

    def __init__(self, fs3d, nx, ny, nz, k, seed, flags=0):
        self.bounds = _split(nz, k)
        self.worlds = [fs3d.VoxelWorld(nx, ny, nz, seed=seed, flags=flags, slab=b) for b in self.bounds]
        for i, w in enumerate(self.worlds):
            w.slab_attach_local(self.worlds[i - 1] if i > 0 else None, self.worlds[i + 1] if i + 1 < k else None)
        self.fuse4 = all(w.slab_can_fuse4() for w in self.worlds)      # what ranks settle with an all-reduce
        for w in self.worlds:
            w.slab_allow_fuse4(self.fuse4)
        self.push()

    def push(self):
        for w in self.worlds:
            w.sync()
        for w in self.worlds:
            w.slab_push_halos()

    def upload(self, g):
        for w, (a, b) in zip(self.worlds, self.bounds):
            w.upload(np.ascontiguousarray(g[a:b]))
        self.push()

    def step(self, n):
        for w in self.worlds:             # every "rank" enqueues all its passes; the kernels pace each other
            w.step(n)

    def download(self):
        return np.concatenate([w.download() for w in self.worlds], axis=0)

    def digest(self):
        return sum(w.digest() for w in self.worlds) & 0xFFFFFFFFFFFFFFFF

    def close(self):
        for w in self.worlds:
            w.close()


@pytest.mark.parametrize("k,dims,scene", [(2, (128, 32, 20), 4), (3, (2048, 48, 26), 4), (3, (4096, 20, 14), 3), (3, (96, 9, 7), 3)])
def test_attached_slab_worlds_step_like_ranks(fs3d, oracle, k, dims, scene):
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, scene, 3)
    r = LocalRanks(fs3d, nx, ny, nz, k, seed=5)
    try:
        r.upload(g)
        t = 0
        for n in (80, 1, 2, 3, 40, 7):                   # 40 fused passes back to back with no host sync, odd mixes after
            r.step(n)
            oracle.run(g, 5, t, n)
            t += n
            assert r.digest() == oracle.digest(g), f"step {t}"
        assert np.array_equal(r.download(), g)
    finally:
        r.close()


def test_attached_slab_worlds_composite_a_frame(fs3d, oracle):
    nx, ny, nz = 64, 40, 30
    g = oracle.generate(nx, ny, nz, 4, 2)
    r = LocalRanks(fs3d, nx, ny, nz, 3, seed=4)
    try:
        r.upload(g)
        r.step(10)
        oracle.run(g, 4, 0, 10)
        W, H = 192, 108
        cam = dict(pos=(0.2, -0.3, -1.3), yaw_deg=20.0, aspect=16.0 / 9.0)
        r.worlds[0].frame_export(W, H, 3)
        for i, w in enumerate(r.worlds):
            w.frame_attach_local(r.worlds[0], i)
        for _ in range(2):                                # the frame is reused from one image to the next
            for w in r.worlds:
                w.raymarch_to_frame(mode=fs3d.RM_VOXELS, **cam)
            for w in r.worlds:
                w.sync()
            img = r.worlds[0].frame_resolve(W, H)
            ref = oracle.raymarch(g, width=W, height=H, mode=1, **cam)
            assert np.array_equal(img, ref)
        # the borrowed frame must not be freed twice: close the borrowers first, then the owner
    finally:
        for w in r.worlds[1:]:
            w.close()
        r.worlds[0].close()


def test_push_watchdog_names_the_stalled_neighbour(fs3d, monkeypatch):
    monkeypatch.setenv("FS3D_PUSH_TIMEOUT_MS", "300")
    nx, ny, nz = 64, 16, 12
    r = LocalRanks(fs3d, nx, ny, nz, 2, seed=1)
    try:
        lo, hi = r.worlds
        lo.generate(fs3d.SCENE_RANDOM, 1)
        hi.generate(fs3d.SCENE_RANDOM, 1)
        r.push()
        lo.step(2)                     # pass 1 waits for nothing
        lo.sync()
        lo.step(2)                     # pass 2 needs the upper neighbour's pass 1, which never runs
        with pytest.raises(fs3d.Fs3dError) as ei:
            lo.sync()                  # returns (no hang) with the stalled side named
        assert ei.value.code == -5 and "its upper neighbour" in str(ei.value)
        with pytest.raises(fs3d.Fs3dError):
            lo.step(2)                 # a failed world refuses to step
        lo.download()                  # but can still be inspected and destroyed
    finally:
        r.close()


# ---- four steps per pass across slabs: two ghost planes per side, delivered by halo4_kernel after every pass ----
@pytest.mark.parametrize("nslabs,dims,scene", [(2, (1024, 16, 12), 3), (3, (2048, 40, 26), 4), (4, (2048, 24, 16), 4), (3, (1024, 70, 30), 4),
                                               (2, (2048, 9, 9), 3), (3, (4096, 20, 26), 4), (2, (4096, 33, 12), 3)])
def test_inprocess_world_four_step_passes_on_one_device(fs3d, oracle, nslabs, dims, scene):
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, scene, 5)
    with fs3d.VoxelWorld(nx, ny, nz, seed=9, devices=[0] * nslabs, flags=fs3d.FLAG_PEER_PUSH_SHARED_DEVICE) as w:
        w.upload(g)
        t = 0
        # four-step passes; a far-ghost refresh after one- and two-step passes and after an edit; long runs without a host sync
        for n in (4, 8, 1, 2, 5, 40, 3, 80, 4):
            ms, launches = w.step_timed(n)
            oracle.run(g, 9, t, n)
            t += n
            assert w.digest() == oracle.digest(g), f"step {t}"
            if n == 40:
                w.set_cell(5, ny - 1, nz // 2, fs3d.SAND)
                g[nz // 2, ny - 1, 5] = 1
        assert np.array_equal(w.download(), g)
    # the first four steps of a fresh world: one far-ghost refresh, one four-step kernel and one delivery per slab
    with fs3d.VoxelWorld(2048, 24, 16, seed=1, devices=[0] * 4, flags=fs3d.FLAG_PEER_PUSH_SHARED_DEVICE) as w:
        w.generate(fs3d.SCENE_MIXED_NOISE, 1)
        assert w.step_timed(4)[1] == 3 * 4
        assert w.step_timed(8)[1] == 2 * 2 * 4          # afterwards kernel + delivery


@pytest.mark.parametrize("k,dims,scene", [(2, (1024, 32, 20), 4), (3, (2048, 48, 26), 4), (3, (1024, 9, 14), 3), (2, (4096, 24, 20), 4)])
def test_attached_slab_worlds_four_step_passes(fs3d, oracle, k, dims, scene):
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, scene, 3)
    r = LocalRanks(fs3d, nx, ny, nz, k, seed=5)
    try:
        assert r.fuse4
        r.upload(g)
        t = 0
        for n in (80, 1, 2, 3, 40, 7, 4):
            r.step(n)
            oracle.run(g, 5, t, n)
            t += n
            assert r.digest() == oracle.digest(g), f"step {t}"
        assert np.array_equal(r.download(), g)
    finally:
        r.close()
