"""Schedule version 2 (SCHEDULE.md §7): eight materials on three bit-planes (FS3D_FLAG_MATERIALS8).

CPU part: the version-2 oracle (scalar C, table driven) against hand-derived behaviour, against the numpy mask-algebra
restatement, against version 1 on version-1 scenes, and against committed golden digests.  GPU part: the Rules3 kernels
(csrc/bitslice3.cuh) against that oracle on every kernel shape."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
E, S, W, X, G, O, H, V = 0, 1, 2, 3, 4, 5, 6, 7          # V = GRAVEL
RANK = {G: 0, E: 1, O: 2, W: 3, H: 4, S: 5, V: 6}


# --------------------------------------------------------------------------------------------- CPU: the oracle
def test_v2_equals_v1_on_version1_materials(oracle):
    for scene, dims in ((1, (64, 40, 24)), (2, (64, 64, 32)), (3, (96, 24, 10)), (4, (128, 48, 16))):
        nx, ny, nz = dims
        a = oracle.generate(nx, ny, nz, scene, 2)
        b = a.copy()
        for t in range(60):
            ea = oracle.step(a, 9, t, version=1)
            eb = oracle.step(b, 9, t, version=2)
            assert ea == eb and np.array_equal(a, b), (scene, t)


def test_numpy_restatement_agrees_v2(oracle):
    from oracle import oracle_np
    for dims, scene in (((32, 9, 5), 5), ((64, 24, 12), 6), ((96, 7, 3), 5), ((32, 1, 2), 5)):
        nx, ny, nz = dims
        g = oracle.generate(nx, ny, nz, scene, 3)
        n = g.copy()
        for t in range(24):
            oracle.step(g, 11, t, version=2)
            oracle_np.step(n, 11, t, version=2)
            assert np.array_equal(g, n), (dims, t)


def test_conservation_and_scene_contents_v2(oracle):
    g = oracle.generate(64, 64, 64, 6, 1)
    h0 = oracle.histogram(g)
    assert all(h0[m] > 0 for m in range(8)) and h0[8:].sum() == 0
    oracle.run(g, 3, 0, 150, version=2)
    assert np.array_equal(oracle.histogram(g), h0)
    r = oracle.generate(64, 64, 64, 5, 1)
    n = 64 ** 3
    hr = oracle.histogram(r)
    assert abs(hr[E] / n - 0.5) < 0.01 and abs(hr[S] / n - 0.125) < 0.01 and abs(hr[G] / n - 0.0625) < 0.005 and hr[X] == 0


def test_gas_rises_and_stays_up(oracle):
    nx, ny, nz = 32, 14, 6
    g = np.zeros((nz, ny, nx), np.uint8)
    g[3, 0, 10] = G
    last = 0
    for t in range(40):
        oracle.step(g, 5, t, version=2)
        zs, ys, xs = np.nonzero(g == G)
        assert len(ys) == 1 and ys[0] >= last          # never sinks
        last = ys[0]
        if t >= ny:
            assert ys[0] == ny - 1                      # two cells per step, then it wanders under the lid


def shaft_scene8(nx=32, ny=40, nz=10, seed=4):
    rng = np.random.default_rng(seed)
    g = np.full((nz, ny, nx), X, np.uint8)
    g[0::2, :, 0::2] = rng.choice(np.array([E, S, W, G, O, H, V], np.uint8), size=(nz // 2, ny, nx // 2),
                                  p=[0.3, 0.12, 0.14, 0.1, 0.12, 0.1, 0.12])
    g[0::2, :, 0::2][rng.random((nz // 2, ny, nx // 2)) < 0.04] = X
    return g


def settle_closed_form8(g):
    """1 x 1 shafts: granular cells (SAND, GRAVEL) never pass each other (neither yields) but sink through everything
    else; the fluids above them end sorted by rank, densest first"""
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
                seg = col[y:y1].copy()
                gran = [m for m in seg if m in (S, V)]
                fluid = sorted([m for m in seg if m not in (S, V)], key=lambda m: -RANK[m])
                col[y:y1] = np.array(gran + fluid, np.uint8)
                y = y1
    return out


def test_settled_shafts_v2_match_closed_form(oracle):
    g0 = shaft_scene8()
    want = settle_closed_form8(g0)
    for seed in (7, 99):
        a = g0.copy()
        t = quiet = 0
        while quiet < 4 and t < 800:
            quiet = quiet + 1 if oracle.step(a, seed, t, version=2) == 0 else 0
            t += 1
        assert quiet == 4
        assert np.array_equal(a, want)


def test_gravel_never_slides_and_sand_does(oracle):
    nx, ny, nz = 32, 10, 8
    for m, slides in ((V, False), (S, True)):
        g = np.zeros((nz, ny, nx), np.uint8)
        g[4, 0:5, 12] = X                 # a 1 x 1 stone pillar
        g[4, 5, 12] = m                   # one grain on top of it
        oracle.run(g, 3, 0, 24, version=2)
        assert (g[4, 5, 12] == m) == (not slides), m
        assert int((g == m).sum()) == 1


def test_honey_spreads_slower_than_water(oracle):
    nx, ny, nz = 64, 6, 1
    ext = {}
    for m in (W, H):
        tot = 0
        for seed in range(12):
            g = np.zeros((nz, ny, nx), np.uint8)
            g[0, 0, :] = X
            g[0, 1:5, 30:34] = m          # a 4 x 4 blob on a floor (2-D world: nz = 1)
            oracle.run(g, seed, 0, 30, version=2)
            xs = np.nonzero((g == m).any(axis=(0, 1)))[0]
            tot += xs.max() - xs.min() + 1
        ext[m] = tot / 12
    assert ext[H] < ext[W] - 1.0, ext


def test_second_coin_is_a_fair_independent_bit(oracle):
    c1 = np.array([oracle.coin(1, 3, 0, x, 5, 7) for x in range(4096)])
    c2 = np.array([oracle.coin2(1, 3, 0, x, 5, 7) for x in range(4096)])
    assert abs(c2.mean() - 0.5) < 0.03 and abs((c1 & c2).mean() - 0.25) < 0.03


def test_golden_digests_v2(oracle):
    with open(os.path.join(GOLDEN, "digests_v2.json")) as f:
        gold = json.load(f)
    assert gold["schedule_version"] == 2
    for case in gold["cases"]:
        nx, ny, nz = case["dims"]
        g = oracle.generate(nx, ny, nz, case["scene"], case["scene_seed"])
        assert oracle.digest(g) == int(case["digest0"], 16)
        t = 0
        for upto, dg in case["digests"]:
            oracle.run(g, case["seed"], t, upto - t, version=2)
            t = upto
            assert oracle.digest(g) == int(dg, 16), f"{case['name']} step {upto}"
        assert [int(v) for v in oracle.histogram(g)[:8]] == case["histogram"]


def test_checkpoint_files_of_version2_worlds(tmp_path, oracle):
    from fallingsand3d_b200 import checkpoint
    g = oracle.generate(64, 12, 6, 6, 2)
    p = str(tmp_path / "v2.fs3d")
    checkpoint.write(p, g, step=7, seed=5, schedule_version=2)
    h, back = checkpoint.read(p)
    assert h["schedule_version"] == 2 and h["encoding"] == 2 and h["payload_bytes"] == g.size // 2 and h["step"] == 7
    assert np.array_equal(back, g)
    with pytest.raises(ValueError):
        checkpoint.write(str(tmp_path / "bad.fs3d"), g, schedule_version=1)      # codes 4-7 do not fit 2 bits


# --------------------------------------------------------------------------------------------- GPU: the kernels
def _run_and_compare(fs3d, oracle, nx, ny, nz, scene, seed, steps, every, flags=0, devices=None):
    g = oracle.generate(nx, ny, nz, scene, 5)
    with fs3d.VoxelWorld(nx, ny, nz, seed=seed, flags=fs3d.FLAG_MATERIALS8 | flags, devices=devices) as w:
        assert w.schedule_version == 2
        w.upload(g)
        assert w.digest() == oracle.digest(g)
        t = 0
        while t < steps:
            n = min(every, steps - t)
            w.step(n)
            oracle.run(g, seed, t, n, version=2)
            t += n
            got = w.download()
            if not np.array_equal(got, g):
                bad = np.argwhere(got != g)
                raise AssertionError(f"{nx}x{ny}x{nz} scene {scene}: mismatch after step {t}: {len(bad)} cells, first "
                                     f"(z,y,x)={bad[0].tolist()} got {got[tuple(bad[0])]} want {g[tuple(bad[0])]}")
        assert np.array_equal(w.histogram(), oracle.histogram(g))


DIMS = [(32, 8, 6), (64, 9, 5), (96, 7, 3), (32, 1, 1), (32, 2, 2), (256, 12, 9), (1024, 6, 5), (1056, 5, 4), (2048, 6, 4),
        (2080, 5, 3), (4096, 4, 3), (3104, 6, 5)]


@pytest.mark.gpu
@pytest.mark.parametrize("dims", DIMS)
def test_v2_kernels_every_step(fs3d, oracle, dims):
    nx, ny, nz = dims
    _run_and_compare(fs3d, oracle, nx, ny, nz, scene=5, seed=7, steps=12, every=1)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", DIMS)
@pytest.mark.parametrize("every", [2, 3, 5])
def test_v2_kernels_fused_passes(fs3d, oracle, dims, every):
    nx, ny, nz = dims
    _run_and_compare(fs3d, oracle, nx, ny, nz, scene=5 if every != 3 else 6, seed=7, steps=30, every=every)


@pytest.mark.gpu
def test_v2_many_warps_segments_and_full_rows(fs3d, oracle):
    _run_and_compare(fs3d, oracle, 64, 200, 40, scene=6, seed=3, steps=8, every=2)
    _run_and_compare(fs3d, oracle, 2048, 64, 10, scene=6, seed=4, steps=8, every=4)
    _run_and_compare(fs3d, oracle, 4096, 72, 14, scene=5, seed=9, steps=8, every=3)
    _run_and_compare(fs3d, oracle, 1024, 512, 8, scene=6, seed=2, steps=6, every=6)


@pytest.mark.gpu
def test_v2_world_on_version1_materials_equals_v1_world(fs3d, oracle):
    nx, ny, nz = 256, 96, 40
    g = oracle.generate(nx, ny, nz, 4, 2)
    with fs3d.VoxelWorld(nx, ny, nz, seed=5) as a, fs3d.VoxelWorld(nx, ny, nz, seed=5, flags=fs3d.FLAG_MATERIALS8) as b:
        a.generate(fs3d.SCENE_MIXED_NOISE, 2)
        b.generate(fs3d.SCENE_MIXED_NOISE, 2)
        for n in (1, 2, 7, 10, 64):
            a.step(n)
            b.step(n)
            assert a.digest() == b.digest()
        oracle.run(g, 5, 0, 84, version=1)
        assert np.array_equal(b.download(), g)


@pytest.mark.gpu
def test_v2_materials_api_and_errors(fs3d, oracle):
    with fs3d.VoxelWorld(64, 16, 8, seed=2) as w:                      # version 1: codes 4-7 stay reserved
        assert w.schedule_version == 1
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.set_cell(0, 0, 0, fs3d.GAS)
        assert ei.value.code == -3
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.generate(fs3d.SCENE_RANDOM8, 1)
        assert ei.value.code == -3
        bad = np.zeros(w.shape, np.uint8)
        bad[1, 2, 3] = fs3d.OIL
        with pytest.raises(fs3d.Fs3dError):
            w.upload(bad)
    with fs3d.VoxelWorld(64, 16, 8, seed=2, flags=fs3d.FLAG_MATERIALS8) as w:
        w.set_cell(3, 4, 5, fs3d.HONEY)
        assert w.get_cell(3, 4, 5) == fs3d.HONEY
        w.fill_box((0, 0, 0), (64, 1, 8), fs3d.GRAVEL)
        w.paint_sphere((30, 8, 4), 2, fs3d.GAS)
        h = w.histogram()
        assert h[fs3d.GRAVEL] == 64 * 8 and h[fs3d.HONEY] == 1 and h[fs3d.GAS] > 0
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.set_cell(0, 0, 0, 8)
        assert ei.value.code == -3
        bad = np.zeros(w.shape, np.uint8)
        bad[1, 2, 3] = 9
        with pytest.raises(fs3d.Fs3dError):
            w.upload(bad)
    for scene in (5, 6):
        for dims in [(64, 64, 64), (96, 40, 24)]:
            nx, ny, nz = dims
            with fs3d.VoxelWorld(nx, ny, nz, flags=fs3d.FLAG_MATERIALS8) as w:
                w.generate(scene, 11)
                assert np.array_equal(w.download(), oracle.generate(nx, ny, nz, scene, 11))


@pytest.mark.gpu
def test_v2_golden_digests_on_gpu(fs3d):
    with open(os.path.join(GOLDEN, "digests_v2.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        nx, ny, nz = case["dims"]
        with fs3d.VoxelWorld(nx, ny, nz, seed=case["seed"], flags=fs3d.FLAG_MATERIALS8) as w:
            w.generate(case["scene"], case["scene_seed"])
            assert w.digest() == int(case["digest0"], 16), case["name"]
            t = 0
            for upto, dg in case["digests"]:
                w.step(upto - t)
                t = upto
                assert w.digest() == int(dg, 16), f"{case['name']} step {upto}"
            assert [int(v) for v in w.histogram()[:8]] == case["histogram"]


@pytest.mark.gpu
@pytest.mark.parametrize("dims,scene,steps", [((128, 100, 40), 6, 200), ((2048, 70, 20), 6, 80), ((96, 33, 17), 5, 100)])
def test_v2_skip_on_equals_skip_off(fs3d, dims, scene, steps):
    nx, ny, nz = dims
    M8 = fs3d.FLAG_MATERIALS8
    with fs3d.VoxelWorld(nx, ny, nz, seed=3, flags=M8) as a, fs3d.VoxelWorld(nx, ny, nz, seed=3, flags=M8 | fs3d.FLAG_SKIP_SETTLED) as b:
        a.generate(scene, 7)
        b.generate(scene, 7)
        for t in range(0, steps, 10):
            a.step(10)
            b.step(10)
            assert a.digest() == b.digest(), f"step {t + 10}"
        assert np.array_equal(a.download(), b.download())


@pytest.mark.gpu
def test_v2_settles_shafts_to_the_closed_form(fs3d):
    g0 = shaft_scene8()
    want = settle_closed_form8(g0)
    nz, ny, nx = g0.shape
    with fs3d.VoxelWorld(nx, ny, nz, seed=7, flags=fs3d.FLAG_MATERIALS8 | fs3d.FLAG_SKIP_SETTLED) as w:
        w.upload(g0)
        for _ in range(60):
            w.step(20)
            if w.activity()[0] == 0:
                break
        assert w.activity()[0] == 0
        assert np.array_equal(w.download(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("nslabs,dims", [(2, (64, 16, 12)), (3, (256, 40, 13)), (3, (2048, 24, 16)), (2, (4096, 12, 10))])
@pytest.mark.parametrize("flags", [0, 8])       # copy path / PUSH kernels on a shared device
def test_v2_multislab_and_push(fs3d, oracle, nslabs, dims, flags):
    nx, ny, nz = dims
    import torch
    k = torch.cuda.device_count()
    devices = [0] * nslabs if flags else [i % k for i in range(nslabs)]
    _run_and_compare(fs3d, oracle, nx, ny, nz, scene=6, seed=9, steps=46, every=23, flags=flags, devices=devices)


@pytest.mark.gpu
def test_v2_checkpoint_resume_and_step_host(fs3d, oracle, tmp_path):
    from fallingsand3d_b200 import checkpoint
    nx, ny, nz = 128, 40, 18
    g = oracle.generate(nx, ny, nz, 6, 4)
    M8 = fs3d.FLAG_MATERIALS8
    p = str(tmp_path / "w.fs3d")
    with fs3d.VoxelWorld(nx, ny, nz, seed=6, flags=M8) as w:
        w.generate(fs3d.SCENE_MIXED8, 4)
        w.step(17)
        oracle.run(g, 6, 0, 17, version=2)
        w.save(p)
        h, cells = checkpoint.read(p)
        assert h["schedule_version"] == 2 and h["encoding"] == 2 and h["step"] == 17 and np.array_equal(cells, g)
        w.step(9)
        d26 = w.digest()
    with fs3d.VoxelWorld(nx, ny, nz, seed=1, flags=M8) as w:
        w.load(p)
        assert w.step_index == 17
        w.step(9)
        assert w.digest() == d26
        host = w.download()
        out = np.empty_like(host)
        oracle.run(g, 6, 17, 9, version=2)
        t = 26
        for n in (2, 1, 1, 2):
            if n == 2 and t % 2:
                n = 1
            w.step_host(host, out, n)
            oracle.run(g, 6, t, n, version=2)
            t += n
            assert np.array_equal(out, g)
            host, out = out, host
    with fs3d.VoxelWorld(nx, ny, nz, seed=1) as w:
        with pytest.raises(fs3d.Fs3dError):
            w.load(p)                                   # a version-1 world refuses a version-2 checkpoint


@pytest.mark.gpu
def test_v2_raymarch_shows_the_new_materials(fs3d, oracle):
    nx, ny, nz = 64, 48, 40
    g = oracle.generate(nx, ny, nz, 6, 3)
    with fs3d.VoxelWorld(nx, ny, nz, seed=1, flags=fs3d.FLAG_MATERIALS8) as w:
        w.upload(g)
        w.step(6)
        oracle.run(g, 1, 0, 6, version=2)
        for cam in (dict(pos=(0.0, 0.0, -1.6), aspect=16.0 / 9.0), dict(pos=(0.3, -0.25, -1.2), aspect=1.5, yaw_deg=17.0)):
            img = w.raymarch(width=160, height=90, mode=fs3d.RM_VOXELS, **cam)
            ref = oracle.raymarch(g, width=160, height=90, mode=1, **cam)
            assert np.array_equal(img, ref)
