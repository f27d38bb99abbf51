"""GPU parity tests: the CUDA step (through the C ABI) against the CPU oracle, bit-exact."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
E, S, W, X = 0, 1, 2, 3


def run_and_compare(fs3d, oracle, nx, ny, nz, scene, seed, steps, every=1, scene_seed=5):
    g = oracle.generate(nx, ny, nz, scene, scene_seed)
    with fs3d.VoxelWorld(nx, ny, nz, seed=seed) as w:
        w.upload(g)
        assert w.digest() == oracle.digest(g)
        t = 0
        while t < steps:
            n = min(every, steps - t)
            w.step(n)
            oracle.run(g, seed, t, n)
            t += n
            got = w.download()
            if not np.array_equal(got, g):
                bad = np.argwhere(got != g)
                raise AssertionError(f"{nx}x{ny}x{nz} scene {scene}: mismatch after step {t}: {len(bad)} cells, "
                                     f"first (z,y,x)={bad[0].tolist()} got {got[tuple(bad[0])]} want {g[tuple(bad[0])]}")
        assert w.step_index == steps
        assert w.digest() == oracle.digest(g)
        assert np.array_equal(w.histogram(), oracle.histogram(g))


# every kernel instantiation: J=1 with row groups (nx < 1024), J=1 full warp, J=2, J=2 x warp pair
# (nx > 2048: full, half-empty second warp, one word in the second warp); odd/even ny and nz; ny, nz = 1
@pytest.mark.parametrize("dims", [(32, 8, 6), (64, 9, 5), (96, 7, 3), (128, 16, 4), (32, 1, 1), (32, 2, 1), (32, 1, 2),
                                  (256, 12, 9), (1024, 6, 5), (1056, 5, 4), (2048, 6, 4), (2080, 4, 3), (4096, 4, 3),
                                  (3072, 5, 4), (4064, 3, 2)])
def test_small_grids_every_step(fs3d, oracle, dims):
    nx, ny, nz = dims
    run_and_compare(fs3d, oracle, nx, ny, nz, scene=3, seed=7, steps=12, every=1)     # single-step passes


@pytest.mark.parametrize("dims", [(32, 8, 6), (64, 9, 5), (96, 7, 3), (32, 1, 1), (32, 2, 2), (32, 3, 1),
                                  (256, 12, 9), (1024, 6, 5), (2048, 6, 4), (2080, 5, 3), (4096, 4, 3), (3104, 6, 5)])
@pytest.mark.parametrize("every", [2, 3, 5])
def test_small_grids_fused_passes(fs3d, oracle, dims, every):
    # step(2) = one fused pass; step(3) at even t = pair + single, at odd t = single + pair; ...
    nx, ny, nz = dims
    run_and_compare(fs3d, oracle, nx, ny, nz, scene=3, seed=7, steps=30, every=every)


def test_fused_equals_unfused(fs3d):
    nx, ny, nz = 256, 96, 40
    with fs3d.VoxelWorld(nx, ny, nz, seed=5) as a, fs3d.VoxelWorld(nx, ny, nz, seed=5, flags=fs3d.FLAG_NO_FUSE) as b:
        a.generate(fs3d.SCENE_MIXED_NOISE, 2)
        b.generate(fs3d.SCENE_MIXED_NOISE, 2)
        for n in (1, 2, 7, 10, 64):
            a.step(n)
            b.step(n)
            assert a.digest() == b.digest()
        assert np.array_equal(a.download(), b.download())


def test_many_warps_and_segments(fs3d, oracle):
    # enough rows that warps split marches into segments with lead-ins
    run_and_compare(fs3d, oracle, 64, 200, 40, scene=3, seed=3, steps=8, every=2)
    run_and_compare(fs3d, oracle, 2048, 64, 10, scene=4, seed=4, steps=8, every=4)
    # warp pairs (nx > 2048): many pairs per CTA, segments starting mid-march, both x-offsets
    run_and_compare(fs3d, oracle, 4096, 72, 14, scene=3, seed=9, steps=8, every=3)


def test_config1_64cubed_sand_block_500_steps(fs3d, oracle):
    # BASELINE config 1, digest compared every step
    nx = ny = nz = 64
    g = oracle.generate(nx, ny, nz, 1, 1)
    with fs3d.VoxelWorld(nx, ny, nz, seed=1) as w:
        w.generate(fs3d.SCENE_SAND_BLOCK, 1)
        assert np.array_equal(w.download(), g)
        for t in range(500):
            w.step(1)
            oracle.step(g, 1, t)
            assert w.digest() == oracle.digest(g), f"step {t + 1}"
        h = w.histogram()
        assert h[E] == 258048 and h[S] == 4096
        assert np.array_equal(w.download(), g)


def test_config2_256cubed_mixed_1000_steps(fs3d, oracle):
    # BASELINE config 2: digest every 10 steps, full compare at the end, histogram invariant
    n = 256
    g = oracle.generate(n, n, n, 2, 1)
    with fs3d.VoxelWorld(n, n, n, seed=1) as w:
        w.generate(fs3d.SCENE_MIXED, 1)
        assert w.digest() == oracle.digest(g)
        h0 = w.histogram()
        for t in range(0, 1000, 10):
            w.step(10)
            oracle.run(g, 1, t, 10)
            assert w.digest() == oracle.digest(g), f"step {t + 10}"
        assert np.array_equal(w.histogram(), h0)
        assert np.array_equal(w.download(), g)


def test_golden_digests_on_gpu(fs3d):
    with open(os.path.join(GOLDEN, "digests.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        nx, ny, nz = case["dims"]
        with fs3d.VoxelWorld(nx, ny, nz, seed=case["seed"]) as w:
            w.generate(case["scene"], case["scene_seed"])
            assert w.digest() == int(case["digest0"], 16), case["name"]
            t = 0
            for upto, dg in case["digests"]:
                w.step(upto - t)
                t = upto
                assert w.digest() == int(dg, 16), f"{case['name']} step {upto}"
            assert [int(v) for v in w.histogram()[:4]] == case["histogram"]


@pytest.mark.parametrize("scene", [1, 2, 3, 4])
def test_device_scene_generators_match_oracle(fs3d, oracle, scene):
    for dims in [(64, 64, 64), (96, 40, 24), (128, 33, 17)]:
        nx, ny, nz = dims
        with fs3d.VoxelWorld(nx, ny, nz) as w:
            w.generate(scene, 11)
            assert np.array_equal(w.download(), oracle.generate(nx, ny, nz, scene, 11))


def test_cell_access_and_errors(fs3d):
    with fs3d.VoxelWorld(64, 16, 8, seed=2) as w:
        assert w.get_cell(3, 4, 5) == E
        w.set_cell(3, 4, 5, S)
        assert w.get_cell(3, 4, 5) == S
        w.fill_box((0, 0, 0), (64, 1, 8), X)
        h = w.histogram()
        assert h[X] == 64 * 8 and h[S] == 1
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.set_cell(64, 0, 0, S)
        assert ei.value.code == -4
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.set_cell(0, 0, 0, 4)
        assert ei.value.code == -3
        bad = np.zeros(w.shape, np.uint8)
        bad[1, 2, 3] = 9
        before = w.download()
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.upload(bad)
        assert ei.value.code == -3
        assert np.array_equal(w.download(), before)   # a rejected upload leaves the world untouched
        v = w.volume_view(0)
        assert v["nx"] == 64 and v["z0"] == 0 and v["z1"] == 8 and v["pitch_z"] == 64 * 16 and v["dev_ptr"]
    with pytest.raises(fs3d.Fs3dError) as ei:
        fs3d.VoxelWorld(48, 8, 8)
    assert ei.value.code == -2


def test_full_size_properties_1024(fs3d):
    # size-independent properties at a BASELINE size the oracle cannot follow step by step:
    # exact conservation, determinism (same seed -> same digest), seed sensitivity
    n = 1024
    digs = []
    for seed in (1, 1, 2):
        with fs3d.VoxelWorld(n, n, n, seed=seed) as w:
            w.generate(fs3d.SCENE_RANDOM, 1)
            h0 = w.histogram()
            w.step(24)
            assert np.array_equal(w.histogram(), h0)
            digs.append(w.digest())
    assert digs[0] == digs[1] and digs[0] != digs[2]


def test_subvolume_of_large_grid_matches_oracle(fs3d, oracle):
    # 2048-wide rows (J = 2 kernel) at full x extent, short in y/z so the oracle finishes in seconds
    run_and_compare(fs3d, oracle, 2048, 96, 24, scene=3, seed=5, steps=16, every=8)


@pytest.mark.parametrize("dims", [(64, 40, 30), (2048, 24, 20), (256, 300, 9), (32, 5, 1)])
def test_step_host_streams_a_host_grid(fs3d, oracle, dims):
    # fs3d_step_host == upload + step + download, chunked over z-pairs with overlapped copies
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, 4, 6)
    host = g.copy()
    out = np.empty_like(host)
    with fs3d.VoxelWorld(nx, ny, nz, seed=11) as w:
        t = 0
        for n in (2, 1, 1, 2, 2, 1):          # 2-step calls only on even steps
            if n == 2 and t % 2:
                n = 1
            w.step_host(host, out, n)
            oracle.run(g, 11, t, n)
            t += n
            assert np.array_equal(out, g), f"after step {t}"
            assert np.array_equal(w.download(), g)      # the device copy is current too
            assert w.step_index == t
            host, out = out, host
        w.step(3)                                        # and ordinary stepping continues from there
        oracle.run(g, 11, t, 3)
        assert np.array_equal(w.download(), g)
        with pytest.raises(fs3d.Fs3dError):
            w.step_host(host, out, 3)
        bad = host.copy()
        bad[0, 0, 0] = 200
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.step_host(bad, out, 1)
        assert ei.value.code == -3
        for call in (lambda: w.slab_step_host_begin(host), lambda: w.slab_step_host(host, out, 1)):
            with pytest.raises(fs3d.Fs3dError) as ei:      # the per-rank variant needs attached neighbours
                call()
            assert ei.value.code == -7


@pytest.mark.parametrize("dims,packed", [((2048, 24, 20), False), ((1024, 40, 70), False), ((4096, 12, 34), False), ((2048, 16, 8), True),
                                         ((1024, 300, 18), True), ((64, 40, 30), False), ((128, 24, 10), True)])
def test_step_host_four_steps_per_call(fs3d, oracle, dims, packed, monkeypatch):
    # n = 4: the grid streams through in chunks of whole bands of the four-step kernel, uploads one plane ahead of the
    # downloads (tiny chunks here, so that every grid is cut into several); grids without that kernel (nx = 64, 128) take
    # two streamed two-step passes
    from fallingsand3d_b200 import checkpoint
    nx, ny, nz = dims
    monkeypatch.setenv("FS3D_HOST_CHUNK_BYTES", str(8 * nx * ny))      # one band (four z-pairs) per chunk
    g = oracle.generate(nx, ny, nz, 4, 6)
    enc = (lambda a: np.ascontiguousarray(checkpoint.pack2(a))) if packed else (lambda a: a.copy())
    dec = (lambda a: checkpoint.unpack2(a, g.size).reshape(g.shape)) if packed else (lambda a: a)
    host = enc(g)
    out = np.empty_like(host)
    with fs3d.VoxelWorld(nx, ny, nz, seed=11) as w:
        call = w.step_host_packed if packed else w.step_host
        t = 0
        for n in (4, 4, 2, 1, 1, 4):
            call(host, out, n)
            oracle.run(g, 11, t, n)
            t += n
            assert np.array_equal(dec(out), g), f"after step {t}"
            assert np.array_equal(w.download(), g)
            assert w.step_index == t
            host, out = out, host
        call(host, host, 4)                              # in place
        oracle.run(g, 11, t, 4)
        assert np.array_equal(dec(host), g)
        w.step(2)
        with pytest.raises(fs3d.Fs3dError):              # step index 22: not a multiple of four
            call(host, out, 4)


def test_paint_sphere_brush(fs3d, oracle):
    import torch
    k = torch.cuda.device_count()
    nx, ny, nz = 64, 40, 30
    zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    for devices in (None, [i % k for i in range(3)]):
        g = oracle.generate(nx, ny, nz, 2, 1)
        with fs3d.VoxelWorld(nx, ny, nz, seed=3, devices=devices) as w:
            w.upload(g)
            for (c, r, m, only_empty) in [((20, 30, 12), 6, S, False), ((-3, 20, 29), 9, W, True), ((40, 5, 10), 4, E, False),
                                          ((63, 39, 0), 0, X, False), ((200, 200, 200), 5, S, False)]:
                w.paint_sphere(c, r, m, only_empty)
                mask = (xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2 <= r * r
                if only_empty:
                    mask &= g == E
                g[mask] = m
                assert np.array_equal(w.download(), g), (c, r, m)
            w.step(12)                      # a painted world steps like an uploaded one (ghost planes were refreshed)
            oracle.run(g, 3, 0, 12)
            assert np.array_equal(w.download(), g)
            with pytest.raises(fs3d.Fs3dError) as ei:
                w.paint_sphere((1, 1, 1), 2, 9)
            assert ei.value.code == -3


def test_full_size_properties_2048(fs3d):
    # BASELINE's headline size: conservation, fused == unfused, skipping == not skipping, determinism
    n = 2048
    digs = {}
    for name, flags in (("fused", 0), ("unfused", fs3d.FLAG_NO_FUSE), ("skip", fs3d.FLAG_SKIP_SETTLED)):
        with fs3d.VoxelWorld(n, n, n, seed=1, flags=flags) as w:
            w.generate(fs3d.SCENE_MIXED_NOISE, 1)
            h0 = w.histogram()
            assert int(h0.sum()) == n ** 3
            w.step(13)                     # odd: pair passes plus a single pass
            assert np.array_equal(w.histogram(), h0)
            digs[name] = w.digest()
    assert digs["fused"] == digs["unfused"] == digs["skip"]


@pytest.mark.parametrize("dims,steps", [((2048, 2048, 4), 6), ((4096, 1024, 4), 6), ((1024, 1024, 8), 8)])
def test_full_extent_planes_match_oracle(fs3d, oracle, dims, steps):
    # full BASELINE x and y extents (every word position of a row, every march segment shape), few planes in z
    nx, ny, nz = dims
    run_and_compare(fs3d, oracle, nx, ny, nz, scene=4, seed=2, steps=steps, every=steps)


def test_headline_grid_2048_equals_oracle_cell_for_cell(fs3d, oracle):
    # north_star: "a 2048^3 mixed sand/water scene stepping bit-exact to the schedule reference".  The FULL headline
    # grid (bench.py's workload: MIXED_NOISE, scene seed 1, coin seed 1), two fused passes = 4 steps, every one of the
    # 8.6 G cells compared with the CPU oracle run on the host (8 GiB per array; ~10 s of oracle time on the box).
    n = 2048
    g = oracle.generate(n, n, n, 4, 1)
    with fs3d.VoxelWorld(n, n, n, seed=1) as w:
        w.generate(fs3d.SCENE_MIXED_NOISE, 1)
        assert w.digest() == oracle.digest(g)
        w.step(4)
        oracle.run(g, 1, 0, 4)
        got = w.download()
        assert w.digest() == oracle.digest(g)
    same = np.array_equal(got, g)
    if not same:
        bad = np.argwhere(got != g)
        raise AssertionError(f"2048^3: {len(bad)} cells differ after 4 steps, first (z,y,x)={bad[0].tolist()}")


def test_headline_grid_matches_oracle_digests(fs3d):
    # tests/golden/bench_digests.json: digests of the full 2048^3 bench workload computed by the CPU oracle (generator
    # beside it) at the step counts bench.py ends on; the GPU must reproduce every one, fused and unfused
    with open(os.path.join(GOLDEN, "bench_digests.json")) as f:
        gold = json.load(f)
    nx, ny, nz = gold["dims"]
    steps = sorted(int(k) for k in gold["digests"])
    for flags in (0, fs3d.FLAG_NO_FUSE):
        with fs3d.VoxelWorld(nx, ny, nz, seed=gold["seed"], flags=flags) as w:
            w.generate(gold["scene"], gold["scene_seed"])
            assert w.digest() == int(gold["digest0"], 16)
            t = 0
            for upto in steps:
                w.step(upto - t)
                t = upto
                assert w.digest() == int(gold["digests"][str(upto)], 16), f"step {upto} flags {flags}"
            assert [int(v) for v in w.histogram()[:4]] == gold["histogram"]


def test_gpu_settles_tie_free_scenes_like_the_sweep(fs3d, oracle):
    # north_star: "reach the same settled configurations as the [stand-in] sweep on deterministic scenes".
    # With settled-tile skipping on, activity() == 0 is the GPU's own statement that nothing can move any more.
    from tests.settle_scenes import basin_scene, settle_closed_form, shaft_scene
    for name, g0, want, limit in (("shafts", shaft_scene(), None, 800), ("basin", basin_scene(), None, 40000)):
        if name == "shafts":
            want = settle_closed_form(g0)
        else:
            want = np.zeros_like(g0)
            want[:, 0, :] = X
            want[:, 1:4, :] = W
        nz, ny, nx = g0.shape
        sweep = g0.copy()
        for i in range(limit):
            if oracle.sweep_step(sweep, with_lateral=1, parity=i) == 0:
                break
        assert np.array_equal(sweep, want), name
        with fs3d.VoxelWorld(nx, ny, nz, seed=7, flags=fs3d.FLAG_SKIP_SETTLED) as w:
            w.upload(g0)
            t = 0
            while t < limit:
                w.step(20)
                t += 20
                if w.activity()[0] == 0:
                    break
            assert w.activity()[0] == 0, f"{name}: GPU world not settled after {t} steps"
            got = w.download()
        assert np.array_equal(got, want), f"{name}: GPU settled state differs from the sweep's / the closed form"


@pytest.mark.parametrize("dims,m8", [((64, 40, 30), False), ((2048, 24, 20), False), ((256, 300, 9), False), ((32, 5, 1), False),
                                     ((128, 40, 18), True), ((2048, 12, 7), True)])
def test_step_host_packed_streams_a_packed_host_grid(fs3d, oracle, dims, m8):
    # the host keeps the grid in the checkpoint encoding (2 or 4 bits per voxel); only those bytes cross PCIe
    from fallingsand3d_b200 import checkpoint
    nx, ny, nz = dims
    version = 2 if m8 else 1
    pack, unpack = (checkpoint.pack4, checkpoint.unpack4) if m8 else (checkpoint.pack2, checkpoint.unpack2)
    g = oracle.generate(nx, ny, nz, 6 if m8 else 4, 6)
    host = np.ascontiguousarray(pack(g))
    out = np.empty_like(host)
    with fs3d.VoxelWorld(nx, ny, nz, seed=11, flags=fs3d.FLAG_MATERIALS8 if m8 else 0) as w:
        t = 0
        for n in (2, 1, 1, 2, 2, 1):
            if n == 2 and t % 2:
                n = 1
            w.step_host_packed(host, out, n)
            oracle.run(g, 11, t, n, version=version)
            t += n
            assert np.array_equal(unpack(out, g.size).reshape(g.shape), g), f"after step {t}"
            assert np.array_equal(w.download(), g)
            host, out = out, host
        w.step_host_packed(host, host, 1)                # in place
        oracle.run(g, 11, t, 1, version=version)
        assert np.array_equal(unpack(host, g.size).reshape(g.shape), g)
        assert np.array_equal(w.download_packed(), host)  # packed download / upload are the same encoding
        w.upload_packed(np.ascontiguousarray(pack(g[::-1].copy())))
        assert np.array_equal(w.download(), g[::-1])
        with pytest.raises(fs3d.Fs3dError):
            w.step_host_packed(host, out, 3)


# ---- four steps per pass (step4_kernel.cuh): rows of 1024 / 2048 / 4096 voxels (one, two, four warps per band), step index
# a multiple of four ----
@pytest.mark.parametrize("dims", [(1024, 6, 5), (2048, 6, 4), (1024, 1, 2), (2048, 2, 2), (1024, 9, 3), (2048, 40, 31), (1024, 70, 26),
                                  (2048, 33, 12), (1024, 300, 4), (4096, 6, 4), (4096, 2, 2), (4096, 37, 21), (4096, 120, 9)])
@pytest.mark.parametrize("every", [4, 8, 5, 13])
def test_four_step_passes_match_oracle(fs3d, oracle, dims, every):
    # every = 4, 8: nothing but four-step passes; 5, 13: four-step passes mixed with two-step and single passes
    nx, ny, nz = dims
    run_and_compare(fs3d, oracle, nx, ny, nz, scene=4 if ny > 8 else 3, seed=7, steps=2 * every + 8, every=every)


def test_four_step_passes_with_many_bands_and_segments(fs3d, oracle):
    # more bands than units and marches cut into segments with warm-up; 13 z-planes: odd count, partial last band
    run_and_compare(fs3d, oracle, 2048, 160, 50, scene=4, seed=3, steps=16, every=8)
    run_and_compare(fs3d, oracle, 1024, 520, 13, scene=3, seed=5, steps=12, every=12)


@pytest.mark.parametrize("dims", [(2048, 160, 50), (2048, 64, 130), (1024, 520, 13), (1024, 96, 120), (4096, 48, 66)])
def test_four_step_passes_with_grouped_bands(fs3d, oracle, dims, monkeypatch):
    # big single slabs give every CTA a span of (group of neighbouring bands x iteration) and stagger its units one
    # y-block each; FS3D_S4_GROUP_SPAN=1 forces that split onto grids the oracle finishes in seconds:
    # partial last groups, spans across group boundaries, units whose shifted segment is empty
    monkeypatch.setenv("FS3D_S4_GROUP_SPAN", "1")
    nx, ny, nz = dims
    run_and_compare(fs3d, oracle, nx, ny, nz, scene=4, seed=11, steps=12, every=4)
    monkeypatch.setenv("FS3D_S4_GROUP_SPAN", "0")       # never grouped: same result
    run_and_compare(fs3d, oracle, nx, ny, nz, scene=4, seed=11, steps=8, every=8)


def test_four_two_and_one_step_passes_agree(fs3d):
    for n in (1024, 2048, 4096):
        digs = {}
        for name, flags in (("four", 0), ("two", fs3d.FLAG_NO_FUSE4), ("one", fs3d.FLAG_NO_FUSE)):
            with fs3d.VoxelWorld(n, 256, 64, seed=5, flags=flags) as w:
                w.generate(fs3d.SCENE_MIXED_NOISE, 2)
                h0 = w.histogram()
                out = []
                for k in (4, 8, 3, 1, 12, 40):
                    ms, launches = w.step_timed(k)
                    out.append((w.digest(), launches))
                assert np.array_equal(w.histogram(), h0)
                digs[name] = out
        assert [d for d, _ in digs["four"]] == [d for d, _ in digs["two"]] == [d for d, _ in digs["one"]]
        # the first call of each world starts at step 0: 4 steps = 1, 2, 4 launches
        assert [digs[k][0][1] for k in ("four", "two", "one")] == [1, 2, 4]
