"""GPU tests: the CUDA ray-march must match the CPU restatement of fs_raymarch.frag pixel for pixel."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_sdf_sphere_mode_matches_reference_shader_model(fs3d, oracle):
    # the shader as shipped: reference camera, 850x450 (window.h:41), aspect 1700/900 (materials.cpp:540)
    with fs3d.VoxelWorld(32, 4, 4) as w:
        img, depth = w.raymarch(mode=fs3d.RM_SDF_SPHERE, with_depth=True)
        ref, dref = oracle.raymarch(None, mode=0, with_depth=True)
        assert np.array_equal(img, ref)
        assert np.array_equal(depth, dref)
        with open(os.path.join(GOLDEN, "fs_raymarch_known_answers.json")) as f:
            tab = json.load(f)
        assert int(np.isfinite(depth).sum()) == tab["frame"]["hit_pixels"]
        for row in tab["hits"]:
            assert img[row["py"], row["px"], 0] == int(np.float32(row["red"]) * 255 + 0.5) or \
                abs(int(img[row["py"], row["px"], 0]) - row["red"] * 255) <= 1
        srgb = w.raymarch(mode=fs3d.RM_SDF_SPHERE | fs3d.RM_SRGB)
        assert np.array_equal(srgb, oracle.raymarch(None, mode=16))


@pytest.mark.parametrize("cam", [
    dict(pos=(0.0, 0.0, -5.0), yaw_deg=0.0, aspect=1700.0 / 900.0, width=850, height=450),
    dict(pos=(0.3, -0.2, -1.4), yaw_deg=12.0, aspect=16.0 / 9.0, width=320, height=180),
    dict(pos=(0.05, 0.02, 0.01), yaw_deg=-140.0, aspect=1.0, width=128, height=128),      # camera inside the volume
    dict(pos=(0.0, -3.0, -0.1), yaw_deg=0.0, aspect=2.0, width=200, height=100),
])
def test_voxel_mode_pixel_exact(fs3d, oracle, cam):
    nx, ny, nz = 96, 64, 80
    g = oracle.generate(nx, ny, nz, 4, 3)
    with fs3d.VoxelWorld(nx, ny, nz, seed=2) as w:
        w.upload(g)
        w.step(20)
        oracle.run(g, 2, 0, 20)
        for mode in (fs3d.RM_VOXELS, fs3d.RM_VOXELS | fs3d.RM_SRGB):
            img, depth = w.raymarch(mode=mode, with_depth=True, **cam)
            ref, dref = oracle.raymarch(g, mode=mode, with_depth=True, **cam)
            assert np.array_equal(depth, dref), f"{(depth != dref).sum()} depth pixels differ"
            assert np.array_equal(img, ref), f"{(img != ref).any(axis=-1).sum()} pixels differ"


def test_palette_and_multislab_raymarch(fs3d, oracle):
    import torch
    nx, ny, nz = 64, 48, 40
    g = oracle.generate(nx, ny, nz, 4, 3)
    pal = np.random.RandomState(1).rand(256, 4).astype(np.float32)
    cam = dict(pos=(0.3, -0.2, -1.4), yaw_deg=12.0, aspect=16.0 / 9.0, width=160, height=90)
    ref = oracle.raymarch(g, mode=1, palette=pal, **cam)
    k = torch.cuda.device_count()
    with fs3d.VoxelWorld(nx, ny, nz, devices=[i % k for i in range(3)]) as w:
        w.upload(g)
        w.set_palette(pal)
        assert np.array_equal(w.raymarch(mode=fs3d.RM_VOXELS, **cam), ref)
    # one rank's slab renders only its planes; min-depth compositing reproduces the whole image
    parts = []
    for lo, hi in [(0, 14), (14, 40)]:
        with fs3d.VoxelWorld(nx, ny, nz, slab=(lo, hi)) as w:
            w.upload(np.ascontiguousarray(g[lo:hi]))
            w.set_palette(pal)
            parts.append(w.raymarch(mode=fs3d.RM_VOXELS, with_depth=True, **cam))
    img = np.zeros_like(ref)
    img[..., 3] = 255
    best = np.full(ref.shape[:2], np.inf, np.float32)
    for im, d in parts:
        closer = d < best
        img[closer] = im[closer]
        best = np.where(closer, d, best)
    assert np.array_equal(img, ref)


def test_frame_path_single_rank_equals_raymarch(fs3d, oracle):
    # fs3d_frame_export / attach / raymarch_to_frame / resolve with one slot == fs3d_raymarch
    # (the multi-rank composite over peer memory is exercised by tests/run_slab_ranks.py)
    nx, ny, nz = 64, 48, 40
    g = oracle.generate(nx, ny, nz, 4, 3)
    cam = dict(pos=(0.3, -0.2, -1.4), yaw_deg=12.0, aspect=16.0 / 9.0)
    with fs3d.VoxelWorld(nx, ny, nz, slab=(0, nz)) as w:
        w.upload(g)
        with pytest.raises(fs3d.Fs3dError):
            w.raymarch_to_frame(**cam)                 # no frame yet
        blob = w.frame_export(160, 90, 1)
        w.frame_attach(blob, 0)
        with pytest.raises(fs3d.Fs3dError):
            w.frame_attach(blob, 1)                    # slot outside the frame
        for mode in (fs3d.RM_VOXELS, fs3d.RM_VOXELS | fs3d.RM_SRGB, fs3d.RM_SDF_SPHERE):
            w.raymarch_to_frame(mode=mode, **cam)
            w.sync()
            img = w.frame_resolve(160, 90)
            assert np.array_equal(img, w.raymarch(mode=mode, width=160, height=90, **cam))
            assert np.array_equal(img, oracle.raymarch(g if mode != fs3d.RM_SDF_SPHERE else None, mode=mode, width=160, height=90, **cam))


def test_one_slab_per_gpu_marches_locally_and_composites(fs3d, oracle):
    # >= 2 GPUs: every device marches its own slab into a frame on the first device (peer stores), which keeps
    # the nearest hit per pixel; image and depth must equal the single-world / oracle result
    import torch
    k = torch.cuda.device_count()
    if k < 2:
        pytest.skip("needs >= 2 GPUs")
    nx, ny, nz = 96, 64, 80
    g = oracle.generate(nx, ny, nz, 4, 3)
    pal = np.random.RandomState(2).rand(256, 4).astype(np.float32)
    with fs3d.VoxelWorld(nx, ny, nz, seed=2, devices=list(range(min(k, 4)))) as w:
        w.upload(g)
        w.step(10)
        oracle.run(g, 2, 0, 10)
        for cam in (dict(pos=(0.3, -0.2, -1.4), yaw_deg=12.0, aspect=16.0 / 9.0, width=320, height=180),
                    dict(pos=(0.05, 0.02, 0.01), yaw_deg=-140.0, aspect=1.0, width=128, height=128)):
            for mode in (fs3d.RM_VOXELS, fs3d.RM_VOXELS | fs3d.RM_SRGB):
                img, depth = w.raymarch(mode=mode, with_depth=True, **cam)
                ref, dref = oracle.raymarch(g, mode=mode, with_depth=True, **cam)
                assert np.array_equal(depth, dref) and np.array_equal(img, ref)
        w.set_palette(pal)
        cam = dict(pos=(0.3, -0.2, -1.4), yaw_deg=12.0, aspect=16.0 / 9.0, width=160, height=90)
        assert np.array_equal(w.raymarch(mode=fs3d.RM_VOXELS, **cam), oracle.raymarch(g, mode=1, palette=pal, **cam))


def test_sdf_mode_equals_the_reference_shader_frames(fs3d):
    # the CUDA kernel against the output of the reference's own fs_raymarch.frag (compiled with its glm; golden
    # frames committed under tests/golden/): same hit pixels, same 8-bit colour, pixel for pixel
    from tests.test_raymarch_oracle import _ref_golden, encode8_linear
    with fs3d.VoxelWorld(32, 4, 4) as w:
        for cam, idx, red in _ref_golden():
            wd, ht = cam["width"], cam["height"]
            img, depth = w.raymarch(mode=fs3d.RM_SDF_SPHERE, with_depth=True, yaw_deg=0.0, **cam)
            assert np.array_equal(np.nonzero(np.isfinite(depth).reshape(-1))[0], idx)
            want = np.zeros(wd * ht, np.uint8)
            want[idx] = encode8_linear(red)
            assert np.array_equal(img[..., 0].reshape(-1), want)
            assert not img[..., 1].any() and not img[..., 2].any() and (img[..., 3] == 255).all()


# ---- the same independent checks (closed-form axis rays, float64 sampled walk) against the CUDA kernel itself ----
def _gpu_render(fs3d, g, seed=1):
    w = fs3d.VoxelWorld(g.shape[2], g.shape[1], g.shape[0], seed=seed)
    w.upload(g)

    def render(pos, width, height, aspect, yaw_deg=0.0):
        return w.raymarch(pos=pos, yaw_deg=yaw_deg, aspect=aspect, width=width, height=height, mode=fs3d.RM_VOXELS, with_depth=True)
    return w, render


def test_cuda_voxel_dda_closed_form_axis_rays(fs3d, oracle):
    from tests.test_raymarch_oracle import _dda_scene, check_axis_rays
    g = _dda_scene(oracle)
    w, render = _gpu_render(fs3d, g)
    try:
        check_axis_rays(render, oracle, g)
    finally:
        w.close()


def test_cuda_voxel_dda_agrees_with_float64_sampling(fs3d, oracle):
    from tests.test_raymarch_oracle import _dda_scene, check_sampled_rays
    g = _dda_scene(oracle)
    w, render = _gpu_render(fs3d, g)
    try:
        for cam in (dict(pos=(0.0, 0.0, -1.6), aspect=16.0 / 9.0), dict(pos=(0.3, -0.25, -1.2), aspect=1.5, yaw_deg=17.0)):
            check_sampled_rays(render, oracle, g, cam, 160, 90)
    finally:
        w.close()


# ---- empty-space skipping: jumps over empty 8^3 bricks and over other ranks' planes must not change a single bit ----
CAMS = [dict(pos=(0.0, 0.0, -1.6), aspect=16.0 / 9.0), dict(pos=(0.3, -0.25, -1.2), aspect=1.5, yaw_deg=17.0),
        dict(pos=(-0.9, 0.2, -0.9), aspect=1.0, yaw_deg=40.0), dict(pos=(0.05, 0.9, -0.2), aspect=1.0, yaw_deg=3.0),
        dict(pos=(0.0, 0.0, 0.0), aspect=1.3, yaw_deg=180.0)]       # the last one sits inside the volume, looking back


@pytest.mark.parametrize("dims,scene", [((64, 48, 40), 2), ((96, 100, 72), 1), ((64, 40, 30), 3), ((256, 60, 36), 4), ((32, 9, 5), 3)])
def test_brick_skipping_is_exact(fs3d, oracle, dims, scene):
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, scene, 3)
    with fs3d.VoxelWorld(nx, ny, nz, seed=1) as w:
        w.upload(g)
        for k, cam in enumerate(CAMS):
            ref, rdepth = oracle.raymarch(g, width=160, height=90, mode=1, with_depth=True, **cam)
            for mode in (fs3d.RM_VOXELS | fs3d.RM_BRICKS, fs3d.RM_VOXELS | fs3d.RM_NO_BRICKS, fs3d.RM_VOXELS):
                img, depth = w.raymarch(width=160, height=90, mode=mode, with_depth=True, **cam)
                assert np.array_equal(img, ref), (k, mode)
                assert np.array_equal(depth, rdepth), (k, mode)
        # the map follows the cells: step, edit, march again
        w.step(7)
        oracle.run(g, 1, 0, 7)
        w.fill_box((8, 2, 3), (24, 7, 5), fs3d.STONE)
        g[3:5, 2:7, 8:24] = 3
        img = w.raymarch(width=160, height=90, mode=fs3d.RM_VOXELS | fs3d.RM_BRICKS, **CAMS[1])
        assert np.array_equal(img, oracle.raymarch(g, width=160, height=90, mode=1, **CAMS[1]))


def test_brick_skipping_on_slabs_and_foreign_planes(fs3d, oracle):
    from tests.test_push_one_gpu import LocalRanks
    nx, ny, nz = 64, 40, 30
    g = oracle.generate(nx, ny, nz, 2, 2)          # MIXED: mostly air
    # one world, three slabs on this device: one kernel marches all of them, each with its own brick map
    with fs3d.VoxelWorld(nx, ny, nz, seed=4, devices=[0, 0, 0]) as w:
        w.upload(g)
        for cam in CAMS:
            ref = oracle.raymarch(g, width=192, height=108, mode=1, **cam)
            for mode in (fs3d.RM_VOXELS | fs3d.RM_BRICKS, fs3d.RM_VOXELS | fs3d.RM_NO_BRICKS):
                assert np.array_equal(w.raymarch(width=192, height=108, mode=mode, **cam), ref)
    # three attached slab worlds: each marches only its own planes and jumps over the others' in one move
    r = LocalRanks(fs3d, nx, ny, nz, 3, seed=4)
    try:
        r.upload(g)
        W, H = 192, 108
        r.worlds[0].frame_export(W, H, 3)
        for i, w in enumerate(r.worlds):
            w.frame_attach_local(r.worlds[0], i)
        for cam in CAMS:
            ref = oracle.raymarch(g, width=W, height=H, mode=1, **cam)
            for mode in (fs3d.RM_VOXELS | fs3d.RM_BRICKS, fs3d.RM_VOXELS | fs3d.RM_NO_BRICKS):
                for w in r.worlds:
                    w.raymarch_to_frame(mode=mode, **cam)
                for w in r.worlds:
                    w.sync()
                assert np.array_equal(r.worlds[0].frame_resolve(W, H), ref)
    finally:
        for w in r.worlds[1:]:
            w.close()
        r.worlds[0].close()


def test_adaptive_bricks_turn_on_for_air_and_off_for_dense_scenes(fs3d):
    # the decision comes from the previous frame's steps per ray; either way the image is the same
    n = 256
    cam = dict(pos=(0.0, 0.0, -1.6), aspect=16.0 / 9.0, width=320, height=180)
    for scene, expect_on in ((fs3d.SCENE_SAND_BLOCK, True), (fs3d.SCENE_RANDOM, False)):
        with fs3d.VoxelWorld(n, n, n, seed=1) as w:
            w.generate(scene, 1)
            forced = w.raymarch(mode=fs3d.RM_VOXELS | fs3d.RM_NO_BRICKS, **cam)
            assert not w.raymarch_bricks_in_use()
            imgs = [w.raymarch(mode=fs3d.RM_VOXELS, **cam) for _ in range(3)]
            assert all(np.array_equal(i, forced) for i in imgs)
            assert w.raymarch_bricks_in_use() == expect_on, scene
