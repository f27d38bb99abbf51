"""CPU tests pinning the ray-march oracle to the reference shader: the known-answer pixels that
SURVEY.md §8c derived from /root/reference/shaders/fs_raymarch.frag (850x450, aspect 1700/900,
camPos (0,0,-5), quad UVs of renderer.cpp:1253-1267).  These are the only reference-derived
fixtures that exist for this repo (the simulation itself has none)."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_table():
    with open(os.path.join(GOLDEN, "fs_raymarch_known_answers.json")) as f:
        return json.load(f)


def test_known_answer_pixels(oracle):
    tab = load_table()
    for row in tab["hits"]:
        d, red, iters = oracle.raymarch_pixel(row["px"], row["py"])
        assert np.allclose(d, row["dir"], atol=2e-6), row
        assert abs(red - row["red"]) < 1e-5, (row, red)
        assert iters == row["iters"], (row, iters)
    for px, py in tab["misses"]:
        _, red, iters = oracle.raymarch_pixel(px, py)
        assert red == 0.0 and iters > 6      # no hit: leaves by the t > 1000 test (fs_raymarch.frag:57-59) or the 64 cap


def test_whole_frame_statistics(oracle):
    tab = load_table()
    img, depth = oracle.raymarch(None, mode=0, with_depth=True)
    assert img.shape == (450, 850, 4)
    hit = np.isfinite(depth)
    assert int(hit.sum()) == tab["frame"]["hit_pixels"]
    # linear red summed over the frame, recomputed per pixel in float (image is 8-bit)
    total = 0.0
    ys, xs = np.nonzero(hit)
    for y, x in zip(ys.tolist(), xs.tolist()):
        total += oracle.raymarch_pixel(x, y)[1]
    assert abs(total - tab["frame"]["sum_red"]) < 0.05
    assert (img[..., 1] == 0).all() and (img[..., 2] == 0).all() and (img[..., 3] == 255).all()
    assert (img[~hit][:, 0] == 0).all()
    # 8-bit linear encode of the centre pixel
    assert img[225, 425, 0] == int(0.540158 * 255 + 0.5)


def test_srgb_encode_thresholds(oracle):
    lin = oracle.raymarch(None, mode=0)
    srgb = oracle.raymarch(None, mode=16)
    v = lin[..., 0].astype(np.float64) / 255.0
    want = np.where(v <= 0.0031308, 12.92 * v, 1.055 * np.power(v, 1 / 2.4) - 0.055) * 255.0
    assert np.abs(srgb[..., 0].astype(np.float64) - want).max() <= 2.0   # within the 8-bit linear quantisation
    assert (srgb[..., 0] >= lin[..., 0]).all()


def test_voxel_mode_slab_compositing(oracle):
    # rendering each z-slab separately and keeping, per pixel, the nearest hit == rendering the whole grid
    g = oracle.generate(64, 48, 40, 4, 3)
    cam = dict(pos=(0.3, -0.2, -1.4), yaw_deg=12.0, aspect=16.0 / 9.0, width=160, height=90, mode=1)
    full, dfull = oracle.raymarch(g, with_depth=True, **cam)
    assert np.isfinite(dfull).sum() > 2000
    parts = []
    for lo, hi in [(0, 14), (14, 26), (26, 40)]:
        parts.append(oracle.raymarch(np.ascontiguousarray(g[lo:hi]), nz_global=40, zlo=lo, with_depth=True, **cam))
    img = np.zeros_like(full)
    img[..., 3] = 255
    best = np.full(dfull.shape, np.inf, np.float32)
    for im, d in parts:
        closer = d < best
        img[closer] = im[closer]
        best = np.where(closer, d, best)
    assert np.array_equal(img, full) and np.array_equal(best, dfull)


def test_voxel_mode_sees_materials(oracle):
    g = np.zeros((32, 32, 32), np.uint8)
    g[10:22, 0:6, 10:22] = 3
    g[12:20, 6:10, 12:20] = 1
    img = oracle.raymarch(g, pos=(0.0, 0.0, -2.0), aspect=1.0, width=64, height=64, mode=1)
    assert img[..., :3].any()
    # grid +y is screen-up for this camera: the sand (higher y) is drawn above the stone
    rows_sand = np.nonzero((img[..., 0] > img[..., 2]).any(axis=1))[0]
    rows_stone = np.nonzero(((img[..., 0] > 0) & (np.abs(img[..., 0].astype(int) - img[..., 1].astype(int)) < 3)).any(axis=1))[0]
    assert rows_sand.size and rows_stone.size and rows_sand.min() < rows_stone.max()


# ---- pinned against the reference shader ITSELF -------------------------------------------------------------
# tests/golden/fs_raymarch_ref_frames.npz is the output of /root/reference/shaders/fs_raymarch.frag compiled as C++
# against the reference's vendored glm (oracle/_ref, tests/golden/make_raymarch_ref_golden.py).

def _ref_golden():
    z = np.load(os.path.join(GOLDEN, "fs_raymarch_ref_frames.npz"))
    for k in range(int(z["n"])):
        c = z[f"cam{k}"]
        cam = dict(pos=(float(c[0]), float(c[1]), float(c[2])), aspect=float(c[3]), width=int(c[4]), height=int(c[5]))
        yield cam, z[f"idx{k}"], z[f"red{k}"]


def encode8_linear(red):
    """float32 linear value -> the uint8 both the oracle and the CUDA kernel store (no sRGB)."""
    r = red.astype(np.float32)
    return np.where(r >= 1.0, 255, np.floor(r * np.float32(255.0) + np.float32(0.5))).astype(np.uint8)


def test_oracle_is_bit_identical_to_the_reference_shader_frames(oracle):
    for cam, idx, red in _ref_golden():
        w, h = cam["width"], cam["height"]
        img, depth = oracle.raymarch(None, mode=0, with_depth=True, **cam)
        hit = np.isfinite(depth).reshape(-1)
        assert np.array_equal(np.nonzero(hit)[0], idx), "hit mask differs from fs_raymarch.frag"
        want = np.zeros(w * h, np.uint8)
        want[idx] = encode8_linear(red)
        assert np.array_equal(img[..., 0].reshape(-1), want)
        assert not img[..., 1].any() and not img[..., 2].any() and (img[..., 3] == 255).all()
        # the linear float value itself, on a sample of the hit pixels: identical bits
        for i in idx[:: max(1, idx.size // 400)]:
            _, r, _ = oracle.raymarch_pixel(int(i % w), int(i // w), pos=cam["pos"], aspect=cam["aspect"], width=w, height=h)
            assert np.float32(r) == red[np.searchsorted(idx, i)]


def test_oracle_matches_live_reference_shader_when_built(oracle):
    # where /root/reference is mounted (or oracle/_ref travelled with the repo): other cameras, whole frames
    if oracle.ref_frame(width=8, height=8) is None:
        pytest.skip("oracle/_ref not available here (needs /root/reference)")
    rng = np.random.RandomState(5)
    for _ in range(4):
        pos = (float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-6, -0.7)))
        cam = dict(pos=pos, aspect=float(rng.uniform(0.8, 2.2)), width=int(rng.randint(40, 300)), height=int(rng.randint(40, 200)))
        f = oracle.ref_frame(**cam)
        img, depth = oracle.raymarch(None, mode=0, with_depth=True, **cam)
        assert np.array_equal(f[..., 0] > 0, np.isfinite(depth))
        assert np.array_equal(img[..., 0], np.where(f[..., 0] > 0, encode8_linear(f[..., 0]), 0))


# ---- independent checks of the voxel DDA (oracle/oracle_np_dda.py): closed-form axis rays and float64 sampling ----
def _dda_scene(oracle):
    g = oracle.generate(64, 48, 40, 4, 3)          # MIXED_NOISE: floor, obstacles, boxes, noise above
    g[:, :, 10:14] = 0                             # some empty columns (misses)
    return g


def check_axis_rays(render, oracle, g, n_cols=60):
    """render(pos, width, height, aspect) -> (img, depth); closed-form answer for the 1 x 1 image on the optical axis"""
    from oracle import oracle_np_dda as dda
    nz, ny, nx = g.shape
    pal = oracle.default_palette()
    rng = np.random.default_rng(1)
    seen_miss = seen_hit = 0
    for _ in range(n_cols):
        i, j = int(rng.integers(nx)), int(rng.integers(ny))
        rgb, t = dda.analytic_axis_ray(g, i, j, pal)
        img, depth = render(dda.column_camera(i, j, nx, ny, nz), 1, 1, 1.0)
        if rgb is None:
            assert np.isinf(depth[0, 0]) and tuple(img[0, 0, :3]) == (0, 0, 0), (i, j)
            seen_miss += 1
            continue
        assert depth[0, 0] == np.float32(t), (i, j, depth[0, 0], t)        # h is a power of two: exact in float32
        want = np.clip(np.floor(rgb * 255.0 + 0.5), 0, 255)
        assert np.all(np.abs(img[0, 0, :3].astype(np.int64) - want) <= 1), (i, j, img[0, 0], want)
        seen_hit += 1
    assert seen_hit > 20 and seen_miss > 0


def check_sampled_rays(render, oracle, g, cam, width, height, n_pix=150):
    """the DDA's hit cell (recovered from its depth) equals the first cell a float64 sampled walk enters"""
    from oracle import oracle_np_dda as dda
    nz, ny, nx = g.shape
    img, depth = render(cam["pos"], width, height, cam["aspect"], cam.get("yaw_deg", 0.0))
    rng = np.random.default_rng(2)
    h, e = dda.box(nx, ny, nz)
    checked = 0
    for _ in range(n_pix * 4):
        px, py = int(rng.integers(width)), int(rng.integers(height))
        d = dda.pixel_ray(px, py, width, height, cam["aspect"], cam.get("yaw_deg", 0.0))
        ref = dda.sample_march(g, cam["pos"], d)
        t = float(depth[py, px])
        if ref is None:
            assert np.isinf(t), (px, py)
            continue
        assert np.isfinite(t), (px, py, ref)
        if dda.entry_point_margin(cam["pos"], d, t, nx, ny, nz) < 0.02:
            continue                                  # grazing a cell edge: either neighbour is a legitimate answer
        # the cell just behind the DDA's entry point
        p = (np.asarray(cam["pos"]) + (t + 1e-4 * h) * d + e) / h
        cell = (int(np.floor(p[0])), ny - 1 - int(np.floor(p[1])), int(np.floor(p[2])))
        assert cell == ref[:3], (px, py, cell, ref)
        assert abs(t - ref[3]) < h / 16, (px, py, t, ref[3])
        assert tuple(img[py, px, :3]) != (0, 0, 0)
        checked += 1
        if checked >= n_pix:
            break
    assert checked >= 40


def _oracle_render(oracle, g):
    def render(pos, width, height, aspect, yaw_deg=0.0):
        return oracle.raymarch(g, pos=pos, yaw_deg=yaw_deg, aspect=aspect, width=width, height=height, mode=1, with_depth=True)
    return render


def test_voxel_dda_closed_form_axis_rays(oracle):
    g = _dda_scene(oracle)
    check_axis_rays(_oracle_render(oracle, g), oracle, g)


@pytest.mark.parametrize("cam", [dict(pos=(0.0, 0.0, -1.6), aspect=16.0 / 9.0), dict(pos=(0.3, -0.25, -1.2), aspect=1.5, yaw_deg=17.0),
                                 dict(pos=(-0.9, 0.2, -0.9), aspect=1.0, yaw_deg=40.0)])
def test_voxel_dda_agrees_with_float64_sampling(oracle, cam):
    g = _dda_scene(oracle)
    check_sampled_rays(_oracle_render(oracle, g), oracle, g, cam, 160, 90)
