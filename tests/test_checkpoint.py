"""Checkpoint format (include/fs3d.h "checkpoint"): the numpy reader/writer on CPU, and fs3d_save /
fs3d_load on the GPU against it.  The reference has no on-disk state."""
import os

import numpy as np
import pytest

from fallingsand3d_b200 import checkpoint as ck


def test_numpy_roundtrip_header_and_layout(tmp_path, oracle):
    nx, ny, nz = 64, 5, 7
    g = oracle.generate(nx, ny, nz, 4, 3)
    path = tmp_path / "a.fs3d"
    ck.write(path, g, step=13, seed=99)
    assert os.path.getsize(path) == 80 + nx * ny * nz // 4
    h, back = ck.read(path)
    assert np.array_equal(back, g)
    assert (h["nx"], h["ny"], h["nz"], h["z_begin"], h["z_end"]) == (nx, ny, nz, 0, nz)
    assert (h["step"], h["seed"], h["format_version"], h["schedule_version"], h["encoding"]) == (13, 99, 1, 1, 1)
    assert h["digest"] == oracle.digest(g) == ck.digest(g, nx, ny)
    # documented bit layout: cell i of the x-fastest stream sits in bits 2(i & 3) of payload byte i >> 2
    raw = open(path, "rb").read()[80:]
    flat = g.reshape(-1)
    for i in (0, 1, 2, 3, 4, 63, 64, 1000, flat.size - 1):
        assert (raw[i >> 2] >> (2 * (i & 3))) & 3 == flat[i]


def test_numpy_slab_digest_uses_global_indices(tmp_path, oracle):
    nx, ny, nz = 32, 6, 10
    g = oracle.generate(nx, ny, nz, 3, 1)
    ck.write(tmp_path / "lo", g[:4], nz=nz, z_begin=0)
    ck.write(tmp_path / "hi", g[4:], nz=nz, z_begin=4)
    hl, _ = ck.read(tmp_path / "lo")
    hh, gh = ck.read(tmp_path / "hi")
    assert (hl["digest"] + hh["digest"]) % 2 ** 64 == oracle.digest(g)     # digests of slabs sum to the world's
    assert hh["z_begin"] == 4 and hh["z_end"] == nz and hh["nz"] == nz and np.array_equal(gh, g[4:])


def test_numpy_reader_rejects_damage(tmp_path):
    g = np.zeros((2, 3, 32), np.uint8)
    g[1, 2, 5] = 2
    p = tmp_path / "c"
    ck.write(p, g)
    raw = bytearray(open(p, "rb").read())
    raw[85] ^= 0x10
    open(p, "wb").write(raw)
    with pytest.raises(ValueError, match="digest"):
        ck.read(p)
    open(p, "wb").write(raw[:-1])
    with pytest.raises(ValueError, match="payload"):
        ck.read(p)
    open(p, "wb").write(b"NOTACKPT" + raw[8:])
    with pytest.raises(ValueError, match="not an fs3d"):
        ck.read(p)
    with pytest.raises(ValueError):
        ck.write(p, np.full((1, 1, 32), 7, np.uint8))          # reserved material codes


GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _golden():
    import json
    with open(os.path.join(GOLDEN, "mixed_noise_64x32x16_t41.json")) as f:
        return json.load(f)


def test_golden_state_file_is_what_the_oracle_produces(oracle):
    # the committed checkpoint (tests/golden/make_golden.py) parses, verifies, and equals a fresh oracle run
    meta = _golden()
    h, g = ck.read(os.path.join(GOLDEN, meta["file"]))
    nx, ny, nz = meta["dims"]
    assert (h["nx"], h["ny"], h["nz"], h["step"], h["seed"]) == (nx, ny, nz, meta["step"], meta["seed"])
    ref = oracle.generate(nx, ny, nz, meta["scene"], meta["scene_seed"])
    oracle.run(ref, meta["seed"], 0, meta["step"])
    assert np.array_equal(g, ref)
    oracle.run(ref, meta["seed"], meta["step"], meta["resume_steps"])
    assert hex(oracle.digest(ref)) == meta["digest_after_resume"]
    assert [int(v) for v in oracle.histogram(ref)[:4]] == meta["histogram"]


@pytest.mark.gpu
def test_gpu_resumes_the_golden_state(fs3d):
    meta = _golden()
    nx, ny, nz = meta["dims"]
    with fs3d.VoxelWorld(nx, ny, nz, seed=1) as w:
        w.load(os.path.join(GOLDEN, meta["file"]))
        assert w.step_index == meta["step"]
        w.step(meta["resume_steps"])
        assert hex(w.digest()) == meta["digest_after_resume"]
        assert [int(v) for v in w.histogram()[:4]] == meta["histogram"]


@pytest.mark.gpu
def test_gpu_save_matches_numpy_reader_and_resume_is_bit_identical(tmp_path, fs3d, oracle):
    nx, ny, nz = 96, 40, 30
    path = str(tmp_path / "w.fs3d")
    with fs3d.VoxelWorld(nx, ny, nz, seed=21) as w:
        w.generate(fs3d.SCENE_MIXED_NOISE, 4)
        w.step(7)                                   # odd step index: the resumed run must keep the phase
        w.save(path)
        h, g = ck.read(path)
        assert np.array_equal(g, w.download())
        assert (h["step"], h["seed"], h["digest"]) == (7, 21, w.digest())
        w.step(9)
        want = w.download()
    with fs3d.VoxelWorld(nx, ny, nz, seed=5) as w2:    # another seed: load must restore the checkpoint's
        w2.load(path)
        assert w2.step_index == 7 and np.array_equal(w2.download(), g)
        w2.step(9)
        assert np.array_equal(w2.download(), want)
    ref = g.copy()
    oracle.run(ref, 21, 7, 9)
    assert np.array_equal(ref, want)


@pytest.mark.gpu
def test_gpu_load_of_numpy_written_state_and_errors(tmp_path, fs3d, oracle):
    nx, ny, nz = 64, 12, 9
    g = oracle.generate(nx, ny, nz, 3, 8)
    path = str(tmp_path / "n.fs3d")
    ck.write(path, g, step=4, seed=3)
    with fs3d.VoxelWorld(nx, ny, nz, seed=1) as w:
        w.load(path)
        assert np.array_equal(w.download(), g) and w.step_index == 4
        w.step(5)
        oracle.run(g, 3, 4, 5)
        assert np.array_equal(w.download(), g)
        before = w.download()
        raw = bytearray(open(path, "rb").read())
        raw[100] ^= 0x04
        bad = str(tmp_path / "bad.fs3d")
        open(bad, "wb").write(raw)
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.load(bad)
        assert ei.value.code == -8                       # digest mismatch
        assert np.array_equal(w.download(), before) and w.step_index == 9      # world untouched
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.load(str(tmp_path / "missing"))
        assert ei.value.code == -8
    with fs3d.VoxelWorld(nx, ny + 1, nz) as w:
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.load(path)
        assert ei.value.code == -1                       # other grid


@pytest.mark.gpu
def test_gpu_multislab_and_slab_checkpoints(tmp_path, fs3d, oracle):
    import torch
    k = torch.cuda.device_count()
    nx, ny, nz = 64, 16, 22
    g = oracle.generate(nx, ny, nz, 4, 2)
    whole = str(tmp_path / "whole.fs3d")
    with fs3d.VoxelWorld(nx, ny, nz, seed=2, devices=[i % k for i in range(3)]) as w:
        w.upload(g)
        w.step(4)
        w.save(whole)
        w.step(6)
        want = w.download()
    h, g4 = ck.read(whole)
    assert h["z_begin"] == 0 and h["z_end"] == nz
    with fs3d.VoxelWorld(nx, ny, nz, seed=9) as w:       # a 3-slab checkpoint loads into a 1-slab world
        w.load(whole)
        w.step(6)
        assert np.array_equal(w.download(), want)
    # a rank's slab world saves / loads only its planes
    part = str(tmp_path / "part.fs3d")
    with fs3d.VoxelWorld(nx, ny, nz, seed=2, slab=(8, 22)) as w:
        w.upload(np.ascontiguousarray(g4[8:]))
        w.save(part)
        hp, gp = ck.read(part)
        assert (hp["z_begin"], hp["z_end"], hp["nz"]) == (8, 22, nz) and np.array_equal(gp, g4[8:])
        w.load(part)
        assert np.array_equal(w.download(), g4[8:])
    with fs3d.VoxelWorld(nx, ny, nz, seed=2, slab=(0, 8)) as w:
        with pytest.raises(fs3d.Fs3dError):
            w.load(part)                                  # other planes than this world holds
