"""Key flags -> camera -> brush over time: the C++ header, the Python mirror and the reference's own
camera lines (renderer.cpp:438-467, compiled into oracle/_ref where /root/reference is mounted; its track is the
committed golden tests/golden/camera_track.json) must produce bit-identical float32 poses for a replayed key script."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _golden():
    with open(os.path.join(GOLDEN, "camera_track.json")) as f:
        g = json.load(f)
    track = [[struct.unpack("<f", bytes.fromhex(h))[0] for h in row] for row in g["track"]]
    return g["frames"], track


def _script():
    with open(os.path.join(GOLDEN, "key_script.txt")) as f:
        return f.read()


def test_python_replay_equals_reference_track():
    from fallingsand3d_b200.input import replay
    frames, gold = _golden()
    track = replay(None, _script(), frames)
    assert len(track) == frames
    for i, (a, b) in enumerate(zip(track, gold)):
        assert [np.float32(v).tobytes() for v in a] == [np.float32(v).tobytes() for v in b], f"frame {i}: {a} != {b}"
    # the script moves every axis and the yaw both ways
    xs = np.array(gold)
    assert xs[:, 0].max() > 0.1 and xs[:, 0].min() < -0.3 and xs[:, 1].min() < 0 < xs[:, 1].max() and xs[:, 2].max() > -4 and xs[:, 3].max() > 4


def test_reference_lines_reproduce_the_golden_track():
    # only where oracle/_ref could be built (here) or travelled with the repo (the GPU box)
    import ctypes as C
    from fallingsand3d_b200.input import KeyFlags, KeyScript
    from oracle import oracle
    if oracle.build_ref() is None or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfs_camera_ref.so")):
        pytest.skip("oracle/_ref not available")
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfs_camera_ref.so"))
    frames, gold = _golden()
    script, flags = KeyScript(_script()), KeyFlags()
    pos, rot = (C.c_float * 3)(), (C.c_float * 3)()
    lib.fs_camera_ref_defaults(pos, rot)
    assert (pos[0], pos[1], pos[2], rot[1]) == (0.0, 0.0, -5.0, 0.0)          # renderer.h:148-149
    for frame in range(frames):
        script.handle_events(frame, flags)
        lib.fs_camera_ref_step((C.c_uint8 * 8)(*[1 if f else 0 for f in flags.as_reference_order()]), pos, rot)
        assert [pos[0], pos[1], pos[2], rot[1]] == gold[frame], f"frame {frame}"


def test_cpp_header_replay_equals_reference_track(tmp_path):
    from fallingsand3d_b200.input import KeyFlags, KeyScript, CameraController, brush_centre
    exe = str(tmp_path / "input_host_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host", "input_host_test.cpp"), "-o", exe], check=True)
    frames, gold = _golden()
    out = subprocess.run([exe, os.path.join(GOLDEN, "key_script.txt"), str(frames)], check=True, capture_output=True,
                         text=True).stdout.splitlines()
    assert len(out) == frames
    script, flags, cam = KeyScript(_script()), KeyFlags(), CameraController()
    for i, line in enumerate(out):
        f = line.split()
        got = [struct.unpack("<f", struct.pack("<I", int(h, 16)))[0] for h in f[:4]]
        assert got == gold[i], f"frame {i}"
        # brush voxel, paint flags and material agree with the Python mirror
        script.handle_events(i, flags)
        cam.integrate(flags)
        assert [int(v) for v in f[4:7]] == list(brush_centre(cam, 4.5, 64, 48, 40)), f"frame {i}"
        assert [int(v) for v in f[7:10]] == [int(flags.holdingPaint), int(flags.holdingErase), flags.material]


@pytest.mark.gpu
def test_replayed_session_paints_and_renders_like_the_oracle(fs3d, oracle):
    # the whole headless loop: key script -> camera + brush -> step -> ray-march from the moved camera; the same edits
    # applied to the oracle grid with numpy must give the same cells and the same frames
    from fallingsand3d_b200.input import replay, brush_centre
    nx, ny, nz, seed = 64, 48, 40, 5
    dist = 3.7                       # puts the brush inside the volume while the script pours sand (frames 60-80)
    g = oracle.generate(nx, ny, nz, 2, 1)
    zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    frames_checked = []
    with fs3d.VoxelWorld(nx, ny, nz, seed=seed) as w:
        w.upload(g)
        state = {"t": 0}

        def on_frame(frame, cam, flags):
            # oracle side of this frame: same brush, then one step
            if flags.holdingPaint or flags.holdingErase:
                c = brush_centre(cam, dist, nx, ny, nz)
                mask = (xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2 <= 9
                if flags.holdingErase:
                    g[mask] = 0
                else:
                    g[mask & (g == 0)] = flags.material
            oracle.step(g, seed, state["t"])
            state["t"] += 1
            if frame % 40 == 39:
                img = w.raymarch(width=170, height=90, mode=fs3d.RM_VOXELS, **cam.camera())
                ref = oracle.raymarch(g, width=170, height=90, mode=1, **cam.camera())
                assert np.array_equal(img, ref), f"frame {frame}"
                frames_checked.append(frame)

        replay(w, _script(), 160, brush_distance=dist, on_frame=on_frame)
        assert np.array_equal(w.download(), g)
        assert len(frames_checked) == 4 and int((g == 1).sum()) > int((oracle.generate(nx, ny, nz, 2, 1) == 1).sum())


@pytest.mark.gpu
def test_cpp_engine_loop_replays_the_session_like_python(fs3d, tmp_path):
    # examples/engine_loop.cpp (C++ wrapper + fs3d_input.hpp through the C ABI) and the Python replay run the same
    # session: same digest, same final camera bits, same ray-marched frames
    from fallingsand3d_b200.input import replay
    lib_dir = os.path.join(ROOT, "fallingsand3d_b200")
    exe = str(tmp_path / "engine_loop")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "engine_loop.cpp"), "-L" + lib_dir, "-lfs3d", "-Wl,-rpath," + lib_dir,
                    "-o", exe], check=True)
    res = subprocess.run([exe, "64", "160", os.path.join(GOLDEN, "key_script.txt"), "3.7"], capture_output=True, text=True,
                         cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout + res.stderr
    words = res.stdout.split()
    frames = {}

    def on_frame(frame, cam, flags):
        if frame % 40 == 39:
            frames[frame + 1] = w.raymarch(width=170, height=90, mode=fs3d.RM_VOXELS, **cam.camera())

    with fs3d.VoxelWorld(64, 48, 40, seed=5) as w:
        w.generate(fs3d.SCENE_MIXED, 1)
        track = replay(w, _script(), 160, brush_distance=3.7, on_frame=on_frame)
        assert words[words.index("digest") + 1] == "%016x" % w.digest()
        assert words[words.index("steps") + 1] == "160"
    cam_bits = [struct.unpack("<f", struct.pack("<I", int(h, 16)))[0] for h in words[words.index("camera") + 1:][:4]]
    assert cam_bits == list(track[-1])
    for k, img in frames.items():
        raw = open(tmp_path / f"frame_{k}.ppm", "rb").read()
        hdr = b"P6\n170 90\n255\n"
        assert raw.startswith(hdr)
        assert np.array_equal(np.frombuffer(raw[len(hdr):], np.uint8).reshape(90, 170, 3), img[:, :, :3]), f"frame {k}"
