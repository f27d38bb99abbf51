"""SURVEY.md §8(f).1, CUDA half: the volume's buffers leave the library as POSIX file descriptors
(FS3D_FLAG_EXPORTABLE + fs3d_volume_export_fd) and a second process imports them through the driver API."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _import_in_child(e, which):
    fd = e["fd"][which]
    cmd = [sys.executable, os.path.join(ROOT, "tests", "vmm_import_child.py"), str(fd), str(e["alloc_bytes"]),
           str(e["first_cell_offset"]), str(e["nx"]), str(e["ny"]), str(e["z0"]), str(e["z1"]), str(e["device"])]
    res = subprocess.run(cmd, pass_fds=[fd], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_exported_buffers_carry_the_volume_into_another_process(fs3d, oracle):
    nx, ny, nz = 256, 48, 40
    g = oracle.generate(nx, ny, nz, 4, 2)
    with fs3d.VoxelWorld(nx, ny, nz, seed=3, flags=fs3d.FLAG_EXPORTABLE) as w:
        w.generate(fs3d.SCENE_MIXED_NOISE, 2)
        for n in (6, 1):                              # after an even and after an odd number of passes: `front` flips
            w.step(n)
            oracle.run(g, 3, w.step_index - n, n)
            w.sync()
            e = w.volume_export_fd(0)
            try:
                assert e["alloc_bytes"] >= (nz + 2) * nx * ny and e["first_cell_offset"] == nx * ny and e["step"] == w.step_index
                got = _import_in_child(e, e["front"])
                assert int(got["digest"], 16) == w.digest() == oracle.digest(g)
                assert got["histogram"][:4] == [int(v) for v in w.histogram()[:4]]
                other = _import_in_child(e, e["front"] ^ 1)          # the back buffer holds an older step
                assert int(other["digest"], 16) != w.digest()
            finally:
                for fd in e["fd"]:
                    os.close(fd)
        assert np.array_equal(w.download(), g)         # exportable worlds step like any other


def test_exportable_multi_slab_world_and_errors(fs3d, oracle):
    nx, ny, nz = 64, 24, 18
    g = oracle.generate(nx, ny, nz, 3, 1)
    with fs3d.VoxelWorld(nx, ny, nz, seed=2, devices=[0, 0, 0],
                         flags=fs3d.FLAG_EXPORTABLE | fs3d.FLAG_PEER_PUSH_SHARED_DEVICE) as w:
        w.upload(g)
        w.step(10)
        oracle.run(g, 2, 0, 10)
        total = 0
        for slab in range(w.num_slabs):
            e = w.volume_export_fd(slab)
            try:
                total += int(_import_in_child(e, e["front"])["digest"], 16)
            finally:
                for fd in e["fd"]:
                    os.close(fd)
        assert total & 0xFFFFFFFFFFFFFFFF == oracle.digest(g)      # the digest sums over slabs
    with fs3d.VoxelWorld(nx, ny, nz) as w:
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.volume_export_fd(0)                      # ordinary allocations cannot be exported
        assert ei.value.code == -7
    with fs3d.VoxelWorld(nx, ny, nz, flags=fs3d.FLAG_EXPORTABLE, slab=(0, 9)) as w:
        with pytest.raises(fs3d.Fs3dError) as ei:
            w.slab_ipc_export()                        # and exportable ones cannot go through CUDA IPC
        assert ei.value.code == -7
