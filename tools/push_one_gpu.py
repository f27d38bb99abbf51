#!/usr/bin/env python
"""PUSH kernels under ncu on ONE device: an in-process world of two slabs sharing the GPU
(FS3D_FLAG_PEER_PUSH_SHARED_DEVICE), 2048 x 2048 x 512 — each slab is one rank's share of 2048^3 at N = 8.
ncu serialises the launches in issue order (pass-major), so every kernel's neighbour data is already there: the
capture shows what the coherent loads, the extra ghost-plane stores and the counter traffic cost, not the spin."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fallingsand3d_b200 as fs3d  # noqa: E402

nx, ny, nz = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (2048, 2048, 512)
cudart = ctypes.CDLL("libcudart.so.12")
with fs3d.VoxelWorld(nx, ny, nz, seed=1, devices=[0, 0], flags=fs3d.FLAG_PEER_PUSH_SHARED_DEVICE) as w:
    w.generate(fs3d.SCENE_MIXED_NOISE, 1)
    w.step(4)
    w.sync()
    cudart.cudaProfilerStart()
    ms, launches = w.step_timed(4)
    cudart.cudaProfilerStop()
    print(f"{nx}x{ny}x{nz} two slabs on one device, PUSH: {ms / 4:.3f} ms/step, {launches} launches, wait stats {w.push_wait_stats()}")
