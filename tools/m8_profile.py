#!/usr/bin/env python
"""Schedule version 2 kernels under ncu: 2048^3 MIXED8, two fused passes (both x-offsets) bracketed by
cudaProfilerStart/Stop (`ncu --profile-from-start off`)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fallingsand3d_b200 as fs3d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
flags = fs3d.FLAG_MATERIALS8 | (fs3d.FLAG_NO_FUSE if "--single" in sys.argv else 0)
cudart = ctypes.CDLL("libcudart.so.12")
with fs3d.VoxelWorld(n, n, n, seed=1, flags=flags) as w:
    w.generate(fs3d.SCENE_MIXED8, 1)
    w.step(8)
    w.sync()
    cudart.cudaProfilerStart()
    ms, launches = w.step_timed(4)
    cudart.cudaProfilerStop()
    print(f"{n}^3 MIXED8 flags={flags}: {ms / 4:.3f} ms/step, {launches} launches")
