"""ad-hoc: time the in-process multi-GPU world (fs3d_create with n_gpus = N: one host thread, N slabs)."""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fallingsand3d_b200 as fs3d
n, size = int(sys.argv[1]), int(sys.argv[2])
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w = fs3d.VoxelWorld(size, size, size, seed=1, devices=list(range(n)), flags=flags)
w.generate(fs3d.SCENE_MIXED_NOISE, 1)
w.step(4)
ms, launches = w.step_timed(100)
print(f"in-process {size}^3 on {n} GPUs flags={flags}: {ms/100:.4f} ms/step  {size**3*100/ms/1e9:.1f} G voxel-updates/s  "
      f"launches={launches} digest={w.digest():#x}")
