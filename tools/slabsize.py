"""ad-hoc: time one GPU on a non-cubic grid (nx ny nz) to separate size effects from multi-GPU effects."""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fallingsand3d_b200 as fs3d
nx, ny, nz = map(int, sys.argv[1:4])
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w = fs3d.VoxelWorld(nx, ny, nz, seed=1, flags=flags)
w.generate(fs3d.SCENE_RANDOM, 1)
w.step(4)
ms, n = w.step_timed(100)
print(f"{nx}x{ny}x{nz} flags={flags}: {ms/100:.4f} ms/step  {nx*ny*nz*100/ms/1e9:.1f} G voxel-updates/s  launches={n}")
