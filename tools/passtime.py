"""ad-hoc: time the fused 2-step passes of one GPU separately by x-offset (OX = 0 / 1 alternate every pass)."""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fallingsand3d_b200 as fs3d
nx, ny, nz = map(int, sys.argv[1:4])
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
per = 1 if flags & fs3d.FLAG_NO_FUSE else 2
w = fs3d.VoxelWorld(nx, ny, nz, seed=1, flags=flags)
w.generate(fs3d.SCENE_RANDOM, 1)
w.step(8)
acc = {}
for i in range(24):
    t = w.step_index
    ms, n = w.step_timed(per)
    acc.setdefault(((t >> 1) & 1, t & 1), []).append(ms)
tag = "PUSH" if os.environ.get("FS3D_DEBUG_FORCE_PUSH") else "plain"
for k in sorted(acc):
    v = sorted(acc[k])
    print(f"{nx}x{ny}x{nz} flags={flags} {tag} OX={k[0]} todd={k[1]}: min {v[0]:.3f} med {v[len(v)//2]:.3f} ms per pass ({per} step)")
