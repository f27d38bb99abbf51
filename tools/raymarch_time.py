"""ad-hoc: where a ray-marched frame's time goes on one GPU (kernel vs copies vs host), with the brick occupancy map
forced on, forced off and adaptive.   python tools/raymarch_time.py [n] [scene]"""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import numpy as np
import fallingsand3d_b200 as fs3d
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
scene = int(sys.argv[2]) if len(sys.argv) > 2 else fs3d.SCENE_RANDOM
W, H = 1920, 1080
cam = dict(pos=(0.0, 0.0, -1.6), yaw_deg=0.0, aspect=W / H)
w = fs3d.VoxelWorld(n, n, n, seed=1, slab=(0, n))
w.generate(scene, 1)
w.step(10)
w.sync()
def t(f, reps=5):
    f(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]
img = np.empty((H, W, 4), np.uint8)
print("scene", scene, "n", n)
blob = w.frame_export(W, H, 1); w.frame_attach(blob, 0)
for name, extra in (("no bricks", fs3d.RM_NO_BRICKS), ("bricks (map cached: static scene)", fs3d.RM_BRICKS), ("adaptive", 0)):
    mode = fs3d.RM_VOXELS | extra
    print(f"{name}: fs3d_raymarch (kernel + D2H into pageable numpy), ms min/med:", t(lambda: w.raymarch(width=W, height=H, mode=mode, **cam)),
          "| raymarch_to_frame + sync (kernel only):", t(lambda: (w.raymarch_to_frame(mode=mode, **cam), w.sync())),
          "| bricks in use:", w.raymarch_bricks_in_use())
def stepped():
    w.step(2); w.sync()
    t0 = time.perf_counter(); w.raymarch_to_frame(mode=fs3d.RM_VOXELS | fs3d.RM_BRICKS, **cam); w.sync()
    return (time.perf_counter() - t0) * 1e3
stepped()
print("bricks, map rebuilt after a step (build + march), ms:", min(stepped() for _ in range(4)))
print("frame_resolve (min over slots + D2H), ms:", t(lambda: w.frame_resolve(W, H, img)))
