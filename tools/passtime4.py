"""ad-hoc: ms per step with four, two and one step(s) per pass on an n^3 MIXED_NOISE grid, and their digests."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fallingsand3d_b200 as fs3d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
VARIANTS = (("four steps per pass", 0), ("two steps per pass", fs3d.FLAG_NO_FUSE4), ("one step per pass", fs3d.FLAG_NO_FUSE))
for name, flags in VARIANTS[:1 if os.environ.get("FS3D_LIB") else 3]:
    with fs3d.VoxelWorld(n, n, n, seed=1, flags=flags) as w:
        w.generate(fs3d.SCENE_MIXED_NOISE, 1)
        w.step(8)
        ms, launches = w.step_timed(40)
        print(f"{n}^3 {name}: {ms / 40:.4f} ms/step  {n ** 3 * 40 / ms / 1e9:.2f} G voxel-updates/s  launches={launches}  digest={w.digest():#x}")
