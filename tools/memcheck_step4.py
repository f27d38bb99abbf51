"""ad-hoc, meant to run under `compute-sanitizer --tool memcheck`: the four-step kernel on small grids of every row width —
plain split, grouped split (FS3D_S4_GROUP_SPAN=1), and through fs3d_step_host with one band per chunk."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fallingsand3d_b200 as fs3d  # noqa: E402

for nx, ny, nz in ((1024, 40, 22), (2048, 24, 20), (4096, 12, 18)):
    digs = []
    for span in ("0", "1"):
        os.environ["FS3D_S4_GROUP_SPAN"] = span
        with fs3d.VoxelWorld(nx, ny, nz, seed=3) as w:
            w.generate(fs3d.SCENE_MIXED_NOISE, 2)
            w.step(8)
            digs.append(w.digest())
    os.environ["FS3D_HOST_CHUNK_BYTES"] = str(8 * nx * ny)
    with fs3d.VoxelWorld(nx, ny, nz, seed=3) as w:
        w.generate(fs3d.SCENE_MIXED_NOISE, 2)
        host = w.download()
        out = np.empty_like(host)
        w.step_host(host, out, 4)
        w.step_host(out, host, 4)
        digs.append(w.digest())
    assert digs[0] == digs[1] == digs[2], digs
    print(f"{nx}x{ny}x{nz}: {digs[0]:#x} three ways")
print("MEMCHECK_RUN_OK")
