#!/usr/bin/env python
"""Settled-tile skipping in its sparse phase (BASELINE config 3, 1024^3 RANDOM): runs until the column has settled
to ~12.5 % live tiles, then times passes and compares with `live fraction x dense time`.

  python tools/skip_sparse.py [settle_steps] [--profile]
--profile brackets the timed passes with cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fallingsand3d_b200 as fs3d  # noqa: E402

settle = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4000
profile = "--profile" in sys.argv
n = 1024
with fs3d.VoxelWorld(n, n, n, seed=1) as w:
    w.generate(fs3d.SCENE_RANDOM, 1)
    w.step(20)
    dense_ms, _ = w.step_timed(100)
    dense = dense_ms / 100
with fs3d.VoxelWorld(n, n, n, seed=1, flags=fs3d.FLAG_SKIP_SETTLED) as w:
    w.generate(fs3d.SCENE_RANDOM, 1)
    w.step(settle)
    w.sync()
    cudart = ctypes.CDLL("libcudart.so.12") if profile else None
    if profile:
        cudart.cudaProfilerStart()
        ms, launches = w.step_timed(4)
        cudart.cudaProfilerStop()
    ms, launches = w.step_timed(200)
    run, total = w.activity()
    frac = run / total
    print(json.dumps({"grid": [n, n, n], "settle_steps": settle, "dense_ms_per_step": dense, "skip_ms_per_step": ms / 200,
                      "active_tile_fraction": frac, "ideal_ms_per_step": frac * dense,
                      "fraction_of_ideal": frac * dense / (ms / 200), "launches_per_pass": launches / 100}))
