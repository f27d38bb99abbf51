"""Generates tests/golden/fs_raymarch_ref_frames.npz from the REFERENCE's own fragment shader.

oracle/_ref/libfs_raymarch_ref.so is /root/reference/shaders/fs_raymarch.frag compiled as C++ against the
reference's vendored glm (oracle/Makefile `ref`, oracle/ref_shader/fs_raymarch_ref.cpp).  It cannot travel to
a box without /root/reference as source, so its output is committed here: for each camera the flat indices of
the pixels the shader hits and their linear red value (green and blue are 0 and alpha 1 on every pixel, which
this script asserts).  Run from the repo root where /root/reference is mounted."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle  # noqa: E402

CAMERAS = [   # pos, aspect, width, height — the first is the engine's default camera and window
    ((0.0, 0.0, -5.0), 1700.0 / 900.0, 850, 450),
    ((0.3, -0.2, -2.0), 16.0 / 9.0, 320, 180),
    ((0.0, 0.1, -0.9), 1.0, 128, 128),
    ((2.0, 1.0, -3.0), 2.0, 200, 100),
]

out = {"n": np.array(len(CAMERAS))}
for k, (pos, aspect, w, h) in enumerate(CAMERAS):
    f = oracle.ref_frame(pos, aspect, w, h)
    assert f is not None, "oracle/_ref is not built: needs /root/reference"
    assert not f[..., 1].any() and not f[..., 2].any() and (f[..., 3] == 1.0).all()
    red = f[..., 0].reshape(-1)
    idx = np.nonzero(red > 0)[0].astype(np.uint32)
    out[f"cam{k}"] = np.array(list(pos) + [aspect, w, h], dtype=np.float64)
    out[f"idx{k}"] = idx
    out[f"red{k}"] = red[idx].astype(np.float32)
    print(f"camera {k}: {idx.size} hit pixels, sum red {red.sum(dtype=np.float64):.4f}")
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fs_raymarch_ref_frames.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
