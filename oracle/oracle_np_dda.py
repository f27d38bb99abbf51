"""Independent checks of the voxel ray-march (mode FS3D_RM_VOXELS).  TEST INFRASTRUCTURE ONLY.

oracle/fs3d_raymarch_oracle.c follows the CUDA kernel's order of operations so that images are bit-identical; a logic
error shared by both would pass.  This module is structured differently on purpose:

  * analytic_axis_ray: a 1 x 1 image puts the only pixel centre on the optical axis (u = v = 1/2 -> direction exactly
    (0, 0, 1), shaders/fs_raymarch.frag:67-75), so a camera placed in front of voxel column (i, j) must hit the first
    non-empty cell k of that column at t = (-ez + k h) - oz, on its -z face, with colour
    palette[m] * max(0.05, dot((0,0,-1), normalize(p - light)))  (fs_raymarch.frag:49-55) — a closed form, no marching;
  * sample_march: float64 brute force — walk the ray in steps of h / 32 and report the first non-empty cell entered.
    No DDA state, no incremental crossing times; used on pixels whose hit is not within a hair of a cell edge.

Camera model: /root/reference/shaders/fs_raymarch.vert:30-37, fs_raymarch.frag:67-75, quad UVs
/root/reference/src/engine/rendering/renderer.cpp:1253-1267.  Volume placement as in csrc/raymarch.cuh: cell edge
h = 1 / max(nx, ny, nz), box centred at the origin, grid +y = world -y."""
import math

import numpy as np

LIGHT = np.array([2.0, 5.0, 3.0])


def box(nx, ny, nz):
    h = 1.0 / max(nx, ny, nz)
    return h, np.array([nx * h / 2, ny * h / 2, nz * h / 2])


def column_camera(i, j, nx, ny, nz, z=-5.0):
    """Camera position whose optical axis runs through the centre of voxel column (i, j)."""
    h, e = box(nx, ny, nz)
    return (-e[0] + (i + 0.5) * h, e[1] - (j + 0.5) * h, z)


def analytic_axis_ray(grid, i, j, palette, z=-5.0):
    """(rgb uint8 triple as floats before rounding, depth) for the single pixel of a 1 x 1 image from column_camera."""
    nz, ny, nx = grid.shape
    h, e = box(nx, ny, nz)
    col = grid[:, j, i]
    ks = np.nonzero(col)[0]
    if len(ks) == 0:
        return None, math.inf
    k = int(ks[0])
    t = (-e[2] + k * h) - z
    o = np.array(column_camera(i, j, nx, ny, nz, z))
    p = o + t * np.array([0.0, 0.0, 1.0])
    l = p - LIGHT
    l /= np.linalg.norm(l)
    diffuse = max(0.05, float(np.dot(np.array([0.0, 0.0, -1.0]), l)))
    return np.asarray(palette[int(col[k])][:3], dtype=np.float64) * diffuse, t


def pixel_ray(px, py, width, height, aspect, yaw_deg=0.0):
    u = 1.0 - (px + 0.5) / width
    v = (py + 0.5) / height
    q = np.array([u * 2 - 1, (v * 2 - 1) / aspect, 1.0])
    d = q / np.linalg.norm(q)
    a = math.radians(yaw_deg)
    return np.array([math.cos(a) * d[0] + math.sin(a) * d[2], d[1], math.cos(a) * d[2] - math.sin(a) * d[0]])


def sample_march(grid, pos, d, oversample=32):
    """First non-empty cell a finely sampled float64 ray enters: (i, j, k, t_entry_estimate) or None."""
    nz, ny, nx = grid.shape
    h, e = box(nx, ny, nz)
    o = np.asarray(pos, dtype=np.float64)
    # parametric interval inside the box
    tmin, tmax = 0.0, math.inf
    for a in range(3):
        if d[a] != 0:
            t0, t1 = (-e[a] - o[a]) / d[a], (e[a] - o[a]) / d[a]
            tmin, tmax = max(tmin, min(t0, t1)), min(tmax, max(t0, t1))
        elif abs(o[a]) > e[a]:
            return None
    if tmin > tmax:
        return None
    ts = np.arange(tmin + 1e-9, tmax, h / oversample)
    pts = o[None, :] + ts[:, None] * d[None, :]
    idx = np.floor((pts + e[None, :]) / h).astype(np.int64)
    ok = np.all((idx >= 0) & (idx < np.array([nx, ny, nz])[None, :]), axis=1)
    idx, ts = idx[ok], ts[ok]
    cells = grid[idx[:, 2], ny - 1 - idx[:, 1], idx[:, 0]]
    hit = np.nonzero(cells)[0]
    if len(hit) == 0:
        return None
    f = int(hit[0])
    return int(idx[f, 0]), int(ny - 1 - idx[f, 1]), int(idx[f, 2]), float(ts[f])


def entry_point_margin(pos, d, t, nx, ny, nz):
    """Distance (in cells) from the hit point to the nearest cell edge on the face it lies in: small -> the ray grazes
    an edge and a float32 walk may legitimately pick the neighbouring cell."""
    h, e = box(nx, ny, nz)
    p = (np.asarray(pos) + t * np.asarray(d) + e) / h
    frac = np.abs(p - np.round(p))
    frac.sort()
    return float(frac[1])            # smallest is ~0 (on the face); the second smallest is the edge distance
