/*
 * fs3d_raymarch_oracle.c — CPU ORACLE of the offscreen ray-march.  TEST INFRASTRUCTURE ONLY.
 *
 * float32 restatement of the reference's full-screen ray-march shader:
 *   /root/reference/shaders/fs_raymarch.vert:30-37  (origin = camPos.xyz, uv, aspect passthrough)
 *   /root/reference/shaders/fs_raymarch.frag:10-12  sphereSDF
 *   /root/reference/shaders/fs_raymarch.frag:19-26  map_the_world = sphere r 0.5 at origin
 *   /root/reference/shaders/fs_raymarch.frag:28-36  calculate_normal (central differences, 0.001)
 *   /root/reference/shaders/fs_raymarch.frag:38-65  ray_march (64 iterations, 0.001 hit, 1000 far, light (2,5,3))
 *   /root/reference/shaders/fs_raymarch.frag:67-81  main (uv*2-1, y /= aspect, normalize(vec3(uv,1)))
 *   /root/reference/src/engine/rendering/renderer.cpp:1253-1267  quad UVs => u = 1-(px+.5)/W, v = (py+.5)/H
 * Pinned: bit-identical (hit mask, linear float colour, 8-bit image) to the reference shader source itself compiled
 * against the reference's vendored glm (oracle/_ref via `make ref`; golden frames tests/golden/fs_raymarch_ref_frames.npz),
 * and to the known-answer pixels of SURVEY.md §8c (tests/test_raymarch_oracle.py).  The voxel-DDA mode has no reference counterpart
 * (fs_raymarch takes no volume input — materials.cpp:520-521 creates it with zero bindings).
 *
 * Compile with -ffp-contract=off: every operation below is a separately rounded float op in the
 * same order as fallingsand3d_b200/csrc/raymarch.cuh.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    float s = ax * bx; float t = ay * by; s = s + t; t = az * bz; return s + t;
}
static float len3(float x, float y, float z) { return sqrtf(dot3(x, y, z, x, y, z)); }
/* GLSL normalize as the reference's vendored glm 0.9.9.7 evaluates it (detail/func_geometric.inl: v * inversesqrt(dot(v, v)),
 * inversesqrt(x) = 1 / sqrt(x)); with this form the frame is bit-identical to fs_raymarch.frag compiled against that glm
 * (oracle/_ref, tests/test_raymarch_oracle.py). */
static float inv_len3(float x, float y, float z) { return 1.0f / sqrtf(dot3(x, y, z, x, y, z)); }
static float sphere_sdf(float x, float y, float z) { return len3(x, y, z) - 0.5f; }

static float diffuse_at(float px, float py, float pz, float nx, float ny, float nz) {
    float lx = px - 2.0f, ly = py - 5.0f, lz = pz - 3.0f;
    float il = inv_len3(lx, ly, lz);
    lx = lx * il; ly = ly * il; lz = lz * il;
    float d = dot3(nx, ny, nz, lx, ly, lz);
    return d > 0.05f ? d : 0.05f;
}

/* crossing time of lattice boundary k: ((k h - e) - o) * (1 / d) */
static float boundary_t(int k, float h, float e, float o, float rcp) {
    float v = (float)k * h; v = v - e; v = v - o; return v * rcp;
}

void fs3d_oracle_srgb_thresholds(float *thr /* 256 */) {
    for (int i = 1; i <= 255; ++i) {
        double c = (i - 0.5) / 255.0;
        double lin = c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4);
        thr[i - 1] = (float)lin;
    }
    thr[255] = INFINITY;
}

static uint8_t encode8(float v, const float *thr) {
    if (!(v > 0.0f)) return 0;
    if (thr) {
        int lo = 0, hi = 255;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (v >= thr[mid]) lo = mid + 1; else hi = mid; }
        return (uint8_t)lo;
    }
    if (v >= 1.0f) return 255;
    float s = v * 255.0f; s = s + 0.5f;
    return (uint8_t)(int)s;
}

/* Linear-light red channel + iteration count of one pixel in SDF mode (for the known-answer table). */
void fs3d_oracle_raymarch_pixel(const float pos[3], float aspect, uint32_t W, uint32_t H, uint32_t px, uint32_t py,
                                float dir_out[3], float *red_out, int *iters_out);

/*
 * grid: planes [zlo, zhi) of an nx×ny×nz world (cells outside read EMPTY) — a rank's slab, or the
 * whole grid with zlo = 0, zhi = nz.  mode: 0 SDF sphere, 1 voxels; | 16 = sRGB encode.
 * palette: 256×4 floats.  depth may be NULL.
 */
void fs3d_oracle_raymarch(const uint8_t *grid, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t zlo, uint32_t zhi,
                          const float pos[3], float yaw_deg, float aspect, uint32_t W, uint32_t H, uint32_t mode,
                          const float *palette, uint8_t *img, float *depth_out) {
    float thr_buf[256];
    const float *thr = NULL;
    if (mode & 16u) { fs3d_oracle_srgb_thresholds(thr_buf); thr = thr_buf; }
    const double yaw = (double)yaw_deg * 3.14159265358979323846 / 180.0;
    const float cs = yaw_deg == 0.0f ? 1.0f : (float)cos(yaw);
    const float sn = yaw_deg == 0.0f ? 0.0f : (float)sin(yaw);
    uint32_t nmax = nx > ny ? nx : ny; if (nz > nmax) nmax = nz;
    const float h = 1.0f / (float)nmax;
    const float e[3] = { (float)nx * h * 0.5f, (float)ny * h * 0.5f, (float)nz * h * 0.5f };
    const int n[3] = { (int)nx, (int)ny, (int)nz };

    for (uint32_t py = 0; py < H; ++py)
        for (uint32_t px = 0; px < W; ++px) {
            float u = ((float)px + 0.5f) / (float)W; u = 1.0f - u;
            float v = ((float)py + 0.5f) / (float)H;
            float qx = u * 2.0f - 1.0f;
            float inv_aspect = 1.0f / aspect;
            float qy = (v * 2.0f - 1.0f) * inv_aspect;
            float iq = inv_len3(qx, qy, 1.0f);
            float dx = qx * iq, dy = qy * iq, dz = 1.0f * iq;
            {
                float a = cs * dx, b = sn * dz; float rx = a + b;
                a = cs * dz; b = sn * dx; float rz = a - b;
                dx = rx; dz = rz;
            }
            float r = 0.f, g = 0.f, b = 0.f, depth = INFINITY;
            if ((mode & 15u) == 0) {
                float t = 0.0f;
                for (int i = 0; i < 64; ++i) {
                    float cx = pos[0] + t * dx, cy = pos[1] + t * dy, cz = pos[2] + t * dz;
                    float d = sphere_sdf(cx, cy, cz);
                    if (d < 0.001f) {
                        const float s = 0.001f;
                        float gx = sphere_sdf(cx + s, cy, cz) - sphere_sdf(cx - s, cy, cz);
                        float gy = sphere_sdf(cx, cy + s, cz) - sphere_sdf(cx, cy - s, cz);
                        float gz = sphere_sdf(cx, cy, cz + s) - sphere_sdf(cx, cy, cz - s);
                        float ig = inv_len3(gx, gy, gz);
                        gx = gx * ig; gy = gy * ig; gz = gz * ig;
                        r = diffuse_at(cx, cy, cz, gx, gy, gz);
                        depth = t;
                        break;
                    } else if (t > 1000.0f) {
                        break;
                    }
                    t = t + d;
                }
            } else {
                float tmin = 0.0f, tmax = INFINITY;
                int miss = 0;
                const float o[3] = { pos[0], pos[1], pos[2] }, d[3] = { dx, dy, dz };
                for (int a = 0; a < 3; ++a) {
                    if (d[a] != 0.0f) {
                        float t0 = (-e[a] - o[a]) / d[a], t1 = (e[a] - o[a]) / d[a];
                        if (t0 > t1) { float s = t0; t0 = t1; t1 = s; }
                        if (t0 > tmin) tmin = t0;
                        if (t1 < tmax) tmax = t1;
                    } else if (o[a] < -e[a] || o[a] > e[a]) {
                        miss = 1;
                    }
                }
                if (!miss && tmin <= tmax) {
                    /* The walk visits voxel after voxel.  The time at which the ray crosses lattice boundary k of axis a is
                     * computed from k itself — ((k h - e) - o) * (1 / d), each operation rounded separately — not accumulated,
                     * so a ray's state is a function of where it is.  (The CUDA kernel relies on that to jump over empty
                     * bricks and other ranks' planes; this oracle never jumps, and the two must still agree bit for bit.) */
                    int idx[3], stp[3], last_axis = -1;
                    float tnext[3], rcp[3];
                    for (int a = 0; a < 3; ++a) {
                        float p0 = o[a] + tmin * d[a];
                        float f = (p0 + e[a]) / h;
                        int i = (int)floorf(f);
                        if (i < 0) i = 0;
                        if (i > n[a] - 1) i = n[a] - 1;
                        idx[a] = i;
                        stp[a] = d[a] > 0.0f ? 1 : (d[a] < 0.0f ? -1 : 0);
                        rcp[a] = stp[a] != 0 ? 1.0f / d[a] : 0.0f;
                        tnext[a] = stp[a] != 0 ? boundary_t(i + (stp[a] > 0 ? 1 : 0), h, e[a], o[a], rcp[a]) : INFINITY;
                    }
                    {
                        float best = -1.0f;
                        for (int a = 0; a < 3; ++a)
                            if (d[a] != 0.0f) {
                                float t0 = (-e[a] - o[a]) / d[a], t1 = (e[a] - o[a]) / d[a];
                                float tn = t0 < t1 ? t0 : t1;
                                if (tn == tmin && tn > best) { best = tn; last_axis = a; }
                            }
                    }
                    float t = tmin;
                    const int max_steps = n[0] + n[1] + n[2] + 3;
                    for (int s = 0; s < max_steps; ++s) {
                        int gx = idx[0], gy = n[1] - 1 - idx[1], gz = idx[2];
                        uint8_t m = 0;
                        if ((uint32_t)gz >= zlo && (uint32_t)gz < zhi)
                            m = grid[(size_t)gx + (size_t)nx * ((size_t)gy + (size_t)ny * ((size_t)gz - zlo))];
                        if (m != 0) {
                            float nrm[3] = { 0.f, 0.f, 0.f };
                            if (last_axis >= 0) nrm[last_axis] = stp[last_axis] > 0 ? -1.0f : 1.0f;
                            float hx = pos[0] + t * dx, hy = pos[1] + t * dy, hz = pos[2] + t * dz;
                            float df = diffuse_at(hx, hy, hz, nrm[0], nrm[1], nrm[2]);
                            const float *c = palette + 4 * (int)m;
                            r = c[0] * df; g = c[1] * df; b = c[2] * df;
                            depth = t;
                            break;
                        }
                        int a = 0;
                        if (tnext[1] < tnext[a]) a = 1;
                        if (tnext[2] < tnext[a]) a = 2;
                        t = tnext[a];
                        idx[a] += stp[a];
                        if (idx[a] < 0 || idx[a] >= n[a]) break;
                        tnext[a] = boundary_t(idx[a] + (stp[a] > 0 ? 1 : 0), h, e[a], o[a], rcp[a]);
                        last_axis = a;
                    }
                }
            }
            size_t pix = (size_t)py * W + px;
            img[4 * pix + 0] = encode8(r, thr);
            img[4 * pix + 1] = encode8(g, thr);
            img[4 * pix + 2] = encode8(b, thr);
            img[4 * pix + 3] = 255;
            if (depth_out) depth_out[pix] = depth;
        }
}

void fs3d_oracle_raymarch_pixel(const float pos[3], float aspect, uint32_t W, uint32_t H, uint32_t px, uint32_t py,
                                float dir_out[3], float *red_out, int *iters_out) {
    float u = ((float)px + 0.5f) / (float)W; u = 1.0f - u;
    float v = ((float)py + 0.5f) / (float)H;
    float qx = u * 2.0f - 1.0f;
    float inv_aspect = 1.0f / aspect;
    float qy = (v * 2.0f - 1.0f) * inv_aspect;
    float iq = inv_len3(qx, qy, 1.0f);
    float dx = qx * iq, dy = qy * iq, dz = 1.0f * iq;
    dir_out[0] = dx; dir_out[1] = dy; dir_out[2] = dz;
    *red_out = 0.0f; *iters_out = 64;
    float t = 0.0f;
    for (int i = 0; i < 64; ++i) {
        float cx = pos[0] + t * dx, cy = pos[1] + t * dy, cz = pos[2] + t * dz;
        float d = sphere_sdf(cx, cy, cz);
        if (d < 0.001f) {
            const float s = 0.001f;
            float gx = sphere_sdf(cx + s, cy, cz) - sphere_sdf(cx - s, cy, cz);
            float gy = sphere_sdf(cx, cy + s, cz) - sphere_sdf(cx, cy - s, cz);
            float gz = sphere_sdf(cx, cy, cz + s) - sphere_sdf(cx, cy, cz - s);
            float ig = inv_len3(gx, gy, gz);
            gx = gx * ig; gy = gy * ig; gz = gz * ig;
            *red_out = diffuse_at(cx, cy, cz, gx, gy, gz);
            *iters_out = i;
            return;
        } else if (t > 1000.0f) {
            *iters_out = i;
            return;
        }
        t = t + d;
    }
}
