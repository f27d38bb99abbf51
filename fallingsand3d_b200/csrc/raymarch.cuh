// raymarch.cuh — offscreen CUDA ray-march of the voxel volume.
//
// Camera and shading follow the reference's full-screen ray-march shader exactly:
//   /root/reference/shaders/fs_raymarch.vert:30-37   origin = camPos.xyz, uv passthrough, aspect
//   /root/reference/shaders/fs_raymarch.frag:67-81   dir = normalize(vec3((uv*2-1) with y /= aspect, 1))
//   /root/reference/shaders/fs_raymarch.frag:38-65   <= 64 sphere-trace iterations, hit d < 0.001, far t > 1000
//   /root/reference/shaders/fs_raymarch.frag:28-36   central-difference normal, step 0.001
//   /root/reference/shaders/fs_raymarch.frag:49-55   light (2,5,3), direction = normalize(p - light) (sic), max(0.05, n·l)
//   /root/reference/src/engine/rendering/renderer.cpp:1253-1267  quad UVs: pixel centre (px,py) ->
//        u = 1 - (px+.5)/W, v = (py+.5)/H  (x mirrored; Vulkan NDC y down)
// Mode FS3D_RM_SDF_SPHERE reproduces the shader as shipped (analytic sphere, no volume input).
// Mode FS3D_RM_VOXELS keeps camera + light and replaces map_the_world by an Amanatides-Woo DDA
// through the uint8 grid with a 256-entry palette (the reference's unused colors[256],
// renderer.cpp:136-393, is the intended shape of that palette).
//
// Every float operation is an explicit round-to-nearest intrinsic in a fixed order (no FMA
// contraction), so the image is bit-identical to oracle/fs3d_raymarch_oracle.c.
#pragma once
#include <cmath>
#include "common.cuh"

struct fs3d_world;

namespace fs3d {

constexpr int RM_MAX_SLABS = 16;

struct RMParams {
    float ox, oy, oz;        // ray origin (camera position)
    float cs, sn;            // cos / sin of yaw (0 -> 1, 0)
    float aspect;
    uint32_t W, H, mode;
    uint32_t nx, ny, nz;     // global grid
    float h;                 // voxel edge in world units: 1 / max(nx, ny, nz)
    float ex, ey, ez;        // half extents of the volume box
    int nslabs;
    const uint8_t *slab_ptr[RM_MAX_SLABS];   // pointer to the slab's first owned plane
    uint32_t slab_z0[RM_MAX_SLABS], slab_z1[RM_MAX_SLABS];
    const float *palette;    // 256 x rgba
    const float *srgb_thr;   // 255 thresholds, or nullptr for linear output
    uint8_t *img;            // W*H*4
    float *depth;            // W*H hit parameter t (inf on miss), may be nullptr
    unsigned long long *frame;   // non-null: store (float bits of t) << 32 | rgba8 per pixel here instead (this
                                 // rank's slot of the compositor's frame, possibly peer memory over NVLink)
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz));
}
__device__ __forceinline__ float len3(float x, float y, float z) { return fsqrt(dot3(x, y, z, x, y, z)); }
// GLSL normalize evaluated the way the reference's vendored glm does (v * (1 / sqrt(dot(v, v)))): with this form the
// oracle — and therefore this kernel — is bit-identical to fs_raymarch.frag compiled against that glm (oracle/_ref)
__device__ __forceinline__ float inv_len3(float x, float y, float z) { return fdiv(1.0f, fsqrt(dot3(x, y, z, x, y, z))); }

__device__ __forceinline__ float sphere_sdf(float x, float y, float z) { return fsub(len3(x, y, z), 0.5f); }

__device__ __forceinline__ float diffuse_at(float px, float py, float pz, float nx, float ny, float nz) {
    // direction_to_light = normalize(p - light_pos), light_pos = (2, 5, 3)   [fs_raymarch.frag:49-53]
    float lx = fsub(px, 2.0f), ly = fsub(py, 5.0f), lz = fsub(pz, 3.0f);
    float il = inv_len3(lx, ly, lz);
    lx = fmul(lx, il); ly = fmul(ly, il); lz = fmul(lz, il);
    float d = dot3(nx, ny, nz, lx, ly, lz);
    return d > 0.05f ? d : 0.05f;
}

__device__ __forceinline__ uint8_t encode8(float v, const float *thr) {
    if (!(v > 0.0f)) return 0;
    if (thr) {   // sRGB: number of thresholds <= v (binary search over 255 ascending floats)
        int lo = 0, hi = 255;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (v >= thr[mid]) lo = mid + 1; else hi = mid; }
        return (uint8_t)lo;
    }
    if (v >= 1.0f) return 255;
    return (uint8_t)(int)fadd(fmul(v, 255.0f), 0.5f);
}

__device__ __forceinline__ uint8_t voxel_at(const RMParams &p, int ix, int iy, int iz) {
#pragma unroll 1
    for (int s = 0; s < p.nslabs; ++s)
        if ((uint32_t)iz >= p.slab_z0[s] && (uint32_t)iz < p.slab_z1[s])
            return p.slab_ptr[s][(size_t)ix + (size_t)p.nx * ((size_t)iy + (size_t)p.ny * ((size_t)iz - p.slab_z0[s]))];
    return 0;   // plane not held by this rank: transparent (composited by depth across ranks)
}

__global__ void raymarch_kernel(const RMParams p) {
    const uint32_t px = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= p.W || py >= p.H) return;

    // camera  [renderer.cpp:1253-1267, fs_raymarch.frag:67-75]
    float u = fsub(1.0f, fdiv(fadd((float)px, 0.5f), (float)p.W));
    float v = fdiv(fadd((float)py, 0.5f), (float)p.H);
    float qx = fsub(fmul(u, 2.0f), 1.0f);
    float qy = fmul(fsub(fmul(v, 2.0f), 1.0f), fdiv(1.0f, p.aspect));
    float iq = inv_len3(qx, qy, 1.0f);
    float dx = fmul(qx, iq), dy = fmul(qy, iq), dz = fmul(1.0f, iq);
    // optional yaw about y (camRot.y, renderer.cpp:460-467); identity when cs = 1, sn = 0
    {
        float rx = fadd(fmul(p.cs, dx), fmul(p.sn, dz));
        float rz = fsub(fmul(p.cs, dz), fmul(p.sn, dx));
        dx = rx; dz = rz;
    }

    float r = 0.f, g = 0.f, b = 0.f, depth = INFINITY;

    if ((p.mode & 15u) == FS3D_RM_SDF_SPHERE) {
        float t = 0.0f;
        for (int i = 0; i < 64; ++i) {
            float cx = fadd(p.ox, fmul(t, dx)), cy = fadd(p.oy, fmul(t, dy)), cz = fadd(p.oz, fmul(t, dz));
            float d = sphere_sdf(cx, cy, cz);
            if (d < 0.001f) {
                const float e = 0.001f;
                float gx = fsub(sphere_sdf(fadd(cx, e), cy, cz), sphere_sdf(fsub(cx, e), cy, cz));
                float gy = fsub(sphere_sdf(cx, fadd(cy, e), cz), sphere_sdf(cx, fsub(cy, e), cz));
                float gz = fsub(sphere_sdf(cx, cy, fadd(cz, e)), sphere_sdf(cx, cy, fsub(cz, e)));
                float ig = inv_len3(gx, gy, gz);
                gx = fmul(gx, ig); gy = fmul(gy, ig); gz = fmul(gz, ig);
                r = diffuse_at(cx, cy, cz, gx, gy, gz);   // vec3(1,0,0) * diffuse
                depth = t;
                break;
            } else if (t > 1000.0f) {
                break;
            }
            t = fadd(t, d);
        }
    } else {
        // volume box [-ex,ex] x [-ey,ey] x [-ez,ez]; voxel (i,j,k) spans x: -ex + i h, y: +ey - (j+1) h
        // (grid +y is world -y: this camera's screen-down is world +y), z: -ez + k h
        float tmin = 0.0f, tmax = INFINITY;
        bool miss = false;
        const float o[3] = {p.ox, p.oy, p.oz}, d[3] = {dx, dy, dz}, e[3] = {p.ex, p.ey, p.ez};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (d[a] != 0.0f) {
                float t0 = fdiv(fsub(-e[a], o[a]), d[a]), t1 = fdiv(fsub(e[a], o[a]), d[a]);
                if (t0 > t1) { float s = t0; t0 = t1; t1 = s; }
                if (t0 > tmin) tmin = t0;
                if (t1 < tmax) tmax = t1;
            } else if (o[a] < -e[a] || o[a] > e[a]) {
                miss = true;
            }
        }
        if (!miss && tmin <= tmax) {
            // entry point in lattice units along each axis (world axis direction, not grid y)
            const int n[3] = {(int)p.nx, (int)p.ny, (int)p.nz};
            int idx[3], stp[3];
            float tnext[3], tdelta[3];
            int last_axis = -1;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float pos = fadd(o[a], fmul(tmin, d[a]));            // world coordinate at entry
                float f = fdiv(fadd(pos, e[a]), p.h);                 // lattice coordinate [0, n]
                int i = (int)floorf(f);
                if (i < 0) i = 0;
                if (i > n[a] - 1) i = n[a] - 1;
                idx[a] = i;
                if (d[a] > 0.0f) {
                    stp[a] = 1;
                    tnext[a] = fdiv(fsub(fsub(fmul((float)(i + 1), p.h), e[a]), o[a]), d[a]);
                    tdelta[a] = fdiv(p.h, d[a]);
                } else if (d[a] < 0.0f) {
                    stp[a] = -1;
                    tnext[a] = fdiv(fsub(fsub(fmul((float)i, p.h), e[a]), o[a]), d[a]);
                    tdelta[a] = fdiv(p.h, -d[a]);
                } else {
                    stp[a] = 0; tnext[a] = INFINITY; tdelta[a] = INFINITY;
                }
            }
            // which face did we enter through?  the axis whose slab entry time equals tmin
            {
                float best = -1.0f;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    if (d[a] != 0.0f) {
                        float t0 = fdiv(fsub(-e[a], o[a]), d[a]), t1 = fdiv(fsub(e[a], o[a]), d[a]);
                        float tn = t0 < t1 ? t0 : t1;
                        if (tn == tmin && tn > best) { best = tn; last_axis = a; }
                    }
                }
            }
            float t = tmin;
            const int max_steps = n[0] + n[1] + n[2] + 3;
            for (int s = 0; s < max_steps; ++s) {
                uint8_t m = voxel_at(p, idx[0], n[1] - 1 - idx[1], idx[2]);
                if (m != 0) {
                    float nrm[3] = {0.f, 0.f, 0.f};
                    if (last_axis >= 0) nrm[last_axis] = stp[last_axis] > 0 ? -1.0f : 1.0f;
                    float hx = fadd(p.ox, fmul(t, dx)), hy = fadd(p.oy, fmul(t, dy)), hz = fadd(p.oz, fmul(t, dz));
                    float df = diffuse_at(hx, hy, hz, nrm[0], nrm[1], nrm[2]);
                    const float *c = p.palette + 4 * (int)m;
                    r = fmul(c[0], df); g = fmul(c[1], df); b = fmul(c[2], df);
                    depth = t;
                    break;
                }
                int a = 0;
                if (tnext[1] < tnext[a]) a = 1;
                if (tnext[2] < tnext[a]) a = 2;
                t = tnext[a];
                idx[a] += stp[a];
                if (idx[a] < 0 || idx[a] >= n[a]) break;
                tnext[a] = fadd(tnext[a], tdelta[a]);
                last_axis = a;
            }
        }
    }

    const size_t pix = (size_t)py * p.W + px;
    if (p.frame) {
        // t >= 0, so its bit pattern orders like the float: the compositor takes the 64-bit minimum over ranks
        const uint32_t rgba = (uint32_t)encode8(r, p.srgb_thr) | ((uint32_t)encode8(g, p.srgb_thr) << 8) |
                              ((uint32_t)encode8(b, p.srgb_thr) << 16) | 0xFF000000u;
        p.frame[pix] = ((unsigned long long)__float_as_uint(depth) << 32) | rgba;
        return;
    }
    p.img[4 * pix + 0] = encode8(r, p.srgb_thr);
    p.img[4 * pix + 1] = encode8(g, p.srgb_thr);
    p.img[4 * pix + 2] = encode8(b, p.srgb_thr);
    p.img[4 * pix + 3] = 255;
    if (p.depth) p.depth[pix] = depth;
}

// Compositor: per pixel the slot with the smallest hit parameter wins (a miss is +inf and black,
// exactly what every rank wrote for it, so the minimum is right for misses too).
__global__ void frame_resolve_kernel(const unsigned long long *frame, uint32_t nslots, size_t npix, uint32_t *rgba,
                                     float *depth /* may be nullptr */) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long best = frame[i];
        for (uint32_t s = 1; s < nslots; ++s) {
            const unsigned long long v = frame[(size_t)s * npix + i];
            best = v < best ? v : best;
        }
        rgba[i] = (uint32_t)best;
        if (depth) depth[i] = __uint_as_float((uint32_t)(best >> 32));
    }
}

// host side, defined in fs3d.cu
int raymarch_world(fs3d_world *w, const fs3d_camera *cam, uint32_t width, uint32_t height, uint32_t mode,
                   uint8_t *host_rgba8, float *host_depth, unsigned long long *frame_slot = nullptr);

// 255 ascending linear-light thresholds: value >= thr[i] encodes to at least i + 1
inline void srgb_thresholds(float *thr) {
    for (int i = 1; i <= 255; ++i) {
        double c = (i - 0.5) / 255.0;
        double lin = c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4);
        thr[i - 1] = (float)lin;
    }
}

}  // namespace fs3d
