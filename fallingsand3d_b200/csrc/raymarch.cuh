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
// Mode FS3D_RM_VOXELS keeps camera + light and replaces map_the_world by an Amanatides-Woo walk
// through the uint8 grid with a 256-entry palette (the reference's unused colors[256],
// renderer.cpp:136-393, is the intended shape of that palette).  Crossing times are recomputed from the integer
// boundary index (boundary_t) rather than accumulated, which makes the walk's state a function of position alone; the
// kernel uses that to jump over empty 8^3 bricks and over planes held by other ranks in one move, exactly.
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
    uint32_t zheld0, zheld1;     // union of the slabs' plane ranges: rays jump over the planes outside it in one move
    const uint32_t *bricks[RM_MAX_SLABS];    // per slab: occupancy bits of its 8 x 8 x 8 bricks (nullptr: not built, march every voxel)
    unsigned long long *steps_out;           // += loop iterations of every ray (the host decides from it whether bricks pay)
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

// ---- empty-space skipping ------------------------------------------------------------------------------------------
// Occupancy of 8 x 8 x 8 bricks (grid coordinates, z bricks aligned to global z), one bit each.  A warp of the build
// kernel takes one (brick row y, brick layer z) of one 1024-voxel x segment: lane l ORs the 32-byte words
// x in [32 l, 32 l + 32) of the brick's (up to) 64 rows — four bricks per lane — and four ballots store the segment's
// 128 bits: brick r = 4 l + k of the segment sits in word k, bit l.
__host__ __device__ inline size_t brick_word(uint32_t nbx, uint32_t nby, uint32_t bx, uint32_t by, uint32_t bzl) {
    const uint32_t nseg = (nbx + 127u) >> 7;
    return (((size_t)bzl * nby + by) * nseg + (bx >> 7)) * 4u + (bx & 3u);
}
__host__ __device__ inline uint32_t brick_bit(uint32_t bx) { return (bx & 127u) >> 2; }
inline size_t brick_words(uint32_t nx, uint32_t ny, uint32_t z0, uint32_t z1) {
    const uint32_t nbx = nx / 8, nby = (ny + 7) / 8, nbz = ((z1 + 7) >> 3) - (z0 >> 3);
    return (size_t)nbz * nby * ((nbx + 127u) >> 7) * 4u;
}
__global__ void brick_build_kernel(const uint8_t *owned, uint32_t nx, uint32_t ny, uint32_t z0, uint32_t z1, uint32_t *bricks) {
    const uint32_t nbx = nx / 8, nby = (ny + 7) / 8, nbz = ((z1 + 7) >> 3) - (z0 >> 3), nseg = (nbx + 127u) >> 7;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t nwork = (uint64_t)nbz * nby * nseg;
    for (uint64_t wk = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5; wk < nwork; wk += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
        const uint32_t seg = (uint32_t)(wk % nseg), by = (uint32_t)((wk / nseg) % nby), bzl = (uint32_t)(wk / ((uint64_t)nseg * nby));
        const uint32_t xw = seg * 32u + lane;                       // 32-voxel word of the row
        uint32_t acc[4] = {0u, 0u, 0u, 0u};                         // OR of bytes [8k, 8k + 8) of the word over the brick's rows
        if (xw * 32u < nx) {
            const uint32_t zb = ((z0 >> 3) + bzl) << 3;
            for (uint32_t dz = 0; dz < 8u; ++dz) {
                const uint32_t z = zb + dz;
                if (z < z0 || z >= z1) continue;
                for (uint32_t dy = 0; dy < 8u; ++dy) {
                    const uint32_t y = by * 8u + dy;
                    if (y >= ny) break;
                    const uint4 *q = reinterpret_cast<const uint4 *>(owned + ((size_t)(z - z0) * ny + y) * nx + (size_t)xw * 32u);
                    const uint4 a = q[0], b = q[1];
                    acc[0] |= a.x | a.y; acc[1] |= a.z | a.w; acc[2] |= b.x | b.y; acc[3] |= b.z | b.w;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t bits = __ballot_sync(0xFFFFFFFFu, acc[k] != 0u);
            if (lane == 0) bricks[(((size_t)bzl * nby + by) * nseg + seg) * 4u + k] = bits;
        }
    }
}

// crossing time of lattice boundary k on axis a: ((k h - e) - o) * (1 / d), every operation rounded separately.  The
// walk below recomputes it from the integer boundary index after every move instead of accumulating t += dt, so the
// state of a ray depends only on WHERE it is, not on how it got there — which is what makes jumping over empty
// bricks (and over the planes another rank holds) exact: after a jump the ray is in the very state the
// voxel-by-voxel walk would have reached (oracle/fs3d_raymarch_oracle.c walks voxel by voxel and must agree).
__device__ __forceinline__ float boundary_t(int k, float h, float e, float o, float rcp) {
    return fmul(fsub(fsub(fmul((float)k, h), e), o), rcp);
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
    uint32_t nsteps = 0;                 // loop iterations of this ray's walk (voxel moves + jumps)

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
            float tnext[3], rcp[3];
            int last_axis = -1;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float pos = fadd(o[a], fmul(tmin, d[a]));            // world coordinate at entry
                float f = fdiv(fadd(pos, e[a]), p.h);                 // lattice coordinate [0, n]
                int i = (int)floorf(f);
                if (i < 0) i = 0;
                if (i > n[a] - 1) i = n[a] - 1;
                idx[a] = i;
                stp[a] = d[a] > 0.0f ? 1 : (d[a] < 0.0f ? -1 : 0);
                rcp[a] = stp[a] != 0 ? fdiv(1.0f, d[a]) : 0.0f;
                tnext[a] = stp[a] != 0 ? boundary_t(i + (stp[a] > 0 ? 1 : 0), p.h, e[a], o[a], rcp[a]) : INFINITY;
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
            // Jump out of an empty axis-aligned box lo <= idx < hi (lattice coordinates) in ONE move, landing in exactly
            // the state the voxel-by-voxel walk reaches when it leaves the box: the exit event is the earliest crossing of
            // a box face (ties: lowest axis, like the walk's choice of axis); on the other axes every lattice boundary is
            // crossed whose time is earlier — or equal, on a lower axis — found from the exit point and corrected with
            // the same boundary_t comparisons the walk makes.  Returns false if the ray leaves the grid.
            auto jump = [&](const int (&lo)[3], const int (&hi)[3]) -> bool {
                float tex[3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
                    tex[a] = stp[a] != 0 ? boundary_t(stp[a] > 0 ? hi[a] : lo[a], p.h, e[a], o[a], rcp[a]) : INFINITY;
                int ax = 0;
                if (tex[1] < tex[ax]) ax = 1;
                if (tex[2] < tex[ax]) ax = 2;
                const float T = tex[ax];
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    if (b == ax) { idx[b] = stp[b] > 0 ? hi[b] : lo[b] - 1; continue; }
                    if (stp[b] == 0) continue;
                    auto crossed = [&](int k) { const float tb = boundary_t(k, p.h, e[b], o[b], rcp[b]); return tb < T || (tb == T && b < ax); };
                    int g = (int)floorf(fdiv(fadd(fadd(o[b], fmul(T, d[b])), e[b]), p.h));
                    if (stp[b] > 0) {          // moving up crosses boundary i to enter voxel i
                        g = g < idx[b] ? idx[b] : (g > hi[b] - 1 ? hi[b] - 1 : g);
                        while (g > idx[b] && !crossed(g)) --g;
                        while (g + 1 <= hi[b] - 1 && crossed(g + 1)) ++g;
                    } else {                   // moving down crosses boundary i to leave voxel i
                        g = g > idx[b] ? idx[b] : (g < lo[b] ? lo[b] : g);
                        while (g < idx[b] && !crossed(g + 1)) ++g;
                        while (g - 1 >= lo[b] && crossed(g)) --g;
                    }
                    idx[b] = g;
                }
                if (idx[ax] < 0 || idx[ax] >= n[ax]) return false;
#pragma unroll
                for (int a = 0; a < 3; ++a)
                    if (stp[a] != 0) tnext[a] = boundary_t(idx[a] + (stp[a] > 0 ? 1 : 0), p.h, e[a], o[a], rcp[a]);
                t = T;
                last_axis = ax;
                return true;
            };
            const int max_steps = n[0] + n[1] + n[2] + 3;
            uint32_t cbx = 0xFFFFFFFFu, cby = 0xFFFFFFFFu, cbz = 0xFFFFFFFFu;     // brick whose occupancy bit is cached ...
            int csl = -1;                                                         // ... of which slab (a brick can straddle two)
            bool cached_occ = true;
            int s = 0;
            for (; s < max_steps; ++s) {
                const int gx = idx[0], gy = n[1] - 1 - idx[1], gz = idx[2];
                if ((uint32_t)gz < p.zheld0 || (uint32_t)gz >= p.zheld1) {
                    // planes no slab of this launch holds (another rank's): transparent here — cross them in one move
                    const int lo[3] = {0, 0, (uint32_t)gz < p.zheld0 ? 0 : (int)p.zheld1};
                    const int hi[3] = {n[0], n[1], (uint32_t)gz < p.zheld0 ? (int)p.zheld0 : n[2]};
                    if (!jump(lo, hi)) break;
                    continue;
                }
                int sl = 0;
#pragma unroll 1
                for (int q = 0; q < p.nslabs; ++q) if ((uint32_t)gz >= p.slab_z0[q] && (uint32_t)gz < p.slab_z1[q]) sl = q;
                if (p.bricks[sl] != nullptr) {
                    const uint32_t bx = (uint32_t)gx >> 3, by = (uint32_t)gy >> 3, bz = (uint32_t)gz >> 3;
                    if (bx != cbx || by != cby || bz != cbz || sl != csl) {
                        cbx = bx; cby = by; cbz = bz; csl = sl;
                        const uint32_t nbx = p.nx >> 3, nby = (p.ny + 7u) >> 3;
                        cached_occ = (p.bricks[sl][brick_word(nbx, nby, bx, by, bz - (p.slab_z0[sl] >> 3))] >> brick_bit(bx)) & 1u;
                    }
                    if (!cached_occ) {
                        // the brick (clipped to the slab's planes) in lattice coordinates: grid rows [8 by, 8 by + 8) are
                        // lattice rows [ny - 8 by - 8, ny - 8 by)
                        const int zlo = max((int)(bz << 3), (int)p.slab_z0[sl]), zhi = min((int)(bz << 3) + 8, (int)p.slab_z1[sl]);
                        const int lo[3] = {(int)(bx << 3), max(0, n[1] - (int)(by << 3) - 8), zlo};
                        const int hi[3] = {(int)(bx << 3) + 8, n[1] - (int)(by << 3), zhi};
                        if (!jump(lo, hi)) break;
                        continue;
                    }
                }
                const uint8_t m = p.slab_ptr[sl][(size_t)gx + (size_t)p.nx * ((size_t)gy + (size_t)p.ny * ((size_t)gz - p.slab_z0[sl]))];
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
                tnext[a] = boundary_t(idx[a] + (stp[a] > 0 ? 1 : 0), p.h, e[a], o[a], rcp[a]);
                last_axis = a;
            }
            nsteps = (uint32_t)s;
        }
    }

    if (p.steps_out) {                   // one atomic per warp (warps at the image edge are partial)
        const unsigned act = __activemask();
        const uint32_t sum = __reduce_add_sync(act, nsteps);
        if (((threadIdx.y * blockDim.x + threadIdx.x) & 31u) == (uint32_t)(__ffs(act) - 1)) atomicAdd(p.steps_out, (unsigned long long)sum);
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
