// step_dispatch.cuh — thread-block shapes and the table of step_kernel instantiations, shared by the translation units
// that hold them: fs3d.cu (schedule version 1, Rules1) and fs3d_v2.cu (schedule version 2, Rules3).  Splitting the
// instantiations over two files lets nvcc compile them in parallel (--threads 2).
#pragma once
#include "step_kernel.cuh"

namespace fs3d {

#ifndef FS3D_STEP_THREADS
#define FS3D_STEP_THREADS 256
#endif
constexpr int STEP_THREADS = FS3D_STEP_THREADS;         // J = 2 kernels (nx > 1024)
#ifndef FS3D_STEP_THREADS_J1
#define FS3D_STEP_THREADS_J1 384   // 12 warps per SM: 4.5 % faster than 256 x 2 CTAs at 1024^3 (profiles/r01d_experiments_xy_pair.txt)
#endif
constexpr int STEP_THREADS_J1 = FS3D_STEP_THREADS_J1;   // J = 1 kernels (nx <= 1024) on grids large enough to be bandwidth-bound
// kernel shapes: jidx 0: J = 1 (nx <= 1024), 1: J = 2 (nx <= 2048), 2: J = 2 x 2 warps (nx <= 4096),
//                3: J = 1 in STEP_THREADS-sized CTAs — small grids are launch/latency-bound and ran 15 % slower in 384-thread CTAs
static int step_threads(int jidx) { return jidx == 0 ? STEP_THREADS_J1 : STEP_THREADS; }
static uint32_t warps_per_pair(int jidx) { return jidx == 2 ? 2u : 1u; }

typedef void (*StepFn)(const StepParams);
#define FS3D_TH(J) ((J) == 1 ? STEP_THREADS_J1 : STEP_THREADS)
#define FS3D_ROW(R, J, XW, SK, PU) \
    {{step_kernel<R, J, XW, 0, 0, SK, 1, PU, FS3D_TH(J)>, step_kernel<R, J, XW, 0, 1, SK, 1, PU, FS3D_TH(J)>}, \
     {step_kernel<R, J, XW, 1, 0, SK, 1, PU, FS3D_TH(J)>, step_kernel<R, J, XW, 1, 1, SK, 1, PU, FS3D_TH(J)>}}
#define FS3D_ROW2(R, J, XW, SK, PU) {step_kernel<R, J, XW, 0, 0, SK, 2, PU, FS3D_TH(J)>, step_kernel<R, J, XW, 1, 0, SK, 2, PU, FS3D_TH(J)>}
#define FS3D_ROW_S(R, J, XW, SK, PU) \
    {{step_kernel<R, J, XW, 0, 0, SK, 1, PU, STEP_THREADS>, step_kernel<R, J, XW, 0, 1, SK, 1, PU, STEP_THREADS>}, \
     {step_kernel<R, J, XW, 1, 0, SK, 1, PU, STEP_THREADS>, step_kernel<R, J, XW, 1, 1, SK, 1, PU, STEP_THREADS>}}
#define FS3D_ROW2_S(R, J, XW, SK, PU) {step_kernel<R, J, XW, 0, 0, SK, 2, PU, STEP_THREADS>, step_kernel<R, J, XW, 1, 0, SK, 2, PU, STEP_THREADS>}
#define FS3D_SHAPES(M, R, SK, PU) {M(R, 1, 1, SK, PU), M(R, 2, 1, SK, PU), M(R, 2, 2, SK, PU), M##_S(R, 1, 1, SK, PU)}
// ns = 1: one step (any parity); ns = 2: steps t, t + 1 fused, t even; push = fused halo push over peer memory;
// version = schedule version of the world (1: Rules1, 2: Rules3)
template <class R>
inline StepFn step_fn_of(int jidx, int ox, int todd, int skip, int ns, int push) {
    static StepFn tab1[2][2][4][2][2] = {
        {FS3D_SHAPES(FS3D_ROW, R, 0, 0), FS3D_SHAPES(FS3D_ROW, R, 1, 0)},
        {FS3D_SHAPES(FS3D_ROW, R, 0, 1), FS3D_SHAPES(FS3D_ROW, R, 1, 1)},
    };
    static StepFn tab2[2][2][4][2] = {
        {FS3D_SHAPES(FS3D_ROW2, R, 0, 0), FS3D_SHAPES(FS3D_ROW2, R, 1, 0)},
        {FS3D_SHAPES(FS3D_ROW2, R, 0, 1), FS3D_SHAPES(FS3D_ROW2, R, 1, 1)},
    };
    return ns == 2 ? tab2[push][skip][jidx][ox] : tab1[push][skip][jidx][ox][todd];
}

StepFn step_fn_v1(int jidx, int ox, int todd, int skip, int ns, int push);   // fs3d.cu
StepFn step_fn_v2(int jidx, int ox, int todd, int skip, int ns, int push);   // fs3d_v2.cu

}  // namespace fs3d
