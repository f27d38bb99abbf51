// fs3d_s4.cu — instantiations and launchers of the four-steps-per-pass kernel and its halo delivery (step4_kernel.cuh).
// A translation unit of its own so that nvcc compiles it beside fs3d.cu and fs3d_v2.cu (--threads).
// No reference counterpart; call site: /root/reference/src/engine/engine.cpp:59-70.
#include "step4_kernel.cuh"

namespace fs3d {

#ifndef FS3D_S4_THREADS
#define FS3D_S4_THREADS 384
#endif
constexpr int STEP4_THREADS = FS3D_S4_THREADS;

template <int XW, int NBR>
static cudaError_t launch_one(const Step4Params &p, unsigned grid, cudaStream_t stream) {
    constexpr uint32_t smem = step4_smem_bytes<STEP4_THREADS>();
    cudaError_t e = cudaFuncSetAttribute(step4_kernel<XW, NBR, STEP4_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    step4_kernel<XW, NBR, STEP4_THREADS><<<grid, STEP4_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

// xw = warps per band (1: nx = 1024, 2: nx = 2048, 4: nx = 4096); grid = CTAs (one per SM: the kernel parks 210 KB of pipeline state in shared memory);
// nbr = the slab has z-neighbours (coherent loads, bounded waits on the arrival counters)
cudaError_t step4_launch(int xw, int nbr, const Step4Params &p, unsigned grid, cudaStream_t stream) {
    if (xw == 1) return nbr ? launch_one<1, 1>(p, grid, stream) : launch_one<1, 0>(p, grid, stream);
    if (xw == 2) return nbr ? launch_one<2, 1>(p, grid, stream) : launch_one<2, 0>(p, grid, stream);
    if (xw == 4) return nbr ? launch_one<4, 1>(p, grid, stream) : launch_one<4, 0>(p, grid, stream);
    return cudaErrorInvalidValue;
}

cudaError_t halo4_launch(const Halo4Params &h, cudaStream_t stream) {
    halo4_kernel<<<HALO4_BLOCKS, 256, 0, stream>>>(h);
    return cudaGetLastError();
}

uint32_t step4_units_per_cta(int xw) { return STEP4_THREADS / 32 / (uint32_t)xw; }

}  // namespace fs3d
