// fs3d_s4.cu — instantiations and launcher of the four-steps-per-pass kernel (step4_kernel.cuh).  A translation unit
// of its own so that nvcc compiles it beside fs3d.cu and fs3d_v2.cu (--threads).
// No reference counterpart; call site: /root/reference/src/engine/engine.cpp:59-70.
#include "step4_kernel.cuh"

namespace fs3d {

constexpr int STEP4_THREADS = 256;

// xw = warps per band (1: nx = 1024, 2: nx = 2048); grid = CTAs (one per SM: the kernel takes 212 KB of shared memory)
cudaError_t step4_launch(int xw, const Step4Params &p, unsigned grid, cudaStream_t stream) {
    cudaError_t e;
    if (xw == 1) {
        constexpr uint32_t smem = step4_smem_bytes<1, STEP4_THREADS>();
        e = cudaFuncSetAttribute(step4_kernel<1, STEP4_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        step4_kernel<1, STEP4_THREADS><<<grid, STEP4_THREADS, smem, stream>>>(p);
    } else {
        constexpr uint32_t smem = step4_smem_bytes<2, STEP4_THREADS>();
        e = cudaFuncSetAttribute(step4_kernel<2, STEP4_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        step4_kernel<2, STEP4_THREADS><<<grid, STEP4_THREADS, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

uint32_t step4_units_per_cta(int xw) { return STEP4_THREADS / 32 / (uint32_t)xw; }

}  // namespace fs3d
