// fs3d_v2.cu — the step kernels of schedule version 2 (eight materials, three bit-planes: Rules3 in step_kernel.cuh,
// rule logic in bitslice3.cuh, spec in SCHEDULE.md §7).  A separate translation unit only so that nvcc builds the two
// families of instantiations in parallel; the host code that launches them is in fs3d.cu.
// No reference counterpart; call site: /root/reference/src/engine/engine.cpp:59-70.
#include "step_dispatch.cuh"

namespace fs3d {

StepFn step_fn_v2(int jidx, int ox, int todd, int skip, int ns, int push) { return step_fn_of<Rules3>(jidx, ox, todd, skip, ns, push); }

}  // namespace fs3d
