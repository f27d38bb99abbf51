"""Builds libfs3d.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

No reference counterpart: the reference's only build is its Vulkan/SDL CMake
(/root/reference/CMakeLists.txt:4-46), which has no CUDA target.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfs3d.so")
# tuning experiments: FS3D_NVCC_EXTRA="-DFOO=1" FS3D_LIB_OUT=/path/libfs3d_foo.so python -m fallingsand3d_b200.build --force
SOURCES = [os.path.join(CSRC, "fs3d.cu"), os.path.join(CSRC, "fs3d_v2.cu"), os.path.join(CSRC, "fs3d_s4.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("common.cuh", "bitslice.cuh", "bitslice3.cuh", "aux_kernels.cuh", "step_kernel.cuh", "step_dispatch.cuh", "step4_kernel.cuh", "raymarch.cuh")] + [
    os.path.join(HERE, "..", "include", "fs3d.h")
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "--threads", "3",   # the two translation units (schedule version 1 / version 2 kernels) compile side by side
    "--fmad=false",  # raymarch parity: no silent FMA contraction anywhere in this TU (integer kernels unaffected)
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.environ.get("FS3D_LIB_OUT", LIB)
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("FS3D_NVCC_EXTRA", "").split() + ["-o", out] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libfs3d.so")
    if out != LIB:
        return out
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
