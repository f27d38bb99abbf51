"""CPU tests of the drop-in boundary: libfs3d.so builds, loads, exports every symbol include/fs3d.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fs3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fs3d_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(fs3d):
    from fallingsand3d_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 25
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_declared_symbol(fs3d):
    from fallingsand3d_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.fs3d_schedule_version() == 1


def test_sass_is_sm100a_with_256bit_accesses(fs3d):
    import shutil
    import subprocess
    from fallingsand3d_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "LDG.E" in out and ".256" in out, "step kernel should use 256-bit global loads"
    assert "STG.E" in out


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_fallback(fs3d):
    with pytest.raises(fs3d.Fs3dError) as ei:
        fs3d.VoxelWorld(32, 8, 8)
    assert ei.value.code in (-5, -6)
    assert "ERROR" in str(ei.value)


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under fallingsand3d_b200/ may import, include,
    # link or dlopen it (comments citing it are fine)
    pkg = os.path.join(ROOT, "fallingsand3d_b200")
    bad = re.compile(r"^\s*(import\s+oracle|from\s+oracle|from\s+\.\.?oracle)|#\s*include\s*[\"<][^\">]*oracle|"
                     r"libfs3d_oracle|oracle\.(step|run|lib)\(", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), os.path.join(dirpath, f)


def _build_engine_loop(tmp_path):
    import subprocess
    exe = str(tmp_path / "engine_loop")
    lib_dir = os.path.join(ROOT, "fallingsand3d_b200")
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "engine_loop.cpp"), "-L" + lib_dir, "-lfs3d", "-Wl,-rpath," + lib_dir, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.skipif(_has_cuda(), reason="only meaningful on a box without a GPU")
def test_cpp_wrapper_builds_and_reports_errors_like_the_engine(fs3d, tmp_path):
    # include/fs3d.hpp + examples/engine_loop.cpp compile as plain C++17 against the C ABI; without a GPU
    # the wrapper logs "ERROR: ..." and throws (util::displayError, debug.cpp:23-27) -> the example exits 2
    import subprocess
    exe = _build_engine_loop(tmp_path)
    res = subprocess.run([exe, "32", "2"], capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 2 and res.stderr.startswith("ERROR: fs3d:")


@pytest.mark.gpu
def test_cpp_engine_loop_runs_on_the_gpu(fs3d, tmp_path):
    import subprocess
    exe = _build_engine_loop(tmp_path)
    res = subprocess.run([exe, "64", "100"], capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout + res.stderr
    assert "steps 100" in res.stdout and "sand 4096 -> 4096" in res.stdout
    assert os.path.getsize(tmp_path / "frame_100.ppm") == len("P6\n850 450\n255\n") + 850 * 450 * 3


def test_header_is_plain_c99(tmp_path):
    # the boundary is a C ABI: include/fs3d.h must compile as C (not only C++), and the constants the
    # file format and the bindings rely on must have the documented values
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text("""
#include <stddef.h>
#include "fs3d.h"
_Static_assert(FS3D_CKPT_HEADER_BYTES == 80, "checkpoint header");
_Static_assert(FS3D_IPC_BLOB_BYTES == 256, "ipc blob");
_Static_assert(sizeof(fs3d_camera) == 20, "camera: pos[3], yaw, aspect");
_Static_assert(offsetof(fs3d_view, pitch_y) == 32 && sizeof(fs3d_view) == 56, "view layout used by _lib.View");
_Static_assert(offsetof(fs3d_desc, seed) == 16 && offsetof(fs3d_desc, devices) == 32 && sizeof(fs3d_desc) == 48, "desc layout used by _lib.Desc");
int use(fs3d_world *w) { uint64_t s = 0; return fs3d_step(w, 1) + fs3d_step_index(w, &s) + fs3d_save(w, "x"); }
""")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                          "-c", str(src), "-o", str(tmp_path / "abi.o")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
