"""CPU test of the kernel's bit-sliced helpers (csrc/bitslice.cuh compiled with g++): unit checks
plus a word-level emulation of a full step, built from the same helpers, against the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "bitslice_host_test.cpp")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host") / "bitslice_host_test")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", out, SRC], check=True)
    return out


def test_helper_units(exe):
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and "bad=0" in res.stdout


@pytest.mark.parametrize("pairfn", [0, 1])
@pytest.mark.parametrize("dims", [(32, 8, 6), (64, 9, 5), (96, 7, 3), (128, 16, 4), (32, 5, 1)])
def test_word_level_emulation_matches_oracle(exe, oracle, dims, pairfn):
    # pairfn = 1: the XY sub-step through xy_pair_substep0/1 (each block evaluated once, both rows of a
    # z-pair per call) — the functions the kernel runs; pairfn = 0: the per-column xy_substep
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, 3, 5)
    for t in range(8):
        kxy, kzy = oracle.key(7, t, 0), oracle.key(7, t, 1)
        out = subprocess.run([exe, str(nx), str(ny), str(nz), str(kxy), str(kzy), str(t), str(pairfn)], input=g.tobytes(),
                             capture_output=True, check=True).stdout
        e = np.frombuffer(out, dtype=np.uint8).reshape(g.shape)
        oracle.step(g, 7, t)
        assert np.array_equal(e, g), f"step {t}"


# ---- schedule version 2 (eight materials, three rank-encoded bit-planes: csrc/bitslice3.cuh) ----
SRC3 = os.path.join(ROOT, "tests", "host", "bitslice3_host_test.cpp")


@pytest.fixture(scope="module")
def exe3(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host3") / "bitslice3_host_test")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", out, SRC3], check=True)
    return out


def test_v2_helper_units(exe3):
    res = subprocess.run([exe3], capture_output=True, text=True)
    assert res.returncode == 0 and "bad=0" in res.stdout, res.stdout


@pytest.mark.parametrize("dims,scene", [((32, 8, 6), 5), ((64, 9, 5), 5), ((96, 7, 3), 6), ((128, 16, 4), 6), ((32, 5, 1), 5)])
def test_v2_word_level_emulation_matches_oracle(exe3, oracle, dims, scene):
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, scene, 5)
    for t in range(12):
        kxy, kzy = oracle.key(7, t, 0), oracle.key(7, t, 1)
        out = subprocess.run([exe3, str(nx), str(ny), str(nz), str(kxy), str(kzy), str(t)], input=g.tobytes(),
                             capture_output=True, check=True).stdout
        e = np.frombuffer(out, dtype=np.uint8).reshape(g.shape)
        oracle.step(g, 7, t, version=2)
        assert np.array_equal(e, g), f"step {t}"
