"""Child process of tests/test_export_gpu.py: imports a slab buffer that ANOTHER process exported as a POSIX file
descriptor (fs3d_volume_export_fd) through the CUDA driver API — what a foreign consumer does — copies the cells out
and prints their digest and histogram.  No libfs3d here: only libcuda and the oracle's digest.

  python vmm_import_child.py FD ALLOC_BYTES FIRST_CELL_OFFSET NX NY Z0 Z1 DEVICE"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402

fd, alloc, off, nx, ny, z0, z1, dev = (int(v) for v in sys.argv[1:9])
cu = C.CDLL("libcuda.so.1")


def chk(r, what):
    if r != 0:
        raise SystemExit(f"{what}: CUresult {r}")


class Loc(C.Structure):
    _fields_ = [("type", C.c_int), ("id", C.c_int)]


class AccessDesc(C.Structure):
    _fields_ = [("location", Loc), ("flags", C.c_int)]


chk(cu.cuInit(0), "cuInit")
device, ctx = C.c_int(), C.c_void_p()
chk(cu.cuDeviceGet(C.byref(device), dev), "cuDeviceGet")
chk(cu.cuDevicePrimaryCtxRetain(C.byref(ctx), device), "cuDevicePrimaryCtxRetain")
chk(cu.cuCtxSetCurrent(ctx), "cuCtxSetCurrent")
handle = C.c_ulonglong()
CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR = 1
chk(cu.cuMemImportFromShareableHandle(C.byref(handle), C.c_void_p(fd), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
    "cuMemImportFromShareableHandle")
ptr = C.c_ulonglong()
chk(cu.cuMemAddressReserve(C.byref(ptr), C.c_size_t(alloc), C.c_size_t(0), C.c_ulonglong(0), C.c_ulonglong(0)), "cuMemAddressReserve")
chk(cu.cuMemMap(ptr, C.c_size_t(alloc), C.c_size_t(0), handle, C.c_ulonglong(0)), "cuMemMap")
acc = AccessDesc(Loc(1, dev), 1)       # CU_MEM_LOCATION_TYPE_DEVICE, CU_MEM_ACCESS_FLAGS_PROT_READ
chk(cu.cuMemSetAccess(ptr, C.c_size_t(alloc), C.byref(acc), C.c_size_t(1)), "cuMemSetAccess")
cells = np.empty((z1 - z0, ny, nx), np.uint8)
chk(cu.cuMemcpyDtoH_v2(cells.ctypes.data_as(C.c_void_p), C.c_ulonglong(ptr.value + off), C.c_size_t(cells.size)), "cuMemcpyDtoH")
print(json.dumps({"digest": hex(oracle.digest(cells, z0)), "histogram": [int(v) for v in oracle.histogram(cells)[:8]]}))
chk(cu.cuMemUnmap(ptr, C.c_size_t(alloc)), "cuMemUnmap")
chk(cu.cuMemAddressFree(ptr, C.c_size_t(alloc)), "cuMemAddressFree")
chk(cu.cuMemRelease(handle), "cuMemRelease")
