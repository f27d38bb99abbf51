"""ctypes binding of libfs3d.so — the C ABI declared in include/fs3d.h.

There is no CPU fallback: if the CUDA library is missing this module raises, it never
substitutes another implementation (the oracle under oracle/ is test infrastructure only).
Reference seams: /root/reference/src/engine/engine.cpp:59-70 (frame loop),
/root/reference/src/engine/rendering/materials.cpp:388-418 (volume hand-off).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FS3D_LIB", os.path.join(HERE, "libfs3d.so"))   # FS3D_LIB: tuning experiments only


class Desc(C.Structure):
    _fields_ = [("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32), ("seed", C.c_uint64),
                ("n_gpus", C.c_int32), ("devices", C.POINTER(C.c_int32)), ("flags", C.c_uint32)]


class View(C.Structure):
    _fields_ = [("dev_ptr", C.c_void_p), ("device", C.c_int32), ("nx", C.c_uint32), ("ny", C.c_uint32),
                ("z0", C.c_uint32), ("z1", C.c_uint32), ("pitch_y", C.c_uint64), ("pitch_z", C.c_uint64),
                ("step", C.c_uint64)]


class Export(C.Structure):
    _fields_ = [("fd", C.c_int32 * 2), ("alloc_bytes", C.c_uint64), ("first_cell_offset", C.c_uint64), ("front", C.c_uint32),
                ("device", C.c_int32), ("nx", C.c_uint32), ("ny", C.c_uint32), ("z0", C.c_uint32), ("z1", C.c_uint32),
                ("pitch_y", C.c_uint64), ("pitch_z", C.c_uint64), ("step", C.c_uint64)]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("yaw_deg", C.c_float), ("aspect", C.c_float)]


class Halo(C.Structure):
    _fields_ = [("send_lo", C.c_void_p), ("send_hi", C.c_void_p), ("recv_lo", C.c_void_p), ("recv_hi", C.c_void_p),
                ("plane_bytes", C.c_uint64), ("stream", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/fs3d.h declares
_W = C.c_void_p
SIGNATURES = {
    "fs3d_create": (C.c_int, [C.POINTER(Desc), C.POINTER(_W)]),
    "fs3d_create_slab": (C.c_int, [C.POINTER(Desc), C.c_uint32, C.c_uint32, C.POINTER(_W)]),
    "fs3d_destroy": (None, [_W]),
    "fs3d_set_cell": (C.c_int, [_W, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8]),
    "fs3d_get_cell": (C.c_int, [_W, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8)]),
    "fs3d_fill_box": (C.c_int, [_W, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint8]),
    "fs3d_paint_sphere": (C.c_int, [_W, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint8, C.c_int]),
    "fs3d_generate": (C.c_int, [_W, C.c_int, C.c_uint64]),
    "fs3d_upload": (C.c_int, [_W, C.c_void_p]),
    "fs3d_download": (C.c_int, [_W, C.c_void_p]),
    "fs3d_step": (C.c_int, [_W, C.c_uint32]),
    "fs3d_sync": (C.c_int, [_W]),
    "fs3d_kernel_launches": (C.c_int, [_W, C.POINTER(C.c_uint64)]),
    "fs3d_step_index": (C.c_int, [_W, C.POINTER(C.c_uint64)]),
    "fs3d_step_timed": (C.c_int, [_W, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "fs3d_step_host": (C.c_int, [_W, C.c_void_p, C.c_void_p, C.c_uint32]),
    "fs3d_upload_packed": (C.c_int, [_W, C.c_void_p]),
    "fs3d_download_packed": (C.c_int, [_W, C.c_void_p]),
    "fs3d_step_host_packed": (C.c_int, [_W, C.c_void_p, C.c_void_p, C.c_uint32]),
    "fs3d_slab_step_host_begin": (C.c_int, [_W, C.c_void_p]),
    "fs3d_slab_step_host": (C.c_int, [_W, C.c_void_p, C.c_void_p, C.c_uint32]),
    "fs3d_histogram": (C.c_int, [_W, C.POINTER(C.c_uint64)]),
    "fs3d_digest": (C.c_int, [_W, C.POINTER(C.c_uint64)]),
    "fs3d_activity": (C.c_int, [_W, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fs3d_num_slabs": (C.c_int, [_W, C.POINTER(C.c_int32)]),
    "fs3d_volume_view": (C.c_int, [_W, C.c_int32, C.POINTER(View)]),
    "fs3d_volume_export_fd": (C.c_int, [_W, C.c_int32, C.POINTER(Export)]),
    "fs3d_set_palette": (C.c_int, [_W, C.POINTER(C.c_float)]),
    "fs3d_raymarch": (C.c_int, [_W, C.POINTER(Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "fs3d_raymarch_bricks_in_use": (C.c_int, [_W, C.c_int32]),
    "fs3d_raymarch_depth": (C.c_int, [_W, C.POINTER(Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "fs3d_slab_halo": (C.c_int, [_W, C.c_int, C.POINTER(Halo)]),
    "fs3d_slab_pass_steps": (C.c_int, [_W, C.c_uint32]),
    "fs3d_slab_step_edges": (C.c_int, [_W]),
    "fs3d_slab_step_interior": (C.c_int, [_W]),
    "fs3d_slab_step_finish": (C.c_int, [_W]),
    "fs3d_slab_ipc_export": (C.c_int, [_W, C.c_void_p, C.c_uint64]),
    "fs3d_slab_ipc_attach": (C.c_int, [_W, C.c_void_p, C.c_void_p]),
    "fs3d_slab_attach_local": (C.c_int, [_W, _W, _W]),
    "fs3d_slab_push_halos": (C.c_int, [_W]),
    "fs3d_slab_can_fuse4": (C.c_int, [_W]),
    "fs3d_slab_allow_fuse4": (C.c_int, [_W, C.c_int]),
    "fs3d_push_wait_stats": (C.c_int, [_W, C.POINTER(C.c_uint64)]),
    "fs3d_frame_export": (C.c_int, [_W, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]),
    "fs3d_frame_attach": (C.c_int, [_W, C.c_void_p, C.c_uint32]),
    "fs3d_frame_attach_local": (C.c_int, [_W, _W, C.c_uint32]),
    "fs3d_raymarch_to_frame": (C.c_int, [_W, C.POINTER(Camera), C.c_uint32]),
    "fs3d_frame_resolve": (C.c_int, [_W, C.c_void_p]),
    "fs3d_save": (C.c_int, [_W, C.c_char_p]),
    "fs3d_load": (C.c_int, [_W, C.c_char_p]),
    "fs3d_last_error": (C.c_char_p, []),
    "fs3d_schedule_version": (C.c_int, []),
    "fs3d_world_schedule_version": (C.c_int, [_W]),
}

_lib = None


def load():
    """Loads libfs3d.so (building is __graft_entry__.build()'s / build.py's job). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m fallingsand3d_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
