"""ctypes loader for ``libmq3d.so`` (the C ABI of include/mq3d.h).

There is deliberately no fallback: if the shared object is missing or a call fails, a
RuntimeError is raised (north star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmq3d.so")

MQ3D_OK = 0
MQ3D_ERR_CUDA = 1
MQ3D_ERR_INVALID = 2
MQ3D_ERR_NO_BLOCK_TOUCHED = 3
MQ3D_ERR_STATE = 4

ATTR_TSDF_WEIGHT = 1
ATTR_COLOR = 2

# every symbol include/mq3d.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "mq3d_last_error", "mq3d_version", "mq3d_selftest_rcp",
    "mq3d_grid_create", "mq3d_grid_destroy", "mq3d_grid_reset", "mq3d_grid_reserve",
    "mq3d_grid_num_blocks", "mq3d_grid_info", "mq3d_grid_pool", "mq3d_grid_export", "mq3d_grid_import",
    "mq3d_grid_set_partition", "mq3d_grid_set_ghost_mode", "mq3d_grid_set_batch_gates", "mq3d_grid_ghost_select",
    "mq3d_grid_ghost_counts", "mq3d_grid_peer_descriptor", "mq3d_grid_ghost_pull",
    "mq3d_depth_prepare", "mq3d_touch", "mq3d_integrate", "mq3d_integrate_sequence",
    "mq3d_color_resample", "mq3d_integrate_sequence_rgbx",
    "mq3d_extract_mesh_count", "mq3d_extract_mesh_fill", "mq3d_extract_mesh", "mq3d_extract_points_count",
    "mq3d_extract_points_fill", "mq3d_extract_mesh_colors", "mq3d_extract_points_colors", "mq3d_confidence",
    "mq3d_mesh_filter", "mq3d_odometry_information",
    "mq3d_scene_create", "mq3d_scene_destroy", "mq3d_scene_add_triangles",
    "mq3d_scene_create_rays_pinhole", "mq3d_scene_cast_rays",
]


class SeqStats(C.Structure):
    _fields_ = [("frames_integrated", C.c_int64), ("block_visits", C.c_int64),
                ("blocks_loaded", C.c_int64), ("num_blocks", C.c_int64), ("batches", C.c_int64),
                ("voxel_updates", C.c_int64), ("touch_ms", C.c_double), ("integrate_ms", C.c_double),
                ("slow_div_batches", C.c_int64)]


class MeshFilterInfo(C.Structure):
    _fields_ = [("components", C.c_int64), ("components_kept", C.c_int64), ("largest_component", C.c_int64),
                ("fallback_largest", C.c_int64), ("input_triangles", C.c_int64), ("removed_triangles", C.c_int64),
                ("non_manifold_edges", C.c_int64)]


class Mq3dError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for this path.")
    l = C.CDLL(LIB_PATH)
    l.mq3d_last_error.restype = C.c_char_p
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    pd = C.POINTER(C.c_double)
    pf = C.POINTER(C.c_float)
    sig = {
        "mq3d_selftest_rcp": [C.c_uint, C.c_uint, C.POINTER(C.c_ulonglong)],
        "mq3d_grid_create": [f32, i32, i64, i32, i32, C.POINTER(vp)],
        "mq3d_grid_destroy": [vp],
        "mq3d_grid_reset": [vp, vp],
        "mq3d_grid_reserve": [vp, i64, vp],
        "mq3d_grid_num_blocks": [vp, C.POINTER(i64), vp],
        "mq3d_grid_info": [vp, C.POINTER(f32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)],
        "mq3d_grid_pool": [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
        "mq3d_grid_export": [vp, vp, vp, vp, vp, vp],
        "mq3d_grid_import": [vp, vp, vp, vp, vp, i64, vp],
        "mq3d_grid_set_partition": [vp, i32, i32, i32],
        "mq3d_grid_set_ghost_mode": [vp, i32],
        "mq3d_grid_set_batch_gates": [vp, vp, i32],
        "mq3d_grid_ghost_select": [vp, i32, C.POINTER(i64), vp, vp, vp, vp, vp],
        "mq3d_grid_ghost_counts": [vp, C.POINTER(i64), vp],
        "mq3d_grid_peer_descriptor": [vp, vp, vp],
        "mq3d_grid_ghost_pull": [vp, vp, C.POINTER(i64), vp],
        "mq3d_depth_prepare": [vp, i32, i32, i32, pd, pd, vp, vp, vp, f64, C.c_int32, vp, vp, vp],
        "mq3d_touch": [vp, vp, i32, i32, pd, pd, f32, f32, f32, vp, C.POINTER(i64), vp],
        "mq3d_integrate": [vp, vp, i64, vp, i32, i32, vp, i32, i32, pd, pd, pd, f32, f32, f32, vp],
        "mq3d_integrate_sequence": [vp, vp, vp, i32, i32, i32, vp, i32, i32, pd, pd, pd, f32, f32, f32, i32,
                                    C.POINTER(SeqStats), vp],
        "mq3d_color_resample": [vp, i32, i32, i32, i32, i32, pd, pd, vp, i32, vp],
        "mq3d_integrate_sequence_rgbx": [vp, vp, vp, i32, i32, i32, vp, pd, pd, f32, f32, f32, i32,
                                         C.POINTER(SeqStats), vp],
        "mq3d_extract_mesh_count": [vp, f32, C.POINTER(i64), C.POINTER(i64), vp],
        "mq3d_extract_mesh_fill": [vp, vp, vp, vp, vp, vp],
        "mq3d_extract_mesh": [vp, f32, vp, vp, vp, vp, vp, i64, i64, C.POINTER(i64), C.POINTER(i64), vp],
        "mq3d_extract_points_count": [vp, f32, C.POINTER(i64), vp],
        "mq3d_extract_points_fill": [vp, vp, vp, vp, vp],
        "mq3d_extract_mesh_colors": [vp, vp, vp],
        "mq3d_extract_points_colors": [vp, vp, vp],
        "mq3d_confidence": [vp, vp, i32, i32, i32, pf, pf, pf, i32, f64, f64, vp, vp, vp],
        "mq3d_odometry_information": [vp, vp, i32, i32, pd, pd, f32, f32, f32, pd, i32, vp],
        "mq3d_mesh_filter": [vp, vp, vp, i64, vp, i64, i64, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(i64),
                             C.POINTER(MeshFilterInfo), i32, vp],
        "mq3d_scene_create": [i32, C.POINTER(vp)],
        "mq3d_scene_destroy": [vp],
        "mq3d_scene_add_triangles": [vp, vp, i64, vp, i64, vp],
        "mq3d_scene_create_rays_pinhole": [pd, pd, i32, i32, vp, vp],
        "mq3d_scene_cast_rays": [vp, vp, i64, vp, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = l
    return l


def check(rc: int):
    if rc != MQ3D_OK:
        msg = lib().mq3d_last_error().decode(errors="replace")
        raise Mq3dError(rc, msg)


def dptr(t) -> C.c_void_p:
    """Device pointer of a (contiguous CUDA) torch tensor, or NULL for None."""
    if t is None:
        return C.c_void_p(None)
    return C.c_void_p(t.data_ptr())


def darr(a):
    """ctypes double* view of a contiguous float64 numpy array."""
    return a.ctypes.data_as(C.POINTER(C.c_double))


def farr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))
