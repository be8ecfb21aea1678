"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- see mq3d_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmq3d_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mq3d_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_grid_create.restype = C.c_void_p
        _lib.orc_grid_create.argtypes = [C.c_float, C.c_int]
        _lib.orc_grid_destroy.argtypes = [C.c_void_p]
        _lib.orc_grid_num_blocks.restype = C.c_int64
        _lib.orc_grid_num_blocks.argtypes = [C.c_void_p]
        _lib.orc_depth_valid.restype = C.c_int
    return _lib


def set_num_threads(n: int) -> int:
    """OpenMP threads of the oracle's parallel loops (explicit: OMP_NUM_THREADS from a launcher is ignored).
    Returns the team size now in effect."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return int(lib().orc_get_max_threads())


def get_max_threads() -> int:
    return int(lib().orc_get_max_threads())


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def depth_to_linear(raw: np.ndarray, near: float, far: float) -> np.ndarray:
    raw = _f32(raw)
    out = np.empty_like(raw)
    lib().orc_depth_to_linear(_p(raw), C.c_int64(raw.size), C.c_double(near), C.c_double(far), _p(out))
    return out


def depth_valid(raw: np.ndarray) -> bool:
    raw = _f32(raw)
    return bool(lib().orc_depth_valid(_p(raw), C.c_int64(raw.size)))


def depth_mask(depth: np.ndarray, conf: np.ndarray, count: np.ndarray, conf_thr: float, count_thr: int):
    depth = _f32(depth).copy()
    conf = _f64(conf)
    count = np.ascontiguousarray(count, dtype=np.int32)
    lib().orc_depth_mask(_p(depth), _p(conf), _p(count), C.c_int64(depth.size), C.c_double(conf_thr),
                         C.c_int32(count_thr))
    return depth


class Grid:
    """Oracle voxel block grid (Open3D-0.19 VoxelBlockGrid semantics, CPU restatement)."""

    def __init__(self, voxel_size: float, with_color: bool = False):
        self.voxel_size = float(np.float32(voxel_size))
        self.with_color = with_color
        self._h = C.c_void_p(lib().orc_grid_create(C.c_float(voxel_size), int(with_color)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_grid_destroy(self._h)
            self._h = None

    @property
    def num_blocks(self) -> int:
        return int(lib().orc_grid_num_blocks(self._h))

    def touch(self, depth, K, E, depth_max, trunc_mult, depth_scale=1.0):
        depth = _f32(depth)
        H, W = depth.shape
        K = _f64(K)
        E = _f64(E)
        keys = np.empty(((W // 4) * (H // 4) * 4, 3), dtype=np.int32)
        n = C.c_int64(0)
        rc = lib().orc_touch(self._h, _p(depth), W, H, _p(K), _p(E), C.c_float(depth_scale),
                             C.c_float(depth_max), C.c_float(trunc_mult), _p(keys), C.byref(n))
        if rc:
            raise RuntimeError("No block is touched in TSDF volume, abort integration.")
        return keys[: n.value].copy()

    def integrate(self, keys, depth, K, E, depth_max, trunc_mult, color=None, Kc=None, depth_scale=1.0):
        depth = _f32(depth)
        H, W = depth.shape
        keys = np.ascontiguousarray(keys, dtype=np.int32)
        K = _f64(K)
        E = _f64(E)
        upd = C.c_int64(0)
        if color is not None:
            color = np.ascontiguousarray(color, dtype=np.uint8)
            CH, CW = color.shape[:2]
            Kc = _f64(Kc)
        else:
            CH = CW = 0
            Kc = K
        lib().orc_integrate(self._h, _p(keys), C.c_int64(len(keys)), _p(depth), W, H, _p(color), CW, CH,
                            _p(K), _p(Kc), _p(E), C.c_float(depth_scale), C.c_float(depth_max),
                            C.c_float(trunc_mult), C.byref(upd))
        return upd.value

    def export(self):
        n = self.num_blocks
        keys = np.empty((n, 3), np.int32)
        tsdf = np.empty((n, 16, 16, 16), np.float32)
        weight = np.empty((n, 16, 16, 16), np.float32)
        color = np.empty((n, 16, 16, 16, 3), np.float32) if self.with_color else None
        lib().orc_grid_export(self._h, _p(keys), _p(tsdf), _p(weight), _p(color))
        return keys, tsdf, weight, color

    def load(self, keys, tsdf, weight, color=None):
        keys = np.ascontiguousarray(keys, np.int32)
        tsdf = _f32(tsdf)
        weight = _f32(weight)
        color = _f32(color) if color is not None else None
        lib().orc_grid_import(self._h, _p(keys), _p(tsdf), _p(weight), _p(color), C.c_int64(len(keys)))

    def extract_mesh(self, weight_threshold: float = 3.0):
        V = C.c_int64(0)
        T = C.c_int64(0)
        lib().orc_extract_mesh(self._h, C.c_float(weight_threshold), None, None, None, None,
                               C.byref(V), C.byref(T))
        verts = np.empty((V.value, 3), np.float32)
        normals = np.empty((V.value, 3), np.float32)
        tris = np.empty((T.value, 3), np.int32)
        vkeys = np.empty((V.value, 4), np.int32)
        lib().orc_extract_mesh(self._h, C.c_float(weight_threshold), _p(verts), _p(normals), _p(tris),
                               _p(vkeys), C.byref(V), C.byref(T))
        return verts, normals, tris, vkeys

    def extract_points(self, weight_threshold: float = 3.0):
        P = C.c_int64(0)
        lib().orc_extract_points(self._h, C.c_float(weight_threshold), None, None, None, C.byref(P))
        pts = np.empty((P.value, 3), np.float32)
        nrm = np.empty((P.value, 3), np.float32)
        pk = np.empty((P.value, 4), np.int32)
        lib().orc_extract_points(self._h, C.c_float(weight_threshold), _p(pts), _p(nrm), _p(pk), C.byref(P))
        return pts, nrm, pk

    def vertex_colors(self, keys: np.ndarray) -> np.ndarray:
        """Colours (float32 [n,3], 0..1) of mesh vertices / points given their (x,y,z,axis) keys."""
        keys = np.ascontiguousarray(keys, np.int32)
        out = np.empty((len(keys), 3), np.float32)
        rc = lib().orc_vertex_colors(self._h, _p(keys), C.c_int64(len(keys)), _p(out))
        if rc != 0:
            raise RuntimeError("orc_vertex_colors: grid has no colour attribute or a key is not allocated")
        return out


def pixel_error_map(K, Ecw, Ecw_inv, ref_idx, ref_depth, tgt_idx, tgt_depth, depth_max):
    K = _f32(K)
    Ecw = _f32(Ecw)
    Ecw_inv = _f32(Ecw_inv)
    ref_depth = _f32(ref_depth)
    tgt_depth = _f32(tgt_depth)
    H, W = ref_depth.shape
    err = np.empty((H, W), np.float32)
    lib().orc_pixel_error_map(_p(K[ref_idx]), _p(K[tgt_idx]), _p(Ecw[ref_idx]), _p(Ecw[tgt_idx]),
                              _p(Ecw_inv[tgt_idx]), _p(ref_depth), _p(tgt_depth), W, H,
                              C.c_double(depth_max), _p(err))
    return err


def confidence(depths, K, Ecw, Ecw_inv, frame_range, depth_max, err_thr, frame_ok=None):
    depths = _f32(depths)
    N, H, W = depths.shape
    K = _f32(K)
    Ecw = _f32(Ecw)
    Ecw_inv = _f32(Ecw_inv)
    ok = None if frame_ok is None else np.ascontiguousarray(frame_ok, np.uint8)
    conf = np.zeros((N, H, W), np.float64)
    count = np.zeros((N, H, W), np.int32)
    lib().orc_confidence(_p(depths), _p(ok), N, W, H, _p(K), _p(Ecw), _p(Ecw_inv), int(frame_range),
                         C.c_double(depth_max), C.c_double(err_thr), _p(conf), _p(count))
    return conf, count


def create_rays_pinhole(K, E, W, H):
    K = _f64(K)
    E = _f64(E)
    rays = np.empty((H, W, 6), np.float32)
    if lib().orc_create_rays_pinhole(_p(K), _p(E), int(W), int(H), _p(rays)):
        raise RuntimeError("singular intrinsic matrix")
    return rays


def cast_rays(verts, tris, rays):
    verts = _f32(verts)
    tris = np.ascontiguousarray(tris, np.int32)
    rays = _f32(rays)
    out = np.empty(rays.shape[:-1], np.float32)
    lib().orc_cast_rays(_p(verts), C.c_int64(len(verts)), _p(tris), C.c_int64(len(tris)), _p(rays),
                        C.c_int64(out.size), _p(out))
    return out


def odometry_information(source_depth, target_depth, K, source_to_target, dist_threshold, depth_scale, depth_max):
    """o3d.t.pipelines.odometry.compute_odometry_information_matrix (make_fragments.py:142-150): float64 [6,6]."""
    s, t = _f32(source_depth), _f32(target_depth)
    H, W = s.shape
    info = np.zeros((6, 6), np.float64)
    lib().orc_odometry_information(_p(s), _p(t), C.c_int(W), C.c_int(H), _p(_f64(K).reshape(9)),
                                   _p(_f64(source_to_target).reshape(16)), C.c_float(dist_threshold), C.c_float(depth_scale),
                                   C.c_float(depth_max), _p(info))
    return info
