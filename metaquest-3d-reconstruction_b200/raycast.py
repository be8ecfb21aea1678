"""RaycastingScene: host-side mirror of ``o3d.t.geometry.RaycastingScene`` for the colour-aligned
depth rendering step (reference: processing/reconstruction/reconstruct_scene.py:197-198,
processing/reconstruction/utils/o3d_utils.py:324-342), backed by the LBVH kernels of libmq3d.so."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .vbg import _as_np, _device_index, _stream


class RaycastingScene:
    def __init__(self, nthreads: int = 0, device="CUDA:0"):
        self.device_index = _device_index(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        _lib.check(_lib.lib().mq3d_scene_create(self.device_index, C.byref(h)))
        self._h = h
        self._keep = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().mq3d_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_triangles(self, mesh_or_vertices, triangles=None) -> int:
        """add_triangles(mesh) or add_triangles(vertex_positions, triangle_indices); returns geometry id 0."""
        if triangles is None:
            v, t = mesh_or_vertices.vertex.positions, mesh_or_vertices.triangle.indices
        else:
            v, t = mesh_or_vertices, triangles
        v = v.torch if hasattr(v, "torch") else v
        t = t.torch if hasattr(t, "torch") else t
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(np.ascontiguousarray(v))
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t))
        v = v.to(self.device).to(torch.float32).contiguous()
        t = t.to(self.device).to(torch.int32).contiguous()
        if v.dim() != 2 or v.shape[1] != 3 or t.dim() != 2 or t.shape[1] != 3:
            raise RuntimeError("vertices must be [V,3] and triangles [T,3]")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_scene_add_triangles(self._h, _lib.dptr(v), int(v.shape[0]), _lib.dptr(t),
                                                           int(t.shape[0]), _stream()))
        return 0

    def create_rays_pinhole(self, intrinsic_matrix, extrinsic_matrix, width_px: int, height_px: int) -> torch.Tensor:
        """float32 [H,W,6] rays (origin, unnormalised direction with camera-z component 1)."""
        K = _as_np(intrinsic_matrix, np.float64, (3, 3))
        E = _as_np(extrinsic_matrix, np.float64, (4, 4))
        rays = torch.empty((int(height_px), int(width_px), 6), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_scene_create_rays_pinhole(_lib.darr(K), _lib.darr(E), int(width_px),
                                                                 int(height_px), _lib.dptr(rays), _stream()))
        return rays

    def cast_rays(self, rays, nthreads: int = 0) -> dict:
        rays = rays.torch if hasattr(rays, "torch") else rays
        if isinstance(rays, np.ndarray):
            rays = torch.from_numpy(np.ascontiguousarray(rays))
        rays = rays.to(self.device).to(torch.float32).contiguous()
        if rays.shape[-1] != 6:
            raise RuntimeError("rays must have a trailing dimension of 6")
        t_hit = torch.empty(rays.shape[:-1], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_scene_cast_rays(self._h, _lib.dptr(rays), int(t_hit.numel()), _lib.dptr(t_hit),
                                                       _stream()))
        return {"t_hit": t_hit}
