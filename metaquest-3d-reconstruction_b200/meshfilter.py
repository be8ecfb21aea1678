"""Device implementation of ``filter_mesh_components`` (reference
processing/reconstruction/utils/o3d_utils.py:241-321): one call into mq3d_mesh_filter, the mesh never leaves HBM.
ops.filter_mesh_components dispatches here for CUDA meshes and prints the reference's messages."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .geometry import TriangleMesh
from .vbg import _stream


def filter_mesh_components_device(mesh: TriangleMesh, min_triangle_count: int):
    """Returns (filtered TriangleMesh on the same device, (components, kept components, triangles removed with the
    small components, input triangles, final triangles, largest component, fallback flag))."""
    v = mesh.vertex.positions
    v = v.torch if hasattr(v, "torch") else v
    t = mesh.triangle.indices
    t = t.torch if hasattr(t, "torch") else t
    if not v.is_cuda:
        raise RuntimeError("filter_mesh_components_device needs a CUDA mesh (use backend='host' for host meshes)")
    dev = v.device
    un = lambda x: None if x is None else (x.torch if hasattr(x, "torch") else x)
    v = v.to(torch.float32).contiguous()
    t = t.to(dev).to(torch.int32).contiguous()
    nrm, col = un(mesh.vertex.normals), un(mesh.vertex.colors)
    nrm = None if nrm is None else nrm.to(dev).to(torch.float32).contiguous()
    col = None if col is None else col.to(dev).to(torch.float32).contiguous()
    V, T = int(v.shape[0]), int(t.shape[0])
    ov, ot = torch.empty_like(v), torch.empty_like(t)
    on = None if nrm is None else torch.empty_like(nrm)
    oc = None if col is None else torch.empty_like(col)
    nV, nT, info = C.c_int64(), C.c_int64(), _lib.MeshFilterInfo()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mq3d_mesh_filter(_lib.dptr(v), _lib.dptr(nrm), _lib.dptr(col), V, _lib.dptr(t), T,
                                               int(min_triangle_count), _lib.dptr(ov), _lib.dptr(on), _lib.dptr(oc),
                                               _lib.dptr(ot), C.byref(nV), C.byref(nT), C.byref(info), int(dev.index or 0),
                                               _stream()))
    out = TriangleMesh(ov[: nV.value], ot[: nT.value], None if on is None else on[: nV.value],
                       None if oc is None else oc[: nV.value])
    if info.non_manifold_edges:
        # edges carried by more than two triangles never come out of marching cubes; Open3D removes the smallest-area
        # triangles of such edges in unordered_map order -- finished with the host routine on this (rare) path
        from . import meshops
        tt = meshops.remove_non_manifold_edges(out.vertex.positions.cpu().numpy().astype(np.float64),
                                               out.triangle.indices.cpu().numpy().astype(np.int64))
        out = TriangleMesh(out.vertex.positions, torch.from_numpy(np.ascontiguousarray(tt, np.int32)).to(dev),
                           out.vertex.normals, out.vertex.colors)
    final = int(out.triangle.indices.shape[0])
    return out, (int(info.components), int(info.components_kept), int(info.removed_triangles), int(info.input_triangles),
                 final, int(info.largest_component), bool(info.fallback_largest))
