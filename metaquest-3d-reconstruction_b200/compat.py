"""``import mq3d_b200.compat as o3d`` -- the subset of the Open3D Python surface that the reference's
hot path touches, mapped onto the sm_100a library.

With this module bound to the name ``o3d`` the reference's own o3d_utils.integrate / load_depth_map /
raycast_in_color_view bodies (processing/reconstruction/utils/o3d_utils.py:109-238,324-342) run
unmodified: o3d.core.Tensor / Device / Dtype, o3d.t.geometry.Image / VoxelBlockGrid / RaycastingScene /
TriangleMesh / PointCloud, o3d.geometry.* (legacy containers) and o3d.io.write_*.  Nothing here
computes: tensors are torch tensors, the work happens behind include/mq3d.h.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import geometry as _geom
from . import io_utils as _io
from . import raycast as _raycast
from . import vbg as _vbg


class Device:
    """o3d.core.Device("CUDA:0").  CPU devices are representable (host-side tensors such as K and E live
    there) but no compute object accepts them."""

    def __init__(self, spec: str = "CPU:0"):
        s = str(spec).upper()
        kind, _, idx = s.partition(":")
        if kind not in ("CPU", "CUDA"):
            raise ValueError(f"Invalid device string {spec}")
        self.kind, self.index = kind, int(idx or 0)

    def __str__(self):
        return f"{self.kind}:{self.index}"

    __repr__ = __str__

    def __eq__(self, other):
        return str(self) == str(other)

    def __hash__(self):
        return hash(str(self))

    @property
    def torch(self) -> torch.device:
        return torch.device("cuda", self.index) if self.kind == "CUDA" else torch.device("cpu")


class Dtype:
    Float32, Float64, Int32, Int64, UInt8, UInt16, Bool = (torch.float32, torch.float64, torch.int32,
                                                           torch.int64, torch.uint8, torch.int32, torch.bool)


class Tensor:
    """o3d.core.Tensor(np_array, dtype=..., device=...) backed by a torch tensor."""

    def __init__(self, data, dtype=None, device=None):
        if isinstance(data, Tensor):
            t = data.torch
        elif isinstance(data, torch.Tensor):
            t = data
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(data)))
        if dtype is not None:
            t = t.to(dtype)
        if device is not None:
            t = t.to(device.torch if isinstance(device, Device) else device)
        self.torch = t

    @property
    def shape(self):
        return tuple(self.torch.shape)

    @property
    def dtype(self):
        return self.torch.dtype

    @property
    def device(self):
        d = self.torch.device
        return Device(f"CUDA:{d.index or 0}" if d.type == "cuda" else "CPU:0")

    def cpu(self):
        return Tensor(self.torch.cpu())

    def cuda(self, index=0):
        return Tensor(self.torch.cuda(index))

    def to(self, x):
        if isinstance(x, Device):
            return Tensor(self.torch.to(x.torch))
        return Tensor(self.torch.to(x))

    def numpy(self):
        return self.torch.detach().cpu().numpy()

    def __len__(self):
        return self.torch.shape[0]

    def __getitem__(self, k):
        return Tensor(self.torch[k])


class Image:
    """o3d.t.geometry.Image(tensor=Tensor)."""

    def __init__(self, tensor=None):
        self._t = tensor if isinstance(tensor, Tensor) else Tensor(tensor)

    def as_tensor(self):
        return self._t

    @property
    def rows(self):
        return self._t.shape[0]

    @property
    def columns(self):
        return self._t.shape[1]


def _dev(device):
    if isinstance(device, Device):
        if device.kind != "CUDA":
            raise RuntimeError(f"device {device} is not a CUDA device; this build has no CPU fallback")
        return device.torch
    return device


def _wrap_mesh(m: _geom.TriangleMesh):
    m.vertex.positions = Tensor(m.vertex.positions)
    if m.vertex.normals is not None:
        m.vertex.normals = Tensor(m.vertex.normals)
    m.triangle.indices = Tensor(m.triangle.indices)
    return m


class TriangleMesh(_geom.TriangleMesh):
    def __init__(self, positions=None, indices=None, normals=None, colors=None):
        un = lambda x: x.torch if isinstance(x, Tensor) else x
        super().__init__(un(positions), un(indices), un(normals), un(colors))

    @classmethod
    def from_legacy(cls, mesh, device=None):
        m = _geom.TriangleMesh.from_legacy(mesh)
        return cls(m.vertex.positions, m.triangle.indices, m.vertex.normals, m.vertex.colors)


class PointCloud(_geom.PointCloud):
    pass


class VoxelBlockGrid(_vbg.VoxelBlockGrid):
    """o3d.t.geometry.VoxelBlockGrid with Open3D's argument conventions (Device objects, Tensor I/O)."""

    def __init__(self, attr_names=("tsdf", "weight"), attr_dtypes=None, attr_channels=None, voxel_size=0.0058,
                 block_resolution=16, block_count=10000, device=Device("CUDA:0")):
        super().__init__(attr_names, attr_dtypes, attr_channels, voxel_size, block_resolution, block_count,
                         _dev(device))

    def compute_unique_block_coordinates(self, depth, intrinsic, extrinsic, depth_scale=1000.0, depth_max=3.0,
                                         trunc_voxel_multiplier=8.0):
        return Tensor(super().compute_unique_block_coordinates(depth, intrinsic, extrinsic, depth_scale, depth_max,
                                                               trunc_voxel_multiplier))

    def extract_triangle_mesh(self, weight_threshold=3.0, estimated_vertex_number=-1):
        v, n, t = self.extract_triangle_mesh_arrays(weight_threshold)
        return TriangleMesh(v, t, n)

    def extract_point_cloud(self, weight_threshold=3.0, estimated_point_number=-1):
        p, n = self.extract_point_cloud_arrays(weight_threshold)
        return PointCloud(p, n)

    @classmethod
    def load(cls, path, device=Device("CUDA:0")):
        g = _vbg.VoxelBlockGrid.load(path, device=_dev(device))
        g.__class__ = cls
        return g


class RaycastingScene(_raycast.RaycastingScene):
    def __init__(self, nthreads=0, device=Device("CUDA:0")):
        super().__init__(nthreads, _dev(device))

    def create_rays_pinhole(self, intrinsic_matrix, extrinsic_matrix, width_px, height_px):
        return Tensor(super().create_rays_pinhole(intrinsic_matrix, extrinsic_matrix, width_px, height_px))

    def cast_rays(self, rays, nthreads=0):
        return {k: Tensor(v) for k, v in super().cast_rays(rays, nthreads).items()}


core = SimpleNamespace(Device=Device, Tensor=Tensor, Dtype=Dtype, float32=torch.float32, float64=torch.float64,
                       int32=torch.int32, int64=torch.int64, uint8=torch.uint8, uint16=torch.int32)
t = SimpleNamespace(geometry=SimpleNamespace(Image=Image, VoxelBlockGrid=VoxelBlockGrid, RaycastingScene=RaycastingScene,
                                             TriangleMesh=TriangleMesh, PointCloud=PointCloud))
geometry = SimpleNamespace(TriangleMesh=_geom.LegacyTriangleMesh, PointCloud=_geom.LegacyPointCloud)


def _write_pcd(filename, pointcloud, write_ascii=False, compressed=False, print_progress=False):
    _io.write_point_cloud(filename, pointcloud)
    return True


def _write_mesh(filename, mesh, write_ascii=False, compressed=False, **kw):
    _io.write_triangle_mesh(filename, mesh)
    return True


io = SimpleNamespace(write_point_cloud=_write_pcd, write_triangle_mesh=_write_mesh)
__version__ = "0.19.0+mq3d_b200"
