"""``sys.modules["open3d"] = mq3d_b200.compat`` -- the part of the Open3D Python surface the reference's
reconstruction modules touch, mapped onto the sm_100a library.

With this module standing in for ``open3d`` the reference's own modules import and run unmodified:
``processing.reconstruction.utils.o3d_utils`` (load_depth_map / integrate / filter_mesh_components /
raycast_in_color_view, :109-342), ``config.reconstruction_config`` (device strings, :6-29),
``dataio.reconstruction_data_io`` (VoxelBlockGrid.save / load, PLY writers, :42-78) and
``processing.reconstruction.reconstruct_scene`` (the stage driver).  See ``install()`` and INTEGRATION.md.

What is real: o3d.core.Tensor / Device / Dtype, o3d.t.geometry.Image / VoxelBlockGrid / RaycastingScene /
TriangleMesh / PointCloud, o3d.geometry.TriangleMesh / PointCloud (legacy containers with the clean-up methods
filter_mesh_components calls), o3d.utility.Vector3dVector / Vector3iVector, o3d.io / o3d.t.io PLY I/O, and
plain containers for o3d.pipelines.registration.PoseGraph* and o3d.camera.PinholeCamera* (the reference's
pose converters only store values in them).  Nothing here computes: tensors are torch tensors, the work happens
behind include/mq3d.h.

What is not: everything else resolves to a placeholder, so that annotations and attribute chains such as
``o3d.t.pipelines.registration.ICPConvergenceCriteria`` import fine, and raises ``RuntimeError`` the moment it is
called -- ICP / pose-graph optimisation, colour-map optimisation and the interactive viewers are outside this
hot-path build and are never silently emulated.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

from . import geometry as _geom
from . import io_utils as _io
from . import raycast as _raycast
from . import vbg as _vbg

__version__ = "0.19.0+mq3d_b200"


class _OutOfScope:
    """Placeholder for an Open3D name outside the hot-path build: usable in annotations and attribute chains,
    raises when called."""

    def __init__(self, path: str):
        object.__setattr__(self, "_path", path)

    def __getattr__(self, name: str):
        if name.startswith("__"):
            raise AttributeError(name)
        return _OutOfScope(f"{self._path}.{name}")

    def __call__(self, *a, **k):
        raise RuntimeError(f"open3d.{self._path} is outside the B200 hot-path build (TSDF integration, marching cubes, "
                           f"confidence, raycast); it is not emulated")

    def __repr__(self):
        return f"<open3d.{self._path}: not part of mq3d_b200.compat>"


class _Namespace(types.ModuleType):
    """A sub-namespace of the stand-in (o3d.core, o3d.t.geometry, ...): unknown names become placeholders."""

    def __init__(self, path: str, **members):
        super().__init__(f"open3d.{path}")
        self.__dict__["_path"] = path
        self.__dict__.update(members)

    def __getattr__(self, name: str):
        if name.startswith("__"):
            raise AttributeError(name)
        return _OutOfScope(f"{self._path}.{name}")


# ---------------------------------------------------------------------------------------------------------
# o3d.core
# ---------------------------------------------------------------------------------------------------------
class Device:
    """o3d.core.Device("CUDA:0").  CPU devices are representable (host-side tensors such as K and E live
    there) but no compute object accepts them."""

    def __init__(self, spec="CPU:0"):
        if isinstance(spec, Device):
            spec = str(spec)
        s = str(spec).upper()
        kind, _, idx = s.partition(":")
        if kind not in ("CPU", "CUDA"):
            raise RuntimeError(f"Invalid device string {spec}")      # Open3D raises RuntimeError here too
        self.kind, self.index = kind, int(idx or 0)

    def __str__(self):
        return f"{self.kind}:{self.index}"

    __repr__ = __str__

    def __eq__(self, other):
        return str(self) == str(other).upper()

    def __hash__(self):
        return hash(str(self))

    def get_type(self):
        return self.kind

    def get_id(self):
        return self.index

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
            t = t.to(Device(device).torch if isinstance(device, (Device, str)) else device)
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

    def to(self, x, *a, **k):
        if isinstance(x, (Device, str)):
            return Tensor(self.torch.to(Device(x).torch))
        return Tensor(self.torch.to(x))

    def clone(self):
        return Tensor(self.torch.clone())

    def numpy(self):
        return self.torch.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.torch.shape[0]

    def __getitem__(self, k):
        return Tensor(self.torch[k])

    def item(self):
        return self.torch.item()

    def __float__(self):
        return float(self.torch)

    def __truediv__(self, o):
        return Tensor(self.torch / _un(o))

    def __mul__(self, o):
        return Tensor(self.torch * _un(o))

    def __add__(self, o):
        return Tensor(self.torch + _un(o))

    def __sub__(self, o):
        return Tensor(self.torch - _un(o))

    def __matmul__(self, o):
        return Tensor(self.torch @ _un(o))

    def __gt__(self, o):
        return Tensor(self.torch > _un(o))

    def __lt__(self, o):
        return Tensor(self.torch < _un(o))

    def __ge__(self, o):
        return Tensor(self.torch >= _un(o))

    def __le__(self, o):
        return Tensor(self.torch <= _un(o))

    def __bool__(self):
        return bool(self.torch)

    def __getattr__(self, name):
        # anything else (is_cuda, detach, ...) is answered by the torch tensor underneath
        if name == "torch" or name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.__dict__["torch"], name)


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

    @property
    def device(self):
        return self._t.device


def _dev(device):
    if isinstance(device, str):
        device = Device(device)
    if isinstance(device, Device):
        if device.kind != "CUDA":
            raise RuntimeError(f"device {device} is not a CUDA device; this build has no CPU fallback")
        return device.torch
    return device


def _un(x):
    return x.torch if isinstance(x, Tensor) else x


# ---------------------------------------------------------------------------------------------------------
# o3d.t.geometry
# ---------------------------------------------------------------------------------------------------------
class _TensorMap(types.SimpleNamespace):
    """mesh.vertex / mesh.triangle / pcd.point: attribute and item access to Tensor-wrapped arrays."""

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def __contains__(self, k):
        return getattr(self, k, None) is not None


def _wrap(t):
    return None if t is None else (t if isinstance(t, Tensor) else Tensor(t))


class TriangleMesh(_geom.TriangleMesh):
    """o3d.t.geometry.TriangleMesh: positions / normals / colours / indices are compat Tensors."""

    def __init__(self, positions=None, indices=None, normals=None, colors=None):
        if positions is None:
            positions = torch.zeros((0, 3), dtype=torch.float32)
        if indices is None:
            indices = torch.zeros((0, 3), dtype=torch.int32)
        super().__init__(_un(positions), _un(indices), _un(normals), _un(colors))

    # the base class keeps raw torch tensors (the library wrappers read them); the Open3D-facing views wrap them
    @property
    def vertex(self):
        v = self._vertex
        return _TensorMap(positions=_wrap(v.positions), normals=_wrap(v.normals), colors=_wrap(v.colors))

    @vertex.setter
    def vertex(self, ns):
        self._vertex = types.SimpleNamespace(positions=_un(ns.positions), normals=_un(getattr(ns, "normals", None)),
                                             colors=_un(getattr(ns, "colors", None)))

    @property
    def triangle(self):
        return _TensorMap(indices=_wrap(self._triangle.indices))

    @triangle.setter
    def triangle(self, ns):
        self._triangle = types.SimpleNamespace(indices=_un(ns.indices))

    @property
    def device(self):
        d = self._vertex.positions.device
        return Device(f"CUDA:{d.index or 0}" if d.type == "cuda" else "CPU:0")

    def _raw(self) -> _geom.TriangleMesh:
        return _geom.TriangleMesh(self._vertex.positions, self._triangle.indices, self._vertex.normals, self._vertex.colors)

    def to(self, device, *a, **k):
        dev = Device(device).torch if isinstance(device, (Device, str)) else device
        m = self._raw().to(dev)
        return TriangleMesh(m.vertex.positions, m.triangle.indices, m.vertex.normals, m.vertex.colors)

    def cpu(self):
        return self.to(Device("CPU:0"))

    def cuda(self, index=0):
        return self.to(Device(f"CUDA:{index}"))

    def to_legacy(self):
        return self._raw().to_legacy()

    @classmethod
    def from_legacy(cls, mesh, *a, device=None, **k):
        m = _geom.TriangleMesh.from_legacy(mesh)
        out = cls(m.vertex.positions, m.triangle.indices, m.vertex.normals, m.vertex.colors)
        return out if device is None else out.to(device)


class PointCloud(_geom.PointCloud):
    """o3d.t.geometry.PointCloud."""

    def __init__(self, positions=None, normals=None, colors=None):
        if positions is None:
            positions = torch.zeros((0, 3), dtype=torch.float32)
        super().__init__(_un(positions), _un(normals), _un(colors))

    @property
    def point(self):
        p = self._point
        return _TensorMap(positions=_wrap(p.positions), normals=_wrap(p.normals), colors=_wrap(p.colors))

    @point.setter
    def point(self, ns):
        self._point = types.SimpleNamespace(positions=_un(ns.positions), normals=_un(getattr(ns, "normals", None)),
                                            colors=_un(getattr(ns, "colors", None)))

    @property
    def device(self):
        d = self._point.positions.device
        return Device(f"CUDA:{d.index or 0}" if d.type == "cuda" else "CPU:0")

    def _raw(self) -> _geom.PointCloud:
        return _geom.PointCloud(self._point.positions, self._point.normals, self._point.colors)

    def to(self, device=None, *a, **k):
        dev = Device(device).torch if isinstance(device, (Device, str)) else device
        m = self._raw().to(dev)
        return PointCloud(m.point.positions, m.point.normals, m.point.colors)

    def cpu(self):
        return self.to(Device("CPU:0"))

    def to_legacy(self):
        return self._raw().to_legacy()

    @classmethod
    def from_legacy(cls, pcd, *a, device=None, **k):
        t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        out = cls(t(pcd.points), t(pcd.normals), t(pcd.colors))
        return out if device is None else out.to(device)


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
        out = self.extract_triangle_mesh_arrays(weight_threshold, with_colors=self.has_color)
        return TriangleMesh(out[0], out[2], out[1], out[3] if self.has_color else None)

    def extract_point_cloud(self, weight_threshold=3.0, estimated_point_number=-1):
        out = self.extract_point_cloud_arrays(weight_threshold, with_colors=self.has_color)
        return PointCloud(out[0], out[1], out[2] if self.has_color else None)

    def hashmap(self):
        return _OutOfScope("t.geometry.VoxelBlockGrid.hashmap")()

    @classmethod
    def load(cls, path, device=Device("CUDA:0")):
        g = _vbg.VoxelBlockGrid.load(path, device=_dev(device))
        g.__class__ = cls
        return g


class RaycastingScene(_raycast.RaycastingScene):
    def __init__(self, nthreads=0, device=Device("CUDA:0")):
        super().__init__(nthreads, _dev(device))

    def add_triangles(self, mesh_or_vertices, triangles=None):
        if isinstance(mesh_or_vertices, _geom.TriangleMesh):
            return super().add_triangles(mesh_or_vertices._vertex.positions, mesh_or_vertices._triangle.indices)
        return super().add_triangles(_un(mesh_or_vertices), _un(triangles))

    def create_rays_pinhole(self, intrinsic_matrix, extrinsic_matrix, width_px, height_px):
        return Tensor(super().create_rays_pinhole(_np(intrinsic_matrix), _np(extrinsic_matrix), int(width_px),
                                                  int(height_px)))

    def cast_rays(self, rays, nthreads=0):
        return {k: Tensor(v) for k, v in super().cast_rays(_un(rays), nthreads).items()}


def _np(x):
    return x.numpy() if isinstance(x, Tensor) else np.asarray(x)


def _odometry_information(source_depth, target_depth, intrinsic, source_to_target, dist_threshold=0.07, depth_scale=1000.0,
                          depth_max=3.0):
    """o3d.t.pipelines.odometry.compute_odometry_information_matrix -> 6x6 Float64 Tensor on the host."""
    from . import ops
    return Tensor(torch.from_numpy(ops.compute_odometry_information_matrix(source_depth, target_depth, _np(intrinsic),
                                                                           _np(source_to_target), dist_threshold,
                                                                           depth_scale, depth_max)))


# ---------------------------------------------------------------------------------------------------------
# o3d.utility, o3d.pipelines.registration, o3d.camera: value containers
# ---------------------------------------------------------------------------------------------------------
def Vector3dVector(a=()):
    """o3d.utility.Vector3dVector: float64 [n,3] (the legacy containers hold plain arrays)."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, 3))


def Vector3iVector(a=()):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1, 3))


class PoseGraphNode:
    """o3d.pipelines.registration.PoseGraphNode(pose): a 4x4 float64 pose (o3d_utils.py:22-35,99-106)."""

    def __init__(self, pose=None):
        self.pose = np.eye(4) if pose is None else np.asarray(pose, dtype=np.float64)


class PoseGraphEdge:
    def __init__(self, source_node_id=-1, target_node_id=-1, transformation=None, information=None, uncertain=False,
                 confidence=1.0):
        self.source_node_id, self.target_node_id = int(source_node_id), int(target_node_id)
        self.transformation = np.eye(4) if transformation is None else np.asarray(transformation, dtype=np.float64)
        self.information = np.eye(6) if information is None else np.asarray(information, dtype=np.float64)
        self.uncertain, self.confidence = bool(uncertain), float(confidence)


class PoseGraph:
    def __init__(self):
        self.nodes, self.edges = [], []


class PinholeCameraIntrinsic:
    def __init__(self, width=-1, height=-1, fx=0.0, fy=0.0, cx=0.0, cy=0.0):
        self.set_intrinsics(width, height, fx, fy, cx, cy)

    def set_intrinsics(self, width, height, fx, fy, cx, cy):
        self.width, self.height = int(width), int(height)
        self.intrinsic_matrix = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float64)

    def get_focal_length(self):
        return self.intrinsic_matrix[0, 0], self.intrinsic_matrix[1, 1]

    def get_principal_point(self):
        return self.intrinsic_matrix[0, 2], self.intrinsic_matrix[1, 2]


class PinholeCameraParameters:
    def __init__(self):
        self.intrinsic = PinholeCameraIntrinsic()
        self.extrinsic = np.eye(4)


class PinholeCameraTrajectory:
    def __init__(self):
        self.parameters = []


# ---------------------------------------------------------------------------------------------------------
# o3d.io / o3d.t.io
# ---------------------------------------------------------------------------------------------------------
def _legacy_mesh(mesh):
    return mesh.to_legacy() if hasattr(mesh, "to_legacy") else mesh


def _write_pcd(filename, pointcloud, write_ascii=False, compressed=False, print_progress=False):
    _io.write_point_cloud(str(filename), _legacy_mesh(pointcloud))
    return True


def _write_mesh(filename, mesh, write_ascii=False, compressed=False, **kw):
    _io.write_triangle_mesh(str(filename), _legacy_mesh(mesh))
    return True


def _read_mesh_legacy(filename, *a, **k):
    v, tris = _io.read_ply(str(filename))
    names = v.dtype.names
    xyz = np.stack([v["x"], v["y"], v["z"]], axis=1)
    nrm = np.stack([v["nx"], v["ny"], v["nz"]], axis=1) if "nx" in names else None
    col = np.stack([v["red"], v["green"], v["blue"]], axis=1) / 255.0 if "red" in names else None
    return _geom.LegacyTriangleMesh(xyz, tris if tris is not None else np.zeros((0, 3), np.int32), nrm, col)


def _read_pcd_legacy(filename, *a, **k):
    m = _read_mesh_legacy(filename)
    return _geom.LegacyPointCloud(m.vertices, m.vertex_normals, m.vertex_colors)


core = _Namespace("core", Device=Device, Tensor=Tensor, Dtype=Dtype, float32=torch.float32, float64=torch.float64,
                  int32=torch.int32, int64=torch.int64, uint8=torch.uint8, uint16=torch.int32, bool=torch.bool)
t = _Namespace("t",
               geometry=_Namespace("t.geometry", Image=Image, VoxelBlockGrid=VoxelBlockGrid, RaycastingScene=RaycastingScene,
                                   TriangleMesh=TriangleMesh, PointCloud=PointCloud),
               io=_Namespace("t.io", write_point_cloud=_write_pcd, write_triangle_mesh=_write_mesh,
                             read_triangle_mesh=lambda f, *a, **k: TriangleMesh.from_legacy(_read_mesh_legacy(f)),
                             read_point_cloud=lambda filename, *a, **k: PointCloud.from_legacy(_read_pcd_legacy(filename))),
               pipelines=_Namespace("t.pipelines",
                                    odometry=_Namespace("t.pipelines.odometry",
                                                        compute_odometry_information_matrix=_odometry_information)))
geometry = _Namespace("geometry", TriangleMesh=_geom.LegacyTriangleMesh, PointCloud=_geom.LegacyPointCloud)
utility = _Namespace("utility", Vector3dVector=Vector3dVector, Vector3iVector=Vector3iVector)
pipelines = _Namespace("pipelines",
                       registration=_Namespace("pipelines.registration", PoseGraph=PoseGraph, PoseGraphNode=PoseGraphNode,
                                               PoseGraphEdge=PoseGraphEdge))
camera = _Namespace("camera", PinholeCameraIntrinsic=PinholeCameraIntrinsic, PinholeCameraParameters=PinholeCameraParameters,
                    PinholeCameraTrajectory=PinholeCameraTrajectory)
io = _Namespace("io", write_point_cloud=_write_pcd, write_triangle_mesh=_write_mesh, read_triangle_mesh=_read_mesh_legacy,
                read_point_cloud=_read_pcd_legacy)
visualization = _Namespace("visualization")


def __getattr__(name: str):
    """Any other top-level Open3D name (o3d.ml, o3d.data, ...) is a placeholder."""
    if name.startswith("__"):
        raise AttributeError(name)
    return _OutOfScope(name)


def install(force: bool = False):
    """Make ``import open3d`` resolve to this module (only if the real package is absent, unless force=True).
    Call it before importing the reference's modules; returns the module now registered as ``open3d``."""
    if not force:
        try:
            import importlib.util
            if "open3d" not in sys.modules and importlib.util.find_spec("open3d") is not None:
                import open3d
                return open3d
        except Exception:
            pass
    me = sys.modules[__name__]
    sys.modules["open3d"] = me
    for sub in ("core", "t", "geometry", "utility", "pipelines", "camera", "io", "visualization"):
        sys.modules[f"open3d.{sub}"] = getattr(me, sub)
    sys.modules["open3d.t.geometry"] = t.geometry
    sys.modules["open3d.t.io"] = t.io
    return me
