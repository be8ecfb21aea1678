"""Minimal tensor/legacy geometry containers with the attribute surface the reference touches.

The reference uses ``mesh.to_legacy()``, ``mesh.cpu()``, ``mesh.device``, ``mesh.to(device)``,
``pcd.point.positions`` and ``len(legacy.triangles)`` (reconstruct_scene.py:90,111,198;
refine_fragment_poses.py:42; o3d_utils.py:258,307).  These classes carry torch tensors (tensor API)
or numpy arrays (legacy API) and nothing else; PLY writers live in ``io_utils``.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class LegacyTriangleMesh:
    """Stand-in for o3d.geometry.TriangleMesh (float64 vertices, int32 triangles) with the methods the reference's
    filter_mesh_components calls (o3d_utils.py:258-301), implemented on the host in meshops.py."""

    def __init__(self, vertices=None, triangles=None, vertex_normals=None, vertex_colors=None):
        self.vertices = np.zeros((0, 3)) if vertices is None else np.asarray(vertices, dtype=np.float64)
        self.triangles = np.zeros((0, 3), np.int32) if triangles is None else np.asarray(triangles, dtype=np.int32)
        self.vertex_normals = None if vertex_normals is None else np.asarray(vertex_normals, dtype=np.float64)
        self.vertex_colors = None if vertex_colors is None else np.asarray(vertex_colors, dtype=np.float64)

    # -- Open3D legacy TriangleMesh methods ------------------------------------------------------------
    def has_triangles(self) -> bool:
        return len(self.triangles) > 0

    def cluster_connected_triangles(self):
        """(cluster index per triangle, triangles per cluster, area per cluster); triangles sharing an edge are
        connected, clusters are numbered by their lowest triangle index."""
        from . import meshops
        t = self.triangles.astype(np.int64)
        if len(t) == 0:
            return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0)
        label, counts = meshops.cluster_connected_triangles(t)
        area = np.bincount(label, weights=meshops.triangle_areas(self.vertices, t), minlength=len(counts))
        return label, counts, area

    def remove_triangles_by_mask(self, mask):
        mask = np.asarray(mask, bool)
        if len(mask) != len(self.triangles):
            raise RuntimeError("remove_triangles_by_mask: mask length does not match the number of triangles")
        self.triangles = self.triangles[~mask]
        return self

    def _set(self, v, attrs, t):
        self.vertices, self.triangles = np.asarray(v, np.float64), np.asarray(t, np.int32).reshape(-1, 3)
        self.vertex_normals, self.vertex_colors = attrs

    def remove_unreferenced_vertices(self):
        from . import meshops
        used = np.zeros(len(self.vertices), bool)
        used[self.triangles.ravel()] = True
        self._set(*meshops.compact_vertices(self.vertices, [self.vertex_normals, self.vertex_colors],
                                            self.triangles.astype(np.int64), used))
        return self

    def remove_degenerate_triangles(self):
        from . import meshops
        self.triangles = meshops.remove_degenerate_triangles(self.triangles)
        return self

    def remove_duplicated_triangles(self):
        from . import meshops
        self.triangles = meshops.remove_duplicated_triangles(self.triangles)
        return self

    def remove_duplicated_vertices(self):
        from . import meshops
        self._set(*meshops.remove_duplicated_vertices(self.vertices, [self.vertex_normals, self.vertex_colors],
                                                      self.triangles.astype(np.int64)))
        return self

    def remove_non_manifold_edges(self):
        from . import meshops
        self.triangles = meshops.remove_non_manifold_edges(self.vertices, self.triangles.astype(np.int64)).astype(np.int32)
        return self

    @staticmethod
    def create_coordinate_frame(size=1.0, origin=(0, 0, 0)):
        raise RuntimeError("interactive visualisation helpers are outside the hot-path build")


class LegacyPointCloud:
    def __init__(self, points=None, normals=None, colors=None):
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64)
        self.normals = None if normals is None else np.asarray(normals, dtype=np.float64)
        self.colors = None if colors is None else np.asarray(colors, dtype=np.float64)


class TriangleMesh:
    """Stand-in for o3d.t.geometry.TriangleMesh holding torch tensors."""

    def __init__(self, positions: torch.Tensor, indices: torch.Tensor, normals: torch.Tensor = None,
                 colors: torch.Tensor = None):
        self.vertex = SimpleNamespace(positions=positions, normals=normals, colors=colors)
        self.triangle = SimpleNamespace(indices=indices)

    @property
    def device(self):
        return self.vertex.positions.device

    def to(self, device):
        if not isinstance(device, torch.device):
            s = str(device).lower()
            device = torch.device(s if s.startswith("cuda") else "cpu")
        mv = lambda t: None if t is None else t.to(device)
        return TriangleMesh(mv(self.vertex.positions), mv(self.triangle.indices), mv(self.vertex.normals),
                            mv(self.vertex.colors))

    def cpu(self):
        return self.to(torch.device("cpu"))

    def to_legacy(self) -> LegacyTriangleMesh:
        npy = lambda t: None if t is None else t.detach().cpu().numpy()
        return LegacyTriangleMesh(npy(self.vertex.positions), npy(self.triangle.indices), npy(self.vertex.normals),
                                  npy(self.vertex.colors))

    @classmethod
    def from_legacy(cls, mesh: LegacyTriangleMesh, device=torch.device("cpu")):
        t = lambda a, dt: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(device)
        return cls(t(mesh.vertices, torch.float32), t(mesh.triangles, torch.int32),
                   t(mesh.vertex_normals, torch.float32), t(mesh.vertex_colors, torch.float32))


class PointCloud:
    def __init__(self, positions: torch.Tensor, normals: torch.Tensor = None, colors: torch.Tensor = None):
        self.point = SimpleNamespace(positions=positions, normals=normals, colors=colors)

    @property
    def device(self):
        return self.point.positions.device

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return PointCloud(mv(self.point.positions), mv(self.point.normals), mv(self.point.colors))

    def cpu(self):
        return self.to(torch.device("cpu"))

    def to_legacy(self) -> LegacyPointCloud:
        npy = lambda t: None if t is None else t.detach().cpu().numpy()
        return LegacyPointCloud(npy(self.point.positions), npy(self.point.normals), npy(self.point.colors))
