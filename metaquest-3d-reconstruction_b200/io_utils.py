"""Binary little-endian PLY writers/readers for the artefacts ``reconstruct_scene`` persists
(reconstruction/colorless.ply, colorless_mesh_raw.ply, colorless_mesh_clean.ply; reference:
dataio/reconstruction_data_io.py:57-78).  Layout follows what Open3D's writers emit: double
coordinates for legacy geometries, optional normals/colours, ``vertex_indices`` face lists."""
from __future__ import annotations

import numpy as np

from .geometry import LegacyPointCloud, LegacyTriangleMesh


def _vertex_block(points, normals, colors):
    props = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if normals is not None:
        props += [("nx", "<f8"), ("ny", "<f8"), ("nz", "<f8")]
    if colors is not None:
        props += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    arr = np.empty(len(points), dtype=props)
    arr["x"], arr["y"], arr["z"] = points[:, 0], points[:, 1], points[:, 2]
    if normals is not None:
        arr["nx"], arr["ny"], arr["nz"] = normals[:, 0], normals[:, 1], normals[:, 2]
    if colors is not None:
        c = np.clip(np.asarray(colors) * 255.0, 0, 255).astype(np.uint8)
        arr["red"], arr["green"], arr["blue"] = c[:, 0], c[:, 1], c[:, 2]
    header = "".join(f"property {'double' if t == '<f8' else 'uchar'} {n}\n" for n, t in props)
    return arr, header


def write_point_cloud(path, pcd: LegacyPointCloud):
    arr, props = _vertex_block(pcd.points, pcd.normals, pcd.colors)
    with open(path, "wb") as f:
        f.write(f"ply\nformat binary_little_endian 1.0\ncomment Created by mq3d_b200\nelement vertex {len(arr)}\n"
                f"{props}end_header\n".encode())
        f.write(arr.tobytes())


def write_triangle_mesh(path, mesh: LegacyTriangleMesh):
    arr, props = _vertex_block(mesh.vertices, mesh.vertex_normals, mesh.vertex_colors)
    faces = np.empty(len(mesh.triangles), dtype=[("n", "u1"), ("v", "<i4", (3,))])
    faces["n"] = 3
    faces["v"] = mesh.triangles
    with open(path, "wb") as f:
        f.write(f"ply\nformat binary_little_endian 1.0\ncomment Created by mq3d_b200\nelement vertex {len(arr)}\n"
                f"{props}element face {len(faces)}\nproperty list uchar uint vertex_indices\nend_header\n".encode())
        f.write(arr.tobytes())
        f.write(faces.tobytes())


def read_ply(path):
    """Reader for the files written above: returns (vertex structured array, triangles or None)."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        n_vert = n_face = 0
        props, in_vertex = [], False
        while True:
            line = f.readline().decode().strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[:2] == ["element", "vertex"]:
                n_vert, in_vertex = int(tok[2]), True
            elif tok[:2] == ["element", "face"]:
                n_face, in_vertex = int(tok[2]), False
            elif tok[0] == "property" and in_vertex:
                props.append((tok[2], "<f8" if tok[1] == "double" else "u1"))
        verts = np.frombuffer(f.read(n_vert * np.dtype(props).itemsize), dtype=props)
        tris = None
        if n_face:
            faces = np.frombuffer(f.read(n_face * 13), dtype=[("n", "u1"), ("v", "<i4", (3,))])
            tris = faces["v"].copy()
    return verts, tris
