"""Host (NumPy / SciPy) implementations of the Open3D legacy-mesh operations that
``filter_mesh_components`` (processing/reconstruction/utils/o3d_utils.py:241-321) chains: cluster_connected_triangles,
remove_triangles_by_mask, remove_unreferenced_vertices, remove_degenerate_triangles, remove_duplicated_triangles,
remove_duplicated_vertices, remove_non_manifold_edges.  They back the methods of geometry.LegacyTriangleMesh (the
``o3d.geometry.TriangleMesh`` stand-in) and serve as the checker of the device implementation in mq3d_meshfilter.cu;
the product path of ops.filter_mesh_components is the device one.
"""
from __future__ import annotations

import numpy as np


def cluster_connected_triangles(t: np.ndarray):
    """Open3D TriangleMesh::ClusterConnectedTriangles: triangles that share an (undirected) EDGE belong to one
    cluster; clusters are numbered by their lowest triangle index.  Returns (cluster id per triangle, sizes)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    n = len(t)
    e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
    tri = np.tile(np.arange(n), 3)
    order = np.lexsort((e[:, 1], e[:, 0]))
    e, tri = e[order], tri[order]
    same = (e[1:] == e[:-1]).all(1)                       # consecutive entries of one edge: link their triangles
    g = coo_matrix((np.ones(int(same.sum()), np.int8), (tri[:-1][same], tri[1:][same])), shape=(n, n))
    _, label = connected_components(g, directed=False)
    return label, np.bincount(label)


def remove_non_manifold_edges(v: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Open3D TriangleMesh::RemoveNonManifoldEdges: while an edge has more than two triangles, delete its
    smallest-area triangles until two are left (edges visited in sorted order; Open3D's order is that of an
    unordered_map)."""
    while len(t):
        e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
        tri = np.tile(np.arange(len(t)), 3)
        order = np.lexsort((e[:, 1], e[:, 0]))
        e, tri = e[order], tri[order]
        start = np.concatenate([[True], (e[1:] != e[:-1]).any(1)])
        group = np.cumsum(start) - 1
        size = np.bincount(group)
        bad = np.nonzero(size > 2)[0]
        if len(bad) == 0:
            break
        p = v[t].astype(np.float64)
        area = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)
        first = np.nonzero(start)[0]
        for gi in bad:
            members = tri[first[gi]: first[gi] + size[gi]]
            alive = members[area[members] > 0]
            for _ in range(len(alive) - 2):
                alive = members[area[members] > 0]
                area[alive[np.argmin(area[alive])]] = -1.0
        if not (area < 0).any():
            break                                          # only zero-area triangles left on the edge
        t = t[area >= 0]
    return t



def triangle_areas(v: np.ndarray, t: np.ndarray) -> np.ndarray:
    p = np.asarray(v, np.float64)[np.asarray(t, np.int64)]
    return 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)


def compact_vertices(v, attrs, t, used):
    """Drop the vertices not flagged in `used`; returns (vertices, attribute arrays, re-indexed triangles)."""
    remap = np.cumsum(used) - 1
    return v[used], [None if a is None else a[used] for a in attrs], remap[t]


def remove_degenerate_triangles(t: np.ndarray) -> np.ndarray:
    """Triangles that repeat a vertex index (Open3D RemoveDegenerateTriangles)."""
    return t[(t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])]


def remove_duplicated_triangles(t: np.ndarray) -> np.ndarray:
    """Equal up to rotation of the index triple; the first occurrence is kept (Open3D RemoveDuplicatedTriangles)."""
    if len(t) == 0:
        return t
    k = np.argmin(t, axis=1)
    rot = np.stack([np.take_along_axis(t, ((k + i) % 3)[:, None], 1)[:, 0] for i in range(3)], axis=1)
    _, first = np.unique(rot, axis=0, return_index=True)
    return t[np.sort(first)]


def remove_duplicated_vertices(v, attrs, t):
    """Vertices with identical coordinates are merged into their first occurrence (Open3D RemoveDuplicatedVertices);
    this also welds the vertices several ranks re-emit on ghost edges."""
    if len(v) == 0:
        return v, attrs, t
    _, first, inverse = np.unique(v, axis=0, return_index=True, return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    if len(first) == len(v):
        return v, attrs, t
    keep_v = np.zeros(len(v), bool)
    keep_v[first] = True
    new_index = (np.cumsum(keep_v) - 1)[first][inverse]      # old vertex -> index of its first occurrence
    return v[keep_v], [None if a is None else a[keep_v] for a in attrs], new_index[t]
