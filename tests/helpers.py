"""Shared helpers for the parity tests: pipeline camera matrices and canonical forms."""
import numpy as np

import mq3d_b200  # noqa: F401
from mq3d_b200 import synth
from mq3d_b200.models import CoordinateSystem, Side


def pipeline_cameras(dataset):
    """(K float32 [N,3,3] with cx mirrored, E_wc float32 [N,4,4], E_cw float32 [N,4,4]) exactly as the
    reference derives them (o3d_utils.py:14-19,185; reconstruct_scene.py:48-51)."""
    tr = dataset.transforms.convert_coordinate_system(CoordinateSystem.OPEN3D, is_camera=True)
    K = dataset.get_intrinsic_matrices()
    K[:, 0, 2] = dataset.widths - K[:, 0, 2]
    return K, tr.extrinsics_wc, tr.extrinsics_cw


_CAPS = {}


def capture(n_frames, side=Side.LEFT, **kw):
    key = (n_frames, side, tuple(sorted(kw.items())))
    if key not in _CAPS:
        _CAPS[key] = synth.make_depth_capture(n_frames, side, **kw)
    return _CAPS[key]


def sort_blocks(keys, *arrays):
    """Sort block arrays by key (lexicographic x,y,z)."""
    keys = np.asarray(keys)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return (keys[order],) + tuple(None if a is None else np.asarray(a)[order] for a in arrays)


def canonical_mesh(verts, tris, vkeys, normals=None):
    """Order-independent form: vertices sorted by their lattice-edge key; triangles re-indexed,
    rotated so the smallest index leads (winding preserved) and row-sorted."""
    vkeys = np.asarray(vkeys)
    order = np.lexsort((vkeys[:, 3], vkeys[:, 2], vkeys[:, 1], vkeys[:, 0]))
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    t = rank[np.asarray(tris, dtype=np.int64)]
    if len(t):
        k = np.argmin(t, axis=1)
        t = np.stack([np.take_along_axis(t, ((k + i) % 3)[:, None], 1)[:, 0] for i in range(3)], axis=1)
        t = t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]
    out = {"keys": vkeys[order], "verts": np.asarray(verts)[order], "tris": t}
    if normals is not None:
        out["normals"] = np.asarray(normals)[order]
    return out


def canonical_points(pts, pkeys, normals=None):
    pkeys = np.asarray(pkeys)
    order = np.lexsort((pkeys[:, 3], pkeys[:, 2], pkeys[:, 1], pkeys[:, 0]))
    out = {"keys": pkeys[order], "pts": np.asarray(pts)[order]}
    if normals is not None:
        out["normals"] = np.asarray(normals)[order]
    return out


def oracle_integrate_sequence(orc, grid, depths, K, Ewc, depth_max, trunc, colors=None, Kc=None, valid=None):
    """Frame-sequential reference loop (o3d_utils.py:231-236) on the oracle grid."""
    visits = updated = 0
    for i in range(len(depths)):
        if valid is not None and not valid[i]:
            continue
        keys = grid.touch(depths[i], K[i], Ewc[i], depth_max, trunc)
        updated += grid.integrate(keys, depths[i], K[i], Ewc[i], depth_max, trunc,
                                  color=None if colors is None else colors[i],
                                  Kc=None if Kc is None else Kc[i])
        visits += len(keys)
    return visits, updated
