"""GPU parity: K5 marching cubes / point cloud against the CPU oracle (canonical forms, bit-exact)."""
import numpy as np
import pytest
import torch

from helpers import (canonical_mesh, canonical_points, capture, oracle_integrate_sequence, pipeline_cameras)

pytestmark = pytest.mark.gpu

DEPTH_MAX = 4.0
TRUNC = 10.0


def _grids(oracle, cuda_device, n_frames=16, voxel=0.02):
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(n_frames)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin = np.stack([oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(len(ds))])
    og = oracle.Grid(voxel)
    oracle_integrate_sequence(oracle, og, lin, K, Ewc, DEPTH_MAX, TRUNC)
    vbg = VoxelBlockGrid(voxel_size=voxel, block_count=4000, device=cuda_device)
    vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC)
    return og, vbg


@pytest.mark.parametrize("thr", [1.5, 3.0, 0.0])
def test_mesh_matches_oracle(cuda_device, oracle, thr):
    og, vbg = _grids(oracle, cuda_device)
    ov, on, ot, ok = og.extract_mesh(thr)
    v, n, t, k = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(thr, with_keys=True)]
    assert len(v) == len(ov) and len(t) == len(ot) and len(ov) > 1000
    a, b = canonical_mesh(ov, ot, ok, on), canonical_mesh(v, t, k, n)
    assert np.array_equal(a["keys"], b["keys"])                       # same lattice edges carry vertices
    assert np.array_equal(a["tris"], b["tris"])                       # identical topology + winding
    assert np.array_equal(a["verts"].view(np.uint32), b["verts"].view(np.uint32))
    assert np.abs(a["verts"] - b["verts"]).max() <= 1e-4              # north-star bar (m)
    assert np.allclose(a["normals"], b["normals"], atol=1e-6)


def test_points_match_oracle(cuda_device, oracle):
    og, vbg = _grids(oracle, cuda_device)
    op, on, ok = og.extract_points(3.0)
    p, n, k = [x.cpu().numpy() for x in vbg.extract_point_cloud_arrays(3.0, with_keys=True)]
    assert len(p) == len(op) > 1000
    a, b = canonical_points(op, ok, on), canonical_points(p, k, n)
    assert np.array_equal(a["keys"], b["keys"])
    assert np.array_equal(a["pts"].view(np.uint32), b["pts"].view(np.uint32))
    assert np.allclose(a["normals"], b["normals"], atol=1e-6)


def test_mesh_fill_requires_count(cuda_device, oracle):
    from mq3d_b200 import _lib
    from mq3d_b200.vbg import VoxelBlockGrid
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=100, device=cuda_device)
    rc = _lib.lib().mq3d_extract_mesh_fill(vbg._h, None, None, None, None, None)
    assert rc == _lib.MQ3D_ERR_STATE
    v, n, t = vbg.extract_triangle_mesh_arrays(1.5)    # empty grid -> empty mesh
    assert v.shape == (0, 3) and t.shape == (0, 3)


def test_single_call_extraction_equals_count_fill(cuda_device, oracle):
    """mq3d_extract_mesh (classification, scan and emission enqueued without a host round trip, sizes bounded by the
    previous extraction) returns the very arrays of the count -> fill pair -- also when the mesh outgrew the bound
    (the kernels then write nothing and the exact-size fill follows), and for an empty result."""
    from mq3d_b200.vbg import VoxelBlockGrid
    og, vbg = _grids(oracle, cuda_device)
    first = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5, with_keys=True)]      # count -> fill
    assert vbg._mesh_cap[0] >= len(first[0])
    second = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5, with_keys=True)]     # single call
    for a, b in zip(first, second):
        assert a.dtype == b.dtype and np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                                                     b.view(np.uint32) if b.dtype == np.float32 else b)
    # a bound that is too small: nothing is written by the capped kernels, the fallback fills exact buffers
    vbg._mesh_cap = (len(first[0]) // 2, len(first[2]) // 2)
    third = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5, with_keys=True)]
    vbg._mesh_cap = (len(first[0]) + 5, len(first[2]) // 2)              # only the triangles do not fit
    fourth = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5, with_keys=True)]
    for got in (third, fourth):
        for a, b in zip(first, got):
            assert np.array_equal(a, b)
    # threshold above every weight: empty mesh through the single call
    v, n, t = vbg.extract_triangle_mesh_arrays(1e9)
    assert v.shape == (0, 3) and t.shape == (0, 3)
    empty = VoxelBlockGrid(voxel_size=0.02, block_count=100, device=cuda_device)
    empty._mesh_cap = (10, 10)
    v, n, t = empty.extract_triangle_mesh_arrays(1.5)
    assert v.shape == (0, 3) and t.shape == (0, 3)
