"""GPU: filter_mesh_components on the device (mq3d_mesh_filter) against the host implementation of the
same Open3D legacy-mesh semantics (meshops.py) -- identical vertex and triangle arrays, attributes included."""
import numpy as np
import pytest
import torch

from helpers import capture, pipeline_cameras

pytestmark = pytest.mark.gpu


def _raw(v, t, nrm=None, col=None, device="cpu"):
    from mq3d_b200.geometry import TriangleMesh
    f = lambda a, dt: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(device)
    return TriangleMesh(f(v, torch.float32), f(t, torch.int32), f(nrm, torch.float32), f(col, torch.float32))


def _both(v, t, nrm, col, min_count, cuda_device):
    from mq3d_b200.ops import filter_mesh_components
    host = filter_mesh_components(_raw(v, t, nrm, col), min_triangle_count=min_count, backend="host")
    dev = filter_mesh_components(_raw(v, t, nrm, col, cuda_device), min_triangle_count=min_count, backend="device")
    assert dev.vertex.positions.is_cuda
    return host, dev


def _same(host, dev):
    for a, b in ((host.vertex.positions, dev.vertex.positions), (host.triangle.indices, dev.triangle.indices),
                 (host.vertex.normals, dev.vertex.normals), (host.vertex.colors, dev.vertex.colors)):
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a.cpu().numpy(), b.cpu().numpy())


def _tetra(offset, scale=1.0):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32) * scale + np.asarray(offset, np.float32)
    t = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]], np.int32)
    return v, t


def test_device_filter_matches_host_on_hand_built_cases(cuda_device, capsys):
    n = 40
    strip_v = np.array([[i // 2, i % 2, 0] for i in range(2 * n + 2)], np.float32)
    strip_t = np.array([[i, i + 1, i + 2] if i % 2 == 0 else [i + 1, i, i + 2] for i in range(2 * n)], np.int32)
    far_v, far_t = _tetra((100, 0, 0))
    v = np.concatenate([strip_v, far_v])
    t = np.concatenate([strip_t, far_t + len(strip_v)])
    nrm, col = -v, np.clip(np.abs(v) / 10.0, 0, 1).astype(np.float32)
    host, dev = _both(v, t, nrm, col, 10, cuda_device)            # the tetrahedron is dropped, its vertices too
    _same(host, dev)
    assert len(dev.triangle.indices) == 2 * n and len(dev.vertex.positions) == len(strip_v)
    capsys.readouterr()
    host, dev = _both(v, t, nrm, col, 10_000, cuda_device)        # nothing qualifies: the largest component alone
    _same(host, dev)
    assert capsys.readouterr().out.count("Returning largest component only") == 2
    # two tetrahedra that share ONE VERTEX only stay two components (edge adjacency)
    va, ta = _tetra((0, 0, 0))
    vb, tb = _tetra((0, 0, 0), scale=-1.0)
    v2 = np.concatenate([va, vb[1:]])
    t2 = np.concatenate([ta, np.where(tb == 0, 0, tb + 3)])
    host, dev = _both(v2, t2, None, None, 5, cuda_device)          # both have 4 < 5 triangles: the FIRST largest stays
    _same(host, dev)
    assert dev.triangle.indices.cpu().numpy().tolist() == ta.tolist()
    # degenerate, rotated-duplicate and opposite-winding triangles; duplicated vertices are welded, order preserved
    v3 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0], [2, 0, 0], [2, 1, 0], [-0.0, 0, 0]], np.float32)
    t3 = np.array([[0, 1, 2], [1, 2, 0], [0, 2, 1], [2, 2, 3], [4, 5, 6], [7, 5, 6]], np.int32)
    host, dev = _both(v3, t3, None, None, 1, cuda_device)
    _same(host, dev)
    assert dev.triangle.indices.cpu().numpy().tolist() == [[0, 1, 2], [0, 2, 1], [1, 4, 5], [0, 4, 5]]
    # three triangles on one edge (never produced by marching cubes): the smallest one goes, like on the host
    v4 = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, -3, 0], [0, 0, 0.5]], np.float32)
    t4 = np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4]], np.int32)
    host, dev = _both(v4, t4, None, None, 1, cuda_device)
    _same(host, dev)
    assert dev.triangle.indices.cpu().numpy().tolist() == [[0, 1, 2], [1, 0, 3]]


def _scene_mesh(oracle, cuda_device, n_frames=24, voxel=0.02, thr=1.5, color=False):
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(n_frames)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin = np.stack([oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(len(ds))])
    vbg = VoxelBlockGrid(voxel_size=voxel, block_count=4000, device=cuda_device)
    vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, 4.0, 10.0)
    return vbg, lin, K, Ewc


@pytest.mark.parametrize("min_count", [1, 200, 5000, 10 ** 7])
def test_device_filter_matches_host_on_a_marching_cubes_mesh(cuda_device, oracle, min_count):
    vbg, *_ = _scene_mesh(oracle, cuda_device)
    v, n, t = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5)]
    assert len(t) > 20000
    host, dev = _both(v, t, n, None, min_count, cuda_device)
    _same(host, dev)
    assert 0 < len(dev.triangle.indices) <= len(t)


def test_device_filter_welds_the_per_rank_meshes_of_a_partitioned_grid(cuda_device, oracle):
    """Per-rank meshes of a hash-partitioned grid re-emit the vertices on ghost edges; after the device filter
    (min_triangle_count = 1: nothing but the weld) the concatenation has exactly the single-grid vertex set and the
    same triangles, coordinate for coordinate."""
    from mq3d_b200.ops import filter_mesh_components
    from mq3d_b200.vbg import VoxelBlockGrid
    full, lin, K, Ewc = _scene_mesh(oracle, cuda_device, n_frames=10)
    fv, fn, ft = full.extract_triangle_mesh_arrays(1.5)
    d = torch.from_numpy(lin).to(cuda_device)
    vs, ts, off = [], [], 0
    for rank in range(3):
        g = VoxelBlockGrid(voxel_size=0.02, block_count=4000, device=cuda_device)
        g.set_partition(rank, 3, 2)
        g.integrate_sequence(d, K, Ewc, 4.0, 10.0)
        v, n, t = g.extract_triangle_mesh_arrays(1.5)
        vs.append(v)
        ts.append(t + off)
        off += int(v.shape[0])
    cat = _raw(torch.cat(vs).cpu().numpy(), torch.cat(ts).cpu().numpy(), device=cuda_device)
    assert cat.vertex.positions.shape[0] > fv.shape[0]                 # duplicates on the seams
    welded = filter_mesh_components(cat, min_triangle_count=1, backend="device")
    single = filter_mesh_components(_raw(fv.cpu().numpy(), ft.cpu().numpy(), device=cuda_device), min_triangle_count=1,
                                    backend="device")
    wv, wt = welded.vertex.positions.cpu().numpy(), welded.triangle.indices.cpu().numpy()
    sv, st = single.vertex.positions.cpu().numpy(), single.triangle.indices.cpu().numpy()
    assert len(wv) == len(sv) and len(wt) == len(st)

    def canon(v, t):          # triangles as sorted rows of vertex coordinates (rotation-normalised)
        p = v[t].reshape(len(t), 9)
        rows = []
        for r in (p, np.roll(p, -3, axis=1), np.roll(p, -6, axis=1)):
            rows.append(r)
        stack = np.stack(rows, axis=1)
        key = np.array([min(map(tuple, s)) for s in stack.tolist()])
        return key[np.lexsort(key.T[::-1])]
    assert np.array_equal(canon(wv, wt), canon(sv, st))
    assert np.array_equal(np.unique(wv, axis=0), np.unique(sv, axis=0))
