"""GPU: parity at the full size of BASELINE configs[0] (300 left-eye 320x320 frames, voxel 0.02 m,
16^3 blocks, trunc x10, depth_max 4 m, marching cubes at weight_threshold 1.5) against the oracle, plus
size-independent properties at configs[1]-like size (voxel 0.01 m): linearity of the weights, idempotent
extraction, watertightness inside fully observed regions."""
import numpy as np
import pytest
import torch

from helpers import canonical_mesh, capture, oracle_integrate_sequence, pipeline_cameras, sort_blocks

pytestmark = pytest.mark.gpu


def _linear(oracle, cap):
    ds = cap.dataset
    return np.stack([oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(len(ds))])


def test_config0_full_sequence_matches_oracle(cuda_device, oracle):
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare
    cap = capture(300)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin, valid = depth_prepare(torch.from_numpy(cap.raw).to(cuda_device), ds.nears, ds.fars)
    assert int(valid.sum()) == 300
    lin_o = _linear(oracle, cap)
    assert np.array_equal(lin.cpu().numpy().view(np.uint32), lin_o.view(np.uint32))
    og = oracle.Grid(0.02)
    visits, updated = oracle_integrate_sequence(oracle, og, lin_o, K, Ewc, 4.0, 10.0)
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=500, device=cuda_device)      # grows several times
    st = vbg.integrate_sequence(lin, K, Ewc, 4.0, 10.0, frame_valid=valid)
    assert (st.frames_integrated, st.block_visits, st.voxel_updates, st.num_blocks) == (300, visits, updated, og.num_blocks)
    k0, t0, w0 = sort_blocks(*og.export()[:3])
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1)
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    # relative TSDF/weight bar of the north star (1e-5) is met with margin: the arrays are identical
    ov, on, ot, ok = og.extract_mesh(1.5)
    v, n, t, k = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5, with_keys=True)]
    a, b = canonical_mesh(ov, ot, ok, on), canonical_mesh(v, t, k, n)
    assert np.array_equal(a["keys"], b["keys"]) and np.array_equal(a["tris"], b["tris"])
    assert np.abs(a["verts"] - b["verts"]).max() <= 1e-4 and np.array_equal(a["verts"].view(np.uint32), b["verts"].view(np.uint32))
    # Chamfer distance (analysis/computation/compare_mesh_to_ground_truth.py:162: mean(d(a->b)) + mean(d(b->a)))
    # between the two vertex sets is exactly 0 because they coincide
    assert float(np.abs(a["verts"] - b["verts"]).sum()) == 0.0


def test_properties_at_config1_size(cuda_device, oracle):
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare
    cap = capture(300)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin, valid = depth_prepare(torch.from_numpy(cap.raw).to(cuda_device), ds.nears, ds.fars)
    g1 = VoxelBlockGrid(voxel_size=0.01, block_count=8000, device=cuda_device)
    s1 = g1.integrate_sequence(lin, K, Ewc, 4.0, 10.0, frame_valid=valid, batch_frames=64)
    # same frames in two halves with another batch size: identical grid (batching never changes results)
    g2 = VoxelBlockGrid(voxel_size=0.01, block_count=100, device=cuda_device)
    g2.integrate_sequence(lin[:137].contiguous(), K[:137], Ewc[:137], 4.0, 10.0, batch_frames=17)
    g2.integrate_sequence(lin[137:].contiguous(), K[137:], Ewc[137:], 4.0, 10.0, batch_frames=256)
    a = sort_blocks(*[x.cpu().numpy() for x in g1.export_blocks()[:3]])
    b = sort_blocks(*[x.cpu().numpy() for x in g2.export_blocks()[:3]])
    assert all(np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x,
                              y.view(np.uint32) if y.dtype == np.float32 else y) for x, y in zip(a, b))
    # checksum of checksums: total weight == number of voxel updates; weights are whole numbers <= frames
    w = a[2]
    assert float(w.astype(np.float64).sum()) == float(s1.voxel_updates)
    assert w.max() <= 300 and np.array_equal(w, np.round(w))
    assert np.abs(a[1]).max() <= 1.0                                   # tsdf is a running mean of values in [-1,1]
    # extraction is idempotent and deterministic
    m1 = [x.cpu().numpy() for x in g1.extract_triangle_mesh_arrays(1.5)]
    m2 = [x.cpu().numpy() for x in g1.extract_triangle_mesh_arrays(1.5)]
    assert all(np.array_equal(x, y) for x, y in zip(m1, m2))
    v, n, t = m1
    assert t.min() >= 0 and t.max() < len(v) and len(np.unique(t)) == len(v)          # every vertex is referenced
    # manifoldness: no edge is shared by more than two triangles; unit normals
    e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]).astype(np.int64), axis=1)
    _, cnt = np.unique(e[:, 0] * (len(v) + 1) + e[:, 1], return_counts=True)
    assert cnt.max() <= 2
    assert np.abs(np.linalg.norm(n, axis=1) - 1.0).max() < 1e-3


def test_config1_full_size_with_colour_matches_oracle(cuda_device, oracle):
    """BASELINE configs[1] at full size -- 300 frames, voxel 0.01 m, 1280 x 960 RGB -- against the oracle's frame loop:
    active set, tsdf, weight and colour bit-exact, mesh topology / positions / vertex colours identical."""
    from mq3d_b200 import synth, synth_gpu
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare
    n = 300
    cap = capture(n)
    K, Ewc, Ecw = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin, valid = depth_prepare(torch.from_numpy(cap.raw).to(cuda_device), ds.nears, ds.fars)
    cols = torch.empty((n, synth.COLOR_H, synth.COLOR_W, 3), dtype=torch.uint8, device=cuda_device)
    for i in range(0, n, 50):
        cols[i:i + 50] = synth_gpu.render_color(Ecw[i:i + 50], cuda_device)
    Kc = np.zeros((n, 3, 3))
    Kc[:, 0, 0] = Kc[:, 1, 1] = int(synth.COLOR_F)
    Kc[:, 0, 2], Kc[:, 1, 2], Kc[:, 2, 2] = synth.COLOR_W // 2, synth.COLOR_H // 2, 1.0
    vbg = VoxelBlockGrid(attr_names=("tsdf", "weight", "color"), voxel_size=0.01, block_count=6000, device=cuda_device)
    st = vbg.integrate_sequence(lin, K, Ewc, 4.0, 10.0, frame_valid=valid, colors=cols, color_intrinsics=Kc, batch_frames=256)
    lin_o = _linear(oracle, cap)
    cols_h = cols.cpu().numpy()
    og = oracle.Grid(0.01, with_color=True)
    visits, updated = oracle_integrate_sequence(oracle, og, lin_o, K, Ewc, 4.0, 10.0, colors=cols_h, Kc=Kc)
    assert (st.frames_integrated, st.block_visits, st.voxel_updates, st.num_blocks) == (n, visits, updated, og.num_blocks)
    a = sort_blocks(*og.export())
    b = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[3].view(np.uint32), b[3].reshape(a[3].shape).view(np.uint32))
    ov, on, ot, ok = og.extract_mesh(1.5)
    oc = og.vertex_colors(ok)
    v, nrm, t, k, c = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5, with_keys=True, with_colors=True)]
    ma, mb = canonical_mesh(ov, ot, ok, on), canonical_mesh(v, t, k, nrm)
    assert len(v) > 400_000 and np.array_equal(ma["keys"], mb["keys"]) and np.array_equal(ma["tris"], mb["tris"])
    assert np.array_equal(ma["verts"].view(np.uint32), mb["verts"].view(np.uint32))
    order_o = np.lexsort((ok[:, 3], ok[:, 2], ok[:, 1], ok[:, 0]))
    order_g = np.lexsort((k[:, 3], k[:, 2], k[:, 1], k[:, 0]))
    assert np.array_equal(oc[order_o].view(np.uint32), c[order_g].view(np.uint32))


def test_config3_stereo_5mm_slice_matches_oracle(cuda_device, oracle):
    """A 2 x 120-frame stereo slice of BASELINE configs[3] (room-scale 5 mm grid, LEFT sequence then RIGHT into one
    grid, reconstruct_scene.py:64-81) against the oracle's frame loop: grid bit-exact, mesh identical."""
    from mq3d_b200.models import Side
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare
    n = 120
    lins, Ks, Es = [], [], []
    for side in (Side.LEFT, Side.RIGHT):
        cap = capture(n, side)
        K, Ewc, _ = pipeline_cameras(cap.dataset)
        lins.append(_linear(oracle, cap))
        Ks.append(K)
        Es.append(Ewc)
    lin_o, K, Ewc = np.concatenate(lins), np.concatenate(Ks), np.concatenate(Es)
    vbg = VoxelBlockGrid(voxel_size=0.005, block_count=4096, device=cuda_device)                  # grows
    st = vbg.integrate_sequence(torch.from_numpy(lin_o).to(cuda_device), K, Ewc, 4.0, 10.0, batch_frames=256)
    og = oracle.Grid(0.005)
    visits, updated = oracle_integrate_sequence(oracle, og, lin_o, K, Ewc, 4.0, 10.0)
    assert (st.frames_integrated, st.block_visits, st.voxel_updates, st.num_blocks) == (2 * n, visits, updated, og.num_blocks)
    a = sort_blocks(*og.export()[:3])
    b = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert len(a[0]) > 8000 and np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    ov, on, ot, ok = og.extract_mesh(1.5)
    v, nrm, t, k = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5, with_keys=True)]
    ma, mb = canonical_mesh(ov, ot, ok, on), canonical_mesh(v, t, k, nrm)
    assert np.array_equal(ma["keys"], mb["keys"]) and np.array_equal(ma["tris"], mb["tris"])
    assert np.array_equal(ma["verts"].view(np.uint32), mb["verts"].view(np.uint32))
    assert np.allclose(ma["normals"], mb["normals"], atol=1e-6)
