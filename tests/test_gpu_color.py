"""GPU parity: colour integration (Open3D colour overload of Integrate) vs the oracle.
There is no reference call site for this overload, so it is pinned to the oracle restatement only."""
import numpy as np
import pytest
import torch

from helpers import capture, oracle_integrate_sequence, pipeline_cameras, sort_blocks

pytestmark = pytest.mark.gpu


def _setup(oracle, n=6, cw=160, ch=120):
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import synth
    cap = capture(n)
    K, Ewc, Ecw = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin = np.stack([oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(n)])
    f = 110.0
    colors = np.stack([synth.make_color_frame(Ecw[i], width=cw, height=ch, f=f) for i in range(n)])
    Kc = np.tile(np.array([[f, 0, cw / 2.0], [0, f, ch / 2.0], [0, 0, 1.0]]), (n, 1, 1))
    return lin, K, Ewc, colors, Kc


@pytest.mark.parametrize("batch", [1, 64])
def test_color_sequence_matches_oracle(cuda_device, oracle, batch):
    from mq3d_b200.vbg import VoxelBlockGrid
    lin, K, Ewc, colors, Kc = _setup(oracle)
    og = oracle.Grid(0.02, with_color=True)
    oracle_integrate_sequence(oracle, og, lin, K, Ewc, 4.0, 10.0, colors=colors, Kc=Kc)
    vbg = VoxelBlockGrid(attr_names=("tsdf", "weight", "color"), voxel_size=0.02, block_count=500, device=cuda_device)
    vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, 4.0, 10.0,
                           colors=torch.from_numpy(colors).to(cuda_device), color_intrinsics=Kc, batch_frames=batch)
    k0, t0, w0, c0 = sort_blocks(*og.export())
    k1, t1, w1, c1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1)
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32))
    assert c0.max() > 100 and (c0.sum(-1) > 0).mean() > 0.2      # colour was actually written


def test_color_per_frame_matches_oracle(cuda_device, oracle):
    from mq3d_b200.vbg import VoxelBlockGrid
    lin, K, Ewc, colors, Kc = _setup(oracle, n=3)
    og = oracle.Grid(0.02, with_color=True)
    vbg = VoxelBlockGrid(attr_names=("tsdf", "weight", "color"), voxel_size=0.02, block_count=500, device=cuda_device)
    for i in range(3):
        keys = og.touch(lin[i], K[i], Ewc[i], 4.0, 10.0)
        og.integrate(keys, lin[i], K[i], Ewc[i], 4.0, 10.0, color=colors[i], Kc=Kc[i])
        vbg.integrate(keys, lin[i], colors[i], K[i].astype(np.float64), Kc[i], Ewc[i].astype(np.float64), 1.0, 4.0, 10.0)
    k0, t0, w0, c0 = sort_blocks(*og.export())
    k1, t1, w1, c1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1)
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32))


def test_mesh_and_point_colors_match_oracle(cuda_device, oracle):
    """Colour branch of extract_triangle_mesh / extract_point_cloud: bit-exact vs the oracle, matched through the
    (voxel, axis) keys; also checks the Open3D-shaped objects carry the colours."""
    from mq3d_b200.vbg import VoxelBlockGrid
    lin, K, Ewc, colors, Kc = _setup(oracle)
    og = oracle.Grid(0.02, with_color=True)
    oracle_integrate_sequence(oracle, og, lin, K, Ewc, 4.0, 10.0, colors=colors, Kc=Kc)
    vbg = VoxelBlockGrid(attr_names=("tsdf", "weight", "color"), voxel_size=0.02, block_count=500, device=cuda_device)
    vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, 4.0, 10.0,
                           colors=torch.from_numpy(colors).to(cuda_device), color_intrinsics=Kc)
    v, n, t, vk, vc = [x.cpu().numpy() for x in vbg.extract_triangle_mesh_arrays(1.5, with_keys=True, with_colors=True)]
    assert len(v) > 1000 and vc.shape == v.shape
    ref = og.vertex_colors(vk)
    assert np.array_equal(ref.view(np.uint32), vc.view(np.uint32))
    assert 0.0 <= vc.min() and vc.max() <= 1.0 and vc.max() > 0.3
    # same vertices, oracle order: positions identify the match, colours follow
    ov, _, _, ovk = og.extract_mesh(1.5)
    order_o, order_g = np.lexsort(ovk.T[::-1]), np.lexsort(vk.T[::-1])
    assert np.array_equal(ovk[order_o], vk[order_g])
    assert np.array_equal(og.vertex_colors(ovk)[order_o].view(np.uint32), vc[order_g].view(np.uint32))
    p, pn, pk, pc = [x.cpu().numpy() for x in vbg.extract_point_cloud_arrays(1.5, with_keys=True, with_colors=True)]
    assert len(p) > 1000
    assert np.array_equal(og.vertex_colors(pk).view(np.uint32), pc.view(np.uint32))
    mesh = vbg.extract_triangle_mesh(1.5)
    assert mesh.vertex.colors is not None and tuple(mesh.vertex.colors.shape) == v.shape
    assert vbg.extract_point_cloud(1.5).point.colors is not None
    # a colourless grid refuses instead of inventing colours
    plain = VoxelBlockGrid(voxel_size=0.02, block_count=500, device=cuda_device)
    plain.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, 4.0, 10.0)
    with pytest.raises(Exception):
        plain.extract_triangle_mesh_arrays(1.5, with_colors=True)


def test_depth_scale_is_applied(cuda_device, oracle):
    """depth_scale != 1 (never used by the reference, supported by Open3D): depth / scale per sample."""
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(2)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin = np.stack([oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(2)])
    mm = (lin * np.float32(1000.0)).astype(np.float32)
    og = oracle.Grid(0.02)
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=500, device=cuda_device)
    for i in range(2):
        okeys = og.touch(mm[i], K[i], Ewc[i], 4.0, 10.0, depth_scale=1000.0)
        keys = vbg.compute_unique_block_coordinates(mm[i], K[i].astype(np.float64), Ewc[i].astype(np.float64), 1000.0, 4.0, 10.0)
        assert {tuple(k) for k in keys.cpu().numpy().tolist()} == {tuple(k) for k in okeys.tolist()}
        og.integrate(okeys, mm[i], K[i], Ewc[i], 4.0, 10.0, depth_scale=1000.0)
        vbg.integrate(keys, mm[i], K[i].astype(np.float64), Ewc[i].astype(np.float64), 1000.0, 4.0, 10.0)
    k0, t0, w0 = sort_blocks(*og.export()[:3])
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32))


def test_color_resample_paths_agree(cuda_device, oracle):
    """Colour on the depth grid: the resampler gives the same image from device memory and from pinned host
    memory read in place (zero-copy), matches a NumPy restatement of the look-up, and integrating the
    pre-resampled frames (or pinned frames directly) is bit-identical to integrating the raw device frames."""
    from mq3d_b200.vbg import VoxelBlockGrid, color_resample
    lin, K, Ewc, colors, Kc = _setup(oracle, n=5, cw=161, ch=119)        # odd sizes: unaligned rows
    n, H, W = lin.shape
    dev_cols = torch.from_numpy(colors).to(cuda_device)
    pin_cols = torch.from_numpy(colors).pin_memory()
    a = color_resample(dev_cols, K, Kc, W, H)
    b = color_resample(pin_cols, K, Kc, W, H, device=cuda_device)
    torch.cuda.synchronize()
    assert a.dtype == torch.int32 and tuple(a.shape) == (n, H, W) and torch.equal(a, b)
    with pytest.raises(RuntimeError):
        color_resample(torch.from_numpy(colors), K, Kc, W, H)             # pageable host memory is refused
    # NumPy restatement (float32 arithmetic of TransformIndexer Unproject / Project, round half away from zero)
    got = a.cpu().numpy().view(np.uint32)
    for f in range(n):
        fx, fy, cx, cy = (np.float32(K[f][0, 0]), np.float32(K[f][1, 1]), np.float32(K[f][0, 2]), np.float32(K[f][1, 2]))
        cfx, cfy, ccx, ccy = (np.float32(Kc[f][0, 0]), np.float32(Kc[f][1, 1]), np.float32(Kc[f][0, 2]), np.float32(Kc[f][1, 2]))
        uf = cfx * ((np.arange(W, dtype=np.float32) - cx) / fx) + ccx
        vf = cfy * ((np.arange(H, dtype=np.float32) - cy) / fy) + ccy
        rnd = lambda x: np.where(x >= 0, np.floor(x + np.float32(0.5)), np.ceil(x - np.float32(0.5))).astype(np.int64)
        u_ok, v_ok = (uf >= 0) & (uf <= colors.shape[2] - 1), (vf >= 0) & (vf <= colors.shape[1] - 1)
        cu, cv = np.clip(rnd(uf), 0, colors.shape[2] - 1), np.clip(rnd(vf), 0, colors.shape[1] - 1)
        px = colors[f][cv[:, None], cu[None, :]].astype(np.uint32)
        want = px[..., 0] | (px[..., 1] << 8) | (px[..., 2] << 16)
        want = np.where(v_ok[:, None] & u_ok[None, :], want, np.uint32(0xFF000000))
        assert np.array_equal(got[f], want), f
    assert (got >> 24 == 0).mean() > 0.3 and (got >> 24 != 0).any()          # both inside and outside occur
    d = torch.from_numpy(lin).to(cuda_device)
    grids = []
    for kw in (dict(colors=dev_cols, color_intrinsics=Kc), dict(colors_rgbx=a), dict(colors=pin_cols, color_intrinsics=Kc)):
        g = VoxelBlockGrid(attr_names=("tsdf", "weight", "color"), voxel_size=0.02, block_count=500, device=cuda_device)
        g.integrate_sequence(d, K, Ewc, 4.0, 10.0, batch_frames=2, **kw)
        grids.append(sort_blocks(*[x.cpu().numpy() for x in g.export_blocks()]))
    for other in grids[1:]:
        for x, y in zip(grids[0], other):
            assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x,
                                  y.view(np.uint32) if y.dtype == np.float32 else y)
