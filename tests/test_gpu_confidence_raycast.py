"""GPU parity: K4 multi-view confidence (vs reference golden vectors and the oracle) and K6 raycast
(vs the oracle's double-precision closest hit)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from helpers import capture, oracle_integrate_sequence, pipeline_cameras

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_k1_k4.npz"))


def test_k1_matches_reference_golden(cuda_device):
    from mq3d_b200.vbg import depth_prepare
    raw = torch.from_numpy(G["k1_raw"]).to(cuda_device)
    for c, (near, far) in enumerate(G["k1_near_far"]):
        out, _ = depth_prepare(raw, np.full(3, near), np.full(3, far))
        assert np.array_equal(out.cpu().numpy().view(np.uint32), G["k1_linear"][c].view(np.uint32)), c
    vin = G["k1_valid_in"]
    finite = np.nan_to_num(vin, nan=0.5)
    _, valid = depth_prepare(torch.from_numpy(vin).to(cuda_device), np.full(5, 0.1), np.full(5, np.inf))
    assert valid.cpu().tolist() == G["k1_valid"].tolist()


def test_confidence_matches_reference_golden(cuda_device):
    from mq3d_b200.vbg import estimate_confidence
    K, Ecw = G["cam_K"], G["cam_Ecw"]
    Einv = np.linalg.inv(Ecw)
    conf, count = estimate_confidence(torch.from_numpy(G["k4_depths"]).to(cuda_device), K, Ecw, Einv, 2, 4.0, 0.08)
    assert conf.dtype == torch.float64 and count.dtype == torch.int32
    assert np.array_equal(count.cpu().numpy(), G["k4_count"])
    mism = (conf.cpu().numpy() != G["k4_conf"]).mean()
    assert mism <= 1e-6, mism


def test_confidence_matches_oracle_bit_exact(cuda_device, oracle):
    from mq3d_b200.vbg import depth_prepare, estimate_confidence
    cap = capture(12)
    K, _, Ecw = pipeline_cameras(cap.dataset)
    Einv = np.linalg.inv(Ecw)
    ds = cap.dataset
    raw = cap.raw.copy()
    raw[5] = 1.0                                    # an unreadable frame: skipped as ref and as target
    lin, valid = depth_prepare(torch.from_numpy(raw).to(cuda_device), ds.nears, ds.fars)
    conf, count = estimate_confidence(lin, K, Ecw, Einv, 3, 4.0, 0.08, frame_valid=valid)
    oc, on = oracle.confidence(lin.cpu().numpy(), K, Ecw, Einv, 3, 4.0, 0.08, frame_ok=valid.cpu().numpy())
    assert np.array_equal(count.cpu().numpy(), on)
    assert np.array_equal(conf.cpu().numpy(), oc)
    assert on[5].max() == 0 and on[4].max() == 5 and on[6].max() == 5   # frame 5 never counted


def _mesh(cuda_device, oracle):
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(16)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin = np.stack([oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(len(ds))])
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=4000, device=cuda_device)
    vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, 4.0, 10.0)
    return vbg.extract_triangle_mesh(1.5), Ewc


def test_raycast_matches_oracle(cuda_device, oracle):
    from mq3d_b200.raycast import RaycastingScene
    mesh, Ewc = _mesh(cuda_device, oracle)
    scene = RaycastingScene(device=cuda_device)
    scene.add_triangles(mesh)
    Kc = np.array([[217.5, 0, 160.0], [0, 217.5, 120.0], [0, 0, 1.0]], np.float32)   # 320x240 colour-like view
    v = mesh.vertex.positions.cpu().numpy()
    t = mesh.triangle.indices.cpu().numpy()
    for i in (0, 7, 15):
        rays = scene.create_rays_pinhole(Kc, Ewc[i], width_px=320, height_px=240)
        want_rays = oracle.create_rays_pinhole(Kc.astype(np.float64), Ewc[i].astype(np.float64), 320, 240)
        assert np.array_equal(rays.cpu().numpy().view(np.uint32), want_rays.view(np.uint32))
        got = scene.cast_rays(rays)["t_hit"].cpu().numpy()
        want = oracle.cast_rays(v, t, want_rays)
        assert got.shape == (240, 320) and got.dtype == np.float32
        hit = np.isfinite(want)
        assert hit.mean() > 0.5
        agree = np.isfinite(got) == hit
        assert (~agree).mean() < 2e-4                      # hit/miss differs only on silhouette pixels
        both = hit & np.isfinite(got)
        err = np.abs(got[both] - want[both])
        assert np.quantile(err, 0.999) < 1e-4              # north-star tolerance 1e-4 m
        assert (err > 1e-3).mean() < 2e-4                  # edge-on triangles at depth discontinuities


def test_raycast_plane_and_empty_scene(cuda_device):
    from mq3d_b200.geometry import TriangleMesh
    from mq3d_b200.raycast import RaycastingScene
    verts = torch.tensor([[-5, -5, 2.0], [5, -5, 2.0], [5, 5, 2.0], [-5, 5, 2.0]], dtype=torch.float32)
    tris = torch.tensor([[0, 1, 2], [0, 2, 3]], dtype=torch.int32)
    scene = RaycastingScene(device=cuda_device)
    scene.add_triangles(TriangleMesh(verts, tris))
    Kc = np.array([[100.0, 0, 32.0], [0, 100.0, 24.0], [0, 0, 1.0]])
    rays = scene.create_rays_pinhole(Kc, np.eye(4), width_px=64, height_px=48)
    t = scene.cast_rays(rays)["t_hit"].cpu().numpy()
    assert np.allclose(t, 2.0, atol=1e-6)
    empty = RaycastingScene(device=cuda_device)
    assert np.isinf(empty.cast_rays(rays)["t_hit"].cpu().numpy()).all()


def test_odometry_information_matches_oracle(cuda_device, oracle):
    """N4: mq3d_odometry_information against the oracle on consecutive frames and on a key-frame pair of a synthetic
    sequence (same correspondences; float64 sums agree to rounding of the summation order)."""
    from helpers import capture, pipeline_cameras
    from mq3d_b200.ops import compute_odometry_information_matrix
    cap = capture(12)
    K, Ewc, Ecw = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin = [oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(12)]
    for i, j in ((0, 1), (3, 4), (0, 10)):
        T = Ewc[j].astype(np.float64) @ Ecw[i].astype(np.float64)
        want = oracle.odometry_information(lin[i], lin[j], K[0], T, 0.07, 1.0, 4.0)
        got = compute_odometry_information_matrix(torch.from_numpy(lin[i]).to(cuda_device), torch.from_numpy(lin[j]).to(cuda_device),
                                                  K[0], T, dist_threshold=0.07, depth_scale=1.0, depth_max=4.0)
        assert got.shape == (6, 6) and got.dtype == np.float64 and want[5, 5] > 1000
        assert got[5, 5] == want[5, 5]                                  # the same pixels correspond
        assert np.allclose(got, want, rtol=1e-9, atol=1e-6)
