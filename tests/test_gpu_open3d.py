"""Parity against the REAL Open3D, wherever the wheel exists.

`open3d==0.19.0` (environment.yml:17) cannot be installed offline in the build container or on the GPU box, so this
file is skipped there -- K2 / K3 / K5 / K6 then stay "parity unpinned" (pinned to the oracle restatement only).  On
any machine that has both a CUDA device and the wheel it is decisive: the reference's VoxelBlockGrid CPU path and the
sm_100a path are run on the same synthetic frames and diffed at the north star's tolerances -- active block set and
per-block mesh topology exact, TSDF / weight within 1e-5 relative, mesh vertices within 1e-4 m with the Chamfer
distance of analysis/computation/compare_mesh_to_ground_truth.py:162 -- and the `colorless_vbg.npz` bridge
(dataio/reconstruction_data_io.py:42-55) is exercised in both directions.
"""
import sys

import numpy as np
import pytest
import torch

from helpers import canonical_mesh, capture, pipeline_cameras, sort_blocks

pytestmark = pytest.mark.gpu


def _real_open3d():
    mod = sys.modules.get("open3d")
    if mod is not None and getattr(mod, "__name__", "") != "open3d":
        del sys.modules["open3d"]           # the stand-in of test_reference_dropin: look for the real package
    return pytest.importorskip("open3d", reason="the real open3d wheel is not installed (offline image)")


DEPTH_MAX, TRUNC, VOXEL, THR = 4.0, 10.0, 0.02, 1.5


def _frames(oracle, n=12):
    cap = capture(n)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin = np.stack([oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(n)])
    return lin, K, Ewc


def _open3d_grid(o3d, lin, K, Ewc):
    """The reference's own loop (o3d_utils.py:171-229) on Open3D's CPU device."""
    dev = o3d.core.Device("CPU:0")
    vbg = o3d.t.geometry.VoxelBlockGrid(attr_names=("tsdf", "weight"), attr_dtypes=(o3d.core.float32, o3d.core.float32),
                                        attr_channels=((1), (1)), voxel_size=VOXEL, block_resolution=16, block_count=2000,
                                        device=dev)
    key_sets = []
    for i in range(len(lin)):
        depth = o3d.t.geometry.Image(tensor=o3d.core.Tensor(lin[i], dtype=o3d.core.Dtype.Float32, device=dev))
        intrinsic = o3d.core.Tensor(K[i], dtype=o3d.core.Dtype.Float64)
        extrinsic = o3d.core.Tensor(Ewc[i], dtype=o3d.core.Dtype.Float64)
        coords = vbg.compute_unique_block_coordinates(depth=depth, intrinsic=intrinsic, extrinsic=extrinsic, depth_scale=1.0,
                                                      depth_max=DEPTH_MAX, trunc_voxel_multiplier=TRUNC)
        key_sets.append({tuple(k) for k in coords.numpy().tolist()})
        vbg.integrate(block_coords=coords, depth=depth, intrinsic=intrinsic, extrinsic=extrinsic, depth_scale=1.0,
                      depth_max=DEPTH_MAX, trunc_voxel_multiplier=TRUNC)
    return vbg, key_sets


def _chamfer(a, b):
    """mean(a -> b) + mean(b -> a), compare_mesh_to_ground_truth.py:139-165."""
    from scipy.spatial import cKDTree
    return cKDTree(b).query(a)[0].mean() + cKDTree(a).query(b)[0].mean()


def test_grid_and_mesh_against_real_open3d(cuda_device, oracle, tmp_path):
    o3d = _real_open3d()
    from mq3d_b200.vbg import VoxelBlockGrid
    lin, K, Ewc = _frames(oracle)
    ref, ref_keys = _open3d_grid(o3d, lin, K, Ewc)
    mine = VoxelBlockGrid(voxel_size=VOXEL, block_count=2000, device=cuda_device)
    for i in range(len(lin)):
        keys = mine.compute_unique_block_coordinates(lin[i], K[i].astype(np.float64), Ewc[i].astype(np.float64), 1.0,
                                                     DEPTH_MAX, TRUNC)
        assert {tuple(k) for k in keys.cpu().numpy().tolist()} == ref_keys[i], f"K2: touched blocks of frame {i}"
        mine.integrate(keys, lin[i], K[i].astype(np.float64), Ewc[i].astype(np.float64), 1.0, DEPTH_MAX, TRUNC)
    # ---- grid through the npz bridge, both directions ----
    ref.save(str(tmp_path / "open3d_vbg.npz"))
    z = np.load(tmp_path / "open3d_vbg.npz")
    loaded = VoxelBlockGrid.load(str(tmp_path / "open3d_vbg.npz"), device=cuda_device)          # Open3D's file, our reader
    k0, t0, w0 = sort_blocks(*[x.cpu().numpy() for x in loaded.export_blocks()[:3]])
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in mine.export_blocks()[:3]])
    assert np.array_equal(k0, k1), "active block set"
    assert np.array_equal(w0, w1), "weights"
    assert np.allclose(t0, t1, rtol=1e-5, atol=1e-7), "tsdf within 1e-5 relative"
    print("tsdf bit-exact vs Open3D:", bool(np.array_equal(t0.view(np.uint32), t1.view(np.uint32))), "npz keys:", sorted(z.files))
    mine.save(str(tmp_path / "mq3d_vbg.npz"))
    back = o3d.t.geometry.VoxelBlockGrid.load(str(tmp_path / "mq3d_vbg.npz"))                   # our file, Open3D's reader
    bm = back.extract_triangle_mesh(weight_threshold=THR, estimated_vertex_number=-1).to_legacy()
    # ---- mesh ----
    rm = ref.extract_triangle_mesh(weight_threshold=THR, estimated_vertex_number=-1).to_legacy()
    rv, rt = np.asarray(rm.vertices), np.asarray(rm.triangles)
    v, n, t, vk = [x.cpu().numpy() for x in mine.extract_triangle_mesh_arrays(THR, with_keys=True)]
    assert len(rv) == len(v) == len(np.asarray(bm.vertices)) and len(rt) == len(t) == len(np.asarray(bm.triangles))
    # Open3D does not expose lattice-edge keys: derive them from the positions (vertex = vs * (g + ratio * e_axis))
    def keys_of(p):
        g = np.floor(p / VOXEL + 1e-4).astype(np.int64)
        frac = p / VOXEL - g
        return np.concatenate([g, np.argmax(frac, axis=1)[:, None]], axis=1)
    a = canonical_mesh(rv.astype(np.float32), rt, keys_of(rv))
    b = canonical_mesh(v, t, vk)
    assert np.array_equal(a["tris"], b["tris"]), "per-block mesh topology"
    assert np.abs(a["verts"] - b["verts"]).max() <= 1e-4, "vertices within 1e-4 m"
    cd = _chamfer(rv, v.astype(np.float64))
    print(f"Chamfer distance vs Open3D CPU: {cd:.3e} m")
    assert cd <= 1e-4
    rn = np.asarray(rm.vertex_normals)
    if len(rn):
        order_r = np.lexsort(keys_of(rv).T[::-1])
        order_m = np.lexsort(vk.T[::-1])
        assert np.allclose(rn[order_r], n[order_m], atol=1e-4)


def test_point_cloud_and_raycast_against_real_open3d(cuda_device, oracle):
    o3d = _real_open3d()
    from mq3d_b200.raycast import RaycastingScene
    from mq3d_b200.vbg import VoxelBlockGrid
    lin, K, Ewc = _frames(oracle)
    ref, _ = _open3d_grid(o3d, lin, K, Ewc)
    mine = VoxelBlockGrid(voxel_size=VOXEL, block_count=2000, device=cuda_device)
    mine.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC)
    rp = np.asarray(ref.extract_point_cloud().to_legacy().points)
    p = mine.extract_point_cloud_arrays(3.0)[0].cpu().numpy()
    assert len(rp) == len(p)
    assert np.abs(np.sort(rp, axis=0) - np.sort(p.astype(np.float64), axis=0)).max() <= 1e-4
    # raycast: Open3D / Embree on the same mesh, colour-view intrinsics as float32 (o3d_utils.py:324-342)
    v, _, t = mine.extract_triangle_mesh_arrays(THR)
    scene = o3d.t.geometry.RaycastingScene()
    scene.add_triangles(o3d.core.Tensor(v.cpu().numpy()), o3d.core.Tensor(t.cpu().numpy().astype(np.uint32)))
    Kc = np.array([[87.0, 0, 64.0], [0, 87.0, 48.0], [0, 0, 1.0]], np.float32)
    rays = scene.create_rays_pinhole(o3d.core.Tensor(Kc), o3d.core.Tensor(Ewc[3].astype(np.float32)), width_px=128, height_px=96)
    want = scene.cast_rays(rays)["t_hit"].numpy()
    my = RaycastingScene(device=cuda_device)
    my.add_triangles(v, t)
    got = my.cast_rays(my.create_rays_pinhole(Kc, Ewc[3].astype(np.float32), 128, 96))["t_hit"].cpu().numpy()
    both = np.isfinite(want) & np.isfinite(got)
    assert (np.isfinite(want) != np.isfinite(got)).mean() < 2e-3
    assert np.quantile(np.abs(want[both] - got[both]), 0.999) < 1e-4
