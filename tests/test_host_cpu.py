"""CPU: host-side logic -- config parser, dataset containers, on-disk formats, synthetic project layout."""
import os

import numpy as np
import pytest

import mq3d_b200  # noqa: F401
from mq3d_b200 import synth
from mq3d_b200.config import PipelineConfigs, ReconstructionConfig
from mq3d_b200.dataio import DESCRIPTOR_COLUMNS, DataIO, depth_camera_params
from mq3d_b200.geometry import LegacyPointCloud, LegacyTriangleMesh
from mq3d_b200.io_utils import read_ply, write_point_cloud, write_triangle_mesh
from mq3d_b200.models import CameraDataset, ConfidenceMap, DepthDataset, Side

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parse_pipeline_config_yml():
    cfg = PipelineConfigs.parse_config_yml(os.path.join(ROOT, "config", "pipeline_config.yml")).reconstruction
    assert cfg.device == "CUDA:0" and cfg.depth_integration.device == "CUDA:0"
    di = cfg.depth_integration
    assert (di.voxel_size, di.block_resolution, di.block_count, di.depth_max, di.trunc_voxel_multiplier) == \
        (0.01, 16, 50000, 4.0, 10.0)
    assert (di.confidence_threshold, di.valid_count_threshold) == (0.02, 2)
    assert cfg.confidence_estimation.target_frame_range == 10 and cfg.confidence_estimation.error_threshold == 0.08
    assert cfg.color_optimization.weight_threshold == 1.5 and cfg.color_optimization.min_triangle_count == 5000
    assert cfg.fragment_pose_refinement.relative_rmses == [1e-6, 1e-6, 1e-6]     # "1e-6" strings coerced
    assert cfg.use_dataset_cache is False and cfg.color_optimization.use_dataset_cache is False


def test_config_defaults_and_device_policy():
    cfg = ReconstructionConfig.parse({})
    assert cfg.depth_integration.depth_max == 1.5 and cfg.depth_integration.trunc_voxel_multiplier == 8.0
    assert cfg.depth_integration.confidence_threshold == 0.05 and cfg.depth_integration.valid_count_threshold == 4
    assert cfg.color_optimization.weight_threshold == 3.0
    assert ReconstructionConfig.parse({"device": "CPU:0"}).device == "CUDA:0"         # remapped with a warning
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ReconstructionConfig.parse({"device": "CPU:0"}, allow_device_override=False)
    assert ReconstructionConfig.parse({"device": "cuda:3", "depth_integration": {"voxel_size": "0.005"}}) \
        .depth_integration.voxel_size == 0.005


def test_dataset_roundtrip_and_indexing(tmp_path):
    cap = synth.make_depth_capture(5, Side.LEFT, width=32, height=32)
    ds = cap.dataset
    p = tmp_path / "dataset" / "left_depth_dataset.npz"
    ds.save(p)
    ds2 = DepthDataset.load(p)
    assert np.array_equal(ds2.timestamps, ds.timestamps) and np.array_equal(ds2.nears, ds.nears)
    assert ds2.transforms.coordinate_system == ds.transforms.coordinate_system
    assert np.array_equal(ds2.transforms.rotations, ds.transforms.rotations)
    sub = ds[1:4]
    assert len(sub) == 3 and isinstance(sub, DepthDataset) and sub.timestamps[0] == ds.timestamps[1]
    assert [len(f) for f in ds.split(2)] == [2, 2, 1]
    k = ds.get_intrinsic_matrices()
    assert k.dtype == np.float32 and k[0, 0, 0] == 16.0 and k[0, 2, 2] == 1.0
    cds = synth.make_color_dataset(3)
    cds.save(tmp_path / "c.npz")
    assert np.array_equal(CameraDataset.load(tmp_path / "c.npz").fx, cds.fx)
    assert cds.fx.dtype.kind == "i"            # colour intrinsics are integer arrays


def test_project_layout_and_descriptor_csv(tmp_path):
    caps = synth.write_project(tmp_path, 3, width=32, height=32)
    io = DataIO(tmp_path)
    for side in Side:
        ds = caps[side].dataset
        header = (tmp_path / f"{side.value}_depth_descriptors.csv").read_text().splitlines()[0].split(",")
        assert header == DESCRIPTOR_COLUMNS
        raw = io.depth.load_raw_depth_map(side, ds.timestamps[1], 32, 32)
        assert raw.dtype == np.float32 and np.array_equal(raw, caps[side].raw[1])
        assert io.depth.load_raw_depth_map(side, 12345, 32, 32) is None
    assert depth_camera_params(1.0, 1.0, 1.0, 1.0, 320, 320) == (160.0, 160.0, 160.0, 160.0)
    cm = ConfidenceMap(np.random.rand(4, 6), np.ones((4, 6), np.int32))
    io.depth.save_confidence_map(Side.LEFT, 77, cm)
    back = io.depth.load_confidence_map(Side.LEFT, 77)
    assert back.confidence_map.dtype == np.float64 and back.valid_count.dtype == np.int32
    assert np.array_equal(back.confidence_map, cm.confidence_map)
    assert io.depth.load_confidence_map(Side.LEFT, 78) is None
    assert (tmp_path / "left_depth_confidence" / "77.npz").exists()
    with pytest.raises(ValueError):
        ConfidenceMap(np.zeros((2, 2)), np.zeros((3, 2), np.int32))


def test_ply_roundtrip(tmp_path):
    v = np.random.rand(10, 3)
    n = np.random.rand(10, 3)
    t = np.random.randint(0, 10, (7, 3)).astype(np.int32)
    write_triangle_mesh(tmp_path / "m.ply", LegacyTriangleMesh(v, t, n))
    vv, tt = read_ply(tmp_path / "m.ply")
    assert np.array_equal(np.stack([vv["x"], vv["y"], vv["z"]], 1), v) and np.array_equal(tt, t)
    assert np.array_equal(vv["nz"], n[:, 2])
    write_point_cloud(tmp_path / "p.ply", LegacyPointCloud(v, n))
    pv, pt = read_ply(tmp_path / "p.ply")
    assert pt is None and np.array_equal(pv["y"], v[:, 1])


def test_synthetic_capture_is_valid_quest_shaped():
    cap = synth.make_depth_capture(300, Side.LEFT, width=16, height=16)
    assert cap.raw.dtype == np.float32 and cap.raw.shape == (300, 16, 16)
    assert (cap.raw >= 0).all() and (cap.raw <= 1).all()      # never closer than the near plane
    assert 0.005 < (cap.raw == 1.0).mean() < 0.05              # ~2 % dropped pixels
    assert np.array_equal(np.diff(cap.dataset.timestamps)[:3], [33, 34, 33])


def test_raw_sequence_loader_batches_and_flags_missing(tmp_path):
    """load_raw_sequence reads a whole side into one [N,H,W] batch (threaded file reads), zero-fills and flags
    frames whose file is missing, and rejects truncated files."""
    n = 20
    caps = synth.write_project(tmp_path, n, width=32, height=32)
    io = DataIO(tmp_path)
    ds = caps[Side.LEFT].dataset                              # (building it from the CSV validates on the GPU)
    raw, present = io.depth.load_raw_sequence(Side.LEFT, ds)
    assert raw.shape == (len(ds), 32, 32) and raw.dtype == np.float32 and present.all()
    lookup = {int(t): i for i, t in enumerate(caps[Side.LEFT].dataset.timestamps)}
    for i, t in enumerate(ds.timestamps):
        assert np.array_equal(raw[i], caps[Side.LEFT].raw[lookup[int(t)]])
    io.depth.depth_map_path(Side.LEFT, ds.timestamps[3]).unlink()
    io.depth._raw_cache.clear()
    raw2, present2 = io.depth.load_raw_sequence(Side.LEFT, ds)
    assert not present2[3] and present2.sum() == len(ds) - 1 and not raw2[3].any()
    assert np.array_equal(raw2[4], raw[4])
    p = io.depth.depth_map_path(Side.LEFT, ds.timestamps[5])
    p.write_bytes(p.read_bytes()[:100])
    io.depth._raw_cache.clear()
    with pytest.raises(RuntimeError, match="expected 1024 float32"):
        io.depth.load_raw_sequence(Side.LEFT, ds)


def _tetra(offset, scale=1.0):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32) * scale + np.asarray(offset, np.float32)
    t = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]], np.int32)
    return v, t


def _mesh(v, t, with_attrs=True):
    import torch
    from mq3d_b200.geometry import TriangleMesh
    v = np.asarray(v, np.float32)
    nrm = torch.from_numpy(-v) if with_attrs else None
    col = torch.from_numpy(np.clip(np.abs(v) / 10.0, 0, 1).astype(np.float32)) if with_attrs else None
    return TriangleMesh(torch.from_numpy(v), torch.from_numpy(np.asarray(t, np.int32)), nrm, col)


def test_filter_mesh_components_follows_open3d_legacy_semantics(capsys):
    """filter_mesh_components (o3d_utils.py:241-321) on CPU tensors: edge-adjacency clusters, size threshold with the
    largest-cluster fallback, degenerate / duplicated triangle removal, vertex welding, non-manifold edges."""
    from mq3d_b200.meshops import cluster_connected_triangles as _cluster_connected_triangles
    from mq3d_b200.ops import filter_mesh_components
    # two tetrahedra that share ONE VERTEX only (same index): edge adjacency keeps them apart
    va, ta = _tetra((0, 0, 0))
    vb, tb = _tetra((0, 0, 0), scale=-1.0)
    v = np.concatenate([va, vb[1:]])                      # vertex 0 is shared
    tb2 = np.where(tb == 0, 0, tb + 3)
    t = np.concatenate([ta, tb2])
    label, counts = _cluster_connected_triangles(t.astype(np.int64))
    assert counts.tolist() == [4, 4] and label.tolist() == [0] * 4 + [1] * 4
    # a big component (two tetrahedra welded along a face would be non-manifold; use a strip) and a small one
    n = 40
    strip_v = np.array([[i // 2, i % 2, 0] for i in range(2 * n + 2)], np.float32)
    strip_t = np.array([[i, i + 1, i + 2] if i % 2 == 0 else [i + 1, i, i + 2] for i in range(2 * n)], np.int32)
    far_v, far_t = _tetra((100, 0, 0))
    v = np.concatenate([strip_v, far_v])
    t = np.concatenate([strip_t, far_t + len(strip_v)])
    out = filter_mesh_components(_mesh(v, t), min_triangle_count=10)
    ov, ot = out.vertex.positions.numpy(), out.triangle.indices.numpy()
    assert len(ot) == 2 * n and len(ov) == len(strip_v) and ov[:, 0].max() < 50          # tetrahedron dropped
    assert np.array_equal(out.vertex.normals.numpy(), -ov) and out.vertex.colors is not None   # attributes follow
    assert np.array_equal(ov[ot], strip_v[strip_t])
    # nothing reaches the threshold: the largest component survives alone
    out = filter_mesh_components(_mesh(v, t), min_triangle_count=10_000)
    assert len(out.triangle.indices) == 2 * n
    assert "Returning largest component only" in capsys.readouterr().out
    # degenerate, rotated-duplicate and opposite-winding triangles; duplicated vertices are welded
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0], [2, 0, 0], [2, 1, 0]], np.float32)  # 4 == 1
    t = np.array([[0, 1, 2], [1, 2, 0], [0, 2, 1], [2, 2, 3], [4, 5, 6]], np.int32)
    out = filter_mesh_components(_mesh(v, t, with_attrs=False), min_triangle_count=1)
    ot = out.triangle.indices.numpy().tolist()
    assert ot == [[0, 1, 2], [0, 2, 1], [1, 4, 5]]          # rotation removed, reversed kept, (4,5,6) -> (1,4,5)
    assert len(out.vertex.positions) == 6                   # welded; unreferenced vertex 3 stays (nothing was filtered)
    # three triangles on one edge: the smallest one goes
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, -3, 0], [0, 0, 0.5]], np.float32)
    t = np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4]], np.int32)
    out = filter_mesh_components(_mesh(v, t, with_attrs=False), min_triangle_count=1)
    assert out.triangle.indices.numpy().tolist() == [[0, 1, 2], [1, 0, 3]]
    # empty mesh is returned unchanged
    empty = _mesh(np.zeros((0, 3)), np.zeros((0, 3)), with_attrs=False)
    assert filter_mesh_components(empty) is empty


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) prints ONE JSON line with the
    contract's keys; under torchrun only rank 0 works, the other ranks exit 0 silently."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-frames", "3", "--workload", "quest300_v20mm"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm's config is the GPU arm's (same keys and values), the bounded sample is described separately
    assert d["config"]["frames"] == 300 and d["config"]["voxel_size"] == 0.02 and "3 of the 300 frames" in d["cpu_baseline"]["sample"]
    # a launcher's OMP_NUM_THREADS=1 (torch.distributed.run exports it) must not shrink the OpenMP team
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d1 = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d1["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) == d["cpu_baseline"]["cores"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_integrate_fragment_point_cloud_contract(monkeypatch, capsys):
    """integrate_fragment_point_cloud (refine_fragment_poses.py:14-58): forwards the fragment config to integrate(),
    returns (side, point cloud), and turns empty clouds / failures into None with the reference's messages."""
    import torch
    from types import SimpleNamespace
    from mq3d_b200 import ops
    from mq3d_b200.config import FragmentPoseRefinementConfig
    from mq3d_b200.geometry import PointCloud
    cfg = FragmentPoseRefinementConfig(device="CUDA:0", voxel_size=0.02, depth_max=2.5, trunc_voxel_multiplier=6.0)
    ds = SimpleNamespace(timestamps=np.array([10, 20, 30]))
    seen = {}

    class FakeGrid:
        def __init__(self, n):
            self.n = n

        def extract_point_cloud(self):
            return PointCloud(torch.zeros((self.n, 3)), torch.zeros((self.n, 3)))

    def fake_integrate(**kw):
        seen.update(kw)
        return FakeGrid(seen.pop("_points", 5))

    monkeypatch.setattr(ops, "integrate", fake_integrate)
    side, pcd = ops.integrate_fragment_point_cloud("io", ds, Side.RIGHT, cfg)
    assert side == Side.RIGHT and pcd.point.positions.shape == (5, 3)
    assert seen["voxel_size"] == 0.02 and seen["depth_max"] == 2.5 and seen["trunc_voxel_multiplier"] == 6.0
    assert seen["block_count"] == 50_000 and seen["device"] == "CUDA:0" and seen["vbg_opt"] is None
    assert seen["dataset"] is ds and seen["depth_data_io"] == "io" and seen["use_confidence_filtered_depth"] is True
    monkeypatch.setattr(ops, "integrate", lambda **kw: FakeGrid(0))
    assert ops.integrate_fragment_point_cloud("io", ds, Side.LEFT, cfg) is None
    assert "is empty (no valid points)" in capsys.readouterr().out

    def boom(**kw):
        raise RuntimeError("No block is touched in TSDF volume")

    monkeypatch.setattr(ops, "integrate", boom)
    assert ops.integrate_fragment_point_cloud("io", ds, Side.LEFT, cfg) is None
    out = capsys.readouterr().out
    assert "integrate_fragment_point_cloud failed for LEFT" in out and "10 - 30" in out


def test_raw_depth_streamer_ring(tmp_path):
    """ingest.RawDepthStreamer, host-only mode: ordered chunks, bounded number of ring slots in use,
    missing files flagged and zero-filled, truncated files raise in the consumer."""
    from mq3d_b200.ingest import RawDepthStreamer
    H, W, n = 6, 8, 37
    rng = np.random.default_rng(0)
    frames = rng.random((n, H, W)).astype("<f4")
    for i in range(n):
        if i != 11:
            frames[i].tofile(tmp_path / f"{i}.raw")
    s = RawDepthStreamer(lambda i: tmp_path / f"{i}.raw", n, H, W, chunk_frames=5, slots=2, device=None, workers=3)
    assert s.n_chunks == 8 and s.host_bytes == 2 * 5 * H * W * 4
    seen = 0
    for f0, f1, chunk, present in s:
        assert f0 == seen and f1 == min(n, f0 + 5) and chunk.shape == (f1 - f0, H, W)
        for i in range(f0, f1):
            if i == 11:
                assert not present[i - f0] and not chunk[i - f0].any()
            else:
                assert present[i - f0] and np.array_equal(chunk[i - f0], frames[i])
        seen = f1
    assert seen == n and 1 <= s.max_slots_in_use <= 2
    (tmp_path / "20.raw").write_bytes((tmp_path / "20.raw").read_bytes()[:40])
    with pytest.raises(RuntimeError, match="expected 48 float32"):
        for _ in RawDepthStreamer(lambda i: tmp_path / f"{i}.raw", n, H, W, chunk_frames=5, slots=3, device=None):
            pass
    assert list(RawDepthStreamer(lambda i: tmp_path / "x.raw", 0, H, W, device=None)) == []
