"""GPU: the reference-facing surface end to end -- DataIO project on disk -> reconstruct_scene()
(confidence -> integrate LEFT then RIGHT -> save -> extract -> raycast), checked against the same
pipeline assembled from the CPU oracle; and the Open3D-shaped compat layer driving the per-frame loop
exactly as the reference's o3d_utils.integrate does."""
import numpy as np
import pytest
import torch

from helpers import pipeline_cameras, sort_blocks

pytestmark = pytest.mark.gpu

N = 10
W = H = 96


def _config():
    from mq3d_b200.config import ReconstructionConfig
    return ReconstructionConfig.parse({
        "device": "CUDA:0", "use_dataset_cache": False, "estimate_depth_confidences": True,
        "optimize_depth_pose": False, "optimize_color_pose": False, "use_colorless_vbg_cache": False,
        "render_color_aligned_depth": True,
        "confidence_estimation": {"target_frame_range": 3, "depth_max": 4.0, "error_threshold": 0.08,
                                  "skip_if_output_dir_exists": False},
        "depth_integration": {"use_confidence_filtered_depth": True, "confidence_threshold": 0.02,
                              "valid_count_threshold": 2, "voxel_size": 0.03, "block_count": 64, "depth_max": 4.0,
                              "trunc_voxel_multiplier": 10.0},
        "color_optimization": {"weight_threshold": 1.5, "min_triangle_count": 50},
        "color_aligned_depth_rendering": {"only_use_optimized_dataset": False},
    })


def test_reconstruct_scene_matches_oracle_pipeline(cuda_device, oracle, tmp_path):
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import synth
    from mq3d_b200.dataio import DataIO
    from mq3d_b200.io_utils import read_ply
    from mq3d_b200.models import CoordinateSystem, Side
    from mq3d_b200.reconstruct import reconstruct_scene
    from mq3d_b200.vbg import VoxelBlockGrid

    caps = synth.write_project(tmp_path, N, width=W, height=H)
    # one corrupt frame: all ones -> is_depth_map_valid rejects it -> dropped from the dataset
    bad_ts = int(caps[Side.LEFT].dataset.timestamps[4])
    np.ones((H, W), "<f4").tofile(tmp_path / "left_depth" / f"{bad_ts}.raw")
    data_io = DataIO(tmp_path)
    cds = synth.make_color_dataset(3, Side.LEFT)
    cds.widths[:], cds.heights[:] = 128, 96
    cds.fx[:], cds.fy[:], cds.cx[:], cds.cy[:] = 87, 87, 64, 48
    cds.transforms = cds.transforms.convert_coordinate_system(CoordinateSystem.OPEN3D, is_camera=True)
    data_io.color.save_color_dataset(Side.LEFT, cds)

    cfg = _config()
    report = reconstruct_scene(data_io, cfg)
    assert report["active_blocks"] > 50 and report["raycast_frames"] == 3

    # ---- the same pipeline from oracle parts ------------------------------------------------------
    og = oracle.Grid(0.03)
    for side in Side:
        ds = data_io.depth.load_depth_dataset(side)
        keep = [i for i in range(N) if not (side == Side.LEFT and i == 4)]
        assert len(ds) == len(keep)
        assert np.array_equal(ds.timestamps, caps[side].dataset.timestamps[keep])
        raw = caps[side].raw[keep]
        lin = np.stack([oracle.depth_to_linear(raw[i], ds.nears[i], ds.fars[i]) for i in range(len(ds))])
        K, Ewc, Ecw = pipeline_cameras(ds) if ds.transforms.coordinate_system == CoordinateSystem.UNITY else (None,) * 3
        if K is None:      # reconstruct_scene converted the cached dataset object in place
            K = ds.get_intrinsic_matrices()
            K[:, 0, 2] = ds.widths - K[:, 0, 2]
            Ewc, Ecw = ds.transforms.extrinsics_wc, ds.transforms.extrinsics_cw
        conf, count = oracle.confidence(lin, K, Ecw, np.linalg.inv(Ecw), 3, 4.0, 0.08)
        for i in range(len(ds)):
            cm = data_io.depth.load_confidence_map(side, ds.timestamps[i])       # files written by the GPU stage
            assert np.array_equal(cm.valid_count, count[i]) and np.array_equal(cm.confidence_map, conf[i])
            d = oracle.depth_mask(lin[i], conf[i], count[i], 0.02, 2)
            keys = og.touch(d, K[i], Ewc[i], 4.0, 10.0)
            og.integrate(keys, d, K[i], Ewc[i], 4.0, 10.0)
    saved = VoxelBlockGrid.load(str(tmp_path / "reconstruction" / "colorless_vbg.npz"), device=cuda_device)
    k0, t0, w0 = sort_blocks(*og.export()[:3])
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in saved.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    # artefacts
    pts, _ = read_ply(tmp_path / "reconstruction" / "colorless.ply")
    assert len(pts) == len(og.extract_points(3.0)[0]) == report["points"]
    mv, mt = read_ply(tmp_path / "reconstruction" / "colorless_mesh_raw.ply")
    ov, _, ot, _ = og.extract_mesh(1.5)
    assert len(mv) == len(ov) and len(mt) == len(ot)
    cv, ct = read_ply(tmp_path / "reconstruction" / "colorless_mesh_clean.ply")
    assert 0 < len(ct) <= len(mt)
    d0 = np.load(tmp_path / "left_color_aligned_depth" / f"{int(cds.timestamps[0])}.npy")
    assert d0.shape == (96, 128) and d0.dtype == np.float32 and np.isfinite(d0).mean() > 0.3
    assert 0.2 < np.median(d0[np.isfinite(d0)]) < 4.0


def test_compat_layer_runs_the_reference_frame_loop(cuda_device, oracle):
    """Body of o3d_utils.integrate (:171-236) written against `o3d` = mq3d_b200.compat."""
    import mq3d_b200.compat as o3d
    from helpers import capture
    cap = capture(6)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    device = o3d.core.Device("CUDA:0")
    vbg = o3d.t.geometry.VoxelBlockGrid(
        attr_names=("tsdf", "weight"), attr_dtypes=(o3d.core.float32, o3d.core.float32), attr_channels=((1), (1)),
        voxel_size=0.02, block_resolution=16, block_count=50, device=device)
    og = oracle.Grid(0.02)
    for index in range(len(ds)):
        depth_np = oracle.depth_to_linear(cap.raw[index], ds.nears[index], ds.fars[index])
        depth_map = o3d.t.geometry.Image(tensor=o3d.core.Tensor(depth_np, dtype=o3d.core.Dtype.Float32, device=device))
        intrinsic = o3d.core.Tensor(K[index], dtype=o3d.core.Dtype.Float64)
        extrinsic = o3d.core.Tensor(Ewc[index], dtype=o3d.core.Dtype.Float64)
        coords = vbg.compute_unique_block_coordinates(depth=depth_map, intrinsic=intrinsic, extrinsic=extrinsic,
                                                      depth_scale=1.0, depth_max=4.0, trunc_voxel_multiplier=10.0)
        vbg.integrate(block_coords=coords, depth=depth_map, intrinsic=intrinsic, extrinsic=extrinsic,
                      depth_scale=1.0, depth_max=4.0, trunc_voxel_multiplier=10.0)
        keys = og.touch(depth_np, K[index], Ewc[index], 4.0, 10.0)
        og.integrate(keys, depth_np, K[index], Ewc[index], 4.0, 10.0)
    k0, t0, w0 = sort_blocks(*og.export()[:3])
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    mesh = vbg.extract_triangle_mesh(weight_threshold=1.5, estimated_vertex_number=-1)
    legacy = mesh.to_legacy()
    assert len(legacy.triangles) == len(og.extract_mesh(1.5)[2]) and legacy.vertices.dtype == np.float64
    pcd = vbg.extract_point_cloud().to_legacy()
    assert len(pcd.points) == len(og.extract_points(3.0)[0])
    scene = o3d.t.geometry.RaycastingScene(device=device)
    scene.add_triangles(mesh.cpu())
    rays = scene.create_rays_pinhole(o3d.core.Tensor(K[0], dtype=o3d.core.Dtype.Float32),
                                     o3d.core.Tensor(Ewc[0], dtype=o3d.core.Dtype.Float32), width_px=80, height_px=80)
    depth = scene.cast_rays(rays)["t_hit"].cpu().numpy()
    assert depth.shape == (80, 80)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        o3d.t.geometry.VoxelBlockGrid(voxel_size=0.02, block_count=10, device=o3d.core.Device("CPU:0"))


@pytest.mark.parametrize("gated", [True, False])
def test_integrate_frames_host_pipeline_matches_direct(cuda_device, oracle, gated):
    """The chunked H2D/compute pipeline (public host-buffer API used by bench.py's e2e number) gives the
    same grid as one fused call on device-resident frames -- both as one call per chunk and as ONE call whose
    batches are gated on the device by the copy stream's events (upload + K1 + colour resampling per chunk)."""
    import mq3d_b200  # noqa: F401
    from helpers import capture
    from mq3d_b200 import synth
    from mq3d_b200.pipeline import IntegrationParams, extract_mesh_to_host, integrate_frames, pin
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare
    n = 11
    cap = capture(n)
    K, Ewc, Ecw = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    colors = np.stack([synth.make_color_frame(Ecw[i], width=160, height=120, f=110.0) for i in range(n)])
    Kc = np.tile(np.array([[110.0, 0, 80.0], [0, 110.0, 60.0], [0, 0, 1.0]]), (n, 1, 1))
    rng = np.random.default_rng(3)
    conf = rng.random(cap.raw.shape)
    count = rng.integers(0, 6, cap.raw.shape).astype(np.int32)
    names = ("tsdf", "weight", "color")
    a = VoxelBlockGrid(attr_names=names, voxel_size=0.02, block_count=300, device=cuda_device)
    p = IntegrationParams(voxel_size=0.02, depth_max=4.0, trunc_voxel_multiplier=10.0, confidence_threshold=0.3,
                          valid_count_threshold=2, batch_frames=4)
    st = integrate_frames(a, pin(cap.raw), ds.nears, ds.fars, K, Ewc, p, conf=pin(conf), count=pin(count),
                          colors_host=pin(colors), Kc=Kc, gated=gated)
    b = VoxelBlockGrid(attr_names=names, voxel_size=0.02, block_count=300, device=cuda_device)
    lin, valid = depth_prepare(torch.from_numpy(cap.raw).to(cuda_device), ds.nears, ds.fars,
                               torch.from_numpy(conf).to(cuda_device), torch.from_numpy(count).to(cuda_device), None, 0.3, 2)
    st2 = b.integrate_sequence(lin, K, Ewc, 4.0, 10.0, frame_valid=valid, colors=torch.from_numpy(colors).to(cuda_device),
                               color_intrinsics=Kc, batch_frames=64)
    assert (st.frames_integrated, st.block_visits, st.voxel_updates, st.num_blocks, st.batches) == \
        (st2.frames_integrated, st2.block_visits, st2.voxel_updates, st2.num_blocks, 3)
    for x, y in zip(sort_blocks(*[t.cpu().numpy() for t in a.export_blocks()]),
                    sort_blocks(*[t.cpu().numpy() for t in b.export_blocks()])):
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, y.view(np.uint32) if y.dtype == np.float32 else y)
    v, nrm, t = extract_mesh_to_host(a, 1.5)
    assert v.dtype == np.float32 and t.dtype == np.int32 and len(t) > 100


def test_streaming_ingest_equals_cached_ingest(cuda_device, oracle, tmp_path):
    """integrate() fed by the bounded pinned ring (ingest.RawDepthStreamer: file reads -> pinned slots
    -> two device buffers, chunk by chunk) builds the grid that the whole-side buffer builds, bit for bit, with a
    missing and a corrupt file in the sequence; build_depth_dataset validates through the same ring."""
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import synth
    from mq3d_b200.dataio import DataIO
    from mq3d_b200.ingest import stream_side
    from mq3d_b200.models import CoordinateSystem, Side
    from mq3d_b200.ops import integrate
    caps = synth.write_project(tmp_path, 23, sides=(Side.LEFT,), width=W, height=H)
    ts = caps[Side.LEFT].dataset.timestamps
    np.ones((H, W), "<f4").tofile(tmp_path / "left_depth" / f"{int(ts[5])}.raw")     # invalid: dropped by the dataset build
    io = DataIO(tmp_path)
    ds = io.depth.load_depth_dataset(Side.LEFT, use_cache=False)
    assert len(ds) == 22 and io.depth.has_raw_cache(Side.LEFT, ds)
    (tmp_path / "left_depth" / f"{int(ts[9])}.raw").unlink()                            # vanishes after the build
    ds.transforms = ds.transforms.convert_coordinate_system(CoordinateSystem.OPEN3D, is_camera=True)
    kw = dict(dataset=ds, depth_data_io=io.depth, side=Side.LEFT, use_confidence_filtered_depth=False,
              confidence_threshold=0.0, valid_count_threshold=0, voxel_size=0.03, block_resolution=16, block_count=64,
              depth_max=4.0, trunc_voxel_multiplier=10.0, device="CUDA:0")
    io.depth._raw_cache.clear()
    a = integrate(**kw, batch_frames=4, streaming=True)            # 6 chunks through a 3-slot ring
    b = integrate(**kw, batch_frames=4, streaming=False)           # whole side in one pinned buffer
    ka, ta, wa = sort_blocks(*[x.cpu().numpy() for x in a.export_blocks()[:3]])
    kb, tb, wb = sort_blocks(*[x.cpu().numpy() for x in b.export_blocks()[:3]])
    assert len(ka) > 50 and np.array_equal(ka, kb) and np.array_equal(wa, wb)
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32))
    assert wa.max() == 21                                           # 22 dataset frames, one file missing
    s = stream_side(io.depth, Side.LEFT, ds, chunk_frames=4, device=cuda_device)
    assert s.host_bytes == 3 * 4 * H * W * 4
    n = sum(f1 - f0 for f0, f1, _, _ in s)
    assert n == 22 and s.max_slots_in_use <= 3


def test_reconstruct_scene_with_colour(cuda_device, oracle, tmp_path):
    """depth_integration.integrate_color (extension key): the stage driver builds a coloured grid -- every depth frame
    takes the colour frame of the same eye nearest in time -- and writes color_mesh.ply; grid and vertex colours equal
    the oracle's colour branch."""
    import cv2
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import synth
    from mq3d_b200.config import ReconstructionConfig
    from mq3d_b200.dataio import DataIO
    from mq3d_b200.io_utils import read_ply
    from mq3d_b200.models import CoordinateSystem, Side
    from mq3d_b200.reconstruct import reconstruct_scene
    from mq3d_b200.vbg import VoxelBlockGrid
    n, cw, ch, f = 8, 160, 120, 110.0
    caps = synth.write_project(tmp_path, n, sides=(Side.LEFT,), width=W, height=H)
    (tmp_path / "right_depth_descriptors.csv").write_text((tmp_path / "left_depth_descriptors.csv").read_text().splitlines()[0] + "\n")
    (tmp_path / "right_depth").mkdir()
    io = DataIO(tmp_path)
    cap = caps[Side.LEFT]
    K, Ewc, Ecw = pipeline_cameras(cap.dataset)
    cds = synth.make_color_dataset(n, Side.LEFT)
    cds.widths[:], cds.heights[:] = cw, ch
    cds.fx[:], cds.fy[:], cds.cx[:], cds.cy[:] = int(f), int(f), cw // 2, ch // 2
    cds.timestamps[:] = cap.dataset.timestamps + 3                  # 3 ms later than the depth frames: nearest wins
    io.color.save_color_dataset(Side.LEFT, cds)
    (tmp_path / "left_camera_rgb").mkdir()
    cols = [synth.make_color_frame(Ecw[i], width=cw, height=ch, f=f) for i in range(n)]
    for i in range(n):
        cv2.imwrite(str(tmp_path / "left_camera_rgb" / f"{int(cds.timestamps[i])}.png"), cv2.cvtColor(cols[i], cv2.COLOR_RGB2BGR))
    cfg = ReconstructionConfig.parse({
        "device": "CUDA:0", "use_dataset_cache": False, "estimate_depth_confidences": False, "optimize_depth_pose": False,
        "optimize_color_pose": False, "use_colorless_vbg_cache": False, "render_color_aligned_depth": False,
        "depth_integration": {"use_confidence_filtered_depth": False, "voxel_size": 0.03, "block_count": 500, "depth_max": 4.0,
                              "trunc_voxel_multiplier": 10.0, "integrate_color": True, "batch_frames": 3},
        "color_optimization": {"weight_threshold": 1.5, "min_triangle_count": 1}})
    report = reconstruct_scene(io, cfg)
    assert report["mesh_colored"][0] > 500
    og = oracle.Grid(0.03, with_color=True)
    Kc = np.array([[int(f), 0, cw // 2], [0, int(f), ch // 2], [0, 0, 1.0]])
    ds = cap.dataset
    for i in range(n):
        d = oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i])
        og.integrate(og.touch(d, K[i], Ewc[i], 4.0, 10.0), d, K[i], Ewc[i], 4.0, 10.0, color=cols[i], Kc=Kc)
    g = VoxelBlockGrid.load(str(tmp_path / "reconstruction" / "colorless_vbg.npz"), device=cuda_device)
    assert g.has_color
    ok, ot, ow, oc = og.export()
    a = sort_blocks(ok, ot, ow, oc)
    b = sort_blocks(*[x.cpu().numpy() for x in g.export_blocks()])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[3].view(np.uint32), b[3].reshape(a[3].shape).view(np.uint32))
    v, _ = read_ply(tmp_path / "reconstruction" / "color_mesh.ply")
    assert "red" in v.dtype.names and len(v) == report["mesh_colored"][0] and v["red"].max() > 100
