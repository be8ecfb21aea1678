"""The drop-in boundary, proven against the reference's OWN modules: with ``mq3d_b200.compat`` registered as
``open3d`` the reference's reconstruction modules import and run unmodified.

The reference sources are never copied into this repository.  They are looked up at /root/reference (the build
container) or at baseline/_ref (git-ignored staging area that tools/stage_reference.sh fills so that the GPU
half of this file can run on the GPU box); the tests skip when neither exists.
"""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _reference_scripts():
    for base in (Path("/root/reference"), ROOT / "baseline" / "_ref"):
        if (base / "scripts" / "processing" / "reconstruction" / "utils" / "o3d_utils.py").exists():
            return base
    return None


REF = _reference_scripts()
needs_reference = pytest.mark.skipif(REF is None, reason="reference sources not present (/root/reference or baseline/_ref)")

MODULES = ["config.reconstruction_config", "config.pipeline_configs", "dataio.reconstruction_data_io", "dataio.data_io",
           "processing.reconstruction.utils.o3d_utils",
           "processing.reconstruction.confidence_estimation.estimate_depth_confidences",
           "processing.reconstruction.depth_optimization.make_fragments",
           "processing.reconstruction.depth_optimization.refine_fragment_poses",
           "processing.reconstruction.reconstruct_scene", "pipeline.pipeline_processor"]


@pytest.fixture(scope="module")
def ref():
    """The reference's modules imported on top of the stand-in (module name -> module)."""
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import compat
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    sys.path.insert(0, str(REF / "scripts"))
    o3d = compat.install(force=True)
    mods = {"open3d": o3d}
    for m in MODULES:
        mods[m] = importlib.import_module(m)
    yield mods
    sys.path[:] = saved_path
    for k in list(sys.modules):
        if k not in saved_mods:
            del sys.modules[k]


@needs_reference
def test_reference_modules_import_on_the_stand_in(ref):
    """every module of the reconstruction stage imports with open3d := mq3d_b200.compat
    (annotations such as o3d.pipelines.registration.PoseGraph and o3d.t.pipelines.registration.* resolve)."""
    o3d = ref["open3d"]
    u = ref["processing.reconstruction.utils.o3d_utils"]
    assert u.o3d is o3d and callable(u.integrate) and callable(u.filter_mesh_components)
    # the reference's own YAML parses through its own parser; the device string becomes the stand-in's Device
    import yaml
    cfg = yaml.safe_load(open(REF / "config" / "pipeline_config.yml"))
    rc = ref["config.reconstruction_config"].ReconstructionConfig.parse(cfg["reconstruction"])
    assert isinstance(rc.device, o3d.core.Device) and str(rc.device) == "CPU:0"
    assert rc.depth_integration.voxel_size == 0.01 and rc.depth_integration.device == rc.device
    assert rc.color_optimization.min_triangle_count == 5000
    # out-of-scope names resolve (annotations, attribute chains) but raise the moment they are called
    with pytest.raises(RuntimeError, match="outside the B200 hot-path build"):
        _ = rc.fragment_pose_refinement.icp_criteria_list
    with pytest.raises(RuntimeError, match="outside the B200 hot-path build"):
        o3d.pipelines.color_map.run_rigid_optimizer(None, None, None, None)
    # a CPU device is representable but no compute object accepts it (north star: no CPU fallback)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        o3d.t.geometry.VoxelBlockGrid(attr_names=("tsdf", "weight"), attr_dtypes=(o3d.core.float32, o3d.core.float32),
                                      attr_channels=((1), (1)), voxel_size=0.01, block_resolution=16, block_count=100,
                                      device=rc.device)


@needs_reference
def test_reference_pose_converters_round_trip(ref):
    """o3d_utils.convert_transforms_to_pose_graph / convert_pose_graph_to_transforms and the trajectory pair
    (:22-106) run on the stand-in's value containers."""
    u = ref["processing.reconstruction.utils.o3d_utils"]
    tr_mod = importlib.import_module("models.transforms")
    rng = np.random.default_rng(3)
    from scipy.spatial.transform import Rotation
    pos = rng.normal(size=(5, 3))
    quat = Rotation.random(5, random_state=1).as_quat()
    tr = tr_mod.Transforms(coordinate_system=tr_mod.CoordinateSystem.OPEN3D, positions=pos, rotations=quat)
    pg = u.convert_transforms_to_pose_graph(tr)
    assert len(pg.nodes) == 5 and pg.nodes[0].pose.shape == (4, 4)
    back = u.convert_pose_graph_to_transforms(pg)
    assert np.allclose(back.positions, pos, atol=1e-6)
    assert np.allclose(np.abs(np.sum(back.rotations * quat, axis=1)), 1.0, atol=1e-6)


@needs_reference
def test_reference_filter_mesh_components_runs_on_the_stand_in(ref, capsys):
    """The reference's own filter_mesh_components (:241-321) -- to_legacy, cluster_connected_triangles,
    remove_triangles_by_mask, remove_unreferenced_vertices, the four clean-up calls, from_legacy, to(device) --
    runs on the stand-in's legacy mesh and agrees with this package's implementation."""
    import torch
    from mq3d_b200 import ops
    from mq3d_b200.geometry import TriangleMesh as RawMesh
    o3d = ref["open3d"]
    u = ref["processing.reconstruction.utils.o3d_utils"]
    # two grids of quads (a large and a small component), one duplicated triangle, one degenerate triangle
    def grid(nx, ny, off):
        xs, ys = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
        v = np.stack([xs.ravel() + off, ys.ravel(), np.zeros(xs.size)], axis=1).astype(np.float32)
        idx = lambda i, j: i * (ny + 1) + j
        t = []
        for i in range(nx):
            for j in range(ny):
                t += [[idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)], [idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)]]
        return v, np.array(t, np.int32)
    v1, t1 = grid(6, 5, 0.0)
    v2, t2 = grid(2, 1, 100.0)
    v = np.concatenate([v1, v2])
    t = np.concatenate([t1, t2 + len(v1), t1[:1][:, [1, 2, 0]], [[3, 3, 4]]]).astype(np.int32)
    mesh = o3d.t.geometry.TriangleMesh(torch.from_numpy(v), torch.from_numpy(t))
    out = u.filter_mesh_components(mesh, min_triangle_count=10)
    text = capsys.readouterr().out
    assert "Removed 1 small component(s)" in text or "Removed 2 small component(s)" in text
    mine = ops.filter_mesh_components(RawMesh(torch.from_numpy(v), torch.from_numpy(t)), min_triangle_count=10, backend="host")
    a = out.to_legacy()
    b = mine.to_legacy()
    assert len(a.triangles) == len(b.triangles) == len(t1) and np.array_equal(a.triangles, b.triangles)
    assert np.array_equal(a.vertices, b.vertices) and str(out.device) == "CPU:0"


# ---------------------------------------------------------------------------------------------------------
# GPU: the reference's own integrate(), raycast_in_color_view() and stage driver, unmodified
# ---------------------------------------------------------------------------------------------------------
@needs_reference
@pytest.mark.gpu
def test_reference_integrate_and_raycast_unmodified(ref, cuda_device, oracle, tmp_path):
    """o3d_utils.integrate (:153-238: per-frame load_depth_map -> compute_unique_block_coordinates -> integrate, with
    the reference's own DepthDataIO reading a synthetic QRC project) builds the same grid as the oracle's frame
    loop, bit for bit; raycast_in_color_view (:324-342) on the filtered mesh gives the oracle's z-depths."""
    import torch
    from helpers import pipeline_cameras, sort_blocks
    from mq3d_b200 import synth
    from mq3d_b200.models import Side
    o3d = ref["open3d"]
    u = ref["processing.reconstruction.utils.o3d_utils"]
    DataIO = ref["dataio.data_io"].DataIO
    side_mod = importlib.import_module("models.side")
    tr_mod = importlib.import_module("models.transforms")
    caps = synth.write_project(tmp_path, 10, sides=(Side.LEFT,), width=160, height=160)
    data_io = DataIO(project_dir=tmp_path)
    ds = data_io.depth.load_depth_dataset(side=side_mod.Side.LEFT, use_cache=False)
    ds.transforms = ds.transforms.convert_coordinate_system(target_coordinate_system=tr_mod.CoordinateSystem.OPEN3D,
                                                            is_camera=True)
    dev = o3d.core.Device("CUDA:0")
    vbg = u.integrate(dataset=ds, depth_data_io=data_io.depth, side=side_mod.Side.LEFT, use_confidence_filtered_depth=False,
                      confidence_threshold=0.0, valid_count_threshold=0, voxel_size=0.02, block_resolution=16,
                      block_count=1000, depth_max=4.0, trunc_voxel_multiplier=10.0, device=dev, show_progress=False)
    # oracle frame loop on the same files
    cap = caps[Side.LEFT]
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    og = oracle.Grid(0.02)
    for i in range(10):
        d = oracle.depth_to_linear(cap.raw[i], cap.dataset.nears[i], cap.dataset.fars[i])
        og.integrate(og.touch(d, K[i], Ewc[i], 4.0, 10.0), d, K[i], Ewc[i], 4.0, 10.0)
    ok, ot, ow, _ = og.export()
    k0, t0, w0 = sort_blocks(ok, ot, ow)
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    # VoxelBlockGrid.save / load through the reference's own ReconstructionDataIO (A.6 layout)
    data_io.reconstruction.save_colorless_vbg(vbg=vbg)
    again = data_io.reconstruction.load_colorless_vbg()
    k2, t2, w2 = sort_blocks(*[x.cpu().numpy() for x in again.export_blocks()[:3]])
    assert np.array_equal(k1, k2) and np.array_equal(t1.view(np.uint32), t2.view(np.uint32))
    # mesh -> the reference's own filter -> scene -> the reference's own raycast loop
    mesh = vbg.extract_triangle_mesh(weight_threshold=1.5, estimated_vertex_number=-1)
    mesh = u.filter_mesh_components(mesh, min_triangle_count=50)
    scene = o3d.t.geometry.RaycastingScene(device=dev)
    scene.add_triangles(mesh.cpu())
    cam_mod = importlib.import_module("models.camera_dataset")
    n, W, H = 2, 96, 72
    cds = cam_mod.CameraDataset(directory_relative_path="left_camera_rgb", image_file_names=np.array(["a.png", "b.png"]),
                                timestamps=ds.timestamps[:n], fx=np.full(n, 60), fy=np.full(n, 60), cx=np.full(n, W // 2),
                                cy=np.full(n, H // 2),
                                transforms=tr_mod.Transforms(coordinate_system=ds.transforms.coordinate_system,
                                                             positions=ds.transforms.positions[:n],
                                                             rotations=ds.transforms.rotations[:n]),
                                widths=np.full(n, W), heights=np.full(n, H))
    depths = list(u.raycast_in_color_view(scene=scene, dataset=cds))
    assert len(depths) == n and depths[0].shape == (H, W) and depths[0].dtype == np.float32
    lm = mesh.to_legacy()
    Kc = u.compute_o3d_intrinsic_matrices(cds)
    for i in range(n):
        rays = oracle.create_rays_pinhole(Kc[i].astype(np.float64), cds.transforms.extrinsics_wc[i].astype(np.float64), W, H)
        want = oracle.cast_rays(lm.vertices.astype(np.float32), lm.triangles, rays.reshape(-1, 6)).reshape(H, W)
        got = depths[i]
        both = np.isfinite(want) & np.isfinite(got)
        assert (np.isfinite(want) != np.isfinite(got)).mean() < 2e-3 and both.sum() > 0.3 * H * W
        assert np.quantile(np.abs(want[both] - got[both]), 0.999) < 1e-4


@needs_reference
@pytest.mark.gpu
def test_reference_fragment_pose_graph_unmodified(ref, cuda_device, oracle, tmp_path):
    """depth_optimization.make_fragments.build_pose_graph_for_fragment (:83-240) -- load_depth_map on the device,
    o3d.t.pipelines.odometry.compute_odometry_information_matrix per consecutive pair and per overlapping key-frame
    pair, PoseGraph / PoseGraphNode / PoseGraphEdge -- runs unmodified on the stand-in; every edge's
    information matrix equals the oracle's.  The pose-graph optimisation that would follow is out of scope."""
    from helpers import pipeline_cameras
    from mq3d_b200 import synth
    from mq3d_b200.models import Side
    o3d = ref["open3d"]
    mf = ref["processing.reconstruction.depth_optimization.make_fragments"]
    cfg_mod = ref["config.reconstruction_config"]
    side_mod = importlib.import_module("models.side")
    tr_mod = importlib.import_module("models.transforms")
    caps = synth.write_project(tmp_path, 12, sides=(Side.LEFT,), width=160, height=160)
    data_io = ref["dataio.data_io"].DataIO(project_dir=tmp_path)
    ds = data_io.depth.load_depth_dataset(side=side_mod.Side.LEFT, use_cache=False)
    ds.transforms = ds.transforms.convert_coordinate_system(target_coordinate_system=tr_mod.CoordinateSystem.OPEN3D,
                                                            is_camera=True)
    cfg = cfg_mod.FragmentGenerationConfig(device=o3d.core.Device("CUDA:0"))
    cfg.use_confidence_filtered_depth = False
    cfg.depth_max, cfg.odometry_loop_interval, cfg.dist_threshold = 4.0, 5, 0.07
    pg = mf.build_pose_graph_for_fragment(frag_dataset=ds, depth_data_io=data_io.depth, side=side_mod.Side.LEFT, config=cfg)
    assert len(pg.nodes) == 12 and len(pg.edges) >= 11
    cap = caps[Side.LEFT]
    K, Ewc, Ecw = pipeline_cameras(cap.dataset)
    lin = [oracle.depth_to_linear(cap.raw[i], cap.dataset.nears[i], cap.dataset.fars[i]) for i in range(12)]
    for e in pg.edges:
        i, j = e.source_node_id, e.target_node_id
        want = oracle.odometry_information(lin[i], lin[j], K[0], np.asarray(e.transformation, np.float64), 0.07, 1.0, 4.0)
        assert e.information.shape == (6, 6) and np.allclose(e.information, want, rtol=1e-9, atol=1e-6), (i, j)
        assert e.uncertain == (j != i + 1)
    with pytest.raises(RuntimeError, match="outside the B200 hot-path build"):
        mf.optimize_dataset_pose(data_io.depth, ds, side_mod.Side.LEFT, cfg)


@needs_reference
@pytest.mark.gpu
def test_reference_stage_driver_unmodified(ref, cuda_device, tmp_path, capsys):
    """processing.reconstruction.reconstruct_scene.reconstruct_scene -- the reference's stage driver -- runs end to
    end on the stand-in: datasets, integrate LEFT then RIGHT into one grid, colorless_vbg.npz, colourless point
    cloud PLY.  Same grid as this package's own driver on the same project."""
    from mq3d_b200 import synth
    from mq3d_b200.models import Side
    from helpers import sort_blocks
    o3d = ref["open3d"]
    synth.write_project(tmp_path, 8, sides=(Side.LEFT, Side.RIGHT), width=160, height=160)
    cfg_mod = ref["config.reconstruction_config"]
    rc = cfg_mod.ReconstructionConfig.parse({
        "device": "CUDA:0", "use_dataset_cache": True, "estimate_depth_confidences": False, "optimize_depth_pose": False,
        "use_colorless_vbg_cache": False, "visualize_colorless_pcd": False, "optimize_color_pose": False,
        "visualize_colored_mesh": False, "sample_point_cloud_from_colored_mesh": False, "render_color_aligned_depth": False,
        "depth_integration": {"use_confidence_filtered_depth": False, "voxel_size": 0.02, "block_resolution": 16,
                              "block_count": 2000, "depth_max": 4.0, "trunc_voxel_multiplier": 10.0}})
    assert str(rc.device) == "CUDA:0" and str(rc.depth_integration.device) == "CUDA:0"
    data_io = ref["dataio.data_io"].DataIO(project_dir=tmp_path)
    ref["processing.reconstruction.reconstruct_scene"].reconstruct_scene(data_io=data_io, config=rc)
    pc = data_io.path_config.reconstruction
    npz = [p for p in Path(tmp_path).rglob("*.npz") if "vbg" in p.name]
    ply = [p for p in Path(tmp_path).rglob("*.ply")]
    assert len(npz) == 1 and len(ply) >= 1, (npz, ply, pc)
    g = o3d.t.geometry.VoxelBlockGrid.load(str(npz[0]))
    # this package's own driver on the same project gives the same grid
    from mq3d_b200.config import ReconstructionConfig
    from mq3d_b200.dataio import DataIO as MyDataIO
    from mq3d_b200.reconstruct import reconstruct_scene as my_reconstruct
    other = tmp_path / "mine"
    synth.write_project(other, 8, sides=(Side.LEFT, Side.RIGHT), width=160, height=160)
    mc = ReconstructionConfig.parse({
        "device": "CUDA:0", "use_dataset_cache": True, "estimate_depth_confidences": False, "optimize_depth_pose": False,
        "use_colorless_vbg_cache": False, "optimize_color_pose": False, "render_color_aligned_depth": False,
        "depth_integration": {"use_confidence_filtered_depth": False, "voxel_size": 0.02, "block_resolution": 16,
                              "block_count": 2000, "depth_max": 4.0, "trunc_voxel_multiplier": 10.0}})
    my_reconstruct(MyDataIO(other), mc)
    from mq3d_b200.vbg import VoxelBlockGrid
    mine = VoxelBlockGrid.load(str(next(Path(other).rglob("*vbg*.npz"))), device=cuda_device)
    a = sort_blocks(*[x.cpu().numpy() for x in g.export_blocks()[:3]])
    b = sort_blocks(*[x.cpu().numpy() for x in mine.export_blocks()[:3]])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
