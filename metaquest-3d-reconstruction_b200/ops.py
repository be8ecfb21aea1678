"""Stage helpers with the reference's names, arguments and error behaviour
(processing/reconstruction/utils/o3d_utils.py: compute_o3d_intrinsic_matrices :14-19, load_depth_map
:109-150, integrate :153-238, filter_mesh_components :241-321, raycast_in_color_view :324-342 and
confidence_estimation/estimate_depth_confidences.py: estimate_depth_confidences :120-154).

``integrate`` keeps the reference signature but runs the whole frame loop as one fused device call
(K1 -> K2 -> K3); per-frame Open3D-shaped calls remain available on VoxelBlockGrid for callers that
drive the loop themselves.
"""
from __future__ import annotations

from typing import Generator, Optional

import numpy as np
import torch

from .geometry import TriangleMesh
from .models import CameraDataset, ConfidenceMap, CoordinateSystem, DepthDataset, Side
from .raycast import RaycastingScene
from .vbg import VoxelBlockGrid, _device_index, depth_prepare, estimate_confidence


def compute_o3d_intrinsic_matrices(dataset: CameraDataset) -> np.ndarray:
    """float32 [N,3,3] with the principal point mirrored: cx' = width - cx (o3d_utils.py:14-19)."""
    k = dataset.get_intrinsic_matrices()
    k[:, 0, 2] = dataset.widths - k[:, 0, 2]
    return k


def _torch_device(device) -> torch.device:
    return torch.device("cuda", _device_index(device))


def _load_confidence_stack(depth_data_io, side, dataset, f0, f1, H, W):
    """(conf float64 [n,H,W], count int32 [n,H,W], has uint8 [n]) of frames f0..f1 from the per-frame npz files; a
    missing map leaves the frame unfiltered with a warning (o3d_utils.py:137-139)."""
    n = f1 - f0
    conf = np.zeros((n, H, W), np.float64)
    count = np.zeros((n, H, W), np.int32)
    has = np.zeros(n, np.uint8)
    for i in range(n):
        cm = depth_data_io.load_confidence_map(side=side, timestamp=dataset.timestamps[f0 + i])
        if cm is None:
            print(f"[Warning] Confidence map not found for timestamp {dataset.timestamps[f0 + i]}")
            continue
        conf[i], count[i], has[i] = cm.confidence_map, cm.valid_count, 1
    return conf, count, has


def load_depth_map(depth_data_io, side: Side, index: int, dataset: DepthDataset, device,
                   use_confidence_filtered_depth: bool, confidence_threshold: float,
                   valid_count_threshold: int) -> Optional[torch.Tensor]:
    """One linear, confidence-masked depth frame as a float32 [H,W] CUDA tensor, or None when the raw
    file is missing / invalid (o3d_utils.py:109-150)."""
    dev = _torch_device(device)
    W, H = int(dataset.widths[index]), int(dataset.heights[index])
    raw = depth_data_io.load_raw_depth_map(side, dataset.timestamps[index], W, H)
    if raw is None:
        return None
    conf = count = None
    if use_confidence_filtered_depth:
        cm = depth_data_io.load_confidence_map(side=side, timestamp=dataset.timestamps[index])
        if cm is None:
            print(f"[Warning] Confidence map not found for timestamp {dataset.timestamps[index]}")
        else:
            conf = torch.from_numpy(np.ascontiguousarray(cm.confidence_map, np.float64))[None]
            count = torch.from_numpy(np.ascontiguousarray(cm.valid_count, np.int32))[None]
    lin, valid = depth_prepare(torch.from_numpy(np.ascontiguousarray(raw))[None].to(dev),
                               [dataset.nears[index]], [dataset.fars[index]], conf, count, None,
                               confidence_threshold, valid_count_threshold)
    return lin[0] if int(valid[0]) else None


def integrate(dataset: DepthDataset, depth_data_io, side: Side, use_confidence_filtered_depth: bool,
              confidence_threshold: float, valid_count_threshold: int, voxel_size: float, block_resolution: int,
              block_count: int, depth_max: float, trunc_voxel_multiplier: float, device, show_progress: bool = False,
              desc: Optional[str] = None, vbg_opt: Optional[VoxelBlockGrid] = None,
              confidence: Optional[tuple] = None, batch_frames: int = 256, streaming: Optional[bool] = None,
              colors=None, color_intrinsics=None) -> VoxelBlockGrid:
    """Drop-in for o3d_utils.integrate (:153-238).  `dataset.transforms` must already be in the OPEN3D
    convention (reconstruct_scene.py:48-51).  `confidence` optionally passes device-resident
    (conf float64 [N,H,W], count int32 [N,H,W]) straight from K4, skipping the npz round trip.

    Ingest: frames that `build_depth_dataset` kept in memory are integrated from there; otherwise
    (streaming=None and no cache, or streaming=True) the `.raw` files stream through a bounded pinned ring and two
    device buffers (ingest.RawDepthStreamer) in chunks of `batch_frames` -- file reads, H2D copies and kernels of
    consecutive chunks overlap, memory does not grow with the capture.  Both routes give the same grid.

    colors / color_intrinsics: Open3D's colour Integrate overload for grids created with the colour attribute
    (north-star row A3c).  `colors` is uint8 [N,CH,CW,3] (pinned host or CUDA tensor) or a callable (f0, f1) -> such a
    tensor for frames f0..f1 (so that a long capture's images are staged chunk by chunk); color_intrinsics [N,3,3]."""
    dev = _torch_device(device)
    vbg = vbg_opt if vbg_opt is not None else VoxelBlockGrid(
        attr_names=("tsdf", "weight"), attr_channels=((1), (1)), voxel_size=voxel_size,
        block_resolution=block_resolution, block_count=block_count, device=dev)
    n = len(dataset.timestamps)
    if n == 0:
        return vbg
    extrinsic_wc = dataset.transforms.extrinsics_wc
    intrinsics = compute_o3d_intrinsic_matrices(dataset)
    H, W = int(dataset.heights[0]), int(dataset.widths[0])
    use_color = colors is not None and vbg.has_color
    if streaming is None:
        streaming = not depth_data_io.has_raw_cache(side, dataset)

    def chunk_confidence(f0, f1):
        if not use_confidence_filtered_depth:
            return None, None, None
        if confidence is not None:
            return confidence[0][f0:f1], confidence[1][f0:f1], None
        c, k, h = _load_confidence_stack(depth_data_io, side, dataset, f0, f1, H, W)
        return torch.from_numpy(c), torch.from_numpy(k), torch.from_numpy(h)

    def run(raw_dev, present, f0, f1):
        conf, count, has = chunk_confidence(f0, f1)
        lin, valid = depth_prepare(raw_dev, dataset.nears[f0:f1], dataset.fars[f0:f1], conf, count, has,
                                   confidence_threshold, valid_count_threshold)
        if not present.all():   # missing files: load_depth_map returns None -> frame skipped (:200-201)
            valid = valid * torch.from_numpy(present.astype(np.int32)).to(dev)
        vbg.integrate_sequence(lin, intrinsics[f0:f1], extrinsic_wc[f0:f1], float(depth_max), float(trunc_voxel_multiplier),
                               1.0, frame_valid=valid, batch_frames=batch_frames,
                               colors=(colors(f0, f1) if callable(colors) else colors[f0:f1]) if use_color else None,
                               color_intrinsics=np.asarray(color_intrinsics)[f0:f1] if use_color else None)

    if streaming:
        from .ingest import stream_side
        with torch.cuda.device(dev):
            for f0, f1, raw_dev, present in stream_side(depth_data_io, side, dataset, chunk_frames=batch_frames, device=dev):
                run(raw_dev, present, f0, f1)
        return vbg
    raw, present = depth_data_io.load_raw_sequence(side, dataset)
    raw_t = torch.from_numpy(raw)
    if torch.cuda.is_available():
        raw_t = raw_t.pin_memory()
    # confidence maps (and colour frames) are staged per chunk: host memory stays bounded
    chunk = batch_frames if use_color else max(batch_frames, 1024)
    for f0 in range(0, n, chunk):
        f1 = min(n, f0 + chunk)
        run(raw_t[f0:f1].to(dev, non_blocking=True), present[f0:f1], f0, f1)
    return vbg


def integrate_fragment_point_cloud(depth_data_io, frag_dataset: DepthDataset, side: Side, config):
    """Drop-in for integrate_fragment_point_cloud (depth_optimization/refine_fragment_poses.py:14-58; SURVEY 8f
    N4): one fragment's frames -> fresh VoxelBlockGrid -> point cloud.  Returns (side, PointCloud), or None when
    the fragment yields no points or fails (the reference logs and skips such fragments; the ICP / pose-graph
    consumers of the point clouds stay out of scope).  `config` is a FragmentPoseRefinementConfig."""
    try:
        vbg = integrate(dataset=frag_dataset, depth_data_io=depth_data_io, side=side,
                        use_confidence_filtered_depth=config.use_confidence_filtered_depth,
                        confidence_threshold=config.confidence_threshold,
                        valid_count_threshold=config.valid_count_threshold, voxel_size=config.voxel_size,
                        block_resolution=config.block_resolution, block_count=config.block_count,
                        depth_max=config.depth_max, trunc_voxel_multiplier=config.trunc_voxel_multiplier,
                        device=config.device, show_progress=False, desc=None, vbg_opt=None)
        pcd = vbg.extract_point_cloud()
        if pcd.point.positions.shape[0] == 0:
            print(f"[Warning] Fragment point cloud for {side.name} is empty (no valid points). "
                  f"Dataset has {len(frag_dataset.timestamps)} frames. "
                  f"This may indicate insufficient depth data or overly strict filtering.")
            return None
        return side, pcd
    except Exception as e:
        n = len(frag_dataset.timestamps)
        print(f"[Error] integrate_fragment_point_cloud failed for {side.name}: {e}")
        print(f"[Error] Fragment dataset info: {n} frames, timestamps range: "
              f"{frag_dataset.timestamps.min() if n > 0 else 'N/A'} - {frag_dataset.timestamps.max() if n > 0 else 'N/A'}")
        return None


def compute_odometry_information_matrix(source_depth, target_depth, intrinsic, source_to_target, dist_threshold: float = 0.07,
                                        depth_scale: float = 1000.0, depth_max: float = 3.0) -> np.ndarray:
    """Drop-in for o3d.t.pipelines.odometry.compute_odometry_information_matrix as build_pose_graph_for_fragment calls
    it (depth_optimization/make_fragments.py:142-150,228-233): float64 [6,6] information matrix of the
    point-to-point residuals between two depth frames under `source_to_target`.  Depth frames: float32 [H,W] CUDA
    tensors (or Images / arrays, moved to the device of the source frame); intrinsic 3x3 and source_to_target 4x4 on
    the host.  The pose-graph optimisation consuming the edges stays out of scope."""
    import ctypes as C
    from . import _lib
    from .vbg import _as_np, _stream, as_depth_tensor
    src = source_depth
    for attr in ("as_tensor", ):
        if hasattr(src, attr):
            src = src.as_tensor()
    src = src.torch if hasattr(src, "torch") else src
    if isinstance(src, np.ndarray):
        src = torch.from_numpy(np.ascontiguousarray(src))
    dev = src.device if src.is_cuda else torch.device("cuda", torch.cuda.current_device())
    s = as_depth_tensor(source_depth, dev)
    t = as_depth_tensor(target_depth, dev)
    if s.shape != t.shape:
        raise RuntimeError(f"source and target depth differ in shape: {tuple(s.shape)} vs {tuple(t.shape)}")
    H, W = s.shape
    K = _as_np(intrinsic, np.float64, (3, 3))
    T = _as_np(source_to_target, np.float64, (4, 4))
    info = np.zeros((6, 6), np.float64)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mq3d_odometry_information(_lib.dptr(s), _lib.dptr(t), W, H, _lib.darr(K), _lib.darr(T),
                                                        C.c_float(dist_threshold), C.c_float(depth_scale),
                                                        C.c_float(depth_max), _lib.darr(info), int(dev.index or 0), _stream()))
    return info


def estimate_depth_confidences(depth_data_io, config, device="CUDA:0", save: bool = True) -> dict:
    """Drop-in for estimate_depth_confidences (estimate_depth_confidences.py:120-154): one K4 launch per
    side instead of a process pool.  Writes `<side>_depth_confidence/<ts>.npz` (keys confidence_map f64,
    valid_count i32) like the reference, and returns {side: (conf, count)} device tensors for direct reuse."""
    dev = _torch_device(device)
    out = {}
    for side in Side:
        if config.skip_if_output_dir_exists and depth_data_io.exists_depth_confidence_map_dir(side=side):
            print(f"[{side.name}] Skipping confidence map estimation: output directory already exists. "
                  f"Set skip_if_output_dir_exists = False to force re-estimation.")
            continue
        dataset = depth_data_io.load_depth_dataset(side=side)
        n = len(dataset)
        if n == 0:
            continue
        K = compute_o3d_intrinsic_matrices(dataset)
        Ecw = dataset.transforms.convert_coordinate_system(CoordinateSystem.OPEN3D, is_camera=True).extrinsics_cw
        Einv = np.linalg.inv(Ecw)
        raw, present = depth_data_io.load_raw_sequence(side, dataset)
        lin, valid = depth_prepare(torch.from_numpy(raw).pin_memory().to(dev, non_blocking=True), dataset.nears,
                                   dataset.fars)
        valid = valid * torch.from_numpy(present.astype(np.int32)).to(dev)
        conf, count = estimate_confidence(lin, K, Ecw, Einv, config.target_frame_range, config.depth_max,
                                          config.error_threshold, frame_valid=valid)
        out[side] = (conf, count)
        if save:
            ch, kh, vh = conf.cpu().numpy(), count.cpu().numpy(), valid.cpu().numpy()
            for i in range(n):
                if not vh[i]:
                    continue   # build_confidence_map returned None: no file is written (:27-31,112-113)
                if depth_data_io.load_confidence_map(side=side, timestamp=dataset.timestamps[i]) is not None:
                    continue   # per-file skip (:94-96)
                depth_data_io.save_confidence_map(side=side, timestamp=dataset.timestamps[i],
                                                  confidence_map=ConfidenceMap(ch[i], kh[i]))
    return out


def filter_mesh_components(mesh: TriangleMesh, min_triangle_count: int = 2000, backend: str = "auto") -> TriangleMesh:
    """Drop-in for o3d_utils.filter_mesh_components (:241-321), step for step on the legacy-mesh semantics of
    Open3D: cluster_connected_triangles (edge adjacency), keep clusters with >= min_triangle_count triangles (or
    the largest one), remove_unreferenced_vertices (only if something was removed), remove_degenerate_triangles
    (repeated vertex index), remove_duplicated_triangles (equal up to rotation), remove_duplicated_vertices
    (identical coordinates; first occurrence kept, which also welds the vertices several ranks re-emit on
    ghost edges) and remove_non_manifold_edges.  Normals and colours follow their vertices.

    backend "device": the whole chain on the GPU (mq3d_mesh_filter: sort-based edge adjacency + pointer-jumping
    union-find, compaction, coordinate weld -- no D2H / H2D round trip of the mesh); "host": the NumPy / SciPy
    implementation in meshops.py (the checker); "auto": device for CUDA meshes."""
    if int(mesh.triangle.indices.shape[0]) == 0:
        print("[Warning] Mesh filtering: Input mesh has no triangles, returning as-is")
        return mesh
    if backend == "auto":
        backend = "device" if mesh.vertex.positions.is_cuda else "host"
    if backend == "device":
        from .meshfilter import filter_mesh_components_device
        out, info = filter_mesh_components_device(mesh, min_triangle_count)
    else:
        out, info = _filter_mesh_components_host(mesh, min_triangle_count)
    n_comp, n_valid, removed_tris, original, final, largest, fallback = info
    if fallback:
        print(f"[Warning] Mesh filtering: No components have >= {min_triangle_count} triangles. "
              f"Largest component has {largest} triangles.")
        print("[Warning] Mesh filtering: Returning largest component only.")
    removed = n_comp - n_valid
    if removed > 0:
        print(f"[Info] Mesh filtering: Found {n_comp} connected component(s)")
        print(f"[Info] Mesh filtering: Removed {removed} small component(s) with < {min_triangle_count} triangles")
        print(f"[Info] Mesh filtering: Removed {removed_tris} triangles from small components")
        print(f"[Info] Mesh filtering: Kept {n_valid} component(s) with >= {min_triangle_count} triangles")
        print(f"[Info] Mesh filtering: Final mesh has {final} triangles (was {original})")
    else:
        print(f"[Info] Mesh filtering: All {n_comp} component(s) have >= {min_triangle_count} triangles, "
              f"no filtering needed")
    return out


def _filter_mesh_components_host(mesh: TriangleMesh, min_triangle_count: int):
    """The reference's chain on a legacy mesh (geometry.LegacyTriangleMesh implements Open3D's methods on the host)."""
    legacy = mesh.to_legacy()
    clusters, counts, _ = legacy.cluster_connected_triangles()
    clusters, counts = np.asarray(clusters), np.asarray(counts)
    original = len(clusters)
    valid = np.nonzero(counts >= min_triangle_count)[0]
    fallback = len(valid) == 0
    if fallback:
        valid = np.array([np.argmax(counts)])
    keep = np.isin(clusters, valid)
    removed_tris = int(original - keep.sum())
    if removed_tris > 0:
        legacy.remove_triangles_by_mask(~keep)
        legacy.remove_unreferenced_vertices()
    legacy.remove_degenerate_triangles()
    legacy.remove_duplicated_triangles()
    legacy.remove_duplicated_vertices()
    legacy.remove_non_manifold_edges()
    out = TriangleMesh.from_legacy(legacy, device=mesh.device)
    return out, (len(counts), len(valid), removed_tris, original, len(legacy.triangles), int(counts.max()), fallback)


def raycast_in_color_view(scene: RaycastingScene, dataset: CameraDataset) -> Generator[np.ndarray, None, None]:
    """Per colour frame: pinhole rays with float32 K (cx mirrored) and world->camera E, closest hit
    `t_hit` as float32 [H,W] host array, inf on miss (o3d_utils.py:324-342)."""
    intrinsics = compute_o3d_intrinsic_matrices(dataset)
    extrinsics = dataset.transforms.extrinsics_wc
    for i in range(len(dataset)):
        rays = scene.create_rays_pinhole(intrinsics[i].astype(np.float32), extrinsics[i].astype(np.float32),
                                         width_px=int(dataset.widths[i]), height_px=int(dataset.heights[i]))
        yield scene.cast_rays(rays)["t_hit"].cpu().numpy()
