"""``DataIO(project_dir)`` facade with ``.depth / .color / .rgbd / .reconstruction`` -- the same
project layout and method names as the reference's scripts/dataio (data_io.py:10-16,
depth_data_io.py:14-280, image_data_io.py:166-297, rgbd_data_io.py:8-32,
reconstruction_data_io.py:9-145, config/project_path_config.py:6-261).

Ingest differences: ``build_depth_dataset`` reads every ``.raw`` once into one pinned
host buffer, validates all frames with K1's fused reduction on the GPU and keeps the buffer, so
``integrate()`` and the confidence stage never re-read or re-convert a file.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional

import numpy as np

from .geometry import LegacyPointCloud, LegacyTriangleMesh
from .io_utils import write_point_cloud, write_triangle_mesh
from .models import CameraDataset, ConfidenceMap, CoordinateSystem, DepthDataset, Side, Transforms

DEPTH_DIR = {Side.LEFT: "left_depth", Side.RIGHT: "right_depth"}
DEPTH_DESCRIPTOR_CSV = {Side.LEFT: "left_depth_descriptors.csv", Side.RIGHT: "right_depth_descriptors.csv"}
DEPTH_CONFIDENCE_DIR = {Side.LEFT: "left_depth_confidence", Side.RIGHT: "right_depth_confidence"}
DEPTH_DATASET_NPZ = {Side.LEFT: "dataset/left_depth_dataset.npz", Side.RIGHT: "dataset/right_depth_dataset.npz"}
OPT_DEPTH_DATASET_NPZ = {Side.LEFT: "dataset/left_depth_dataset_optimized.npz",
                         Side.RIGHT: "dataset/right_depth_dataset_optimized.npz"}
RGB_DIR = {Side.LEFT: "left_camera_rgb", Side.RIGHT: "right_camera_rgb"}
COLOR_DATASET_NPZ = {Side.LEFT: "dataset/left_camera_dataset.npz", Side.RIGHT: "dataset/right_camera_dataset.npz"}
OPT_COLOR_DATASET_NPZ = {Side.LEFT: "dataset/left_camera_dataset_optimized.npz",
                         Side.RIGHT: "dataset/right_camera_dataset_optimized.npz"}
COLOR_ALIGNED_DEPTH_DIR = {Side.LEFT: "left_color_aligned_depth", Side.RIGHT: "right_color_aligned_depth"}

DESCRIPTOR_COLUMNS = ["timestamp_ms", "width", "height", "near_z", "far_z", "fov_left_angle_tangent",
                      "fov_right_angle_tangent", "fov_top_angle_tangent", "fov_down_angle_tangent",
                      "create_pose_location_x", "create_pose_location_y", "create_pose_location_z",
                      "create_pose_rotation_x", "create_pose_rotation_y", "create_pose_rotation_z",
                      "create_pose_rotation_w"]


def depth_camera_params(left, right, top, bottom, width, height):
    """utils/depth_utils.py:4-18."""
    return (width / (right + left), height / (top + bottom), width * right / (right + left),
            height * top / (top + bottom))


class DepthDataIO:
    def __init__(self, project_dir: Path):
        self.project_dir = Path(project_dir)
        self.depth_datasets: dict[Side, DepthDataset] = {}
        self._raw_cache: dict[Side, tuple[np.ndarray, np.ndarray]] = {}   # side -> (timestamps, raw frames)

    # -- paths ----------------------------------------------------------------------------------------
    def depth_dir(self, side):
        return self.project_dir / DEPTH_DIR[side]

    def depth_map_path(self, side, timestamp):
        return self.depth_dir(side) / f"{int(timestamp)}.raw"

    def confidence_dir(self, side):
        return self.project_dir / DEPTH_CONFIDENCE_DIR[side]

    def confidence_map_path(self, side, timestamp):
        return self.confidence_dir(side) / f"{int(timestamp)}.npz"

    def exists_depth_confidence_map_dir(self, side) -> bool:
        return self.confidence_dir(side).exists()

    # -- raw frames -------------------------------------------------------------------------------------
    RAW_CACHE_BUDGET_BYTES = int(os.environ.get("MQ3D_RAW_CACHE_MB", "2048")) << 20

    def has_raw_cache(self, side, dataset: DepthDataset) -> bool:
        """True when every frame of `dataset` is still in memory from `build_depth_dataset` (small captures)."""
        cached = self._raw_cache.get(side)
        if cached is None or len(cached[0]) == 0:
            return False
        have = set(int(t) for t in cached[0])
        return all(int(t) in have for t in dataset.timestamps)

    def load_raw_depth_map(self, side, timestamp, width, height) -> Optional[np.ndarray]:
        """`<ts>.raw` little-endian float32 [H,W] (depth_data_io.py:41-46), or None if missing."""
        p = self.depth_map_path(side, timestamp)
        if not p.exists():
            return None
        return np.fromfile(p, dtype="<f4").reshape((int(height), int(width)))

    def load_raw_sequence(self, side, dataset: DepthDataset):
        """(raw float32 [N,H,W] host array, present bool [N]) for the dataset's frames; frames whose file
        is missing are zero-filled and flagged (they fail validation, like the reference's None)."""
        cached = self._raw_cache.get(side)
        ts = np.asarray(dataset.timestamps)
        n = len(ts)
        if n == 0:
            return np.zeros((0, 0, 0), np.float32), np.zeros((0,), bool)
        H, W = int(dataset.heights[0]), int(dataset.widths[0])
        if not (np.all(dataset.heights == H) and np.all(dataset.widths == W)):
            raise RuntimeError("all depth frames of a side must share one resolution")
        raw = np.zeros((n, H, W), np.float32)
        present = np.zeros(n, bool)
        lookup = {}
        if cached is not None:
            lookup = {int(t): i for i, t in enumerate(cached[0])}
        todo = []
        for i, t in enumerate(ts):
            j = lookup.get(int(t))
            if j is not None:
                raw[i] = cached[1][j]
                present[i] = True
            else:
                todo.append(i)

        def read(i):           # straight into row i of the batch; file reads release the GIL
            p = self.depth_map_path(side, ts[i])
            if not p.exists():
                return i, False
            a = np.fromfile(p, dtype="<f4")
            if a.size != H * W:
                raise RuntimeError(f"{p}: expected {H * W} float32 values, found {a.size}")
            raw[i] = a.reshape(H, W)
            return i, True

        if len(todo) > 8:      # the file reads of a whole side overlap instead of queueing
            import os
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
                done = list(pool.map(read, todo))
        else:
            done = [read(i) for i in todo]
        for i, ok in done:
            present[i] = ok
        return raw, present

    # -- confidence maps ----------------------------------------------------------------------------------
    def load_confidence_map(self, side, timestamp) -> Optional[ConfidenceMap]:
        p = self.confidence_map_path(side, timestamp)
        if p.exists():
            try:
                d = np.load(p)
                return ConfidenceMap(confidence_map=d["confidence_map"], valid_count=d["valid_count"])
            except Exception as e:  # mirrors depth_data_io.py:102-103
                print(f"[Error] Failed to load confidence map for {side.name} at timestamp {timestamp}: {e}")
        return None

    def save_confidence_map(self, side, timestamp, confidence_map: ConfidenceMap) -> None:
        p = self.confidence_map_path(side, timestamp)
        p.parent.mkdir(parents=True, exist_ok=True)
        np.savez(p, confidence_map=confidence_map.confidence_map, valid_count=confidence_map.valid_count)

    # -- datasets -------------------------------------------------------------------------------------------
    def load_depth_dataset(self, side, use_cache: bool = True) -> DepthDataset:
        if side in self.depth_datasets:
            return self.depth_datasets[side]
        p = self.project_dir / DEPTH_DATASET_NPZ[side]
        if use_cache and p.exists():
            try:
                ds = DepthDataset.load(p)
                self.depth_datasets[side] = ds
                return ds
            except Exception as e:
                print(f"[Error] Depth dataset cache is corrupted or invalid. Rebuilding cache from the original source...\n{e}")
        ds = self.build_depth_dataset(side)
        self.depth_datasets[side] = ds
        ds.save(p)
        return ds

    def load_optimized_depth_dataset(self, side) -> Optional[DepthDataset]:
        p = self.project_dir / OPT_DEPTH_DATASET_NPZ[side]
        if p.exists():
            try:
                return DepthDataset.load(p)
            except Exception as e:
                print(f"[Error] Depth dataset cache is corrupted or invalid.\n{e}")
        return None

    def save_optimized_depth_dataset(self, side, dataset: DepthDataset):
        dataset.save(self.project_dir / OPT_DEPTH_DATASET_NPZ[side])

    def build_depth_dataset(self, side) -> DepthDataset:
        """depth_data_io.py:166-261: one row per descriptor whose `.raw` exists and is valid.  Validation
        (all-0 / all-1 / NaN / negative) runs as K1's fused reduction over the whole side at once."""
        import pandas as pd
        import torch
        from .vbg import depth_prepare
        df = pd.read_csv(self.project_dir / DEPTH_DESCRIPTOR_CSV[side])
        rows = []
        for _, r in df.iterrows():
            ts, w, h = int(r["timestamp_ms"]), int(r["width"]), int(r["height"])
            if self.depth_map_path(side, ts).exists():
                rows.append((ts, w, h, r))
        n = len(rows)
        keep = np.zeros(n, bool)
        raws = None
        if n:
            # every file is read ONCE, through the bounded ring of ingest.RawDepthStreamer; is_depth_map_valid runs as
            # K1's fused reduction on the device chunk by chunk; the frames are kept in memory for integrate() only
            # when the whole side fits the cache budget (MQ3D_RAW_CACHE_MB, default 2 GiB)
            from .ingest import RawDepthStreamer
            H, W = rows[0][2], rows[0][1]
            for ts, w, h, _ in rows:
                if (w, h) != (W, H):
                    raise RuntimeError("all depth frames of a side must share one resolution")
            dev = torch.device("cuda", torch.cuda.current_device())
            nears = np.array([float(r["near_z"]) for *_, r in rows])
            fars = np.array([float(r["far_z"]) for *_, r in rows])
            cache = n * H * W * 4 <= self.RAW_CACHE_BUDGET_BYTES
            raws = np.zeros((n, H, W), np.float32) if cache else None
            streamer = RawDepthStreamer(lambda i: self.depth_map_path(side, rows[i][0]), n, H, W, chunk_frames=64, device=dev)
            for f0, f1, raw_dev, present in streamer:
                _, valid = depth_prepare(raw_dev, nears[f0:f1], fars[f0:f1])
                keep[f0:f1] = valid.cpu().numpy().astype(bool) & present
                if cache:
                    raws[f0:f1] = streamer.host_view
        sel = [rows[i] for i in range(n) if keep[i]]
        if raws is not None:
            self._raw_cache[side] = (np.array([s[0] for s in sel], dtype=np.int64), raws[keep])
        else:
            self._raw_cache.pop(side, None)
        cols = {k: [] for k in ("ts", "w", "h", "near", "far", "fx", "fy", "cx", "cy", "pos", "rot")}
        for ts, w, h, r in sel:
            fx, fy, cx, cy = depth_camera_params(float(r["fov_left_angle_tangent"]), float(r["fov_right_angle_tangent"]),
                                                 float(r["fov_top_angle_tangent"]), float(r["fov_down_angle_tangent"]), w, h)
            cols["ts"].append(ts); cols["w"].append(w); cols["h"].append(h)
            cols["near"].append(float(r["near_z"])); cols["far"].append(float(r["far_z"]))
            cols["fx"].append(fx); cols["fy"].append(fy); cols["cx"].append(cx); cols["cy"].append(cy)
            cols["pos"].append([r["create_pose_location_x"], r["create_pose_location_y"], r["create_pose_location_z"]])
            cols["rot"].append([r["create_pose_rotation_x"], r["create_pose_rotation_y"], r["create_pose_rotation_z"],
                                r["create_pose_rotation_w"]])
        return DepthDataset(
            directory_relative_path=DEPTH_DIR[side],
            image_file_names=np.array([f"{t}.raw" for t in cols["ts"]]),
            timestamps=np.array(cols["ts"], dtype=np.int64),
            fx=np.array(cols["fx"]), fy=np.array(cols["fy"]), cx=np.array(cols["cx"]), cy=np.array(cols["cy"]),
            transforms=Transforms(CoordinateSystem.UNITY, np.array(cols["pos"], dtype=np.float64).reshape(-1, 3),
                                  np.array(cols["rot"], dtype=np.float64).reshape(-1, 4)),
            widths=np.array(cols["w"], dtype=np.int64), heights=np.array(cols["h"], dtype=np.int64),
            nears=np.array(cols["near"]), fars=np.array(cols["far"]))


class ImageDataIO:
    """Colour camera datasets.  Only the cached-dataset path is on the hot path (reconstruct_scene.py:211
    loads with use_cache=True); building from hmd_poses.csv + camera characteristics is upstream ingest."""

    def __init__(self, project_dir: Path):
        self.project_dir = Path(project_dir)

    def color_dataset_path(self, side):
        return self.project_dir / COLOR_DATASET_NPZ[side]

    def load_color_dataset(self, side, use_cache: bool = True) -> CameraDataset:
        p = self.color_dataset_path(side)
        if p.exists():
            return CameraDataset.load(p)
        raise FileNotFoundError(f"Colour dataset cache {p} not found (the HMD-pose interpolation that builds it "
                                f"is upstream of this hot-path build; see DESIGN.md)")

    def save_color_dataset(self, side, dataset: CameraDataset):
        dataset.save(self.color_dataset_path(side))

    def load_optimized_color_dataset(self, side) -> Optional[CameraDataset]:
        p = self.project_dir / OPT_COLOR_DATASET_NPZ[side]
        return CameraDataset.load(p) if p.exists() else None

    def load_rgb(self, side, timestamp) -> np.ndarray:
        import cv2
        p = self.project_dir / RGB_DIR[side] / f"{int(timestamp)}.png"
        bgr = cv2.imread(str(p))
        if bgr is None:
            raise FileNotFoundError(f"Image file not found or cannot be read: {p}")
        return cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)


class RGBDDataIO:
    def __init__(self, project_dir: Path):
        self.project_dir = Path(project_dir)

    def color_aligned_depth_path(self, side, timestamp):
        return self.project_dir / COLOR_ALIGNED_DEPTH_DIR[side] / f"{int(timestamp)}.npy"

    def load_color_aligned_depth(self, side, timestamp) -> np.ndarray:
        p = self.color_aligned_depth_path(side, timestamp)
        if not p.exists():
            raise FileNotFoundError(f"Color-aligned depth file not found: {p}")
        return np.load(p)

    def save_color_aligned_depth(self, depth_map: np.ndarray, side, timestamp):
        p = self.color_aligned_depth_path(side, timestamp)
        p.parent.mkdir(parents=True, exist_ok=True)
        np.save(p, depth_map)


class ReconstructionDataIO:
    def __init__(self, project_dir: Path):
        self.project_dir = Path(project_dir)

    def _p(self, name):
        p = self.project_dir / "reconstruction" / name
        p.parent.mkdir(parents=True, exist_ok=True)
        return p

    def get_colorless_vbg_path(self):
        return self.project_dir / "reconstruction" / "colorless_vbg.npz"

    def load_colorless_vbg(self, device="CUDA:0"):
        from .vbg import VoxelBlockGrid
        p = self.get_colorless_vbg_path()
        return VoxelBlockGrid.load(str(p), device=device) if p.exists() else None

    def save_colorless_vbg(self, vbg):
        vbg.save(str(self._p("colorless_vbg.npz")))

    def save_colorless_pcd_legacy(self, pcd: LegacyPointCloud):
        write_point_cloud(self._p("colorless.ply"), pcd)

    def save_colorless_mesh_raw_legacy(self, mesh: LegacyTriangleMesh):
        write_triangle_mesh(self._p("colorless_mesh_raw.ply"), mesh)

    def save_colorless_mesh_clean_legacy(self, mesh: LegacyTriangleMesh):
        write_triangle_mesh(self._p("colorless_mesh_clean.ply"), mesh)

    def save_colored_mesh_legacy(self, mesh: LegacyTriangleMesh):
        write_triangle_mesh(self._p("color_mesh.ply"), mesh)


class DataIO:
    def __init__(self, project_dir: Path):
        self.project_dir = Path(project_dir).resolve()
        self.color = ImageDataIO(self.project_dir)
        self.depth = DepthDataIO(self.project_dir)
        self.rgbd = RGBDDataIO(self.project_dir)
        self.reconstruction = ReconstructionDataIO(self.project_dir)
