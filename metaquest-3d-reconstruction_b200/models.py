"""Host-side data model of the reference pipeline (O(N_frames) NumPy, stays on the CPU).

Mirrors the names and numeric behaviour of the reference's ``scripts/models``:
``Side`` (side.py:4-6), ``CoordinateSystem``/``Transforms`` (transforms.py:8-288),
``CameraDataset``/``DepthDataset`` (camera_dataset.py:13-238) and ``ConfidenceMap``
(confidence_map.py:7-32).  Dtype contracts the kernels depend on:
intrinsics and extrinsics are float32, world->camera is ``np.linalg.inv`` of the float32
camera->world matrices.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from pathlib import Path

import numpy as np
from scipy.spatial.transform import Rotation


class Side(Enum):
    LEFT = "left"
    RIGHT = "right"


class CoordinateSystem(Enum):
    UNITY = "Unity"            # world Y-up left-handed; camera X-right Y-up Z-forward
    OPEN3D = "Open3D"          # world Y-up right-handed; camera X-right Y-down Z-forward
    NERFSTUDIO = "NerfStudio"  # world Z-up right-handed; camera X-right Y-up Z-backward
    COLMAP = "COLMAP"          # world Y-down right-handed; camera X-right Y-down Z-forward


# world-axis bases relative to Unity (transforms.py:134-151) and camera bases (:154-161)
_WORLD_BASIS = {
    CoordinateSystem.UNITY: np.eye(3),
    CoordinateSystem.OPEN3D: np.diag((1.0, 1.0, -1.0)),
    CoordinateSystem.NERFSTUDIO: np.array([[1.0, 0, 0], [0, 0, 1], [0, 1, 0]]),
    CoordinateSystem.COLMAP: np.diag((1.0, -1.0, 1.0)),
}
_CAMERA_BASIS = {
    CoordinateSystem.UNITY: np.eye(3),
    CoordinateSystem.OPEN3D: np.diag((1.0, -1.0, -1.0)),
    CoordinateSystem.NERFSTUDIO: np.array([[1.0, 0, 0], [0, 0, 1], [0, -1, 0]]),
    CoordinateSystem.COLMAP: np.eye(3),
}


@dataclass
class Transforms:
    """Camera poses: positions (N,3) and camera->world quaternions (N,4) in xyzw order."""
    coordinate_system: CoordinateSystem
    positions: np.ndarray
    rotations: np.ndarray

    def _matrices(self, inverse: bool) -> np.ndarray:
        n = len(self.positions)
        m = np.zeros((n, 4, 4), dtype=np.float32)
        m[:, :3, :3] = Rotation.from_quat(self.rotations).as_matrix()
        m[:, :3, 3] = self.positions
        m[:, 3, 3] = 1.0
        return np.linalg.inv(m) if inverse else m

    @property
    def extrinsics_cw(self) -> np.ndarray:
        """(N,4,4) float32 camera->world (transforms.py:205-220)."""
        return self._matrices(inverse=False)

    @property
    def extrinsics_wc(self) -> np.ndarray:
        """(N,4,4) float32 world->camera = inv(float32 camera->world) (transforms.py:216)."""
        return self._matrices(inverse=True)

    def convert_coordinate_system(self, target_coordinate_system: CoordinateSystem,
                                  is_camera: bool = False, skip_rotation: bool = False) -> "Transforms":
        """transforms.py:164-202."""
        src = self.coordinate_system
        if src == target_coordinate_system:
            return self
        conv = _WORLD_BASIS[target_coordinate_system] @ _WORLD_BASIS[src].T
        pos = (conv @ self.positions.T).T
        if skip_rotation:
            return Transforms(target_coordinate_system, pos, self.rotations)
        rot = Rotation.from_quat(self.rotations).as_matrix()
        if is_camera:
            rot = rot @ _CAMERA_BASIS[src].T
        rot = conv @ rot @ conv.T
        if is_camera:
            rot = rot @ _CAMERA_BASIS[target_coordinate_system]
        return Transforms(target_coordinate_system, pos, Rotation.from_matrix(rot).as_quat())


@dataclass
class ConfidenceMap:
    confidence_map: np.ndarray  # float64 [H,W]
    valid_count: np.ndarray     # int32 [H,W]

    def __post_init__(self):
        if self.confidence_map.shape != self.valid_count.shape:
            raise ValueError("Confidence map and valid mask must have the same shape.")
        if self.confidence_map.ndim != 2:
            raise ValueError("Confidence map must be a 2D array.")

    @property
    def shape(self):
        return self.confidence_map.shape


_ARRAY_FIELDS = ("image_file_names", "timestamps", "fx", "fy", "cx", "cy", "widths", "heights")


@dataclass
class CameraDataset:
    """Struct-of-arrays camera sequence persisted as npz (camera_dataset.py:13-214)."""
    directory_relative_path: str
    image_file_names: np.ndarray
    timestamps: np.ndarray
    fx: np.ndarray
    fy: np.ndarray
    cx: np.ndarray
    cy: np.ndarray
    transforms: Transforms
    widths: np.ndarray
    heights: np.ndarray

    _extra_fields = ()

    def __len__(self) -> int:
        return len(self.timestamps)

    def find_nearest_index(self, timestamp: int) -> int:
        """Index of the frame closest in time (models/camera_dataset.py:79-90; ties go to the earlier frame)."""
        i = int(np.searchsorted(self.timestamps, timestamp, side="left"))
        if i == len(self.timestamps):
            return i - 1
        if i == 0:
            return 0
        return i if abs(self.timestamps[i] - timestamp) < abs(self.timestamps[i - 1] - timestamp) else i - 1

    def get_intrinsic_matrices(self) -> np.ndarray:
        """(N,3,3) float32 (camera_dataset.py:93-104)."""
        k = np.zeros((len(self.fx), 3, 3), dtype=np.float32)
        k[:, 0, 0] = self.fx
        k[:, 1, 1] = self.fy
        k[:, 0, 2] = self.cx
        k[:, 1, 2] = self.cy
        k[:, 2, 2] = 1.0
        return k

    def to_dict(self) -> dict:
        d = {"directory_relative_path": self.directory_relative_path}
        for f in _ARRAY_FIELDS + self._extra_fields:
            d[f] = getattr(self, f)
        d["coordinate_system"] = self.transforms.coordinate_system.name
        d["positions"] = self.transforms.positions
        d["rotations"] = self.transforms.rotations
        return d

    @classmethod
    def from_dict(cls, data: dict):
        data = dict(data)
        if "coordinate_system" in data:
            data["transforms"] = Transforms(
                CoordinateSystem[str(data.pop("coordinate_system"))],
                data.pop("positions"), data.pop("rotations"))
        data["directory_relative_path"] = str(data["directory_relative_path"])
        return cls(**data)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            out = {f: getattr(self, f)[idx] for f in _ARRAY_FIELDS + self._extra_fields}
            out["positions"] = self.transforms.positions[idx]
            out["rotations"] = self.transforms.rotations[idx]
            return out
        d = self.to_dict()
        for k, v in d.items():
            if isinstance(v, np.ndarray) and v.ndim > 0:
                d[k] = v[idx]
        return self.__class__.from_dict(d)

    def split(self, fragment_size: int):
        return [self[i:i + fragment_size] for i in range(0, len(self), fragment_size)]

    def save(self, path: Path):
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        np.savez(path, **self.to_dict())

    @classmethod
    def load(cls, path: Path):
        return cls.from_dict(dict(np.load(path, allow_pickle=False)))


@dataclass
class DepthDataset(CameraDataset):
    nears: np.ndarray = None
    fars: np.ndarray = None

    _extra_fields = ("nears", "fars")
