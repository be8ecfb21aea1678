#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE's own NumPy code (K1, K4, camera glue).

Runs only in the build container (needs /root/reference); the resulting small .npz fixtures are
committed and are what tests/ read.  open3d is absent here, so it is replaced by a MagicMock for the
import of estimate_depth_confidences.py (which only needs it for type annotations / config defaults).

  python tests/golden/make_golden.py
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/scripts"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.modules.setdefault("open3d", MagicMock())

from utils.depth_utils import convert_depth_to_linear, compute_depth_camera_params  # noqa: E402
from dataio.depth_data_io import DepthDataIO  # noqa: E402
from models.transforms import Transforms as RefTransforms, CoordinateSystem as RefCS  # noqa: E402
from models.camera_dataset import DepthDataset as RefDepthDataset  # noqa: E402
from processing.reconstruction.confidence_estimation.compute_pixel_error_map import compute_pixel_error_map  # noqa: E402
from processing.reconstruction.confidence_estimation.estimate_depth_confidences import build_confidence_map  # noqa: E402

import mq3d_b200  # noqa: E402,F401
from mq3d_b200 import synth  # noqa: E402
from mq3d_b200.models import Side  # noqa: E402


def ref_dataset(cap):
    ds = cap.dataset
    return RefDepthDataset(
        directory_relative_path=ds.directory_relative_path, image_file_names=ds.image_file_names,
        timestamps=ds.timestamps, fx=ds.fx, fy=ds.fy, cx=ds.cx, cy=ds.cy,
        transforms=RefTransforms(RefCS.UNITY, ds.transforms.positions, ds.transforms.rotations),
        widths=ds.widths, heights=ds.heights, nears=ds.nears, fars=ds.fars)


class FakeDepthIO:
    """load_depth_map_by_index exactly as DepthDataIO does (depth_data_io.py:56-77), from memory."""

    def __init__(self, raw):
        self.raw = raw
        self.io = DepthDataIO(depth_path_config=None)

    def load_depth_map_by_index(self, side, dataset, index):
        if index < 0 or index >= len(dataset.timestamps):
            return None
        arr = self.raw[index]
        if not self.io.is_depth_map_valid(depth_map=arr):
            return None
        return convert_depth_to_linear(arr, dataset.nears[index], dataset.fars[index])


def main():
    out = {}
    # ---- K1 -------------------------------------------------------------------------------------
    cap = synth.make_depth_capture(3, Side.LEFT, width=64, height=48)
    raw = cap.raw.copy()
    raw[2, :5, :7] = 0.0
    raw[2, 40:, 60:] = 1.0
    cases = [(np.float64(0.1), np.float64(np.inf)), (np.float64(0.1), np.float64(5.0)),
             (np.float64(0.25), np.float64(0.1)), (np.float64(0.1), np.float64(100.0))]
    out["k1_raw"] = raw
    out["k1_near_far"] = np.array(cases, dtype=np.float64)
    out["k1_linear"] = np.stack([np.stack([convert_depth_to_linear(raw[i], n, f) for i in range(3)]) for n, f in cases])
    io = DepthDataIO(depth_path_config=None)
    valid_in = np.stack([raw[0], np.zeros_like(raw[0]), np.ones_like(raw[0]), raw[1], raw[2]])
    valid_in[3, 3, 3] = np.nan
    valid_in[4, 7, 7] = -1e-3
    out["k1_valid_in"] = valid_in
    out["k1_valid"] = np.array([io.is_depth_map_valid(depth_map=v) for v in valid_in], dtype=np.int32)
    out["k1_cam_params"] = np.array(compute_depth_camera_params(0.9, 1.1, 1.05, 0.95, 320, 320), dtype=np.float64)

    # ---- camera glue: Transforms / intrinsics as the pipeline derives them -----------------------
    cap = synth.make_depth_capture(8, Side.RIGHT, width=80, height=80, sweep_frames=120)
    rds = ref_dataset(cap)
    tr = rds.transforms.convert_coordinate_system(target_coordinate_system=RefCS.OPEN3D, is_camera=True)
    K = rds.get_intrinsic_matrices()
    K[:, 0, 2] = rds.widths - K[:, 0, 2]
    Ecw = tr.extrinsics_cw
    Ewc = tr.extrinsics_wc
    Einv = np.linalg.inv(Ecw)
    out["cam_positions"] = cap.dataset.transforms.positions
    out["cam_rotations"] = cap.dataset.transforms.rotations
    out["cam_K"] = K
    out["cam_Ecw"] = Ecw
    out["cam_Ewc"] = Ewc

    # ---- K4 -------------------------------------------------------------------------------------
    fio = FakeDepthIO(cap.raw)
    depths = np.stack([fio.load_depth_map_by_index(None, rds, i) for i in range(8)])
    out["k4_raw"] = cap.raw
    out["k4_depths"] = depths
    pairs = [(0, 1), (3, 1), (4, 7), (7, 6)]
    out["k4_pairs"] = np.array(pairs, dtype=np.int32)
    out["k4_err"] = np.stack([compute_pixel_error_map(K, Ecw, Einv, r, depths[r], t, depths[t], depth_max=4.0)
                              for r, t in pairs])
    confs, counts = [], []
    for r in range(8):
        cm = build_confidence_map(depth_data_io=fio, dataset=rds, intrinsic_matrices=K, extrinsic_matrices=Ecw,
                                  extrinsic_matrices_inv=Einv, side=None, ref_frame_idx=r, target_frame_range=2,
                                  depth_max=4.0, error_threshold=0.08)
        confs.append(cm.confidence_map)
        counts.append(cm.valid_count)
    out["k4_conf"] = np.stack(confs)
    out["k4_count"] = np.stack(counts)
    assert out["k4_conf"].dtype == np.float64 and out["k4_count"].dtype == np.int32
    assert out["k1_linear"].dtype == np.float32 and out["k4_err"].dtype == np.float32
    # the float64 promotion in convert_depth_to_linear depends on NumPy's scalar promotion rules (NEP 50, NumPy >= 2)
    out["meta_numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(HERE, "reference_k1_k4.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
