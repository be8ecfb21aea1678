"""CPU: pin the oracle (and the host-side camera glue) against golden vectors produced by the
REFERENCE's own NumPy code (tests/golden/make_golden.py -> reference_k1_k4.npz)."""
import os

import numpy as np
import pytest

import mq3d_b200  # noqa: F401
from mq3d_b200.models import CoordinateSystem, Transforms
from mq3d_b200.synth import depth_intrinsics

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_k1_k4.npz"))


def test_k1_linear_depth_bit_exact(oracle):
    raw = G["k1_raw"]
    for c, (near, far) in enumerate(G["k1_near_far"]):
        for i in range(raw.shape[0]):
            got = oracle.depth_to_linear(raw[i], near, far)
            assert np.array_equal(got.view(np.uint32), G["k1_linear"][c, i].view(np.uint32)), (c, i)


def test_k1_validity_flags(oracle):
    got = [int(oracle.depth_valid(v)) for v in G["k1_valid_in"]]
    assert got == G["k1_valid"].tolist() == [1, 0, 0, 0, 0]


def test_depth_camera_params_match_reference():
    fx, fy, cx, cy = depth_intrinsics(320, 320, tangents=(0.9, 1.1, 1.05, 0.95))
    assert np.array_equal(np.array([fx, fy, cx, cy]), G["k1_cam_params"])


def test_transforms_and_intrinsics_match_reference():
    tr = Transforms(CoordinateSystem.UNITY, G["cam_positions"], G["cam_rotations"])
    o3d = tr.convert_coordinate_system(CoordinateSystem.OPEN3D, is_camera=True)
    assert o3d.extrinsics_cw.dtype == np.float32
    assert np.array_equal(o3d.extrinsics_cw, G["cam_Ecw"])
    assert np.array_equal(o3d.extrinsics_wc, G["cam_Ewc"])


def test_k4_pixel_error_map(oracle):
    K, Ecw = G["cam_K"], G["cam_Ecw"]
    Einv = np.linalg.inv(Ecw)
    depths = G["k4_depths"]
    for (r, t), want in zip(G["k4_pairs"], G["k4_err"]):
        got = oracle.pixel_error_map(K, Ecw, Einv, r, depths[r], t, depths[t], 4.0)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        m = ~np.isnan(want)
        assert m.sum() > 1000
        # float64 chain rounded to float32: identical up to BLAS-vs-scalar summation order
        assert np.allclose(got[m], want[m], rtol=0, atol=1e-6)
        assert (got[m].view(np.uint32) != want[m].view(np.uint32)).mean() < 0.02


def test_k4_confidence_maps(oracle):
    K, Ecw = G["cam_K"], G["cam_Ecw"]
    Einv = np.linalg.inv(Ecw)
    conf, count = oracle.confidence(G["k4_depths"], K, Ecw, Einv, 2, 4.0, 0.08)
    assert conf.dtype == np.float64 and count.dtype == np.int32
    assert np.array_equal(count, G["k4_count"])
    mism = (conf != G["k4_conf"]).mean()
    assert mism <= 1e-6, mism          # threshold flips only on last-ulp ties
    assert count.max() == 4 and count[0].max() == 2   # window clipped at the sequence ends
