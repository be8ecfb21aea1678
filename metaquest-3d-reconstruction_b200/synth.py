"""Synthetic Quest-shaped captures.

Analytic scene (Unity world, Y up): room 5.0 x 3.0 x 4.0 m centred at (0,1.5,0), a sphere
r=0.5 m at (0.8,1.0,-0.5) and a 1.0 x 0.8 x 0.6 m box standing on the floor.  The head moves on a
circle (r=0.5 m, height 1.5 m) sweeping yaw with a +-15 deg sinusoidal pitch; eyes sit +-31.5 mm
along head-right.  Depth frames are 320x320 raw NDC (`d = 1 - near/z`, far = inf) sampled at
integer pixel coordinates, with multiplicative Gaussian noise and 2 % dropped pixels (raw 1.0 ->
linear 0).  Poses are emitted in Unity convention so they traverse the same
``convert_coordinate_system(OPEN3D, is_camera=True)`` the reference applies
(reconstruct_scene.py:48-51).

`write_project` lays the capture out exactly as the reference's DepthDataIO expects
(dataio/depth_data_io.py:166-261: `<side>_depth/<ts>.raw` + `<side>_depth_descriptors.csv`).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np
from scipy.spatial.transform import Rotation

from .models import CoordinateSystem, DepthDataset, CameraDataset, Side, Transforms

ROOM_MIN = np.array([-2.5, 0.0, -2.0])
ROOM_MAX = np.array([2.5, 3.0, 2.0])
SPHERE_C = np.array([0.8, 1.0, -0.5])
SPHERE_R = 0.5
BOX_MIN = np.array([-1.5, 0.0, 0.5])
BOX_MAX = np.array([-0.5, 0.8, 1.1])

DEPTH_W = DEPTH_H = 320
NEAR = 0.1
FAR = float("inf")
IPD_HALF = 0.0315
HEAD_RADIUS = 0.5
COLOR_W, COLOR_H = 1280, 960
COLOR_F = 870.0


@dataclass
class SyntheticCapture:
    """One eye's depth sequence, already in the dtypes the pipeline produces."""
    dataset: DepthDataset            # Unity-convention poses, as build_depth_dataset would emit
    raw: np.ndarray                  # [N,H,W] float32 raw NDC depth


def head_trajectory(n_frames: int, sweep_frames: int = 900):
    """Unity-convention head positions (N,3) and camera->world quaternions (N,4 xyzw)."""
    i = np.arange(n_frames, dtype=np.float64)
    yaw = 2.0 * np.pi * i / sweep_frames
    pitch = np.deg2rad(15.0) * np.sin(2.0 * np.pi * i / 120.0)
    # head centre travels on a circle, looking outward-ish (yaw follows the angle)
    # r = 0.5 m (SURVEY 8d proposed 0.8 m, but that path grazes the sphere: depths below the near
    # plane make raw NDC negative and is_depth_map_valid rejects ~50 of every 300 frames)
    pos = np.stack([HEAD_RADIUS * np.sin(yaw), np.full_like(yaw, 1.5), HEAD_RADIUS * np.cos(yaw)], axis=1)
    # Unity: +Y up, left-handed; yaw about +Y then pitch about local +X
    rot = Rotation.from_euler("y", yaw[:, None]) * Rotation.from_euler("x", pitch[:, None])
    return pos, rot


def eye_poses(n_frames: int, side: Side, sweep_frames: int = 900):
    pos, rot = head_trajectory(n_frames, sweep_frames)
    off = np.array([-IPD_HALF if side == Side.LEFT else IPD_HALF, 0.0, 0.0])
    return pos + rot.apply(off), rot.as_quat()


def timestamps_ms(n_frames: int, t0: int = 1_700_000_000_000) -> np.ndarray:
    return t0 + np.round(33.333 * np.arange(n_frames)).astype(np.int64)


def _scene_o3d():
    """Scene primitives in Open3D world (Unity world with z negated)."""
    f = np.array([1.0, 1.0, -1.0])
    rmin, rmax = np.minimum(ROOM_MIN * f, ROOM_MAX * f), np.maximum(ROOM_MIN * f, ROOM_MAX * f)
    bmin, bmax = np.minimum(BOX_MIN * f, BOX_MAX * f), np.maximum(BOX_MIN * f, BOX_MAX * f)
    return rmin, rmax, SPHERE_C * f, SPHERE_R, bmin, bmax


def trace_scene(origin: np.ndarray, dirs: np.ndarray):
    """Closest hit of rays origin + t*dirs (dirs [...,3], unnormalised) with the scene.

    Returns (t, surface_id) with surface_id 0..5 room walls (-x,+x,-y,+y,-z,+z), 6 sphere, 7 box.
    """
    rmin, rmax, sc, sr, bmin, bmax = _scene_o3d()
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
        # room (we are inside): exit distance per axis
        t_hi = np.where(dirs > 0, (rmax - origin) * inv, (rmin - origin) * inv)
        t_hi = np.where(dirs == 0, np.inf, t_hi)
        ax = np.argmin(t_hi, axis=-1)
        t = np.take_along_axis(t_hi, ax[..., None], axis=-1)[..., 0]
        sid = 2 * ax + (np.take_along_axis(dirs, ax[..., None], axis=-1)[..., 0] > 0)
        # sphere
        oc = origin - sc
        a = np.sum(dirs * dirs, axis=-1)
        b = 2.0 * np.sum(dirs * oc, axis=-1)
        c = float(np.dot(oc, oc) - sr * sr)
        disc = b * b - 4 * a * c
        ts = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
        ts = np.where(ts > 0, ts, np.inf)
        hit = ts < t
        t = np.where(hit, ts, t)
        sid = np.where(hit, 6, sid)
        # box (from outside): slab entry
        t0 = (bmin - origin) * inv
        t1 = (bmax - origin) * inv
        tn = np.max(np.minimum(t0, t1), axis=-1)
        tf = np.min(np.maximum(t0, t1), axis=-1)
        tb = np.where((tn <= tf) & (tn > 0), tn, np.inf)
        hit = tb < t
        t = np.where(hit, tb, t)
        sid = np.where(hit, 7, sid)
    return t, sid


def depth_intrinsics(width=DEPTH_W, height=DEPTH_H, tangents=(1.0, 1.0, 1.0, 1.0)):
    """fx,fy,cx,cy as utils/depth_utils.py:4-18 computes them from the FOV tangents (L,R,T,D)."""
    left, right, top, bottom = tangents
    fx = width / (right + left)
    fy = height / (top + bottom)
    cx = width * right / (right + left)
    cy = height * top / (top + bottom)
    return fx, fy, cx, cy


def make_depth_capture(n_frames: int, side: Side = Side.LEFT, noise: float = 0.005,
                       dropout: float = 0.02, width: int = DEPTH_W, height: int = DEPTH_H,
                       sweep_frames: int = 900, seed_offset: int = 0) -> SyntheticCapture:
    pos_u, quat_u = eye_poses(n_frames, side, sweep_frames)
    ts = timestamps_ms(n_frames)
    fx, fy, cx, cy = depth_intrinsics(width, height)
    tr_u = Transforms(CoordinateSystem.UNITY, pos_u, quat_u)
    e_cw = tr_u.convert_coordinate_system(CoordinateSystem.OPEN3D, is_camera=True).extrinsics_cw
    # pipeline intrinsics: cx' = W - cx (o3d_utils.py:14-19)
    u = np.arange(width, dtype=np.float64)
    v = np.arange(height, dtype=np.float64)
    dx = (u - (width - cx)) / fx
    dy = (v - cy) / fy
    dirs_cam = np.stack(np.broadcast_arrays(dx[None, :], dy[:, None], np.ones((height, width))), axis=-1)
    raw = np.empty((n_frames, height, width), dtype=np.float32)
    for i in range(n_frames):
        e = e_cw[i].astype(np.float64)
        dirs = dirs_cam @ e[:3, :3].T
        z, _ = trace_scene(e[:3, 3], dirs)          # dirs have camera-z = 1 -> t is z-depth
        rng = np.random.Generator(np.random.PCG64(int(ts[i]) + seed_offset + (0 if side == Side.LEFT else 7)))
        if noise > 0:
            z = z * (1.0 + noise * rng.standard_normal(z.shape))
        d = 1.0 - NEAR / z
        if dropout > 0:
            d = np.where(rng.random(z.shape) < dropout, 1.0, d)
        raw[i] = d.astype(np.float32)
    n = n_frames
    ds = DepthDataset(
        directory_relative_path=f"{side.value}_depth",
        image_file_names=np.array([f"{t}.raw" for t in ts]),
        timestamps=ts,
        fx=np.full(n, fx), fy=np.full(n, fy), cx=np.full(n, cx), cy=np.full(n, cy),
        transforms=tr_u,
        widths=np.full(n, width), heights=np.full(n, height),
        nears=np.full(n, NEAR), fars=np.full(n, FAR),
    )
    return SyntheticCapture(dataset=ds, raw=raw)


_HUES = np.array([[200, 80, 80], [80, 200, 80], [90, 90, 210], [210, 210, 90],
                  [210, 90, 210], [90, 210, 210], [240, 160, 60], [150, 110, 70]], dtype=np.float64)


def make_color_frame(e_cw: np.ndarray, width: int = COLOR_W, height: int = COLOR_H,
                     f: float = COLOR_F) -> np.ndarray:
    """Procedural RGB u8 [H,W,3] seen from pose e_cw (Open3D convention): per-surface hue times a
    25 cm checker."""
    u = np.arange(width, dtype=np.float64)
    v = np.arange(height, dtype=np.float64)
    cx, cy = width / 2.0, height / 2.0
    dirs_cam = np.stack(np.broadcast_arrays(((u - cx) / f)[None, :], ((v - cy) / f)[:, None],
                                            np.ones((height, width))), axis=-1)
    e = e_cw.astype(np.float64)
    dirs = dirs_cam @ e[:3, :3].T
    t, sid = trace_scene(e[:3, 3], dirs)
    p = e[:3, 3] + t[..., None] * dirs
    checker = (np.floor(p[..., 0] * 4) + np.floor(p[..., 1] * 4) + np.floor(p[..., 2] * 4)) % 2
    rgb = _HUES[sid] * (0.55 + 0.45 * checker[..., None])
    return np.clip(rgb, 0, 255).astype(np.uint8)


def write_project(project_dir: Path, n_frames: int, sides=(Side.LEFT, Side.RIGHT), **kw) -> dict:
    """Write a QRC-layout project directory; returns {side: SyntheticCapture}."""
    project_dir = Path(project_dir)
    out = {}
    for side in sides:
        cap = make_depth_capture(n_frames, side, **kw)
        ddir = project_dir / f"{side.value}_depth"
        ddir.mkdir(parents=True, exist_ok=True)
        ds = cap.dataset
        rows = ["timestamp_ms,width,height,near_z,far_z,fov_left_angle_tangent,fov_right_angle_tangent,"
                "fov_top_angle_tangent,fov_down_angle_tangent,create_pose_location_x,"
                "create_pose_location_y,create_pose_location_z,create_pose_rotation_x,"
                "create_pose_rotation_y,create_pose_rotation_z,create_pose_rotation_w"]
        for i in range(len(ds)):
            cap.raw[i].astype("<f4").tofile(ddir / f"{int(ds.timestamps[i])}.raw")
            p, q = ds.transforms.positions[i], ds.transforms.rotations[i]
            rows.append(",".join([str(int(ds.timestamps[i])), str(int(ds.widths[i])), str(int(ds.heights[i])),
                                  repr(NEAR), "inf", "1.0", "1.0", "1.0", "1.0"]
                                 + [repr(float(x)) for x in p] + [repr(float(x)) for x in q]))
        (project_dir / f"{side.value}_depth_descriptors.csv").write_text("\n".join(rows) + "\n")
        out[side] = cap
    return out


def make_color_dataset(n_frames: int, side: Side = Side.LEFT, sweep_frames: int = 900) -> CameraDataset:
    """Colour camera sequence co-located with the depth eye (integer intrinsics)."""
    pos_u, quat_u = eye_poses(n_frames, side, sweep_frames)
    ts = timestamps_ms(n_frames)
    n = n_frames
    return CameraDataset(
        directory_relative_path=f"{side.value}_camera_rgb",
        image_file_names=np.array([f"{t}.png" for t in ts]),
        timestamps=ts,
        fx=np.full(n, int(COLOR_F)), fy=np.full(n, int(COLOR_F)),
        cx=np.full(n, COLOR_W // 2), cy=np.full(n, COLOR_H // 2),
        transforms=Transforms(CoordinateSystem.UNITY, pos_u, quat_u),
        widths=np.full(n, COLOR_W), heights=np.full(n, COLOR_H),
    )
