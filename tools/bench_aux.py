#!/usr/bin/env python
"""Satellite-kernel measurements (not the driver's bench line): K4 multi-view confidence (BASELINE
config #3 shape: +-5 neighbours, both eyes) and K6 colour-aligned raycast (1280x960 views of the
config #2 mesh), each next to the CPU oracle on a bounded sample.  Prints one JSON line per kernel."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import mq3d_b200  # noqa: E402,F401
from mq3d_b200 import synth, synth_gpu  # noqa: E402
from mq3d_b200.models import CoordinateSystem, Side, Transforms  # noqa: E402
from mq3d_b200.raycast import RaycastingScene  # noqa: E402
from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare, estimate_confidence  # noqa: E402


def cams(F, side):
    pos, quat = synth.eye_poses(F, side)
    tr = Transforms(CoordinateSystem.UNITY, pos, quat).convert_coordinate_system(CoordinateSystem.OPEN3D, True)
    fx, fy, cx, cy = synth.depth_intrinsics()
    K = np.zeros((F, 3, 3), np.float32)
    K[:, 0, 0], K[:, 1, 1], K[:, 2, 2] = fx, fy, 1.0
    K[:, 0, 2], K[:, 1, 2] = synth.DEPTH_W - cx, cy
    return K, tr.extrinsics_wc, tr.extrinsics_cw


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--range", type=int, default=5)
    ap.add_argument("--views", type=int, default=60)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on a bounded sample")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    # ---- K4 ---------------------------------------------------------------------------------------
    F, R = args.frames, args.range
    K, Ewc, Ecw = cams(F, Side.LEFT)
    raw = synth_gpu.render_depth(Ecw, dev)
    lin, valid = depth_prepare(raw, np.full(F, synth.NEAR), np.full(F, synth.FAR))
    Einv = np.linalg.inv(Ecw)
    ms, (conf, count) = timed(lambda: estimate_confidence(lin, K, Ecw, Einv, R, 4.0, 0.08, frame_valid=valid), 5, 2)
    pairs = sum(min(F, i + R + 1) - max(0, i - R) - 1 for i in range(F))
    px = 320 * 320
    line = {"kernel": "k_confidence", "frames": F, "range": R, "pairs": pairs, "ms": ms, "pairs_per_s": pairs / ms * 1e3,
            "gpx_pairs_per_s": pairs * px / ms * 1e-6, "hbm_min_gbs": (F * px * 16) / ms * 1e-6,
            "hbm_frac_of_measured_peak": (F * px * 16) / ms * 1e-6 / peak,
            "note": "minimum traffic = each depth frame read once + f64 conf + i32 count written (16 B/px/frame)"}
    if args.cpu:
        from oracle import oracle as orc
        n = 24
        t0 = time.perf_counter()
        oc, on = orc.confidence(lin[:n].cpu().numpy(), K[:n], Ecw[:n], Einv[:n], R, 4.0, 0.08)
        dt = time.perf_counter() - t0
        cp = sum(min(n, i + R + 1) - max(0, i - R) - 1 for i in range(n))
        line["cpu_oracle_pairs_per_s"] = cp / dt
        line["cpu_cores"] = os.cpu_count()
        line["cpu_sample"] = f"first {n} frames, OpenMP"
        c2, n2 = estimate_confidence(lin[:n].contiguous(), K[:n], Ecw[:n], Einv[:n], R, 4.0, 0.08)
        line["count_mismatch_vs_oracle"] = int((n2.cpu().numpy() != on).sum())
        line["conf_mismatch_vs_oracle"] = int((c2.cpu().numpy() != oc).sum())
    print(json.dumps(line))

    # ---- K6 ---------------------------------------------------------------------------------------
    F2 = 300
    K2, Ewc2, Ecw2 = cams(F2, Side.LEFT)
    raw2 = synth_gpu.render_depth(Ecw2, dev)
    lin2, valid2 = depth_prepare(raw2, np.full(F2, synth.NEAR), np.full(F2, synth.FAR))
    vbg = VoxelBlockGrid(voxel_size=0.01, block_count=60000, device=dev)
    vbg.integrate_sequence(lin2, K2, Ewc2, 4.0, 10.0, frame_valid=valid2)
    mesh = vbg.extract_triangle_mesh(1.5)
    torch.cuda.synchronize()
    builds = []
    for _ in range(4):          # first build pays module load + first-touch allocations
        t0 = time.perf_counter()
        scene = RaycastingScene(device=dev)
        scene.add_triangles(mesh)
        torch.cuda.synchronize()
        builds.append((time.perf_counter() - t0) * 1e3)
    build_ms = min(builds)
    Kc = np.array([[870.0, 0, 640.0], [0, 870.0, 480.0], [0, 0, 1.0]], np.float32)
    views = list(range(0, F2, max(1, F2 // args.views)))[: args.views]

    def render():
        out = None
        for i in views:
            rays = scene.create_rays_pinhole(Kc, Ewc2[i], width_px=1280, height_px=960)
            out = scene.cast_rays(rays)["t_hit"]
        return out
    ms, t_hit = timed(render, 3, 1)
    nrays = len(views) * 1280 * 960
    line = {"kernel": "k_cast_rays(+k_rays_pinhole)", "triangles": int(mesh.triangle.indices.shape[0]),
            "views": len(views), "ms_per_view": ms / len(views), "mrays_per_s": nrays / ms * 1e-3, "bvh_build_ms": build_ms, "bvh_build_ms_first_call": builds[0],
            "hit_fraction_last_view": float(torch.isfinite(t_hit).float().mean())}
    if args.cpu:
        from oracle import oracle as orc
        v = mesh.vertex.positions.cpu().numpy()
        t = mesh.triangle.indices.cpu().numpy()
        rays = orc.create_rays_pinhole(Kc.astype(np.float64), Ewc2[views[0]].astype(np.float64), 1280, 960)
        t0 = time.perf_counter()
        ref = orc.cast_rays(v, t, rays)
        dt = time.perf_counter() - t0
        line["cpu_oracle_mrays_per_s"] = 1280 * 960 / dt * 1e-6
        line["cpu_cores"] = os.cpu_count()
        got = scene.cast_rays(scene.create_rays_pinhole(Kc, Ewc2[views[0]], width_px=1280, height_px=960))["t_hit"].cpu().numpy()
        both = np.isfinite(ref) & np.isfinite(got)
        line["hit_miss_disagreement"] = float((np.isfinite(ref) != np.isfinite(got)).mean())
        line["t_hit_abs_err_p999"] = float(np.quantile(np.abs(ref[both] - got[both]), 0.999))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
