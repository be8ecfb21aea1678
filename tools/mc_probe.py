#!/usr/bin/env python
"""Per-kernel device times of the marching-cubes path on a bench workload (MQ3D_TRACE events inside libmq3d).
Usage: python tools/mc_probe.py [--workload NAME] [--frames N] [--repeat K]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=5)
    a = ap.parse_args()
    import torch
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare
    cfg = dict(bench.WORKLOADS[a.workload])
    if a.frames:
        cfg["frames"] = a.frames
    dev = torch.device("cuda", 0)
    wl = bench.build_workload(cfg, dev, 0, 1)
    attrs = ("tsdf", "weight", "color") if cfg["color"] else ("tsdf", "weight")
    vbg = VoxelBlockGrid(attr_names=attrs, voxel_size=cfg["voxel"], block_count=cfg["block_count"], device=dev)
    lin, valid = depth_prepare(wl["raw"], wl["nears"], wl["fars"])
    st = vbg.integrate_sequence(lin, wl["K"], wl["Ewc"], cfg["depth_max"], cfg["trunc"], 1.0, frame_valid=valid,
                                colors=wl["colors"], color_intrinsics=wl["Kc"], batch_frames=256)
    out = vbg.extract_triangle_mesh_arrays(cfg["weight_thr"], with_colors=cfg["color"])
    print(f"blocks {st.num_blocks} vertices {out[0].shape[0]} triangles {out[2].shape[0]}", file=sys.stderr)
    os.environ["MQ3D_TRACE"] = "1"
    for _ in range(a.repeat):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = vbg.extract_triangle_mesh_arrays(cfg["weight_thr"], with_colors=cfg["color"])
        e1.record()
        torch.cuda.synchronize()
        print(f"[probe] extract_triangle_mesh_arrays {e0.elapsed_time(e1):.3f} ms", file=sys.stderr)


if __name__ == "__main__":
    main()
