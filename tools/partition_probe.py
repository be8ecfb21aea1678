#!/usr/bin/env python
"""One rank's share of a multi-GPU run on a single GPU: the grid is partitioned as rank 0 of `world` (owned blocks
only), so k_integrate sees the few-items-per-launch regime of N GPUs.  Prints the integrate time per shape variant
(MQ3D_INTEG_VARIANT).  Usage: python tools/partition_probe.py [world] [frames] [variants...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    variants = sys.argv[3:] or ["0", "32"]
    cfg = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
    cfg["frames"] = frames
    dev = torch.device("cuda", 0)
    wl = bench.build_workload(cfg, dev, 0, 1)
    lin, valid = depth_prepare(wl["raw"], wl["nears"], wl["fars"])
    for v in variants:
        os.environ["MQ3D_INTEG_VARIANT"] = v
        vbg = VoxelBlockGrid(voxel_size=cfg["voxel"], block_count=cfg["block_count"], device=dev)
        vbg.set_partition(0, world, 1, integrate_ghosts=False)
        ms, tot = [], []
        for _ in range(3):
            vbg.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st = vbg.integrate_sequence(lin, wl["K"], wl["Ewc"], cfg["depth_max"], cfg["trunc"], 1.0, frame_valid=valid,
                                        batch_frames=256)
            e1.record()
            torch.cuda.synchronize()
            ms.append(st.integrate_ms)
            tot.append(e0.elapsed_time(e1))
        print(f"world {world} variant {v}: k_integrate {min(ms):.3f} ms, whole sequence call {min(tot):.3f} ms, touch "
              f"{st.touch_ms:.3f} ms, {st.num_blocks} blocks, {st.block_visits} block visits, {st.blocks_loaded} residencies",
              file=sys.stderr)
        vbg.close()


if __name__ == "__main__":
    main()
