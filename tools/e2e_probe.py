#!/usr/bin/env python
"""Where the end-to-end step of the default workload spends its time (host wall clock with synchronisation at phase
boundaries): uploads alone, integrate_frames, extraction, mesh read-back."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from mq3d_b200.pipeline import IntegrationParams, extract_mesh_to_host, integrate_frames, to_host
    from mq3d_b200.vbg import VoxelBlockGrid
    cfg = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
    if len(sys.argv) > 1:
        cfg["frames"] = int(sys.argv[1])
    gated = None if len(sys.argv) <= 2 else bool(int(sys.argv[2]))
    dev = torch.device("cuda", 0)
    wl = bench.build_workload(cfg, dev, 0, 1)
    vbg = VoxelBlockGrid(voxel_size=cfg["voxel"], block_count=cfg["block_count"], device=dev)
    raw_host = wl["raw"].cpu().pin_memory()
    params = IntegrationParams(voxel_size=cfg["voxel"], block_count=cfg["block_count"], depth_max=cfg["depth_max"],
                               trunc_voxel_multiplier=cfg["trunc"], use_confidence_filtered_depth=False, batch_frames=256)

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for it in range(4):
        t0 = sync()
        vbg.reset()
        t1 = sync()
        st = integrate_frames(vbg, raw_host, wl["nears"], wl["fars"], wl["K"], wl["Ewc"], params, gated=gated)
        t2 = sync()
        out = vbg.extract_triangle_mesh_arrays(cfg["weight_thr"])
        t3 = sync()
        host = to_host(out)
        t4 = sync()
        print(f"[e2e] reset {1e3*(t1-t0):.2f} ms  integrate_frames {1e3*(t2-t1):.2f} ms (k_integrate {st.integrate_ms:.2f}, "
              f"touch {st.touch_ms:.2f})  extract {1e3*(t3-t2):.2f} ms  to_host {1e3*(t4-t3):.2f} ms "
              f"({sum(a.nbytes for a in host)/1e6:.0f} MB)", file=sys.stderr)
    # uploads alone, chunked as integrate_frames does
    t0 = sync()
    keep = [raw_host[i:i + 256].to(dev, non_blocking=True) for i in range(0, raw_host.shape[0], 256)]
    t1 = sync()
    print(f"[e2e] chunked uploads alone: {1e3*(t1-t0):.2f} ms for {raw_host.numel()*4/1e9:.2f} GB", file=sys.stderr)


if __name__ == "__main__":
    main()
