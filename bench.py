#!/usr/bin/env python
"""Hot-path benchmark: Quest depth (+RGB) frames -> TSDF voxel-block grid -> marching cubes.

Contract (driver): python bench.py --gpus N --steps K --warmup W [--impl reference]
prints ONE JSON line on rank 0.  A "step" = one pass of the hot path over the whole synthetic
sequence of the workload: grid reset -> K1 (linearise) -> K2+K3 (fused touch/integrate, colour) ->
K5 (marching cubes).  Workload at every N: BASELINE.json configs[1] (300 left-eye 320x320 depth
frames + 1280x960 RGB, voxel 0.01 m, 16^3 blocks, trunc x10, depth_max 4 m, weight_threshold 1.5).
N > 1: voxel blocks are hash-partitioned across ranks; every rank holds the frames (NCCL
broadcast at set-up), integrates the blocks it owns, fetches the one-block ghost shell from the owners'
pools over NVLink peer memory (or by NCCL exchange / redundant integration, --ghosts), extracts its part
of the mesh, and rank 0 gathers vertices / normals / colours / triangles -- strong scaling of one capture.

value      = depth frames/s with inputs resident in HBM (device-timed, CUDA events, max over ranks)
e2e        = same metric through the public host-buffer API: pinned host frames -> depth by DMA, colour
             read in place by the resampler (N > 1: 1/N per rank + NCCL all-gather) -> ... -> mesh on the host
roofline   = dominant kernel (k_integrate): algorithmic bytes (40 B/voxel-visit with
             colour, 16 B without, + frame images + keys) / its device time (CUDA events recorded
             inside libmq3d on the launching stream) vs MEASURED_PEAKS.json
cpu_baseline / --impl reference = the CPU oracle ("Open3D-0.19 semantics, CPU restatement"; open3d is
             not installable offline) on the host cores, bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]
    "quest300_rgb_v10mm": dict(frames=300, voxel=0.01, color=True, depth_max=4.0, trunc=10.0, weight_thr=1.5,
                               block_count=60000, sides=1),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case)
    "quest300_v20mm": dict(frames=300, voxel=0.02, color=False, depth_max=4.0, trunc=10.0, weight_thr=1.5,
                           block_count=20000, sides=1),
    # BASELINE.json configs[3] scaled to what a synthetic generator can hold: stereo, 5 mm
    "room_stereo_v5mm": dict(frames=1000, voxel=0.005, color=False, depth_max=4.0, trunc=10.0, weight_thr=1.5,
                             block_count=400000, sides=2),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        self.index = index
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def build_workload(cfg, device, rank, world):
    """Synthetic sequence on the device (rank 0 renders; other ranks receive it by NCCL broadcast)."""
    import torch
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import synth, synth_gpu
    from mq3d_b200.models import CoordinateSystem, Side, Transforms

    F = cfg["frames"]
    sides = [Side.LEFT, Side.RIGHT][: cfg["sides"]]
    Ks, Es, Ecws = [], [], []
    for side in sides:
        pos, quat = synth.eye_poses(F, side)
        tr = Transforms(CoordinateSystem.UNITY, pos, quat).convert_coordinate_system(CoordinateSystem.OPEN3D, True)
        fx, fy, cx, cy = synth.depth_intrinsics()
        K = np.zeros((F, 3, 3), np.float32)
        K[:, 0, 0], K[:, 1, 1], K[:, 2, 2] = fx, fy, 1.0
        K[:, 0, 2], K[:, 1, 2] = synth.DEPTH_W - cx, cy      # cx' = W - cx (o3d_utils.py:14-19)
        Ks.append(K)
        Es.append(tr.extrinsics_wc)
        Ecws.append(tr.extrinsics_cw)
    K = np.concatenate(Ks)
    Ewc = np.concatenate(Es)
    Ecw = np.concatenate(Ecws)
    n = len(K)
    raw = torch.empty((n, synth.DEPTH_H, synth.DEPTH_W), dtype=torch.float32, device=device)
    colors = torch.empty((n, synth.COLOR_H, synth.COLOR_W, 3), dtype=torch.uint8, device=device) if cfg["color"] else None
    if rank == 0:
        raw.copy_(synth_gpu.render_depth(Ecw, device))
        if colors is not None:
            for i in range(0, n, 50):
                colors[i:i + 50] = synth_gpu.render_color(Ecw[i:i + 50], device)
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(raw, 0)                       # the north star's "frame broadcast" (NCCL/NVLink)
        if colors is not None:
            dist.broadcast(colors, 0)
    Kc = None
    if colors is not None:
        Kc = np.zeros((n, 3, 3), np.float64)
        Kc[:, 0, 0] = Kc[:, 1, 1] = int(synth.COLOR_F)
        Kc[:, 0, 2], Kc[:, 1, 2], Kc[:, 2, 2] = synth.COLOR_W // 2, synth.COLOR_H // 2, 1.0
    nears = np.full(n, synth.NEAR)
    fars = np.full(n, synth.FAR)
    return dict(raw=raw, colors=colors, K=K, Ewc=Ewc, Kc=Kc, nears=nears, fars=fars, n=n)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import mq3d_b200  # noqa: F401
    from mq3d_b200.pipeline import IntegrationParams, extract_mesh_to_host, integrate_frames
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg = dict(WORKLOADS[args.workload])
    if args.frames:
        cfg["frames"] = args.frames
    wl = build_workload(cfg, device, rank, world)
    n = wl["n"]
    color = cfg["color"]
    vbg = VoxelBlockGrid(attr_names=("tsdf", "weight", "color") if color else ("tsdf", "weight"),
                         voxel_size=cfg["voxel"], block_count=cfg["block_count"], device=device)
    if world > 1:
        vbg.set_partition(rank, world, args.tile, integrate_ghosts=(args.ghosts == "integrate"))
    params = IntegrationParams(voxel_size=cfg["voxel"], block_count=cfg["block_count"], depth_max=cfg["depth_max"],
                               trunc_voxel_multiplier=cfg["trunc"], use_confidence_filtered_depth=False,
                               batch_frames=args.batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        """inputs resident in HBM"""
        vbg.reset()
        lin, valid = depth_prepare(wl["raw"], wl["nears"], wl["fars"])
        st = vbg.integrate_sequence(lin, wl["K"], wl["Ewc"], cfg["depth_max"], cfg["trunc"], 1.0, frame_valid=valid,
                                    colors=wl["colors"], color_intrinsics=wl["Kc"], batch_frames=args.batch)
        t_a = time.perf_counter()
        if world > 1 and args.ghosts == "exchange":
            from mq3d_b200.dist import exchange_ghosts
            exchange_ghosts(vbg, rank, world, timings=xch_phases)   # owners -> ghost shells, once, before extraction
            torch.cuda.synchronize()
        elif world > 1 and args.ghosts == "pull":
            from mq3d_b200.dist import fill_ghost_shell
            # ghost shells read straight from the owners' pools (NVLink peer memory); falls back to the NCCL
            # exchange on all ranks if CUDA IPC is unavailable
            ghost_mode_used[0] = fill_ghost_shell(vbg, rank, world, "pull")
            torch.cuda.synchronize()
        t_b = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        # Open3D's extract_triangle_mesh on a coloured grid also yields vertex colours
        out = vbg.extract_triangle_mesh_arrays(cfg["weight_thr"], with_colors=color)
        v, nrm, t, vc = out[0], out[1], out[2], (out[3] if color else None)
        e1.record()
        t_c = time.perf_counter()
        if world > 1:            # the north star's "final gather": per-rank meshes -> rank 0 over NCCL
            from mq3d_b200.dist import gather_mesh
            gather_mesh(v, nrm, t, dst=0, colors=vc)
            torch.cuda.synchronize()
        mgpu_ms["exchange"].append((t_b - t_a) * 1e3)
        mgpu_ms["gather"].append((time.perf_counter() - t_c) * 1e3)
        return st, (v, nrm, t, vc), (e0, e1)

    mgpu_ms = {"exchange": [], "gather": []}
    ghost_mode_used = [args.ghosts]
    xch_phases = {} if os.environ.get("MQ3D_TRACE") else None      # per-phase exchange times (diagnostics)

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats, mc_events = [], []
    ev0.record()
    mesh = None
    for _ in range(args.steps):
        mesh = None          # release the previous step's mesh before the next one is allocated
        st, mesh, mce = step_device()
        stats.append(st)
        mc_events.append(mce)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    mc_list = [a.elapsed_time(b) for a, b in mc_events]
    if os.environ.get("MQ3D_TRACE"):
        print("[bench] per-step mc ms:", [round(x, 3) for x in mc_list], file=sys.stderr)
        if xch_phases:
            print(f"[bench r{rank}] exchange phases, median ms/step:",
                  {k: round(float(np.median(v)), 3) for k, v in xch_phases.items()}, file=sys.stderr)
    mc_ms = float(np.mean(mc_list))
    st = stats[-1]
    integ_ms = float(np.mean([s.integrate_ms for s in stats]))
    touch_ms = float(np.mean([s.touch_ms for s in stats]))
    V, T = int(mesh[0].shape[0]), int(mesh[2].shape[0])

    # ---- end to end through the host-buffer API ---------------------------------------------------
    raw_host = wl["raw"].cpu().pin_memory()
    col_host = wl["colors"].cpu().pin_memory() if color else None

    def step_e2e():
        vbg.reset()
        if world == 1:
            integrate_frames(vbg, raw_host, wl["nears"], wl["fars"], wl["K"], wl["Ewc"], params, colors_host=col_host,
                             Kc=wl["Kc"])
            return extract_mesh_to_host(vbg, cfg["weight_thr"], with_colors=color)
        # N > 1: every rank uploads 1/N of each chunk over its own PCIe link, NCCL all-gather completes the
        # chunk over NVLink; per-rank meshes are gathered on rank 0, which reads the whole mesh back
        from mq3d_b200.dist import exchange_ghosts, gather_mesh
        integrate_frames(vbg, raw_host, wl["nears"], wl["fars"], wl["K"], wl["Ewc"], params, colors_host=col_host,
                         Kc=wl["Kc"], shard=(rank, world))
        if args.ghosts == "exchange":
            exchange_ghosts(vbg, rank, world)
        elif args.ghosts == "pull":
            from mq3d_b200.dist import fill_ghost_shell
            fill_ghost_shell(vbg, rank, world, "pull")
        out = vbg.extract_triangle_mesh_arrays(cfg["weight_thr"], with_colors=color)
        got = gather_mesh(out[0], out[1], out[2], dst=0, colors=out[3] if color else None)
        if rank != 0:
            torch.cuda.synchronize()
            return (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))
        from mq3d_b200.pipeline import to_host
        return to_host((got[0], got[1], got[2]) + ((got[4],) if color else ()))

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(e2e_steps):
        host_mesh = step_e2e()
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1) / e2e_steps
    h2d = raw_host.numel() * 4
    h2d_note = "depth frames by DMA"
    if color and world == 1:
        # zero-copy colour: the resampler reads the pinned frames in place; what crosses PCIe are the 32-byte
        # sectors of the sampled pixels.  Sampled columns are < 32 B apart, so count every touched row in full.
        from mq3d_b200 import synth
        fy, cy = float(wl["K"][0][1, 1]), float(wl["K"][0][1, 2])
        vf = wl["Kc"][0][1, 1] * ((np.arange(synth.DEPTH_H) - cy) / fy) + wl["Kc"][0][1, 2]
        rows = np.unique(np.round(vf[(vf >= 0) & (vf <= synth.COLOR_H - 1)]))
        h2d += n * len(rows) * synth.COLOR_W * 3
        h2d_note += (f" + colour read in place from pinned host memory by k_color_resample: {len(rows)} of "
                     f"{synth.COLOR_H} rows per frame touched (full frames would be {col_host.numel()} B)")
    elif color:
        h2d += col_host.numel()
        h2d_note += " + full colour frames, 1/N per rank by DMA, completed by NCCL all-gather"
    hv, ht = host_mesh[0], host_mesh[2]
    d2h = sum(a.nbytes for a in host_mesh)

    # ---- reduce over ranks (max time; sums of per-rank work) --------------------------------------
    t = torch.tensor([total_ms, e2e_ms, integ_ms, touch_ms, mc_ms], dtype=torch.float64, device=device)
    w = torch.tensor([st.block_visits, st.voxel_updates, st.blocks_loaded, st.num_blocks, V, T], dtype=torch.float64,
                     device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, integ_ms, touch_ms, mc_ms = t.tolist()
    visits_blocks, updates, loaded, nblocks, V, T = [int(x) for x in w.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if (hv.shape[0], ht.shape[0]) != (V, T):
        raise SystemExit(f"end-to-end mesh {hv.shape[0]}/{ht.shape[0]} differs from the device-resident run {V}/{T}")
    ms_per_step = total_ms / args.steps
    frames_per_s = n / (ms_per_step * 1e-3)
    bytes_per_visit = 40 if color else 16
    visits = visits_blocks * 4096
    img_bytes = n * (320 * 320 * 4 + (1280 * 960 * 3 if color else 0))
    algo_bytes = bytes_per_visit * visits + img_bytes + 12 * visits_blocks
    batched_bytes = bytes_per_visit * loaded * 4096 + img_bytes + 12 * loaded
    peak, peak_src = peaks()
    achieved = algo_bytes / (integ_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "integrate_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    out = {
        "metric": "depth frames/s (TSDF integrate + marching cubes)", "value": frames_per_s, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "frames": n, "depth": "320x320 f32 raw NDC",
                   "color": "1280x960 u8 RGB" if color else None, "voxel_size": cfg["voxel"], "block": "16^3",
                   "trunc_voxel_multiplier": cfg["trunc"], "depth_max": cfg["depth_max"],
                   "weight_threshold": cfg["weight_thr"], "batch_frames": args.batch,
                   "partition": f"tile-hash T={args.tile}, ghost shell by {ghost_mode_used[0]}" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (depth+RGB sequence > 126 MB); no explicit flush"},
        "gvoxel_updates_per_s": visits / (ms_per_step * 1e-3) / 1e9,
        "gvoxel_visits_per_s_integrate_kernel": visits / (integ_ms * 1e-3) / 1e9,
        "updated_voxel_fraction": updates / max(visits, 1),
        "mc_ms": mc_ms, "mesh": {"vertices": V, "triangles": T}, "active_blocks": nblocks,
        "kernel_ms": {"k_integrate": integ_ms, "k_touch": touch_ms, "mc_count+fill": mc_ms},
        "multi_gpu_ms_rank0": {k: float(np.median(v[-args.steps:])) for k, v in mgpu_ms.items()} if world > 1 else None,
        "roofline": {"bound": "hbm", "kernel": "k_integrate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                     "bytes_per_voxel_visit": bytes_per_visit,
                     "note": "algorithmic bytes = Open3D per-frame accounting (every voxel visit moves its "
                             "attributes once); the fused kernel keeps a block resident for a whole batch, so "
                             "DRAM traffic is lower than this and frac can exceed 1",
                     "achieved_batched": batched_bytes / (integ_ms * 1e-3) / 1e9,
                     "frac_batched": batched_bytes / (integ_ms * 1e-3) / 1e9 / peak,
                     "block_residencies": loaded, "block_visits": visits_blocks,
                     "launches_per_step": int(st.batches),
                     "algorithmic_bytes_per_launch": algo_bytes / max(int(st.batches), 1)},
        "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "h2d_note": h2d_note},
        # per step on each rank: reset fill (1) + K1 prepare/finalize (2) + per batch [colour resample] + touch +
        # sort + integrate + bitmap clear + MC neighbours/classify/scan/emit [+ colours]
        "gpu_launches": int(args.steps * (3 + st.batches * (5 if color else 4) + (5 if color else 4))),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(cfg, args.workload, sample_frames=args.cpu_frames)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_sample(cfg, sample_frames):
    """First `sample_frames` frames of the workload as host arrays (NumPy generator)."""
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import synth
    from mq3d_b200.models import CoordinateSystem, Side
    cap = synth.make_depth_capture(sample_frames, Side.LEFT)
    ds = cap.dataset
    tr = ds.transforms.convert_coordinate_system(CoordinateSystem.OPEN3D, is_camera=True)
    K = ds.get_intrinsic_matrices()
    K[:, 0, 2] = ds.widths - K[:, 0, 2]
    Ewc, Ecw = tr.extrinsics_wc, tr.extrinsics_cw
    colors = Kc = None
    if cfg["color"]:
        colors = np.stack([synth.make_color_frame(Ecw[i]) for i in range(sample_frames)])
        Kc = np.zeros((sample_frames, 3, 3))
        Kc[:, 0, 0] = Kc[:, 1, 1] = int(synth.COLOR_F)
        Kc[:, 0, 2], Kc[:, 1, 2], Kc[:, 2, 2] = synth.COLOR_W // 2, synth.COLOR_H // 2, 1.0
    return cap, K, Ewc, colors, Kc


def cpu_step(orc, cfg, sample):
    cap, K, Ewc, colors, Kc = sample
    ds = cap.dataset
    g = orc.Grid(cfg["voxel"], with_color=cfg["color"])
    for i in range(len(ds)):
        raw = cap.raw[i]
        if not orc.depth_valid(raw):
            continue
        d = orc.depth_to_linear(raw, ds.nears[i], ds.fars[i])
        keys = g.touch(d, K[i], Ewc[i], cfg["depth_max"], cfg["trunc"])
        g.integrate(keys, d, K[i], Ewc[i], cfg["depth_max"], cfg["trunc"],
                    color=None if colors is None else colors[i], Kc=None if Kc is None else Kc[i])
    mesh = g.extract_mesh(cfg["weight_thr"])
    if cfg["color"]:
        g.vertex_colors(mesh[3])          # the GPU step extracts vertex colours too
    return g


def cpu_baseline(cfg, workload, sample_frames=24, steps=1):
    from oracle import oracle as orc
    orc.build()
    sample = cpu_sample(cfg, sample_frames)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(orc, cfg, sample)
    dt = (time.perf_counter() - t0) / steps
    return {"value": sample_frames / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"first {sample_frames} of {cfg['frames']} frames of {workload} "
                      f"(integrate + marching cubes), OpenMP over {os.cpu_count()} host threads",
            "label": "Open3D-0.19 semantics, CPU restatement (open3d not installable offline)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    cfg = dict(WORKLOADS[args.workload])
    sample = cpu_sample(cfg, args.cpu_frames)
    for _ in range(args.warmup):
        cpu_step(orc, cfg, sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(orc, cfg, sample)
    dt = (time.perf_counter() - t0) / args.steps
    v = args.cpu_frames / dt
    sample_txt = (f"first {args.cpu_frames} of {cfg['frames']} frames of {args.workload} per step "
                  f"(integrate + marching cubes), OpenMP over {os.cpu_count()} host threads")
    print(json.dumps({
        "impl": "reference", "metric": "depth frames/s (TSDF integrate + marching cubes)", "value": v,
        "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "frames": cfg["frames"], "voxel_size": cfg["voxel"],
                   "color": "1280x960 u8 RGB" if cfg["color"] else None},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": sample_txt,
                         "label": "Open3D-0.19 semantics, CPU restatement (open3d not installable offline)"},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="quest300_rgb_v10mm", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override the number of frames per side")
    ap.add_argument("--batch", type=int, default=64, help="frames per block residency (<= 256)")
    ap.add_argument("--tile", type=int, default=1, help="partition super-tile edge in blocks (N > 1)")
    ap.add_argument("--ghosts", default="pull", choices=["pull", "exchange", "integrate"],
                    help="N > 1: 'integrate' = every rank also integrates its ghost shell (no exchange, the "
                         "north-star scheme); 'exchange' = owned blocks only + one packed NCCL ghost-block "
                         "exchange before MC; 'pull' = owned blocks only + ghost blocks copied straight from the "
                         "owners' pools over NVLink peer memory (falls back to 'exchange' without CUDA IPC)")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
