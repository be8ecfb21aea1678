#!/usr/bin/env python
"""Hot-path benchmark: Quest depth (+RGB) frames -> TSDF voxel-block grid -> marching cubes.

Contract (driver): python bench.py --gpus N --steps K --warmup W [--impl reference]
prints ONE JSON line on rank 0.  A "step" = one pass of the hot path over the whole synthetic
sequence of the workload: grid reset -> K1 (linearise) -> K2+K3 (fused touch/integrate) -> K5 (marching
cubes).  Default workload at every N: BASELINE.json configs[3] at full size -- room-scale 5 mm grid, 5000
stereo frames = 10 000 depth images 320x320, LEFT sequence then RIGHT into one grid
(reconstruct_scene.py:64-81), 16^3 blocks, trunc x10, depth_max 4 m, weight_threshold 1.5.
N > 1: voxel blocks are hash-partitioned across ranks; every rank holds the frames (each rank
renders 1/N, NCCL all-gather at set-up = the north star's frame broadcast), integrates the blocks it owns,
fetches the one-block ghost shell from the owners' pools over NVLink peer memory (or by NCCL exchange /
redundant integration, --ghosts), extracts its part of the mesh, and rank 0 gathers vertices / normals /
triangles -- strong scaling of one capture.  Other workloads (--workload): configs[1] quest300_rgb_v10mm,
configs[0] quest300_v20mm, configs[4] batch_v10mm (one independent 2000-frame capture per GPU, weak scaling).

value      = depth frames/s with inputs resident in HBM (device-timed, CUDA events, max over ranks)
e2e        = same metric through the public host-buffer API: pinned host frames -> chunked DMA on a copy
             stream (N > 1: 1/N per rank + NCCL all-gather) -> ... -> mesh on the host
roofline   = dominant kernel (k_integrate).  The kernel keeps a block resident for a whole batch of frames,
             so its DRAM traffic is one load + one store per block residency, not per voxel visit: `achieved`
             is that batched algorithmic traffic over the kernel's device time (CUDA events recorded inside
             libmq3d on the launching stream) and `frac` its share of the measured HBM peak -- a small
             number, because the kernel is bound by instruction issue (`limiter`, `issue_slot_frac` from the
             committed ncu capture).  Open3D's per-visit accounting (16 B / 40 B per voxel visit)
             is kept as `throughput_equiv_gbs` only.
roofline_mc= marching cubes (classify + scan + emit), bytes per SURVEY 8d, vs the same HBM peak
aux        = (N = 1) K4 confidence, K6 raycast and BASELINE configs #0 / #1 / #4, each with its own figures
cpu_baseline / --impl reference = the CPU oracle ("Open3D-0.19 semantics, CPU restatement"; open3d is
             not installable offline) on the host cores, OpenMP team size set explicitly, on a bounded
             strided sample of the SAME workload.
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

METRIC = "depth frames/s (TSDF integrate + marching cubes)"
CPU_LABEL = "Open3D-0.19 semantics, CPU restatement (open3d not installable offline)"

WORKLOADS = {
    # BASELINE.json configs[3], full size: 5000 stereo frames (10 000 depth images), room-scale 5 mm grid
    "room5000_stereo_v5mm": dict(frames=5000, voxel=0.005, color=False, depth_max=4.0, trunc=10.0, weight_thr=1.5,
                                 block_count=1 << 17, sides=2, cpu_frames=48),
    # BASELINE.json configs[1]
    "quest300_rgb_v10mm": dict(frames=300, voxel=0.01, color=True, depth_max=4.0, trunc=10.0, weight_thr=1.5,
                               block_count=60000, sides=1, cpu_frames=24),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case)
    "quest300_v20mm": dict(frames=300, voxel=0.02, color=False, depth_max=4.0, trunc=10.0, weight_thr=1.5,
                           block_count=20000, sides=1, cpu_frames=150),
    # BASELINE.json configs[4]: independent captures, one per GPU (batch_run_full_pipeline.py:195-208), weak scaling
    "batch_v10mm": dict(frames=2000, voxel=0.01, color=False, depth_max=4.0, trunc=10.0, weight_thr=1.5,
                        block_count=60000, sides=1, cpu_frames=64, replicas=True),
}
DEFAULT_WORKLOAD = "room5000_stereo_v5mm"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_figure(name, workload):
    """Per-workload figure derived from a committed ncu capture (profiles/<name>.json), or None."""
    p = os.path.join(ROOT, "profiles", name + ".json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


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


def cameras(cfg, capture=0):
    """Pipeline cameras of the whole sequence, LEFT frames then RIGHT frames: K float32 [n,3,3] with the
    mirrored principal point (o3d_utils.py:14-19), world->camera and camera->world float32 [n,4,4].
    `capture` shifts the start of the head trajectory (independent captures of the batch workload)."""
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import synth
    from mq3d_b200.models import CoordinateSystem, Side, Transforms
    F = cfg["frames"]
    Ks, Es, Ecws = [], [], []
    for side in [Side.LEFT, Side.RIGHT][: cfg["sides"]]:
        pos, quat = synth.eye_poses(F + 113 * capture, side)
        pos, quat = pos[113 * capture:], quat[113 * capture:]
        tr = Transforms(CoordinateSystem.UNITY, pos, quat).convert_coordinate_system(CoordinateSystem.OPEN3D, True)
        fx, fy, cx, cy = synth.depth_intrinsics()
        K = np.zeros((F, 3, 3), np.float32)
        K[:, 0, 0], K[:, 1, 1], K[:, 2, 2] = fx, fy, 1.0
        K[:, 0, 2], K[:, 1, 2] = synth.DEPTH_W - cx, cy      # cx' = W - cx (o3d_utils.py:14-19)
        Ks.append(K)
        Es.append(tr.extrinsics_wc)
        Ecws.append(tr.extrinsics_cw)
    return np.concatenate(Ks), np.concatenate(Es), np.concatenate(Ecws)


def color_intrinsics(n):
    from mq3d_b200 import synth
    Kc = np.zeros((n, 3, 3), np.float64)
    Kc[:, 0, 0] = Kc[:, 1, 1] = int(synth.COLOR_F)
    Kc[:, 0, 2], Kc[:, 1, 2], Kc[:, 2, 2] = synth.COLOR_W // 2, synth.COLOR_H // 2, 1.0
    return Kc


def build_workload(cfg, device, rank, world, capture=0, shared=True):
    """Synthetic sequence on the device.  shared: one capture for all ranks -- every rank renders a contiguous
    1/world of the frames and an NCCL all-gather completes the sequence on every rank (the north star's frame
    broadcast).  Not shared (batch workload): this rank's own capture."""
    import torch
    from mq3d_b200 import synth, synth_gpu
    K, Ewc, Ecw = cameras(cfg, capture)
    n = len(K)
    color = cfg["color"]
    if shared and world > 1:
        import torch.distributed as dist
        per = -(-n // world)
        raw = torch.zeros((per * world, synth.DEPTH_H, synth.DEPTH_W), dtype=torch.float32, device=device)
        a, b = min(n, rank * per), min(n, (rank + 1) * per)
        if b > a:
            synth_gpu.render_depth(Ecw[a:b], device, first_frame=a, out=raw[a:b])
        dist.all_gather_into_tensor(raw, raw[rank * per:(rank + 1) * per].clone())
        raw = raw[:n]
        colors = None
        if color:
            colors = torch.zeros((per * world, synth.COLOR_H, synth.COLOR_W, 3), dtype=torch.uint8, device=device)
            for i in range(a, b, 50):
                colors[i:min(b, i + 50)] = synth_gpu.render_color(Ecw[i:min(b, i + 50)], device)
            dist.all_gather_into_tensor(colors, colors[rank * per:(rank + 1) * per].clone())
            colors = colors[:n]
    else:
        raw = synth_gpu.render_depth(Ecw, device, seed=1234 + 100003 * capture)
        colors = None
        if color:
            colors = torch.empty((n, synth.COLOR_H, synth.COLOR_W, 3), dtype=torch.uint8, device=device)
            for i in range(0, n, 50):
                colors[i:i + 50] = synth_gpu.render_color(Ecw[i:i + 50], device)
    return dict(raw=raw, colors=colors, K=K, Ewc=Ewc, Ecw=Ecw, Kc=color_intrinsics(n) if color else None,
                nears=np.full(n, synth.NEAR), fars=np.full(n, synth.FAR), n=n)


_P = (1 << 31) - 1


def block_checksum(vbg, owned_by=None):
    """(blocks, checksum mod 2^31-1) over the grid's blocks -- position-weighted sum of the tsdf and weight bit
    patterns of every voxel, mixed with the block key; order-independent.  owned_by=(rank, world, tile): only
    the blocks that rank owns.  Summed over the ranks of a partition it must equal the single-grid value."""
    import torch
    from mq3d_b200.dist import tile_owner_np
    keys, tsdf, weight, _ = vbg.export_blocks()
    if owned_by is not None:
        rank, world, tile = owned_by
        own = torch.from_numpy(tile_owner_np(keys.cpu().numpy(), world, tile) == rank).to(keys.device)
        keys, tsdf, weight = keys[own], tsdf[own], weight[own]
    n = int(keys.shape[0])
    dev = keys.device
    idx = torch.arange(4096, dtype=torch.int64, device=dev)
    m1 = (idx * 2654435761 + 1) % _P
    m2 = (idx * 40503 + 977) % _P
    k64 = keys.to(torch.int64)
    kmix = (((k64[:, 0] * 73856093) ^ (k64[:, 1] * 19349663) ^ (k64[:, 2] * 83492791)) % _P + 1) % _P
    total = torch.zeros((), dtype=torch.int64, device=dev)
    for i in range(0, n, 4096):
        t = (tsdf[i:i + 4096].reshape(-1, 4096).view(torch.int32).to(torch.int64) & 0xFFFFFFFF) % _P
        w = (weight[i:i + 4096].reshape(-1, 4096).view(torch.int32).to(torch.int64) & 0xFFFFFFFF) % _P
        s = (((t * m1) % _P).sum(1) + ((w * m2) % _P).sum(1)) % _P
        total = (total + ((s * kmix[i:i + 4096]) % _P).sum()) % _P
    return n, int(total.item())


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import mq3d_b200  # noqa: F401
    from mq3d_b200.pipeline import IntegrationParams, extract_mesh_to_host, integrate_frames, to_host
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
    replicas = bool(cfg.get("replicas"))           # independent captures: no partition, no collectives
    sharded = world > 1 and not replicas
    wl = build_workload(cfg, device, rank, world, capture=rank if replicas else 0, shared=not replicas)
    n = wl["n"]
    color = cfg["color"]
    attrs = ("tsdf", "weight", "color") if color else ("tsdf", "weight")
    vbg = VoxelBlockGrid(attr_names=attrs, voxel_size=cfg["voxel"], block_count=cfg["block_count"], device=device)
    if sharded:
        vbg.set_partition(rank, world, args.tile, integrate_ghosts=(args.ghosts == "integrate"))
    params = IntegrationParams(voxel_size=cfg["voxel"], block_count=cfg["block_count"], depth_max=cfg["depth_max"],
                               trunc_voxel_multiplier=cfg["trunc"], use_confidence_filtered_depth=False,
                               batch_frames=args.batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ghost_mode_used = [args.ghosts]

    def fill_ghosts():
        if not sharded or args.ghosts == "integrate":
            return
        from mq3d_b200.dist import exchange_ghosts, fill_ghost_shell
        if args.ghosts == "exchange":
            exchange_ghosts(vbg, rank, world)       # owners -> ghost shells, once, before extraction
        else:
            # ghost shells read straight from the owners' pools (NVLink peer memory); falls back to the NCCL
            # exchange on all ranks if CUDA IPC is unavailable
            ghost_mode_used[0] = fill_ghost_shell(vbg, rank, world, "pull")

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step_device():
        """inputs resident in HBM"""
        e = [ev()]
        vbg.reset()
        lin, valid = depth_prepare(wl["raw"], wl["nears"], wl["fars"])
        e.append(ev())
        st = vbg.integrate_sequence(lin, wl["K"], wl["Ewc"], cfg["depth_max"], cfg["trunc"], 1.0, frame_valid=valid,
                                    colors=wl["colors"], color_intrinsics=wl["Kc"], batch_frames=args.batch)
        del lin
        e.append(ev())
        fill_ghosts()
        e.append(ev())
        # Open3D's extract_triangle_mesh on a coloured grid also yields vertex colours
        out = vbg.extract_triangle_mesh_arrays(cfg["weight_thr"], with_colors=color)
        v, nrm, t, vc = out[0], out[1], out[2], (out[3] if color else None)
        e.append(ev())
        if sharded:            # the north star's "final gather": per-rank meshes -> rank 0 over NCCL
            from mq3d_b200.dist import gather_mesh
            gather_mesh(v, nrm, t, dst=0, colors=vc)
        e.append(ev())
        return st, (v, nrm, t, vc), e

    # ---- multi-GPU parity: the partitioned grids, summed, are the single-GPU grid ---------------------
    parity = None
    if sharded:
        from mq3d_b200.dist import gather_mesh
        st0, mesh0, _ = step_device()
        nb, cs = block_checksum(vbg, owned_by=(rank, world, args.tile))
        red = torch.tensor([nb, cs, int(mesh0[2].shape[0])], dtype=torch.int64, device=device)
        dist.all_reduce(red)
        whole = gather_mesh(mesh0[0], mesh0[1], mesh0[2], dst=0)          # the concatenated per-rank meshes on rank 0
        del mesh0
        if rank == 0:
            from mq3d_b200.geometry import TriangleMesh
            from mq3d_b200.meshfilter import filter_mesh_components_device
            ref = VoxelBlockGrid(attr_names=attrs, voxel_size=cfg["voxel"], block_count=cfg["block_count"], device=device)
            lin, valid = depth_prepare(wl["raw"], wl["nears"], wl["fars"])
            ref.integrate_sequence(lin, wl["K"], wl["Ewc"], cfg["depth_max"], cfg["trunc"], 1.0, frame_valid=valid,
                                   colors=wl["colors"], color_intrinsics=wl["Kc"], batch_frames=args.batch)
            del lin
            nb1, cs1 = block_checksum(ref)
            m1 = ref.extract_triangle_mesh_arrays(cfg["weight_thr"])
            v1, t1 = int(m1[0].shape[0]), int(m1[2].shape[0])
            # vertices on ghost edges are emitted by every rank that references them: the device mesh filter's weld
            # (identical coordinates -> first occurrence) turns the concatenation into the single-GPU vertex set
            welded, _ = filter_mesh_components_device(TriangleMesh(whole[0], whole[2], whole[1]), 1)
            vw, tw = int(welded.vertex.positions.shape[0]), int(welded.triangle.indices.shape[0])
            # (the single-grid mesh goes through the same weld: a zero crossing exactly on a voxel corner yields the same
            # position on up to three lattice edges)
            single, _ = filter_mesh_components_device(TriangleMesh(m1[0], m1[2], m1[1]), 1)
            v1w = int(single.vertex.positions.shape[0])
            ref.close()
            del ref, m1, welded, single
            parity = {"blocks_single": nb1, "blocks_sharded": int(red[0]), "checksum_single": cs1,
                      "checksum_sharded": int(red[1]) % _P, "triangles_single": t1, "triangles_sharded": int(red[2]),
                      "vertices_single": v1, "vertices_single_welded": v1w, "vertices_gathered": int(whole[0].shape[0]),
                      "vertices_welded": vw,
                      "triangles_welded": tw,
                      "what": "tsdf + weight bit patterns of every owned block (position- and key-weighted sum "
                              "mod 2^31-1) summed over ranks vs one unpartitioned grid on rank 0; gathered mesh "
                              "welded by the device mesh filter vs the single-grid mesh"}
            parity["equal"] = (parity["blocks_single"] == parity["blocks_sharded"]
                               and parity["checksum_single"] == parity["checksum_sharded"]
                               and parity["triangles_single"] == parity["triangles_sharded"] == tw and v1w == vw)
            if not parity["equal"]:
                raise SystemExit(f"multi-GPU parity check failed: {parity}")
        del whole
        torch.cuda.empty_cache()

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats, phase_events = [], []
    ev0.record()
    mesh = None
    for _ in range(args.steps):
        mesh = None          # release the previous step's mesh before the next one is allocated
        st, mesh, pe = step_device()
        stats.append(st)
        phase_events.append(pe)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    phases = np.array([[a.elapsed_time(b) for a, b in zip(pe[:-1], pe[1:])] for pe in phase_events]).mean(0)
    k1_ms, seq_ms, ghost_ms, mc_ms, gather_ms = [float(x) for x in phases]
    st = stats[-1]
    integ_ms = float(np.mean([s.integrate_ms for s in stats]))
    touch_ms = float(np.mean([s.touch_ms for s in stats]))
    V, T = int(mesh[0].shape[0]), int(mesh[2].shape[0])

    # ---- end to end through the host-buffer API ---------------------------------------------------
    raw_host = wl["raw"].cpu().pin_memory()
    col_host = wl["colors"].cpu().pin_memory() if color else None

    def step_e2e():
        vbg.reset()
        if not sharded:
            integrate_frames(vbg, raw_host, wl["nears"], wl["fars"], wl["K"], wl["Ewc"], params, colors_host=col_host,
                             Kc=wl["Kc"])
            return extract_mesh_to_host(vbg, cfg["weight_thr"], with_colors=color)
        # N > 1: every rank uploads 1/N of each chunk over its own PCIe link, NCCL all-gather completes the
        # chunk over NVLink; per-rank meshes are gathered on rank 0, which reads the whole mesh back
        from mq3d_b200.dist import gather_mesh
        integrate_frames(vbg, raw_host, wl["nears"], wl["fars"], wl["K"], wl["Ewc"], params, colors_host=col_host,
                         Kc=wl["Kc"], shard=(rank, world))
        fill_ghosts()
        out = vbg.extract_triangle_mesh_arrays(cfg["weight_thr"], with_colors=color)
        got = gather_mesh(out[0], out[1], out[2], dst=0, colors=out[3] if color else None)
        if rank != 0:
            torch.cuda.synchronize()
            return (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))
        return to_host((got[0], got[1], got[2]) + ((got[4],) if color else ()))

    host_mesh = None
    for _ in range(min(args.warmup, 3)):
        host_mesh = None      # the previous step's mesh is released first, as in the device-resident loop: the pinned
        host_mesh = step_e2e()    # staging of the read-back is then reused instead of being allocated a second time
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_marks = [f0]
    for _ in range(e2e_steps):
        host_mesh = None
        host_mesh = step_e2e()
        e2e_marks.append(ev())
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1) / e2e_steps
    e2e_each = [round(a.elapsed_time(b), 3) for a, b in zip(e2e_marks[:-1], e2e_marks[1:])]
    h2d = raw_host.numel() * 4
    h2d_note = "depth frames by DMA from pinned memory, chunked on a copy stream under K1/K2/K3"
    if color and not sharded:
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
    t = torch.tensor([total_ms, e2e_ms, integ_ms, touch_ms, mc_ms, k1_ms, seq_ms, ghost_ms, gather_ms],
                     dtype=torch.float64, device=device)
    w = torch.tensor([st.block_visits, st.voxel_updates, st.blocks_loaded, st.num_blocks, V, T, integ_ms],
                     dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, integ_ms, touch_ms, mc_ms, k1_ms, seq_ms, ghost_ms, gather_ms = t.tolist()
    visits_blocks, updates, loaded, nblocks, V, T = [int(x) for x in w.tolist()[:6]]
    integ_mean = w.tolist()[6] / world
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if not replicas and (hv.shape[0], ht.shape[0]) != (V, T):
        raise SystemExit(f"end-to-end mesh {hv.shape[0]}/{ht.shape[0]} differs from the device-resident run {V}/{T}")
    ms_per_step = total_ms / args.steps
    frames_job = n * (world if replicas else 1)
    frames_per_s = frames_job / (ms_per_step * 1e-3)
    bytes_per_visit = 40 if color else 16
    visits = visits_blocks * 4096
    from mq3d_b200 import synth
    depth_bytes = n * synth.DEPTH_W * synth.DEPTH_H * 4
    rgbx_bytes = n * synth.DEPTH_W * synth.DEPTH_H * 4 if color else 0     # colour resampled onto the depth grid
    batched_bytes = bytes_per_visit * loaded * 4096 + (depth_bytes + rgbx_bytes) * (world if replicas else 1) + 12 * loaded
    visit_bytes = bytes_per_visit * visits + (depth_bytes + rgbx_bytes) * (world if replicas else 1) + 12 * visits_blocks
    peak, peak_src = peaks()
    # per GPU: the kernel time is the slowest rank's, the bytes are the whole job's
    achieved = batched_bytes / world / (integ_ms * 1e-3) / 1e9
    launches = max(int(st.batches), 1)
    mc_bytes = 8 * nblocks * 4096 * (17 / 16) ** 3 + 24 * V + 12 * T + (12 * V if color else 0)
    mc_achieved = mc_bytes / world / (mc_ms * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": frames_per_s, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak" if replicas else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(args, cfg, world, ghost_mode_used[0]),
        "gvoxel_updates_per_s": visits / (ms_per_step * 1e-3) / 1e9,
        "gvoxel_visits_per_s_integrate_kernel": visits / world / (integ_ms * 1e-3) / 1e9,
        "updated_voxel_fraction": updates / max(visits, 1),
        "mc_ms": mc_ms, "mesh": {"vertices": V, "triangles": T}, "active_blocks": nblocks,
        "kernel_ms": {"k_depth_prepare": k1_ms, "sequence_call": seq_ms, "k_integrate": integ_ms, "k_touch": touch_ms,
                      "mc_count+fill": mc_ms},
        "multi_gpu": {"ghost_shell_ms": ghost_ms, "gather_mesh_ms": gather_ms, "k_integrate_ms_max": integ_ms,
                      "k_integrate_ms_mean": integ_mean, "parity": parity} if sharded else None,
        "roofline": {"bound": "hbm", "kernel": "k_integrate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src,
                     "traffic": profile_figure("integrate_traffic", args.workload),
                     "limiter": "FP32 pipe and instruction issue (not HBM): a block stays resident in registers for a "
                                "whole batch of frames, so DRAM moves one load + store per block residency; the "
                                "packed-FP32 body keeps the FMA pipe busy ~70 % of the cycles at ~76 % issue-slot use",
                     "issue_slot_frac": profile_figure("integrate_issue", args.workload),
                     "fma_pipe_frac": profile_figure("integrate_fma", args.workload),
                     "algorithmic_bytes": "batched: bytes_per_voxel x 4096 x block residencies + frame images read "
                                          "once + 12 B keys per residency",
                     "bytes_per_voxel": bytes_per_visit, "block_residencies": loaded, "block_visits": visits_blocks,
                     "launches_per_step": launches, "algorithmic_bytes_per_launch": batched_bytes / world / launches,
                     "throughput_equiv_gbs": visit_bytes / world / (integ_ms * 1e-3) / 1e9,
                     "throughput_equiv_note": "Open3D's per-visit accounting (SURVEY 8d: every voxel visit moves its "
                                              "attributes once); NOT DRAM traffic of this kernel"},
        "roofline_mc": {"bound": "hbm", "kernel": "k_mc_classify + k_scan_counts + k_mc_emit" + (" + k_mc_colors" if color else ""),
                        "achieved": mc_achieved, "peak": peak, "unit": "GB/s", "frac": mc_achieved / peak,
                        "algorithmic_bytes": mc_bytes / world,
                        "bytes": "8 B x active voxels x (17/16)^3 + 24 B/vertex + 12 B/triangle (SURVEY 8d)",
                        "traffic": profile_figure("mc_traffic", args.workload), "ms": mc_ms},
        "e2e": {"value": frames_job / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "ms_each_step_rank0": e2e_each,
                "h2d_note": h2d_note},
        # per step on each rank: reset fill (1) + K1 prepare / finalize (2) + per batch [colour resample, one launch per
        # 64 frames] + touch + list + sort + integrate launches (unguarded / guarded division x, depth-only on a
        # partitioned grid, the two shapes chosen on the device by batch size: all but one return at once) + bitmap
        # clear; MC neighbours / rows /
        # classify / scan x 2 / emit [+ colours]
        "gpu_launches": int(args.steps * (3 + st.batches * ((4 + 2 + -(-min(args.batch, n) // 64)) if color else
                                                            (4 + (4 if sharded else 2))) + (7 if color else 6))),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_measure(cfg, args.workload, args.cpu_frames or cfg["cpu_frames"], steps=1, warmup=0)[0]
    if world == 1 and not args.no_aux and args.workload == DEFAULT_WORKLOAD:
        del wl, raw_host, host_mesh, mesh
        vbg.close()
        torch.cuda.empty_cache()
        out["aux"] = aux_block(device, args)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_config(args, cfg, world, ghost_mode):
    """`config` of the JSON line: identical keys and values in both arms."""
    n = cfg["frames"] * cfg["sides"]
    replicas = bool(cfg.get("replicas"))
    return {"workload": args.workload, "frames": n, "sides": cfg["sides"], "depth": "320x320 f32 raw NDC",
            "color": "1280x960 u8 RGB" if cfg["color"] else None, "voxel_size": cfg["voxel"], "block": "16^3",
            "trunc_voxel_multiplier": cfg["trunc"], "depth_max": cfg["depth_max"],
            "weight_threshold": cfg["weight_thr"], "batch_frames": args.batch,
            "partition": ("independent capture per GPU" if replicas else
                          f"block hash, tile {args.tile}, ghost shell by {ghost_mode}") if world > 1 else "single GPU",
            "l2": "inputs larger than L2 (the depth sequence is far above 126 MB); no explicit flush"}


# ---------------------------------------------------------------------------------------------------
# aux: the satellite kernels and the other BASELINE configs, N = 1
# ---------------------------------------------------------------------------------------------------
def aux_block(device, args):
    import torch
    from mq3d_b200.ops import filter_mesh_components
    from mq3d_b200.geometry import TriangleMesh
    from mq3d_b200.raycast import RaycastingScene
    from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare, estimate_confidence
    from mq3d_b200 import synth
    peak, _ = peaks()
    aux = {}

    def timed(fn, steps=3, warmup=1):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps, r

    def device_step(cfg, wl, vbg):
        vbg.reset()
        lin, valid = depth_prepare(wl["raw"], wl["nears"], wl["fars"])
        st = vbg.integrate_sequence(lin, wl["K"], wl["Ewc"], cfg["depth_max"], cfg["trunc"], 1.0, frame_valid=valid,
                                    colors=wl["colors"], color_intrinsics=wl["Kc"], batch_frames=args.batch)
        return st, vbg.extract_triangle_mesh_arrays(cfg["weight_thr"], with_colors=cfg["color"])

    # configs[0] and configs[1]: device-resident frames/s; configs[1] adds the colour-aligned depth raycast (K6)
    for name, key in (("quest300_v20mm", "config0_quest300_v20mm"), ("quest300_rgb_v10mm", "config1_quest300_rgb_v10mm")):
        cfg = dict(WORKLOADS[name])
        wl = build_workload(cfg, device, 0, 1)
        attrs = ("tsdf", "weight", "color") if cfg["color"] else ("tsdf", "weight")
        vbg = VoxelBlockGrid(attr_names=attrs, voxel_size=cfg["voxel"], block_count=cfg["block_count"], device=device)
        ms, (st, mesh) = timed(lambda: device_step(cfg, wl, vbg), steps=5, warmup=2)
        row = {"frames": wl["n"], "ms_per_step": ms, "frames_per_s": wl["n"] / (ms * 1e-3),
               "k_integrate_ms": st.integrate_ms, "active_blocks": st.num_blocks,
               "mesh": {"vertices": int(mesh[0].shape[0]), "triangles": int(mesh[2].shape[0])}}
        if not cfg["color"]:
            # the per-frame call shape of the level-1 drop-in (o3d_utils.py:212-229: compute_unique_block_coordinates ->
            # integrate, one frame at a time, synchronous like Open3D's) on the same frames
            lin, valid = depth_prepare(wl["raw"], wl["nears"], wl["fars"])

            def per_frame():
                vbg.reset()
                for i in range(wl["n"]):
                    keys = vbg.compute_unique_block_coordinates(lin[i], wl["K"][i], wl["Ewc"][i], 1.0, cfg["depth_max"],
                                                                cfg["trunc"])
                    vbg.integrate(keys, lin[i], wl["K"][i], wl["Ewc"][i], 1.0, cfg["depth_max"], cfg["trunc"])
                return vbg.num_blocks()
            pf_ms, nb_pf = timed(per_frame, steps=2, warmup=1)
            row["per_frame_api"] = {"ms_per_step": pf_ms, "frames_per_s": wl["n"] / (pf_ms * 1e-3), "active_blocks": int(nb_pf),
                                    "what": "VoxelBlockGrid.compute_unique_block_coordinates + .integrate per frame "
                                            "(two synchronous C-ABI calls per frame; HBM / launch bound)"}
            del lin
        if cfg["color"]:
            # K6 (o3d_utils.py:324-342): filtered mesh -> LBVH -> 1280x960 pinhole rays per colour frame
            tm = filter_mesh_components(TriangleMesh(mesh[0], mesh[2], mesh[1]), min_triangle_count=5000)
            scene = RaycastingScene(device=device)
            t0 = time.perf_counter()
            scene.add_triangles(tm)
            torch.cuda.synchronize()
            build_ms = (time.perf_counter() - t0) * 1e3
            Kc = np.asarray(wl["Kc"], np.float32)
            views = list(range(0, wl["n"], 10))

            def cast_all():
                acc = None
                for i in views:
                    rays = scene.create_rays_pinhole(Kc[i], wl["Ewc"][i], width_px=synth.COLOR_W, height_px=synth.COLOR_H)
                    acc = scene.cast_rays(rays)["t_hit"]
                return acc
            rc_ms, th = timed(cast_all, steps=3, warmup=1)
            n_rays = len(views) * synth.COLOR_W * synth.COLOR_H
            row["k6_raycast"] = {"views": len(views), "rays": n_rays, "ms": rc_ms, "mrays_per_s": n_rays / (rc_ms * 1e-3) / 1e6,
                                 "bvh_build_ms": build_ms, "triangles": int(tm.triangle.indices.shape[0]),
                                 "hit_fraction": float(torch.isfinite(th).float().mean()),
                                 "t_hit_write_gbs": 4 * n_rays / (rc_ms * 1e-3) / 1e9,
                                 "bound": "traversal latency (no roofline claim, SURVEY 8d)"}
        aux[key] = row
        vbg.close()
        del wl, vbg, mesh
        torch.cuda.empty_cache()

    # configs[2]: K4 multi-view confidence, +-5 neighbours, both eyes, 1000 frames each
    cfg = dict(frames=1000, sides=2, color=False)
    wl = build_workload(cfg, device, 0, 1)
    lin, valid = depth_prepare(wl["raw"], wl["nears"], wl["fars"])
    R = 5
    pairs = 0
    for s in range(2):
        for i in range(1000):
            pairs += min(999, i + R) - max(0, i - R)
    Einv = np.linalg.inv(wl["Ecw"])

    def k4():
        out = None
        for s in range(2):
            sl = slice(1000 * s, 1000 * (s + 1))
            out = estimate_confidence(lin[sl], wl["K"][sl], wl["Ecw"][sl], Einv[sl], R, 4.0, 0.08, frame_valid=valid[sl])
        return out
    ms, (conf, cnt) = timed(k4, steps=3, warmup=1)
    px = synth.DEPTH_W * synth.DEPTH_H
    aux["config2_confidence_1000x2_r5"] = {
        "pairs": pairs, "ms": ms, "pairs_per_s": pairs / (ms * 1e-3),
        "min_bytes_gbs": 16 * px * 2000 / (ms * 1e-3) / 1e9, "min_bytes_frac_hbm": 16 * px * 2000 / (ms * 1e-3) / 1e9 / peak,
        "pair_bytes_gbs": 8 * px * pairs / (ms * 1e-3) / 1e9,
        "mean_confidence": float(conf.mean()), "bound": "fp64 issue (SURVEY 8d: report vs HBM and fp64)"}
    del wl, lin, conf, cnt
    torch.cuda.empty_cache()

    # configs[4]: one replica of the batch workload (2000 frames, 1 cm), end to end from pinned host memory
    from mq3d_b200.pipeline import IntegrationParams, extract_mesh_to_host, integrate_frames
    cfg = dict(WORKLOADS["batch_v10mm"])
    wl = build_workload(cfg, device, 0, 1, shared=False)
    vbg = VoxelBlockGrid(voxel_size=cfg["voxel"], block_count=cfg["block_count"], device=device)
    raw_host = wl["raw"].cpu().pin_memory()
    params = IntegrationParams(voxel_size=cfg["voxel"], block_count=cfg["block_count"], depth_max=cfg["depth_max"],
                               trunc_voxel_multiplier=cfg["trunc"], use_confidence_filtered_depth=False,
                               batch_frames=args.batch)

    def capture():
        vbg.reset()
        integrate_frames(vbg, raw_host, wl["nears"], wl["fars"], wl["K"], wl["Ewc"], params)
        return extract_mesh_to_host(vbg, cfg["weight_thr"])
    ms, m = timed(capture, steps=3, warmup=1)
    aux["config4_batch_capture_2000_v10mm"] = {
        "frames": wl["n"], "e2e_ms_per_capture": ms, "captures_per_hour_per_gpu": 3600e3 / ms,
        "frames_per_s": wl["n"] / (ms * 1e-3), "mesh": {"vertices": int(m[0].shape[0]), "triangles": int(m[2].shape[0])},
        "note": "replicas only: N GPUs run N such captures concurrently (bench.py --workload batch_v10mm --gpus N)"}
    vbg.close()
    return aux


# ---------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------
def cpu_sample(cfg, sample_frames):
    """`sample_frames` frames spread evenly over the whole sequence (LEFT then RIGHT), as host arrays.  A strided
    sample sees the whole room, so the grid marching cubes runs on is close to the full job's."""
    import torch
    from mq3d_b200 import synth, synth_gpu
    K, Ewc, Ecw = cameras(cfg)
    n = len(K)
    idx = np.unique(np.linspace(0, n - 1, min(sample_frames, n)).round().astype(np.int64))
    raw = np.stack([synth_gpu.render_depth(Ecw[i:i + 1], "cpu", first_frame=int(i))[0].numpy() for i in idx])
    colors = Kc = None
    if cfg["color"]:
        colors = np.stack([synth.make_color_frame(Ecw[i]) for i in idx])
        Kc = color_intrinsics(len(idx))
    return dict(raw=raw, K=K[idx], Ewc=Ewc[idx], colors=colors, Kc=Kc, n_total=n, n=len(idx),
                nears=np.full(len(idx), synth.NEAR), fars=np.full(len(idx), synth.FAR))


def cpu_step(orc, cfg, s):
    """(seconds integrating the sample, seconds in marching cubes, blocks)"""
    t0 = time.perf_counter()
    g = orc.Grid(cfg["voxel"], with_color=cfg["color"])
    for i in range(s["n"]):
        raw = s["raw"][i]
        if not orc.depth_valid(raw):
            continue
        d = orc.depth_to_linear(raw, s["nears"][i], s["fars"][i])
        keys = g.touch(d, s["K"][i], s["Ewc"][i], cfg["depth_max"], cfg["trunc"])
        g.integrate(keys, d, s["K"][i], s["Ewc"][i], cfg["depth_max"], cfg["trunc"],
                    color=None if s["colors"] is None else s["colors"][i], Kc=None if s["Kc"] is None else s["Kc"][i])
    t1 = time.perf_counter()
    mesh = g.extract_mesh(cfg["weight_thr"])
    if cfg["color"]:
        g.vertex_colors(mesh[3])          # the GPU step extracts vertex colours too
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, g.num_blocks


def cpu_measure(cfg, workload, sample_frames, steps, warmup):
    """Times the oracle on the sample.  Whole-job frames/s = n_total / (n_total / n_sample x integrate time of
    the sample + marching-cubes time): integration scales with the frame count, extraction is charged once per
    job, as in the GPU arm.  Returns (cpu_baseline dict, seconds per step)."""
    from oracle import oracle as orc
    orc.build()
    threads = orc.set_num_threads(len(os.sched_getaffinity(0)))   # launchers export OMP_NUM_THREADS=1
    s = cpu_sample(cfg, sample_frames)
    for _ in range(warmup):
        cpu_step(orc, cfg, s)
    ti, tm, dt = [], [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        a, b, blocks = cpu_step(orc, cfg, s)
        dt.append(time.perf_counter() - t0)
        ti.append(a)
        tm.append(b)
    t_int, t_mc = float(np.mean(ti)), float(np.mean(tm))
    job_s = s["n_total"] / s["n"] * t_int + t_mc
    return {"value": s["n_total"] / job_s, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{s['n']} of the {s['n_total']} frames of {workload}, evenly strided over the sequence: "
                      f"integration {t_int:.2f} s, marching cubes on the resulting {blocks}-block grid {t_mc:.2f} s; "
                      f"whole-job rate = frames / (frames / {s['n']} x integration + marching cubes); "
                      f"OpenMP team of {threads} threads (omp_get_max_threads)",
            "integrate_frames_per_s": s["n"] / t_int, "mc_s": t_mc, "label": CPU_LABEL}, float(np.mean(dt))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = dict(WORKLOADS[args.workload])
    if args.frames:
        cfg["frames"] = args.frames
    world = int(os.environ.get("WORLD_SIZE", 1))
    base, dt = cpu_measure(cfg, args.workload, args.cpu_frames or cfg["cpu_frames"], steps=args.steps, warmup=args.warmup)
    v = base["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak" if cfg.get("replicas") else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, cfg, world, args.ghosts),
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override the number of frames per side")
    ap.add_argument("--batch", type=int, default=256, help="frames per block residency (<= 256)")
    ap.add_argument("--tile", type=int, default=1, help="partition super-tile edge in blocks (N > 1)")
    ap.add_argument("--ghosts", default="pull", choices=["pull", "exchange", "integrate"],
                    help="N > 1: 'integrate' = every rank also integrates its ghost shell (no exchange, the "
                         "north-star scheme); 'exchange' = owned blocks only + one packed NCCL ghost-block "
                         "exchange before MC; 'pull' = owned blocks only + ghost blocks copied straight from the "
                         "owners' pools over NVLink peer memory (falls back to 'exchange' without CUDA IPC)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the bounded CPU sample (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the aux block (K4, K6, configs #0/#1/#4)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
