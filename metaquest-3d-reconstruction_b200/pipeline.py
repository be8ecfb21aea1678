"""Array-level public API of the hot path: host buffers in, grid / mesh out.

This is what ``integrate()`` / ``reconstruct_scene()`` (o3d_utils.py:153-238, reconstruct_scene.py:60-108)
boil down to once the frames are in memory: H2D copy of the raw Quest depth (and colour), K1 linearise
+ mask, fused K2/K3 integration, K5 extraction.  bench.py's end-to-end number times exactly this.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .vbg import SequenceStats, VoxelBlockGrid, color_resample, depth_prepare


@dataclass
class IntegrationParams:
    """Mirror of the reference's IntegrationConfig (config/reconstruction_config.py:99-109)."""
    voxel_size: float = 0.01
    block_resolution: int = 16
    block_count: int = 50_000
    depth_max: float = 1.5
    trunc_voxel_multiplier: float = 8.0
    use_confidence_filtered_depth: bool = True
    confidence_threshold: float = 0.05
    valid_count_threshold: int = 4
    batch_frames: int = 64


def pin(a: np.ndarray) -> torch.Tensor:
    """Pinned host tensor sharing the layout of `a` (copy)."""
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.pin_memory() if torch.cuda.is_available() else t


def integrate_frames(vbg: VoxelBlockGrid, raw_host: torch.Tensor, nears, fars, K, E_wc, params: IntegrationParams,
                     conf: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None,
                     has_conf: Optional[torch.Tensor] = None, colors_host: Optional[torch.Tensor] = None,
                     Kc=None, shard: Optional[tuple] = None, gated: Optional[bool] = None) -> SequenceStats:
    """Integrate [F,H,W] raw NDC depth frames (host, ideally pinned) into `vbg`.

    conf/count (float64 / int32 [F,H,W], host or device) enable the reference's confidence mask
    (o3d_utils.py:131-142).  colors_host: uint8 [F,CH,CW,3] enables Open3D's colour overload.  Colour never
    travels as whole frames when it is pinned: the copy stream resamples each chunk onto the depth pixel grid
    (mq3d_color_resample) reading the pinned frames in place, so only the W x H sampled pixels per frame
    cross PCIe; unpinned colour is copied to the device first.

    shard=(rank, world): multi-GPU upload.  Every rank integrates every frame into its own blocks, but each
    rank moves only 1/world of each chunk over its PCIe link and the chunk is completed by an NCCL
    all-gather over NVLink (the frame broadcast of SURVEY 8e, sharded), still on the copy stream.

    gated (default: single-GPU uploads whose whole-sequence linear depth fits a quarter of the free device memory): the copy
    stream uploads every chunk AND runs K1 on it (into one whole-sequence buffer), recording one event per chunk;
    the fused K2/K3 path is then ONE call whose batches wait for those events on the device
    (mq3d_grid_set_batch_gates) -- no host synchronisation per chunk, so the device never idles while Python
    enqueues the next chunk.  gated=False keeps one call per chunk (memory bounded by two chunks)."""
    dev = vbg.device
    F, H, W = (int(x) for x in raw_host.shape)
    if shard is not None and int(shard[1]) > 1:
        import torch.distributed as dist
        s_rank, s_world = int(shard[0]), int(shard[1])

        def upload(host, f0, f1):
            per = -(-(f1 - f0) // s_world)
            full = torch.empty((per * s_world,) + tuple(host.shape[1:]), dtype=host.dtype, device=dev)
            a = min(f1, f0 + s_rank * per)
            b = min(f1, a + per)
            mine = full[s_rank * per:(s_rank + 1) * per]
            if b > a:
                mine[: b - a].copy_(host[a:b], non_blocking=True)
            dist.all_gather_into_tensor(full, mine)
            return full[: f1 - f0]
    else:
        def upload(host, f0, f1):
            return host[f0:f1].to(dev, non_blocking=True)
    mask = params.use_confidence_filtered_depth and conf is not None
    use_color = colors_host is not None and vbg.has_color
    chunk = max(1, int(params.batch_frames))
    K = np.asarray(K)
    E_wc = np.asarray(E_wc)
    nears = np.asarray(nears, dtype=np.float64)
    fars = np.asarray(fars, dtype=np.float64)
    # Chunked pipeline: a copy stream uploads chunk k+1 (pinned host -> HBM) while the compute stream
    # runs K1 + fused K2/K3 on chunk k; one event per chunk orders the two streams.
    main = torch.cuda.current_stream(dev)
    copy = _copy_stream(dev)
    if gated is None and os.environ.get("MQ3D_E2E_GATED") in ("0", "1"):      # A/B switch for measurements
        gated = os.environ["MQ3D_E2E_GATED"] == "1"
    if gated is None:
        # (sharded uploads keep one call per chunk: with the NCCL all-gather AND K1 of every chunk queued behind the
        # persistent integrate kernels of all ranks the gated form was no faster at N = 2 and noisier at N = 8)
        with torch.cuda.device(dev):
            free_b, _ = torch.cuda.mem_get_info()
        sharded_upload = shard is not None and int(shard[1]) > 1
        gated = (not sharded_upload) and F * H * W * 4 * (2 if use_color else 1) <= free_b // 4
    if gated:
        return _integrate_frames_gated(vbg, raw_host, nears, fars, K, E_wc, params, conf, count, has_conf, colors_host,
                                       Kc, shard, upload, mask, use_color, chunk, main, copy)
    copy.wait_stream(main)
    staged = []
    for f0 in range(0, F, chunk):
        f1 = min(F, f0 + chunk)
        with torch.cuda.stream(copy):
            part = {"raw": upload(raw_host, f0, f1)}
            if use_color:
                pinned = colors_host.device.type == "cpu" and colors_host.is_pinned()
                Kc_np = np.asarray(Kc)
                if pinned and shard is not None and int(shard[1]) > 1:
                    # each rank resamples its 1/world of the chunk straight from its pinned frames; the
                    # all-gather then moves W x H x 4 B per frame over NVLink instead of whole colour frames
                    per = -(-(f1 - f0) // s_world)
                    full = torch.empty((per * s_world, H, W), dtype=torch.int32, device=dev)
                    a = min(f1, f0 + s_rank * per)
                    b = min(f1, a + per)
                    mine = full[s_rank * per:(s_rank + 1) * per]
                    if b > a:
                        color_resample(colors_host[a:b], K[a:b], Kc_np[a:b], W, H, device=dev, out=mine[: b - a])
                    dist.all_gather_into_tensor(full, mine)
                    part["rgbx"] = full[: f1 - f0]
                else:
                    src = colors_host[f0:f1] if pinned else upload(colors_host, f0, f1)
                    part["rgbx"] = color_resample(src.contiguous(), K[f0:f1], Kc_np[f0:f1], W, H, device=dev)
            if mask:
                for name, src in (("conf", conf), ("count", count)):
                    part[name] = upload(src, f0, f1) if src.device.type == "cpu" else src[f0:f1].to(dev)
            ev = torch.cuda.Event()
            ev.record(copy)
        staged.append((f0, f1, part, ev))
    total = None
    for f0, f1, part, ev in staged:
        main.wait_event(ev)
        for t in part.values():
            t.record_stream(main)
        lin, valid = depth_prepare(part["raw"], nears[f0:f1], fars[f0:f1], part.get("conf"), part.get("count"),
                                   None if (has_conf is None or not mask) else has_conf[f0:f1],
                                   params.confidence_threshold, params.valid_count_threshold)
        st = vbg.integrate_sequence(lin, K[f0:f1], E_wc[f0:f1], params.depth_max, params.trunc_voxel_multiplier, 1.0,
                                    frame_valid=valid, colors_rgbx=part.get("rgbx"), batch_frames=chunk)
        if total is None:
            total = st
        else:
            total = SequenceStats(total.frames_integrated + st.frames_integrated, total.block_visits + st.block_visits,
                                  total.blocks_loaded + st.blocks_loaded, st.num_blocks, total.batches + st.batches,
                                  total.voxel_updates + st.voxel_updates, total.touch_ms + st.touch_ms,
                                  total.integrate_ms + st.integrate_ms,
                                  total.slow_div_batches + st.slow_div_batches)
    return total


def _integrate_frames_gated(vbg, raw_host, nears, fars, K, E_wc, params, conf, count, has_conf, colors_host, Kc, shard,
                            upload, mask, use_color, chunk, main, copy) -> SequenceStats:
    """Copy stream: upload + K1 (+ colour resampling) per chunk into whole-sequence buffers, one event per chunk;
    compute stream: one fused K2/K3 call gated by those events."""
    dev = vbg.device
    F, H, W = (int(x) for x in raw_host.shape)
    lin = torch.empty((F, H, W), dtype=torch.float32, device=dev)
    valid = torch.empty((F,), dtype=torch.int32, device=dev)
    rgbx = torch.empty((F, H, W), dtype=torch.int32, device=dev) if use_color else None
    for t in (lin, valid, rgbx):
        if t is not None:
            t.record_stream(copy)
    copy.wait_stream(main)          # the buffers above are allocated on the compute stream's pool
    Kc_np = np.asarray(Kc) if use_color else None
    sharded = shard is not None and int(shard[1]) > 1
    events = []
    for f0 in range(0, F, chunk):
        f1 = min(F, f0 + chunk)
        with torch.cuda.stream(copy):
            raw = upload(raw_host, f0, f1)
            cpart = upload(conf, f0, f1) if mask and conf.device.type == "cpu" else (conf[f0:f1].to(dev) if mask else None)
            npart = upload(count, f0, f1) if mask and count.device.type == "cpu" else (count[f0:f1].to(dev) if mask else None)
            depth_prepare(raw, nears[f0:f1], fars[f0:f1], cpart, npart,
                          None if (has_conf is None or not mask) else has_conf[f0:f1],
                          params.confidence_threshold, params.valid_count_threshold, out=lin[f0:f1], valid_out=valid[f0:f1])
            if use_color:
                pinned = colors_host.device.type == "cpu" and colors_host.is_pinned()
                if pinned and sharded:
                    import torch.distributed as dist
                    s_rank, s_world = int(shard[0]), int(shard[1])
                    per = -(-(f1 - f0) // s_world)
                    full = torch.empty((per * s_world, H, W), dtype=torch.int32, device=dev)
                    a = min(f1, f0 + s_rank * per)
                    b = min(f1, a + per)
                    mine = full[s_rank * per:(s_rank + 1) * per]
                    if b > a:
                        color_resample(colors_host[a:b], K[a:b], Kc_np[a:b], W, H, device=dev, out=mine[: b - a])
                    dist.all_gather_into_tensor(full, mine)
                    rgbx[f0:f1].copy_(full[: f1 - f0])
                else:
                    src = colors_host[f0:f1] if pinned else upload(colors_host, f0, f1)
                    color_resample(src.contiguous(), K[f0:f1], Kc_np[f0:f1], W, H, device=dev, out=rgbx[f0:f1])
            ev = torch.cuda.Event()
            ev.record(copy)
        events.append(ev)
    return vbg.integrate_sequence(lin, K, E_wc, params.depth_max, params.trunc_voxel_multiplier, 1.0, frame_valid=valid,
                                  colors_rgbx=rgbx, batch_frames=chunk, batch_events=events)


_COPY_STREAMS: dict = {}


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.type, dev.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev, priority=-1)    # uploads / resampling get free SM slots first
    return _COPY_STREAMS[key]


def to_host(tensors):
    """Device tensors -> numpy arrays through pinned staging (torch's caching host allocator reuses the blocks),
    all copies in flight together, one synchronisation."""
    outs = []
    for t in tensors:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        outs.append(h)
    torch.cuda.current_stream().synchronize()
    return tuple(h.numpy() for h in outs)


def extract_mesh_to_host(vbg: VoxelBlockGrid, weight_threshold: float, with_colors: bool = False):
    """K5 + D2H: (vertices f32 [V,3], normals f32 [V,3], triangles i32 [T,3][, colors f32 [V,3]]) as numpy
    arrays; with_colors needs a grid with the colour attribute."""
    with torch.cuda.device(vbg.device):
        return to_host(vbg.extract_triangle_mesh_arrays(weight_threshold, with_colors=with_colors))
