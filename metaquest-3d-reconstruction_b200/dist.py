"""Multi-GPU plumbing: voxel blocks are partitioned by a spatial hash of their super-tile,
every rank sees every frame and integrates the blocks it owns; the one-block ghost shell marching cubes
needs is either integrated redundantly (no communication), exchanged as packed blocks over NCCL
(exchange_ghosts) or -- the default -- read straight out of the owners' pools through CUDA-IPC peer
memory over NVLink (pull_ghosts / fill_ghost_shell); all three give bit-identical grids.  Meshes are
extracted per rank for owned cubes only and gathered on rank 0 (gather_mesh).  torch.distributed (NCCL
on the box, gloo in the CPU tests) carries the frame broadcast / sharded upload all-gather, the
512-byte pool descriptors and the final gather.

The Python functions below restate the device-side ownership rule (csrc/mq3d_common.cuh:
mq3d_tile_owner / mq3d_block_needed) so that tests can check the partition without a GPU.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

_BIAS = 1 << 20
_M64 = (1 << 64) - 1


def _hash64(k: int) -> int:
    k &= _M64
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & _M64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & _M64
    k ^= k >> 33
    return k & 0xFFFFFFFF


def _pack(x: int, y: int, z: int) -> int:
    return (((x + _BIAS) & 0xFFFFFFFF) << 42) | (((y + _BIAS) & 0xFFFFFFFF) << 21) | ((z + _BIAS) & 0xFFFFFFFF)


def tile_owner(bx: int, by: int, bz: int, world: int, tile_blocks: int = 8) -> int:
    """Owner rank of the block (bx,by,bz): hash of its super-tile modulo the world size."""
    s = tile_blocks.bit_length() - 1
    return _hash64(_pack(bx >> s, by >> s, bz >> s)) % world


def tile_owner_np(keys: np.ndarray, world: int, tile_blocks: int = 8) -> np.ndarray:
    """Vectorised tile_owner for int block keys [n,3] -> int64 [n]."""
    s = np.uint64(tile_blocks.bit_length() - 1)
    k = (np.asarray(keys, dtype=np.int64) >> np.int64(s)) + _BIAS
    k = k.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    h = (k[:, 0] << np.uint64(42)) | (k[:, 1] << np.uint64(21)) | k[:, 2]
    with np.errstate(over="ignore"):
        h ^= h >> np.uint64(33)
        h *= np.uint64(0xFF51AFD7ED558CCD)
        h ^= h >> np.uint64(33)
        h *= np.uint64(0xC4CEB9FE1A85EC53)
        h ^= h >> np.uint64(33)
    return ((h & np.uint64(0xFFFFFFFF)) % np.uint64(world)).astype(np.int64)


def block_needed(bx: int, by: int, bz: int, rank: int, world: int, tile_blocks: int = 8) -> bool:
    """True when the block is owned by `rank` or lies in the one-block shell around an owned tile."""
    if world <= 1:
        return True
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if tile_owner(bx + dx, by + dy, bz + dz, world, tile_blocks) == rank:
                    return True
    return False


def broadcast_frames(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Frame broadcast (raw depth / colour batches) from `src` to every rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src)
    return t


def _pack_ghosts(vbg, dest: int, count: Optional[int] = None):
    """(flat float32 payload, block count, has_color) for `dest`; layout = vbg.ghost_packed_len/ghost_views."""
    from .vbg import ghost_packed_len, ghost_views
    if hasattr(vbg, "ghost_select_packed"):
        return vbg.ghost_select_packed(dest) if count is None else vbg.ghost_select_packed(dest, count)
    k, t, w, c = vbg.ghost_select(dest)
    m, has_color = int(k.shape[0]), c is not None
    buf = torch.zeros(ghost_packed_len(m, has_color), dtype=torch.float32, device=k.device)
    if m:
        for dst, src in zip(ghost_views(buf, m, has_color), (k, t, w, c)):
            if dst is not None:
                dst.copy_(src.reshape(dst.shape))
    return buf, m, has_color


def exchange_ghosts(vbg, rank: Optional[int] = None, world: Optional[int] = None,
                    timings: Optional[dict] = None) -> int:
    """Owned-only integration mode: fetch the ghost shell from the owners.  One all-gather of the
    [world x world] block-count matrix, then one packed send/recv per rank pair (keys | tsdf | weight |
    colour in a single buffer), then one import.  `vbg` needs ghost_select(dest) (or ghost_select_packed)
    and import_blocks(keys, tsdf, weight, color).  Returns the number of ghost blocks received.  The payload
    is the owners' block values, so afterwards the local grid is bit-identical to one that integrated its
    ghosts redundantly."""
    import torch.distributed as dist
    from .vbg import ghost_packed_len, ghost_views
    if rank is None:
        rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return 0
    import time as _time
    t_last = [_time.perf_counter()]

    def lap(name):          # optional phase timing (synchronises; diagnostics only)
        if timings is not None:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            now = _time.perf_counter()
            timings.setdefault(name, []).append((now - t_last[0]) * 1e3)
            t_last[0] = now

    if hasattr(vbg, "ghost_counts"):
        # one counting kernel + one sync for all destinations; the count matrix travels while the
        # payloads are being packed (asynchronous selects, no further host syncs)
        counts = vbg.ghost_counts()
        lap("counts")
        dev = torch.device(vbg.device)
        send_n = torch.tensor(counts, dtype=torch.int64).to(dev, non_blocking=True)
        matrix = torch.empty((world, world), dtype=torch.int64, device=dev)
        pending = dist.all_gather_into_tensor(matrix, send_n, async_op=True)
        outgoing = {d: _pack_ghosts(vbg, d, counts[d]) for d in range(world) if d != rank}
        has_color = next(iter(outgoing.values()))[2]
        lap("pack")
        pending.wait()
        recv_n = matrix[:, rank].cpu().tolist()           # matrix[src][dst] = blocks src sends to dst
    else:
        outgoing = {d: _pack_ghosts(vbg, d) for d in range(world) if d != rank}
        first = next(iter(outgoing.values()))
        dev, has_color = first[0].device, first[2]
        send_n = torch.zeros(world, dtype=torch.int64, device=dev)
        for d, (_, m, _) in outgoing.items():
            send_n[d] = m
        matrix = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(matrix, send_n)
        recv_n = [int(matrix[p][rank]) for p in range(world)]
    total = sum(recv_n[p] for p in range(world) if p != rank)
    lap("count_matrix")
    incoming, ops = {}, []
    for p in range(world):
        if p == rank:
            continue
        buf, m, _ = outgoing[p]
        if m:
            ops.append(dist.P2POp(dist.isend, buf, p))
        if recv_n[p]:
            incoming[p] = torch.empty(ghost_packed_len(recv_n[p], has_color), dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, incoming[p], p))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    lap("p2p")
    if not incoming:
        return 0
    parts = [ghost_views(buf, recv_n[p], has_color) for p, buf in incoming.items()]
    if len(parts) == 1:
        k, t, w, c = parts[0]
    else:
        k, t, w = (torch.cat([x[i] for x in parts]) for i in range(3))
        c = torch.cat([x[3] for x in parts]) if has_color else None
    vbg.import_blocks(k, t, w, c)
    lap("import")
    return total


def pull_ghosts(vbg, rank: Optional[int] = None, world: Optional[int] = None, fence: bool = True,
                count: bool = True) -> int:
    """Owned-only integration mode, peer-memory variant of exchange_ghosts: all-gather the 512-byte pool
    descriptors (CUDA IPC handles + block counts; the collective is stream-ordered after integration, so it is also
    the "everyone has finished integrating" barrier), then each rank copies the blocks of its ghost shell straight
    out of the owners' pools over NVLink (vbg.ghost_pull: scan, activation and copy kernels enqueued back to back,
    no staging, no send/recv, no import).  A one-word all-reduce afterwards keeps any rank's later work (grid reset,
    next integration) from overtaking a peer that is still reading its pool.

    Host synchronisation: one, to read the gathered descriptors.  The all-reduce also carries a status word, read on
    the host only while the set of exported pools is new to this process (first pull, or some pool was re-allocated
    -- every rank sees the same table, so all ranks agree on that): mapping a peer's pool is the step that can fail,
    and then every rank raises PeerPullError together (the grids are still valid for exchange_ghosts).  A rank that
    cannot export its own pool publishes an empty descriptor, which all ranks see at once.  count=False skips the
    read-back of the number of blocks fetched (returns -1).  fence=False never reads the status."""
    import torch.distributed as dist
    if rank is None:
        rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return 0
    dev = torch.device(vbg.device)
    err, n = None, -1
    import os as _os
    import time as _time
    trace = _os.environ.get("MQ3D_TRACE") is not None
    t_start = _time.perf_counter()
    if trace:
        torch.cuda.synchronize()
        t_start = _time.perf_counter()
    try:
        mine = torch.from_numpy(vbg.peer_descriptor()).to(dev, non_blocking=True)
    except Exception as e:                      # e.g. CUDA IPC not permitted in this container
        err = e
        mine = torch.zeros(PEER_DESC_BYTES, dtype=torch.uint8, device=dev)
    table = torch.empty(world * mine.numel(), dtype=torch.uint8, device=dev)     # flat: valid for NCCL and gloo
    dist.all_gather_into_tensor(table, mine)
    descs = table.cpu().numpy().reshape(world, mine.numel())
    t_gather = _time.perf_counter()
    if not descs[:, :4].any(axis=1).all():      # some rank could not export (magic == 0): nobody pulls
        raise PeerPullError(f"peer-memory ghost pull unavailable on some rank (this rank: {err})")
    # identity of the exported pools: everything but the per-step block count (bytes 32..40)
    ident = descs[:, :32].tobytes() + descs[:, 40:336].tobytes()
    known = getattr(vbg, "_peer_pools_seen", None) == ident
    ok = 1
    try:
        n = vbg.ghost_pull(descs, want_count=count)
    except Exception as e:
        ok, err = 0, e
    # the fence doubles as the status exchange: every rank learns whether every pull was issued
    tok = _fence_token(dev)
    tok.fill_(ok)
    dist.all_reduce(tok, op=dist.ReduceOp.MIN)
    if not ok or (fence and not known):
        if int(tok.item()) == 0:
            raise PeerPullError(f"peer-memory ghost pull failed on some rank (this rank: {err})")
        vbg._peer_pools_seen = ident
    if trace:
        torch.cuda.synchronize()
        import sys as _sys
        print(f"[mq3d r{rank}] pull_ghosts: descriptors+wait {1e3 * (t_gather - t_start):.3f} ms, pull+fence "
              f"{1e3 * (_time.perf_counter() - t_gather):.3f} ms", file=_sys.stderr)
    return n


class PeerPullError(RuntimeError):
    """Raised on EVERY rank when the peer-memory pull could not be issued on some rank."""


PEER_DESC_BYTES = 512
_PULL_USABLE = [True]


def fill_ghost_shell(vbg, rank: Optional[int] = None, world: Optional[int] = None, mode: str = "pull") -> str:
    """Owned-only integration mode: fetch the ghost shell before extraction.  mode "pull" uses peer memory
    (pull_ghosts) and falls back -- consistently on all ranks, for the rest of the process -- to the packed
    NCCL exchange (exchange_ghosts) if CUDA IPC is unavailable; mode "exchange" forces the latter.  Returns
    the mode that was used."""
    if mode == "pull" and _PULL_USABLE[0]:
        try:
            pull_ghosts(vbg, rank, world, count=False)
            return "pull"
        except PeerPullError:
            _PULL_USABLE[0] = False
    exchange_ghosts(vbg, rank, world)
    return "exchange"


_FENCE: dict = {}


def _fence_token(dev: torch.device) -> torch.Tensor:
    key = (dev.type, dev.index)
    if key not in _FENCE:
        _FENCE[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return _FENCE[key]


def gather_mesh(vertices: torch.Tensor, normals: Optional[torch.Tensor], triangles: torch.Tensor, dst: int = 0,
                colors: Optional[torch.Tensor] = None):
    """Concatenate per-rank meshes on `dst`: vertex arrays are appended in rank order and triangle
    indices rebased by the exclusive scan of the per-rank vertex counts.  Returns (vertices, normals,
    triangles, counts [world,2]) on dst and (None, None, None, counts) elsewhere; with `colors` (per-vertex
    float32 [V,3]) a fifth element carries the gathered colours.  Vertices on edges owned by ghost blocks are
    emitted by every rank that references them (no welding)."""
    import torch.distributed as dist
    extra = () if colors is None else (colors,)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        counts = torch.tensor([[vertices.shape[0], triangles.shape[0]]], dtype=torch.int64)
        return (vertices, normals, triangles, counts) + extra
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = vertices.device
    mine = torch.tensor([vertices.shape[0], triangles.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.empty((world, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine) if dev.type == "cuda" else dist.all_gather(list(counts.unbind(0)), mine)
    counts_h = counts.cpu()
    nv, nt = counts_h[:, 0].tolist(), counts_h[:, 1].tolist()
    # exact-size point-to-point transfers straight into slices of the destination arrays (one NCCL group,
    # no padding, no concatenation); every rank must agree on which per-vertex arrays travel
    per_vertex = [vertices, normals, colors]           # None entries stay None in the result
    if rank != dst:
        ops = [dist.P2POp(dist.isend, a.contiguous(), dst) for a in per_vertex if a is not None and nv[rank]]
        if nt[rank]:
            ops.append(dist.P2POp(dist.isend, triangles.contiguous(), dst))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return (None, None, None, counts_h) + ((None,) if colors is not None else ())
    voff = [0] * (world + 1)
    toff = [0] * (world + 1)
    for r in range(world):
        voff[r + 1], toff[r + 1] = voff[r] + nv[r], toff[r] + nt[r]
    out = [None if a is None else torch.empty((voff[world], 3), dtype=a.dtype, device=dev) for a in per_vertex]
    t = torch.empty((toff[world], 3), dtype=triangles.dtype, device=dev)
    ops = []
    for r in range(world):
        vs, ts = slice(voff[r], voff[r + 1]), slice(toff[r], toff[r + 1])
        if r == rank:
            for o, a in zip(out, per_vertex):
                if o is not None:
                    o[vs] = a
            t[ts] = triangles
            continue
        if nv[r]:
            ops += [dist.P2POp(dist.irecv, o[vs], r) for o in out if o is not None]
        if nt[r]:
            ops.append(dist.P2POp(dist.irecv, t[ts], r))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for r in range(1, world):
        if nt[r] and voff[r]:
            t[toff[r]:toff[r + 1]] += voff[r]
    return (out[0], out[1], t, counts_h) + ((out[2],) if colors is not None else ())
