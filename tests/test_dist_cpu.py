"""CPU: the N>1 host path on the gloo backend (world_size 2) -- frame broadcast + final mesh gather --
and the ownership rule of the hash partition."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

import mq3d_b200  # noqa: F401
from mq3d_b200.dist import block_needed, tile_owner


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import mq3d_b200  # noqa: F401
    from mq3d_b200.dist import broadcast_frames, gather_mesh
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4) if rank == 0 else torch.zeros(2, 3, 4)
    broadcast_frames(frames, 0)
    assert torch.equal(frames, torch.arange(24, dtype=torch.float32).reshape(2, 3, 4))
    nv = 3 + 2 * rank                                     # ragged sizes; rank 1 also tests a non-empty tail
    v = torch.full((nv, 3), float(rank)) + torch.arange(nv)[:, None]
    n = -v
    t = torch.tensor([[0, 1, 2]] * (1 + rank), dtype=torch.int32)
    gv, gn, gt, counts = gather_mesh(v, n, t, dst=0)
    assert counts.tolist() == [[3, 1], [5, 2]]
    if rank == 0:
        assert gv.shape == (8, 3) and gt.shape == (3, 3)
        assert gt.tolist() == [[0, 1, 2], [3, 4, 5], [3, 4, 5]]          # rank-1 indices rebased by 3
        assert torch.equal(gv[3:], torch.full((5, 3), 1.0) + torch.arange(5)[:, None]) and torch.equal(gn, -gv)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert gv is None and gt is None
    dist.destroy_process_group()


def test_broadcast_and_gather_world2(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_partition_ownership_rule():
    world, T = 4, 4
    keys = [(x, y, z) for x in range(-9, 9) for y in range(-3, 5) for z in range(-6, 7)]
    owners = np.array([tile_owner(*k, world, T) for k in keys])
    assert set(owners.tolist()) == {0, 1, 2, 3}                        # every rank owns something
    assert np.bincount(owners).min() > 0.1 * len(keys)                 # no rank starves
    for k in keys[::17]:
        o = tile_owner(*k, world, T)
        assert tile_owner(k[0] - k[0] % T, k[1] - k[1] % T, k[2] - k[2] % T, world, T) == o   # tile-constant
        needed = [r for r in range(world) if block_needed(*k, r, world, T)]
        assert o in needed
        nb_owners = {tile_owner(k[0] + dx, k[1] + dy, k[2] + dz, world, T)
                     for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)}
        assert set(needed) == nb_owners                                # ghost shell = owners of the 26 neighbours
    assert block_needed(1, 2, 3, 0, 1, T)                              # world 1: everything is local


class _FakeGrid:
    """Stands in for VoxelBlockGrid in the gloo test: blocks are rows of CPU tensors."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.blocks = {}                       # key tuple -> (tsdf, weight)
        for i in range(3 + rank):              # rank r owns 3+r blocks with recognisable values
            self.blocks[(rank, i, 0)] = (torch.full((4096,), float(10 * rank + i)), torch.full((4096,), float(i)))

    def ghost_select(self, dest):
        keys = [k for k in self.blocks if k[0] == self.rank and (k[1] + dest) % 2 == 0]   # arbitrary rule
        k = torch.tensor(keys, dtype=torch.int32).reshape(-1, 3)
        t = torch.stack([self.blocks[x][0] for x in keys]) if keys else torch.zeros((0, 4096))
        w = torch.stack([self.blocks[x][1] for x in keys]) if keys else torch.zeros((0, 4096))
        return k, t, w, None

    def import_blocks(self, keys, tsdf, weight, color=None):
        for i, key in enumerate(keys.tolist()):
            self.blocks[tuple(key)] = (tsdf[i].clone(), weight[i].clone())


def _ghost_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import mq3d_b200  # noqa: F401
    from mq3d_b200.dist import exchange_ghosts
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = _FakeGrid(rank, world)
    got = exchange_ghosts(g)
    other = 1 - rank
    want = [(other, i, 0) for i in range(3 + other) if (i + rank) % 2 == 0]
    assert got == len(want)
    for key in want:
        assert key in g.blocks and float(g.blocks[key][0][7]) == 10 * other + key[1]
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_ghost_exchange_world2(tmp_path):
    mp.spawn(_ghost_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


class _NoIpcGrid(_FakeGrid):
    """A grid whose peer-memory path is unavailable on ONE rank only (e.g. CUDA IPC refused in a container)."""
    device = "cpu"

    def peer_descriptor(self):
        if self.rank == 1:
            raise RuntimeError("cudaIpcGetMemHandle -> operation not supported")
        return np.full(512, 7, np.uint8)

    def ghost_pull(self, descs, want_count=True):
        raise AssertionError("must not be reached: rank 1 published an empty descriptor, which every rank sees in the "
                             "gathered table before anybody pulls")


class _BadPullGrid(_FakeGrid):
    """Descriptors are fine, but opening a peer's pool fails on rank 0."""
    device = "cpu"

    def peer_descriptor(self):
        return np.full(512, self.rank + 1, np.uint8)

    def ghost_pull(self, descs, want_count=True):
        assert descs.shape == (self.world, 512) and descs[0, 0] == 1 and descs[1, 0] == 2   # rank order
        if self.rank == 0:
            raise RuntimeError("cudaIpcOpenMemHandle -> invalid device context")
        return 0


def _fallback_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import dist as mqd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    other = 1 - rank
    want = [(other, i, 0) for i in range(3 + other) if (i + rank) % 2 == 0]
    # (1) a failing pull on one rank raises PeerPullError on BOTH ranks (nobody is left waiting in a collective)
    g = _BadPullGrid(rank, world)
    try:
        mqd.pull_ghosts(g, rank, world)
        raise AssertionError("pull_ghosts should have raised")
    except mqd.PeerPullError:
        pass
    # (2) fill_ghost_shell then falls back to the NCCL/gloo exchange on every rank, and remembers it
    mqd._PULL_USABLE[0] = True
    g = _NoIpcGrid(rank, world) if rank == 1 else _NoIpcGridRank0(rank, world)
    assert mqd.fill_ghost_shell(g, rank, world, "pull") == "exchange"
    for key in want:
        assert key in g.blocks and float(g.blocks[key][0][7]) == 10 * other + key[1]
    assert mqd._PULL_USABLE[0] is False
    g2 = _FakeGrid(rank, world)                       # no peer_descriptor at all: would fail if pull were retried
    assert mqd.fill_ghost_shell(g2, rank, world, "pull") == "exchange"
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


class _NoIpcGridRank0(_NoIpcGrid):
    pass


def test_ghost_pull_failure_falls_back_on_all_ranks(tmp_path):
    mp.spawn(_fallback_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
