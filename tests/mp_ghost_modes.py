"""Multi-process check of the three ghost-shell modes (run under torchrun, one rank per GPU; launched by
tests/test_gpu_multiproc.py).  Every rank builds the same synthetic capture, integrates it with
  A  integrate_ghosts=True                      (redundant integration, no communication)
  B  owned-only + dist.exchange_ghosts          (packed NCCL send/recv + import)
  C  owned-only + dist.pull_ghosts              (CUDA-IPC peer memory, copy kernel reads the owners' pools)
and requires bit-identical grids and meshes from all three, over several reset/integrate rounds (handle
caching, fences) and with a pool that has to grow during the pull.  Also checks the sharded upload of
pipeline.integrate_frames against the plain one, and gather_mesh against the per-rank meshes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import mq3d_b200  # noqa: E402,F401
from helpers import capture, pipeline_cameras, sort_blocks  # noqa: E402
from mq3d_b200 import synth  # noqa: E402
from mq3d_b200.dist import exchange_ghosts, gather_mesh, pull_ghosts  # noqa: E402
from mq3d_b200.pipeline import IntegrationParams, integrate_frames  # noqa: E402
from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare  # noqa: E402


def same(a, b):
    for x, y in zip(a, b):
        if x is None and y is None:
            continue
        if x.dtype == np.float32:
            x, y = x.view(np.uint32), y.view(np.uint32)
        if not np.array_equal(x, y):
            return False
    return True


def blocks(g):
    return sort_blocks(*[x.cpu().numpy() if x is not None else None for x in g.export_blocks()])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n, tile = 12, 2
    cap = capture(n)
    K, Ewc, Ecw = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    raw = torch.from_numpy(cap.raw).to(dev)
    lin, valid = depth_prepare(raw, ds.nears, ds.fars)
    f, cw, ch = 110.0, 160, 120
    cols_np = np.stack([synth.make_color_frame(Ecw[i], width=cw, height=ch, f=f) for i in range(n)])
    cols = torch.from_numpy(cols_np).to(dev)
    Kc = np.tile(np.array([[f, 0, cw / 2.0], [0, f, ch / 2.0], [0, 0, 1.0]]), (n, 1, 1))
    attrs = ("tsdf", "weight", "color")

    def make(integrate_ghosts, block_count):
        g = VoxelBlockGrid(attr_names=attrs, voxel_size=0.02, block_count=block_count, device=dev)
        g.set_partition(rank, world, tile, integrate_ghosts=integrate_ghosts)
        return g

    def run(g):
        g.reset()
        g.integrate_sequence(lin, K, Ewc, 4.0, 10.0, frame_valid=valid, colors=cols, color_intrinsics=Kc)

    ga, gb, gc = make(True, 4000), make(False, 4000), make(False, 64)
    run(ga)
    ref = blocks(ga)
    ref_mesh = [x.cpu().numpy() for x in ga.extract_triangle_mesh_arrays(1.5, with_colors=True)]
    for it in range(3):
        run(gb)
        nb = exchange_ghosts(gb, rank, world)
        run(gc)
        nc = pull_ghosts(gc, rank, world)
        assert nb == nc and nc > 0, (nb, nc)
        assert same(ref, blocks(gb)), f"rank {rank} round {it}: NCCL exchange differs from redundant integration"
        assert same(ref, blocks(gc)), f"rank {rank} round {it}: peer pull differs from redundant integration"
        mesh = [x.cpu().numpy() for x in gc.extract_triangle_mesh_arrays(1.5, with_colors=True)]
        assert len(mesh[0]) == len(ref_mesh[0]) and len(mesh[2]) == len(ref_mesh[2])
    # sharded upload == plain upload
    params = IntegrationParams(voxel_size=0.02, depth_max=4.0, trunc_voxel_multiplier=10.0,
                               use_confidence_filtered_depth=False, batch_frames=5)
    raw_host, col_host = torch.from_numpy(cap.raw).pin_memory(), torch.from_numpy(cols_np).pin_memory()
    gd = make(True, 4000)
    integrate_frames(gd, raw_host, ds.nears, ds.fars, K, Ewc, params, colors_host=col_host, Kc=Kc, shard=(rank, world))
    assert same(ref, blocks(gd)), f"rank {rank}: sharded upload differs"
    # pageable colour takes the copy-then-resample route; the mesh read-back goes through pinned staging
    ge = make(True, 4000)
    integrate_frames(ge, raw_host, ds.nears, ds.fars, K, Ewc, params, colors_host=torch.from_numpy(cols_np), Kc=Kc,
                     shard=(rank, world))
    assert same(ref, blocks(ge)), f"rank {rank}: sharded upload of pageable colour differs"
    from mq3d_b200.pipeline import extract_mesh_to_host
    host_mesh = extract_mesh_to_host(ge, 1.5, with_colors=True)
    # (block order, hence vertex order, depends on the allocation order: compare as sets keyed by position)
    oa, ob = np.lexsort(host_mesh[0].T[::-1]), np.lexsort(ref_mesh[0].T[::-1])
    for i in (0, 1, 3):
        assert np.array_equal(host_mesh[i][oa].view(np.uint32), ref_mesh[i][ob].view(np.uint32)), i
    assert host_mesh[2].shape == ref_mesh[2].shape
    # gather: rank 0 ends up with every rank's vertices / triangles, indices rebased
    v, nrm, t = ga.extract_triangle_mesh_arrays(1.5)
    gv, gn, gt, counts = gather_mesh(v, nrm, t, dst=0)
    assert int(counts[rank, 0]) == v.shape[0] and int(counts[rank, 1]) == t.shape[0]
    if rank == 0:
        assert gv.shape[0] == int(counts[:, 0].sum()) and gt.shape[0] == int(counts[:, 1].sum())
        assert torch.equal(gv[: v.shape[0]], v) and torch.equal(gt[: t.shape[0]], t)
        assert int(gt.max()) < gv.shape[0] and int(gt[t.shape[0]:].min()) >= v.shape[0]
        tri = gv[gt.long()]
        assert torch.isfinite(tri).all()
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"MP_GHOST_MODES_OK world={world} ghosts={nc}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
