"""Array-level public API of the hot path: host buffers in, grid / mesh out.

This is what ``integrate()`` / ``reconstruct_scene()`` (o3d_utils.py:153-238, reconstruct_scene.py:60-108)
boil down to once the frames are in memory: H2D copy of the raw Quest depth (and colour), K1 linearise
+ mask, fused K2/K3 integration, K5 extraction.  bench.py's end-to-end number times exactly this.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .vbg import SequenceStats, VoxelBlockGrid, depth_prepare


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
                     Kc=None) -> SequenceStats:
    """Integrate [F,H,W] raw NDC depth frames (host, ideally pinned) into `vbg`.

    conf/count (float64 / int32 [F,H,W], host or device) enable the reference's confidence mask
    (o3d_utils.py:131-142).  colors_host: uint8 [F,CH,CW,3] enables Open3D's colour overload."""
    dev = vbg.device
    raw = raw_host.to(dev, non_blocking=True)
    mask = params.use_confidence_filtered_depth and conf is not None
    lin, valid = depth_prepare(raw, nears, fars,
                               conf.to(dev, non_blocking=True) if mask else None,
                               count.to(dev, non_blocking=True) if mask else None,
                               has_conf if mask else None,
                               params.confidence_threshold, params.valid_count_threshold)
    colors = colors_host.to(dev, non_blocking=True) if colors_host is not None and vbg.has_color else None
    return vbg.integrate_sequence(lin, K, E_wc, params.depth_max, params.trunc_voxel_multiplier, 1.0,
                                  frame_valid=valid, colors=colors, color_intrinsics=Kc,
                                  batch_frames=params.batch_frames)


def extract_mesh_to_host(vbg: VoxelBlockGrid, weight_threshold: float):
    """K5 + D2H: (vertices f32 [V,3], normals f32 [V,3], triangles i32 [T,3]) as numpy arrays."""
    v, n, t = vbg.extract_triangle_mesh_arrays(weight_threshold)
    return v.cpu().numpy(), n.cpu().numpy(), t.cpu().numpy()
