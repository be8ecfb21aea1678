"""B200-native volumetric reconstruction hot path for Meta Quest captures.

Drop-in for the Open3D VoxelBlockGrid / RaycastingScene calls made by the reference's
``reconstruct_scene.py`` (see DESIGN.md / INTEGRATION.md).  All arithmetic runs in hand-written
sm_100a CUDA kernels behind the C-ABI declared in ``include/mq3d.h``; there is no CPU fallback.
"""
from .models import (CameraDataset, ConfidenceMap, CoordinateSystem, DepthDataset, Side,
                     Transforms)

__all__ = ["CameraDataset", "ConfidenceMap", "CoordinateSystem", "DepthDataset", "Side", "Transforms"]
