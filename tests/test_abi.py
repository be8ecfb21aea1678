"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/mq3d.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import mq3d_b200  # noqa: F401
from mq3d_b200 import _lib
from mq3d_b200.build import build_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mq3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mq3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(build_lib())
    declared = _declared()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared, "python binding list out of sync with include/mq3d.h"


def test_errors_without_gpu_are_loud():
    import torch
    lib = _lib.lib()
    assert lib.mq3d_version() >= 100
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = lib.mq3d_grid_create(ctypes.c_float(0.01), 16, 100, 1, 0, ctypes.byref(h))
    assert rc != 0 and lib.mq3d_last_error()          # no CPU fallback: creation fails
    import pytest
    from mq3d_b200.vbg import VoxelBlockGrid
    with pytest.raises(RuntimeError):
        VoxelBlockGrid(voxel_size=0.01, block_count=10, device="CUDA:0")
    with pytest.raises(RuntimeError):
        VoxelBlockGrid(voxel_size=0.01, block_count=10, device="CPU:0")
