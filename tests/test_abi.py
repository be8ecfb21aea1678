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


def test_ctypes_signatures_match_the_header():
    """Every prototype in include/mq3d.h has a ctypes signature with the same number of parameters and the same
    pointer / integer / floating kind per parameter (a mismatch would only show up as garbage on a GPU box)."""
    src = open(os.path.join(ROOT, "include", "mq3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|const char \*)\s*(mq3d_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    assert len(protos) == len(_declared())
    lib = _lib.lib()

    def kind_of_c(param):
        p = " ".join(param.split())
        if "*" in p or "[" in p:
            return "ptr"
        if re.match(r"^(const )?(float|double)\b", p):
            return "fp"
        return "int"

    def kind_of_ctypes(t):
        if t in (ctypes.c_float, ctypes.c_double):
            return "fp"
        if t in (ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint, ctypes.c_uint32):
            return "int"
        return "ptr"

    checked = 0
    for name, params in protos:
        params = [p for p in (q.strip() for q in params.split(",")) if p and p != "void"]
        argtypes = getattr(lib, name).argtypes
        if argtypes is None:                              # mq3d_last_error / mq3d_version take no arguments
            assert not params, name
            continue
        assert len(argtypes) == len(params), (name, len(argtypes), params)
        for i, (p, t) in enumerate(zip(params, argtypes)):
            assert kind_of_c(p) == kind_of_ctypes(t), (name, i, p, t)
            if kind_of_c(p) == "fp":                      # float vs double must agree as well
                assert ("double" in p) == (t is ctypes.c_double), (name, i, p, t)
            if kind_of_c(p) == "int":                     # 32 vs 64 bit
                assert ("int64_t" in p) == (ctypes.sizeof(t) == 8), (name, i, p, t)
        checked += 1
    assert checked >= 30
