"""In-tree build of the sm_100a C-ABI library (``libmq3d.so``) with nvcc.

The shared object is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmq3d.so")
OBJ_DIR = os.path.join(HERE, "build")

SOURCES = ["mq3d_grid.cu", "mq3d_depth.cu", "mq3d_integrate.cu", "mq3d_mesh.cu", "mq3d_confidence.cu",
           "mq3d_raycast.cu", "mq3d_peer.cu", "mq3d_meshfilter.cu", "mq3d_odometry.cu"]
HEADERS = ["mq3d_common.cuh", "mc_tables.h", os.path.join("..", "..", "include", "mq3d.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (no CPU fallback exists)")


def _mtime(p: str) -> float:
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdr_time = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that lack a working spec dir; use the system compiler
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    objs, relink = [], force or not os.path.exists(LIB)
    procs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _mtime(obj) < max(_mtime(src), hdr_time):
            extra = os.environ.get("MQ3D_NVCC_EXTRA", "").split()          # experiments, e.g. -DMQ3D_TWO_PHASE
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-ccbin", ccbin] if ccbin else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((s, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
            relink = True
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{out.decode()}")
        if verbose and out:
            print(out.decode(), file=sys.stderr)
    if relink or any(_mtime(o) > _mtime(LIB) for o in objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + (["-ccbin", ccbin] if ccbin else []) + \
              ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
