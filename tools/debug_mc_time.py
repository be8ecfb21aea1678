import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mq3d_b200
from mq3d_b200 import synth, synth_gpu, _lib
from mq3d_b200.models import *
from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare, _stream
dev = torch.device('cuda', 0)
F = 300
voxel = float(sys.argv[1]) if len(sys.argv) > 1 else 0.005
pos, quat = synth.eye_poses(F, Side.LEFT)
tr = Transforms(CoordinateSystem.UNITY, pos, quat).convert_coordinate_system(CoordinateSystem.OPEN3D, True)
fx, fy, cx, cy = synth.depth_intrinsics()
K = np.zeros((F, 3, 3), np.float32); K[:, 0, 0], K[:, 1, 1], K[:, 2, 2] = fx, fy, 1.0
K[:, 0, 2], K[:, 1, 2] = synth.DEPTH_W - cx, cy
raw = synth_gpu.render_depth(tr.extrinsics_cw, dev)
lin, valid = depth_prepare(raw, np.full(F, synth.NEAR), np.full(F, synth.FAR))
vbg = VoxelBlockGrid(voxel_size=voxel, block_count=400000, device=dev)
st = vbg.integrate_sequence(lin, K, tr.extrinsics_wc, 4.0, 10.0, frame_valid=valid)
print(st)
lib = _lib.lib()
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    V, T = C.c_int64(), C.c_int64()
    _lib.check(lib.mq3d_extract_mesh_count(vbg._h, C.c_float(1.5), C.byref(V), C.byref(T), _stream()))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    verts = torch.empty((V.value, 3), dtype=torch.float32, device=dev)
    normals = torch.empty((V.value, 3), dtype=torch.float32, device=dev)
    tris = torch.empty((T.value, 3), dtype=torch.int32, device=dev)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    _lib.check(lib.mq3d_extract_mesh_fill(vbg._h, _lib.dptr(verts), _lib.dptr(normals), _lib.dptr(tris), None, _stream()))
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f'iter {it}: count {1e3*(t1-t0):.3f} ms  alloc {1e3*(t2-t1):.3f} ms  fill {1e3*(t3-t2):.3f} ms  V={V.value} T={T.value}')
    del verts, normals, tris
