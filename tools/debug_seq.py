import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mq3d_b200
from mq3d_b200 import synth, synth_gpu
from mq3d_b200.models import *
from mq3d_b200.vbg import VoxelBlockGrid, depth_prepare
dev = torch.device('cuda', 0)
F = 300
pos, quat = synth.eye_poses(F, Side.LEFT)
tr = Transforms(CoordinateSystem.UNITY, pos, quat).convert_coordinate_system(CoordinateSystem.OPEN3D, True)
fx, fy, cx, cy = synth.depth_intrinsics()
K = np.zeros((F, 3, 3), np.float32); K[:, 0, 0], K[:, 1, 1], K[:, 2, 2] = fx, fy, 1.0
K[:, 0, 2], K[:, 1, 2] = synth.DEPTH_W - cx, cy
raw = synth_gpu.render_depth(tr.extrinsics_cw, dev)
lin, valid = depth_prepare(raw, np.full(F, synth.NEAR), np.full(F, synth.FAR))
print('valid sum', int(valid.sum()), 'lin range', float(lin.min()), float(lin.max()), 'frame 280 center', float(lin[280,160,160]))
for lo, hi in [(0, 64), (256, 300), (0, 300)]:
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=20000, device=dev)
    st = vbg.integrate_sequence(lin[lo:hi].contiguous(), K[lo:hi], tr.extrinsics_wc[lo:hi], 4.0, 10.0, frame_valid=valid[lo:hi].contiguous())
    print(lo, hi, st)
