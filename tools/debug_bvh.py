import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mq3d_b200
from mq3d_b200 import _lib
from mq3d_b200.raycast import RaycastingScene
dev = torch.device('cuda', 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
xs = np.linspace(-1, 1, N + 1)
verts = np.array([[x, y, 2.0] for y in xs for x in xs], np.float32)
tris = []
for j in range(N):
    for i in range(N):
        a = j * (N + 1) + i
        tris += [[a, a + 1, a + N + 2], [a, a + N + 2, a + N + 1]]
tris = np.array(tris, np.int32)
scene = RaycastingScene(device=dev)
scene.add_triangles(torch.from_numpy(verts), torch.from_numpy(tris))
K = np.array([[100.0, 0, 32.0], [0, 100.0, 24.0], [0, 0, 1.0]])
rays = scene.create_rays_pinhole(K, np.eye(4), width_px=64, height_px=48)
t = scene.cast_rays(rays)['t_hit'].cpu().numpy()
print('tris', len(tris), 'hit frac', np.isfinite(t).mean(), 'tmin', t.min())
buf = np.zeros((len(tris), 16), np.float32)
n = _lib.lib().mq3d_scene_debug_nodes(scene._h, buf.ctypes.data_as(C.c_void_p), len(tris))
ib = buf.view(np.int32)
for k in range(min(n, 12)):
    print(k, 'L', buf[k, 0:6].round(2), 'R', buf[k, 6:12].round(2), 'lc rc', ib[k, 12], ib[k, 13])
