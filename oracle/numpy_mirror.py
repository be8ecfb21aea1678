"""Second, independent CPU restatement (NumPy, float32 ufuncs) of K2 and K3 -- TEST INFRASTRUCTURE ONLY.

SURVEY 8c item 10 asks for a cross-restatement: the C oracle (oracle/mq3d_oracle.c), this NumPy mirror and the
CUDA kernels must agree bit for bit on random scenes.  This file is written from the specification in SURVEY
Appendix A.1-A.3 (Open3D 0.19 TransformIndexer / DepthTouch / Integrate as reached from
processing/reconstruction/utils/o3d_utils.py:212-229), not from the C code, so that a misreading of the
specification in one of the two shows up as a mismatch.  Like everything under oracle/ it may only be
imported by tests/; "parity unpinned" applies (no real Open3D was available to check either restatement).

Every arithmetic step is a separate float32 NumPy ufunc call, i.e. IEEE round-to-nearest without fused
multiply-add -- the same evaluation the oracle gets from `-ffp-contract=off` and the kernels from `__f*_rn`.
"""
from __future__ import annotations

import numpy as np

RES = 16
F = np.float32


def _camera(K, E):
    """TransformIndexer storage: float32 casts of the float64 inputs (A.1)."""
    K = np.asarray(K, np.float64)
    E = np.asarray(E, np.float64)
    return (F(K[0, 0]), F(K[1, 1]), F(K[0, 2]), F(K[1, 2])), E[:3, :4].astype(np.float32)


def _rigid(e, scale, x, y, z):
    x, y, z = x * scale, y * scale, z * scale
    return (x * e[0, 0] + y * e[0, 1] + z * e[0, 2] + e[0, 3],
            x * e[1, 0] + y * e[1, 1] + z * e[1, 2] + e[1, 3],
            x * e[2, 0] + y * e[2, 1] + z * e[2, 2] + e[2, 3])


def _inverse_transformation(E):
    """[R^T | -R^T t] in float64 (A.1)."""
    E = np.asarray(E, np.float64)
    P = np.eye(4)
    P[:3, :3] = E[:3, :3].T
    for i in range(3):
        P[i, 3] = -(P[i, 0] * E[0, 3] + P[i, 1] * E[1, 3] + P[i, 2] * E[2, 3])
    return P


def _round_half_away(x):
    """roundf for x >= 0 without the x + 0.5 rounding hazard."""
    r = np.floor(x)
    return (r + ((x - r) >= F(0.5))).astype(np.int64)


def touch(depth, K, E, voxel_size, depth_max, trunc_mult, depth_scale=1.0):
    """A.2 DepthTouch: set of (x, y, z) block keys; raises like Open3D when nothing is touched."""
    depth = np.asarray(depth, np.float32)
    H, W = depth.shape
    vs = F(voxel_size)
    trunc = vs * F(trunc_mult)
    block_size = vs * F(RES)
    dmax = F(depth_max)
    (fx, fy, cx, cy), e = _camera(K, _inverse_transformation(E))
    ys, xs = np.meshgrid(np.arange(0, (H // 4) * 4, 4), np.arange(0, (W // 4) * 4, 4), indexing="ij")
    d = depth[ys, xs] / F(depth_scale)
    ok = (d > 0) & (d < dmax)
    xf, yf = xs.astype(np.float32)[ok], ys.astype(np.float32)[ok]
    d = d[ok]
    one = F(1.0)
    px, py, pz = (xf - cx) * one / fx, (yf - cy) * one / fy, np.full_like(xf, one)
    gx, gy, gz = _rigid(e, one, px, py, pz)
    ox, oy, oz = e[0, 3], e[1, 3], e[2, 3]
    dx, dy, dz = gx - ox, gy - oy, gz - oz
    t_min = np.maximum(d - trunc, F(0))
    t_max = np.minimum(d + trunc, dmax)
    t_step = (t_max - t_min) / F(3)
    keys = set()
    t = t_min.copy()
    for _ in range(4):
        kx = np.floor((ox + t * dx) / block_size).astype(np.int32)
        ky = np.floor((oy + t * dy) / block_size).astype(np.int32)
        kz = np.floor((oz + t * dz) / block_size).astype(np.int32)
        keys.update(zip(kx.tolist(), ky.tolist(), kz.tolist()))
        t = t + t_step
    if not keys:
        raise RuntimeError("No block is touched in TSDF volume")
    return keys


def integrate(grid: dict, keys, depth, K, E, voxel_size, depth_max, trunc_mult, color=None, Kc=None, depth_scale=1.0):
    """A.3 Integrate.  `grid` maps key -> [tsdf f32 [16,16,16], weight f32 [16,16,16], color f32 [16,16,16,3]]
    (z, y, x index order; blocks are created zero-filled on first use).  Returns the number of voxel updates."""
    depth = np.asarray(depth, np.float32)
    H, W = depth.shape
    vs = F(voxel_size)
    trunc = vs * F(trunc_mult)
    dmax = F(depth_max)
    (fx, fy, cx, cy), e = _camera(K, E)
    if color is not None:
        (cfx, cfy, ccx, ccy), _ = _camera(Kc, np.eye(4))
        CH, CW = color.shape[:2]
    zv, yv, xv = np.meshgrid(np.arange(RES), np.arange(RES), np.arange(RES), indexing="ij")
    updated = 0
    for key in keys:
        key = tuple(int(k) for k in key)
        blk = grid.setdefault(key, [np.zeros((RES,) * 3, np.float32), np.zeros((RES,) * 3, np.float32),
                                    np.zeros((RES,) * 3 + (3,), np.float32)])
        tsdf, weight, col = blk
        x = (key[0] * RES + xv).astype(np.float32)
        y = (key[1] * RES + yv).astype(np.float32)
        z = (key[2] * RES + zv).astype(np.float32)
        xc, yc, zc = _rigid(e, vs, x, y, z)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            inv_z = F(1.0) / zc
            u = fx * xc * inv_z + cx
            v = fy * yc * inv_z + cy
        inb = (v >= 0) & (u >= 0) & (v <= F(H) - F(1)) & (u <= F(W) - F(1))
        ui = np.where(inb, u, 0).astype(np.int32)          # (int) truncation
        vi = np.where(inb, v, 0).astype(np.int32)
        d = depth[vi, ui] / F(depth_scale)
        sdf = d - zc
        ok = inb & ~(d <= 0) & ~(d > dmax) & ~(zc <= 0) & ~(sdf < -trunc)
        s = np.where(sdf < trunc, sdf, trunc) / trunc
        inv = F(1.0) / (weight + F(1))
        tsdf[...] = np.where(ok, (weight * tsdf + s) * inv, tsdf)
        if color is not None:
            px = (ui.astype(np.float32) - cx) * F(1.0) / fx
            py = (vi.astype(np.float32) - cy) * F(1.0) / fy
            uf = cfx * px * F(1.0) + ccx                    # Project with z = 1: inv_z = 1
            vf = cfy * py * F(1.0) + ccy
            cin = ok & (vf >= 0) & (uf >= 0) & (vf <= F(CH) - F(1)) & (uf <= F(CW) - F(1))
            cu = np.clip(_round_half_away(np.where(cin, uf, 0)), 0, CW - 1)
            cv = np.clip(_round_half_away(np.where(cin, vf, 0)), 0, CH - 1)
            rgb = color[cv, cu].astype(np.float32)
            new = (weight[..., None] * col + rgb * F(1.0)) * inv[..., None]
            col[...] = np.where(cin[..., None], new, col)
        weight[...] = np.where(ok, weight + F(1), weight)
        updated += int(ok.sum())
    return updated


# ------------------------------------------------------------------------------------------------------------
# A.4 extract_triangle_mesh, dense formulation over the bounding box of the allocated blocks
# ------------------------------------------------------------------------------------------------------------
def extract_mesh(grid: dict, voxel_size, weight_threshold):
    """(vertices f32 [V,3], normals f32 [V,3], triangles i64 [T,3], vertex_keys i32 [V,4]) for `grid`
    (key -> [tsdf, weight, ...] as used by integrate).  Vertex key = (global voxel x, y, z, axis) of the lattice
    edge that carries the vertex; order of vertices / triangles is arbitrary (compare in canonical form)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    from mc_tables_data import CORNERS, EDGE_SHIFTS, EDGE_TABLE, TRI_TABLE   # public-domain Bourke tables

    if not grid:
        return (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int64),
                np.zeros((0, 4), np.int32))
    vs, thr, PAD = F(voxel_size), F(weight_threshold), 3
    keys = np.array(list(grid), np.int64)
    lo = keys.min(0)
    dims = (keys.max(0) - lo + 1) * RES + 2 * PAD
    T = np.zeros((dims[2], dims[1], dims[0]), np.float32)
    Wt = np.zeros_like(T)
    EX = np.zeros(T.shape, bool)
    for key, blk in grid.items():
        o = (np.array(key) - lo) * RES + PAD
        sl = (slice(o[2], o[2] + RES), slice(o[1], o[1] + RES), slice(o[0], o[0] + RES))
        T[sl], Wt[sl], EX[sl] = blk[0], blk[1], True

    def sh(a, dx, dy, dz):                     # a[z + dz, y + dy, x + dx] (PAD keeps the wrap-around unused)
        return np.roll(a, (-dz, -dy, -dx), axis=(0, 1, 2))

    ok = (EX & (Wt > thr))                      # corner usable: exists and w > threshold (w <= thr rejects)
    valid = np.ones(T.shape, bool)
    table = np.zeros(T.shape, np.int32)
    for i, (dx, dy, dz) in enumerate(CORNERS):
        valid &= sh(ok, dx, dy, dz)
        table |= (sh(T, dx, dy, dz) < 0).astype(np.int32) << i
    valid &= EX
    table = np.where(valid, table, 0)
    edge_bits = np.asarray(EDGE_TABLE, np.int32)[table]
    mark = np.zeros((3,) + T.shape, bool)      # mark[axis][owner voxel]
    for i, (dx, dy, dz, ax) in enumerate(EDGE_SHIFTS):
        has = ((edge_bits >> i) & 1).astype(bool)
        mark[ax] |= sh(has, -dx, -dy, -dz)     # owner = cube + shift  <=>  cube = owner - shift
    unit = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]

    def grad(a):                               # (value, available) of the central difference along axis a
        d = unit[a]
        return sh(T, *d) - sh(T, -d[0], -d[1], -d[2]), sh(EX, *d) & sh(EX, -d[0], -d[1], -d[2])

    g_val, g_ok = zip(*[grad(a) for a in range(3)])
    no = [np.where(g_ok[a], g_val[a], F(0)) for a in range(3)]          # gradient at the owner voxel
    ne = [np.zeros(T.shape, np.float32) for _ in range(3)]              # stale across the edges of one voxel
    zi, yi, xi = np.meshgrid(*[np.arange(n) for n in T.shape], indexing="ij")
    gx = (xi - PAD + lo[0] * RES).astype(np.int32)
    gy = (yi - PAD + lo[1] * RES).astype(np.int32)
    gz = (zi - PAD + lo[2] * RES).astype(np.int32)
    verts, norms, vkeys = [], [], []
    with np.errstate(divide="ignore", invalid="ignore"):
        for e in range(3):
            m = mark[e]
            d = unit[e]
            for a in range(3):                 # gradient at the far end of the edge, where available
                ne[a] = np.where(m & sh(g_ok[a], *d), sh(g_val[a], *d), ne[a])
            t_o, t_e = T[m], sh(T, *d)[m]
            ratio = (F(0) - t_o) / (t_e - t_o)
            g = [gx[m].astype(np.float32), gy[m].astype(np.float32), gz[m].astype(np.float32)]
            pos = np.stack([vs * (g[a] + ratio * F(1.0 if a == e else 0.0)) for a in range(3)], axis=1)
            n3 = [(F(1) - ratio) * no[a][m] + ratio * ne[a][m] for a in range(3)]
            norm = (np.sqrt(n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2]).astype(np.float64) + 1e-5).astype(np.float32)
            verts.append(pos)
            norms.append(np.stack([n3[a] / norm for a in range(3)], axis=1))
            vkeys.append(np.stack([gx[m], gy[m], gz[m], np.full(m.sum(), e, np.int32)], axis=1))
    verts, norms, vkeys = np.concatenate(verts), np.concatenate(norms), np.concatenate(vkeys)
    index = {tuple(k): i for i, k in enumerate(vkeys.tolist())}
    tris = []
    for z, y, x in zip(*np.nonzero((table != 0) & (table != 255))):
        row = TRI_TABLE[table[z, y, x]]
        for k in range(0, 15, 3):
            if row[k] < 0:
                break
            ids = []
            for edge in row[k:k + 3]:
                dx, dy, dz, ax = EDGE_SHIFTS[edge]
                ids.append(index[(int(gx[z, y, x]) + dx, int(gy[z, y, x]) + dy, int(gz[z, y, x]) + dz, ax)])
            tris.append(ids[::-1])             # winding reversed: tri[2 - k] = id_k
    return verts, norms, np.array(tris, np.int64).reshape(-1, 3), vkeys.astype(np.int32)


def extract_points(grid: dict, voxel_size, weight_threshold):
    """A.5 extract_point_cloud: (points f32 [P,3], normals f32 [P,3], keys i32 [P,4]); a voxel with w > thr emits a
    point on each +axis edge whose far voxel exists, has w > thr and tsdf_far * tsdf_own < 0 (the float32 product)."""
    if not grid:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 4), np.int32)
    vs, thr, PAD = F(voxel_size), F(weight_threshold), 3
    keys = np.array(list(grid), np.int64)
    lo = keys.min(0)
    dims = (keys.max(0) - lo + 1) * RES + 2 * PAD
    T = np.zeros((dims[2], dims[1], dims[0]), np.float32)
    Wt = np.zeros_like(T)
    EX = np.zeros(T.shape, bool)
    for key, blk in grid.items():
        o = (np.array(key) - lo) * RES + PAD
        sl = (slice(o[2], o[2] + RES), slice(o[1], o[1] + RES), slice(o[0], o[0] + RES))
        T[sl], Wt[sl], EX[sl] = blk[0], blk[1], True

    def sh(a, dx, dy, dz):
        return np.roll(a, (-dz, -dy, -dx), axis=(0, 1, 2))

    unit = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    ok = EX & (Wt > thr)
    g_val = [sh(T, *d) - sh(T, -d[0], -d[1], -d[2]) for d in unit]
    g_ok = [sh(EX, *d) & sh(EX, -d[0], -d[1], -d[2]) for d in unit]
    no = [np.where(g_ok[a], g_val[a], F(0)) for a in range(3)]
    ni = [np.zeros(T.shape, np.float32) for _ in range(3)]
    zi, yi, xi = np.meshgrid(*[np.arange(n) for n in T.shape], indexing="ij")
    gx = (xi - PAD + lo[0] * RES).astype(np.int32)
    gy = (yi - PAD + lo[1] * RES).astype(np.int32)
    gz = (zi - PAD + lo[2] * RES).astype(np.int32)
    pts, norms, pkeys = [], [], []
    with np.errstate(divide="ignore", invalid="ignore", under="ignore"):
        for e in range(3):
            d = unit[e]
            m = ok & sh(ok, *d) & ((sh(T, *d) * T) < 0)
            for a in range(3):
                ni[a] = np.where(m & sh(g_ok[a], *d), sh(g_val[a], *d), ni[a])
            t_o, t_i = T[m], sh(T, *d)[m]
            ratio = (F(0) - t_o) / (t_i - t_o)
            g = [gx[m].astype(np.float32), gy[m].astype(np.float32), gz[m].astype(np.float32)]
            pts.append(np.stack([vs * (g[a] + ratio * F(1.0 if a == e else 0.0)) for a in range(3)], axis=1))
            n3 = [(F(1) - ratio) * no[a][m] + ratio * ni[a][m] for a in range(3)]
            norm = (np.sqrt(n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2]).astype(np.float64) + 1e-5).astype(np.float32)
            norms.append(np.stack([n3[a] / norm for a in range(3)], axis=1))
            pkeys.append(np.stack([gx[m], gy[m], gz[m], np.full(m.sum(), e, np.int32)], axis=1))
    return np.concatenate(pts), np.concatenate(norms), np.concatenate(pkeys).astype(np.int32)


def odometry_information(source_depth, target_depth, K, T, dist_threshold, depth_scale, depth_max):
    """Independent dense restatement of compute_odometry_information_matrix: float32 geometry as
    Open3D's vertex maps and TransformIndexer, float64 accumulation.  Returns float64 [6,6]."""
    f32 = np.float32
    H, W = source_depth.shape
    fx, fy, cx, cy = f32(K[0][0]), f32(K[1][1]), f32(K[0][2]), f32(K[1][2])
    R, t = np.asarray(T, np.float64)[:3, :3].astype(f32), np.asarray(T, np.float64)[:3, 3].astype(f32)

    def vertex_map(depth):
        d = (depth.astype(f32) / f32(depth_scale)).astype(f32)
        ok = (d > 0) & (d < f32(depth_max))
        u, v = np.meshgrid(np.arange(W, dtype=f32), np.arange(H, dtype=f32))
        x = ((u - cx) * d / fx).astype(f32)
        y = ((v - cy) * d / fy).astype(f32)
        return np.stack([x, y, d], -1), ok
    sv, sok = vertex_map(source_depth)
    tv, tok = vertex_map(target_depth)
    q = np.stack([(R[i, 0] * sv[..., 0] + R[i, 1] * sv[..., 1]).astype(f32) for i in range(3)], -1)
    q = np.stack([((q[..., i] + R[i, 2] * sv[..., 2]).astype(f32) + t[i]).astype(f32) for i in range(3)], -1)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv_z = (f32(1.0) / q[..., 2]).astype(f32)
        u = np.round(((fx * q[..., 0]).astype(f32) * inv_z + cx).astype(f32))
        v = np.round(((fy * q[..., 1]).astype(f32) * inv_z + cy).astype(f32))
    # np.round is half-to-even, roundf is half-away-from-zero: redo the ties
    with np.errstate(divide="ignore", invalid="ignore"):
        uf = ((fx * q[..., 0]).astype(f32) * inv_z + cx).astype(f32)
        vf = ((fy * q[..., 1]).astype(f32) * inv_z + cy).astype(f32)
    u = np.where(np.isfinite(uf), np.sign(uf) * np.floor(np.abs(uf) + f32(0.5)), uf)
    v = np.where(np.isfinite(vf), np.sign(vf) * np.floor(np.abs(vf) + f32(0.5)), vf)
    ok = sok & ~(q[..., 2] < 0) & (u >= 0) & (v >= 0) & (u <= W - 1) & (v <= H - 1)
    ui, vi = np.where(ok, u, 0).astype(np.int64), np.where(ok, v, 0).astype(np.int64)
    ok &= tok[vi, ui]
    r = (q - tv[vi, ui]).astype(f32)
    r2 = ((r[..., 0] * r[..., 0] + r[..., 1] * r[..., 1]).astype(f32) + r[..., 2] * r[..., 2]).astype(f32)
    ok &= r2 <= f32(dist_threshold) * f32(dist_threshold)
    p = q[ok].astype(np.float64)
    z, o = np.zeros(len(p)), np.ones(len(p))
    J = np.stack([np.stack([z, p[:, 2], -p[:, 1], o, z, z], 1), np.stack([-p[:, 2], z, p[:, 0], z, o, z], 1),
                  np.stack([p[:, 1], -p[:, 0], z, z, z, o], 1)], 1)       # [n,3,6]
    return np.einsum("nki,nkj->ij", J, J)
