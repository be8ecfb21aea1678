"""CPU: analytic known-answer tests that pin the oracle's Open3D-0.19 restatement without Open3D
(SURVEY.md section 8c, items 2-9).  "Parity unpinned" caveat: these check the restated semantics
against closed forms, not against a real Open3D build."""
import numpy as np
import pytest

K = np.array([[160.0, 0, 160.0], [0, 160.0, 160.0], [0, 0, 1.0]])
I4 = np.eye(4)
VS = 0.02
TRUNC_MULT = 10.0
TRUNC = np.float32(VS) * np.float32(TRUNC_MULT)


def wall(z0, shape=(320, 320)):
    return np.full(shape, z0, dtype=np.float32)


def grid_dict(g):
    keys, tsdf, weight, _ = g.export()
    return {tuple(k): (tsdf[i], weight[i]) for i, k in enumerate(keys.tolist())}


def test_front_parallel_wall_closed_form(oracle):
    """Identity pose, constant depth z0: tsdf = min(z0 - z, trunc)/trunc where z0 - z >= -trunc and
    the voxel projects inside the image with z > 0; weight = number of frames."""
    z0 = 1.0
    g = oracle.Grid(VS)
    d = wall(z0)
    keys = g.touch(d, K, I4, 4.0, TRUNC_MULT)
    # 4-sample rule: z in [z0 - trunc, z0 + trunc] -> block rows floor(z / 0.32) = 2, 3
    assert set(keys[:, 2].tolist()) == {2, 3}
    n_frames = 3
    for _ in range(n_frames):
        g.integrate(keys, d, K, I4, 4.0, TRUNC_MULT)
    vs = np.float32(VS)
    for key, (tsdf, weight) in grid_dict(g).items():
        zz, yy, xx = np.meshgrid(*[np.arange(16, dtype=np.int32)] * 3, indexing="ij")
        gx = (key[0] * 16 + xx).astype(np.float32) * vs
        gy = (key[1] * 16 + yy).astype(np.float32) * vs
        gz = (key[2] * 16 + zz).astype(np.float32) * vs
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.float32(1.0) / gz
            u = np.float32(160.0) * gx * inv + np.float32(160.0)
            v = np.float32(160.0) * gy * inv + np.float32(160.0)
        inb = (u >= 0) & (v >= 0) & (u <= 319) & (v <= 319)
        sdf = np.float32(z0) - gz
        upd = inb & (gz > 0) & (sdf >= -TRUNC)
        want_t = np.where(upd, np.minimum(sdf, TRUNC) / TRUNC, np.float32(0))
        assert np.array_equal(weight, np.where(upd, np.float32(n_frames), np.float32(0)))
        assert np.allclose(tsdf, want_t, rtol=0, atol=2e-7)   # running mean of equal samples


def test_depth_max_inequalities(oracle):
    """touch uses d < depth_max (strict), integrate rejects d > depth_max (so d == depth_max integrates)."""
    g = oracle.Grid(VS)
    with pytest.raises(RuntimeError, match="No block is touched"):
        g.touch(wall(1.0), K, I4, 1.0, TRUNC_MULT)          # d == depth_max -> not touched
    keys = g.touch(wall(1.0), K, I4, 1.0001, TRUNC_MULT)
    assert g.integrate(keys, wall(1.0), K, I4, 1.0, TRUNC_MULT) > 0     # d == depth_max is integrated
    g2 = oracle.Grid(VS)
    assert g2.integrate(keys, wall(1.0), K, I4, 0.9999, TRUNC_MULT) == 0


def test_running_average_order(oracle):
    """Same voxel, k frames with different depths == a scalar float32 loop in frame order."""
    g = oracle.Grid(VS)
    depths = [1.00, 1.07, 0.95, 1.02, 1.11]
    key = np.array([[0, 0, 3]], np.int32)
    for z in depths:
        g.integrate(key, wall(z), K, I4, 4.0, TRUNC_MULT)
    tsdf, weight = grid_dict(g)[(0, 0, 3)]
    zv = 3 * 16 + 2                        # voxel (0,0,2) of block (0,0,3): on the optical axis
    zc = np.float32(zv) * np.float32(VS)
    t, w = np.float32(0), np.float32(0)
    for z in depths:
        sdf = np.float32(z) - zc
        assert sdf >= -TRUNC
        s = min(sdf, TRUNC) / TRUNC
        inv = np.float32(1.0) / (w + np.float32(1))
        t = (w * t + np.float32(s)) * inv
        w = w + np.float32(1)
    assert weight[2, 0, 0] == w == 5
    assert tsdf[2, 0, 0].view(np.uint32) == np.float32(t).view(np.uint32)


def _plane_grid(oracle, c, nb=3, w=5.0):
    """Fully allocated nb^3 block cube whose tsdf is the signed distance to the plane z = c (in
    voxels, scaled), weight w everywhere."""
    g = oracle.Grid(VS)
    keys = np.array([[x, y, z] for x in range(nb) for y in range(nb) for z in range(nb)], np.int32)
    zz = np.arange(16, dtype=np.float32)[:, None, None] * np.ones((16, 16, 16), np.float32)
    tsdf = np.stack([np.clip((c - (k[2] * 16 + zz)) / 10.0, -1, 1).astype(np.float32) for k in keys])
    g.load(keys, tsdf, np.full_like(tsdf, w))
    return g, keys


def test_mc_plane(oracle):
    c = 20.4
    g, keys = _plane_grid(oracle, c)
    v, n, t, vk = g.extract_mesh(3.0)
    assert len(v) > 0
    assert np.allclose(v[:, 2], np.float32(VS) * np.float32(c), atol=2e-6)      # all vertices on the plane
    assert set(vk[:, 3].tolist()) == {2} and set(vk[:, 2].tolist()) == {20}     # z-edges of layer 20
    # cubes need all 8 corners inside the 48^3 region: 47 x 47 crossed columns, 2 triangles each
    assert len(t) == 2 * 47 * 47 and len(v) == 48 * 48
    assert np.allclose(np.abs(n[:, 2]), 1.0, atol=1e-3) and np.allclose(n[:, :2], 0.0, atol=1e-6)
    # winding/orientation consistent: every triangle normal has the same z sign
    tri = v[t]
    nz = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])[:, 2]
    assert (nz > 0).all() or (nz < 0).all()


def test_mc_plane_colors(oracle):
    """Colour branch: a colour ramp along z is interpolated at the crossing with the vertex's own ratio and
    scaled by 1/255."""
    c = 20.4
    g0, keys = _plane_grid(oracle, c)
    k, t, w, _ = g0.export()
    zz = np.arange(16, dtype=np.float32)[:, None, None] * np.ones((16, 16, 16), np.float32)
    col = np.zeros(t.shape + (3,), np.float32)
    for i, key in enumerate(k):
        col[i, ..., 0] = 10.0 * (key[2] * 16 + zz)
        col[i, ..., 2] = 255.0
    g = oracle.Grid(VS, with_color=True)
    g.load(k, t, w, col)
    v, n, tr, vk = g.extract_mesh(3.0)
    vc = g.vertex_colors(vk)
    assert vc.shape == v.shape
    assert np.allclose(vc[:, 0], (0.6 * 200.0 + 0.4 * 210.0) / 255.0, atol=1e-5)
    assert np.array_equal(vc[:, 1], np.zeros(len(vc), np.float32)) and np.allclose(vc[:, 2], 1.0, atol=1e-6)
    p, pn, pk = g.extract_points(3.0)
    assert np.allclose(g.vertex_colors(pk)[:, 0], 204.0 / 255.0, atol=1e-5)
    with pytest.raises(RuntimeError):
        g0.vertex_colors(vk)                      # colourless grid


def test_mc_weight_threshold_is_strict(oracle):
    g, _ = _plane_grid(oracle, 20.4, w=3.0)
    assert len(g.extract_mesh(3.0)[0]) == 0          # w == threshold is rejected (w <= thr)
    assert len(g.extract_mesh(2.999)[0]) == 48 * 48
    assert len(g.extract_points(3.0)[0]) == 0
    assert len(g.extract_points(2.999)[0]) == 48 * 48


def test_mc_missing_neighbour_block(oracle):
    """Dropping one block removes every cube that needs a corner from it."""
    g_full, keys = _plane_grid(oracle, 20.4)
    keep = ~((keys[:, 0] == 1) & (keys[:, 1] == 1) & (keys[:, 2] == 1))
    kf, tf, wf, _ = g_full.export()
    g = oracle.Grid(VS)
    g.load(kf[keep], tf[keep], wf[keep])
    v, n, t, vk = g.extract_mesh(3.0)
    # plane z=20.4 crosses block layer z=1; cubes whose 8 corners touch block (1,1,1) vanish:
    # cube x,y in [15, 31] (17 x 17) at z = 20
    assert len(t) == 2 * (47 * 47 - 17 * 17)


def test_mc_sphere_is_closed(oracle):
    g = oracle.Grid(VS)
    nb = 4
    keys = np.array([[x, y, z] for x in range(nb) for y in range(nb) for z in range(nb)], np.int32)
    ax = np.arange(16, dtype=np.float32)
    tsdf = []
    for k in keys:
        zz, yy, xx = np.meshgrid(k[2] * 16 + ax, k[1] * 16 + ax, k[0] * 16 + ax, indexing="ij")
        r = np.sqrt((xx - 31.3) ** 2 + (yy - 32.1) ** 2 + (zz - 30.7) ** 2)
        tsdf.append(np.clip((r - 20.0) / 10.0, -1, 1).astype(np.float32))
    tsdf = np.stack(tsdf)
    g.load(keys, tsdf, np.full_like(tsdf, 9.0))
    v, n, t, vk = g.extract_mesh(3.0)
    V, T = len(v), len(t)
    e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
    ue, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()                    # watertight
    assert V - len(ue) + T == 2                # Euler characteristic of a sphere
    rad = np.linalg.norm(v / np.float32(VS) - np.array([31.3, 32.1, 30.7]), axis=1)
    assert np.abs(rad - 20.0).max() < 0.05
    # Open3D's normal = interpolated tsdf gradient: points outward (towards +tsdf)
    dirs = v / np.float32(VS) - np.array([31.3, 32.1, 30.7])
    assert ((n * dirs).sum(1) > 0).all()


def test_raycast_plane(oracle):
    """Rays vs an analytic plane: t_hit is the z-depth at pixel centres (x+0.5, y+0.5)."""
    verts = np.array([[-5, -5, 2.0], [5, -5, 2.0], [5, 5, 2.0], [-5, 5, 2.0]], np.float32)
    tris = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    Kc = np.array([[100.0, 0, 32.0], [0, 100.0, 24.0], [0, 0, 1.0]])
    rays = oracle.create_rays_pinhole(Kc, I4, 64, 48)
    assert rays.shape == (48, 64, 6) and np.allclose(rays[..., :3], 0) and np.allclose(rays[..., 5], 1)
    assert np.isclose(rays[10, 20, 3], (20.5 - 32.0) / 100.0)
    t = oracle.cast_rays(verts, tris, rays)
    assert np.allclose(t, 2.0, atol=1e-6)
    # miss -> inf ; camera translated by (0,0,1): world->camera E has t = -R*C
    E = np.eye(4)
    E[2, 3] = -1.0                     # camera at z = +1 looking along +z
    t2 = oracle.cast_rays(verts, tris, oracle.create_rays_pinhole(Kc, E, 64, 48))
    assert np.allclose(t2, 1.0, atol=1e-6)
    E[2, 3] = -3.0                     # camera behind the plane looking away
    assert np.isinf(oracle.cast_rays(verts, tris, oracle.create_rays_pinhole(Kc, E, 64, 48))).all()


def test_confidence_static_camera(oracle):
    """Static camera + identical frames: err = 0, conf = 1, valid_count = window size (clipped)."""
    N, R = 7, 2
    d = np.full((N, 40, 40), 1.5, np.float32)
    d[:, :3, :] = 0.0            # invalid band
    Kf = np.tile(np.array([[40.0, 0, 20.0], [0, 40.0, 20.0], [0, 0, 1.0]], np.float32), (N, 1, 1))
    E = np.tile(np.eye(4, dtype=np.float32), (N, 1, 1))
    conf, count = oracle.confidence(d, Kf, E, E.copy(), R, 4.0, 0.05)
    inner = (slice(None), slice(4, 38), slice(1, 38))
    assert (conf[inner] == 1.0).all()
    want = np.array([min(N, i + R + 1) - max(0, i - R) - 1 for i in range(N)])
    assert (count[inner] == want[:, None, None]).all()
    assert (count[:, :3, :] == 0).all() and (conf[:, :3, :] == 0).all()
    # bilinear taps need u1 < W, v1 < H: the last row/column never validates
    assert (count[:, -1, :] == 0).all() and (count[:, :, -1] == 0).all()


def test_color_branch_closed_form(oracle):
    """Colour branch of Integrate: the colour of a voxel that projects to depth pixel
    (ui, vi) = (int u, int v) comes from colour pixel round(colourK . depthK^-1 . (ui, vi, 1)) -- a function of
    the depth pixel alone -- and is averaged with the voxel's weight like the tsdf."""
    z0, CW, CH, fc = 1.0, 200, 150, 110.0
    Kc = np.array([[fc, 0, CW / 2.0], [0, fc, CH / 2.0], [0, 0, 1.0]])
    uu, vv = np.meshgrid(np.arange(CW), np.arange(CH))
    img = np.stack([uu % 251, vv % 241, (uu + vv) % 239], axis=-1).astype(np.uint8)      # position-coding colours
    g = oracle.Grid(VS, with_color=True)
    d = wall(z0)
    keys = g.touch(d, K, I4, 4.0, TRUNC_MULT)
    g.integrate(keys, d, K, I4, 4.0, TRUNC_MULT, color=img, Kc=Kc)
    k2, t2, w2, c2 = g.export()
    vs = np.float32(VS)
    checked = 0
    for i, key in enumerate(k2.tolist()):
        zz, yy, xx = np.meshgrid(*[np.arange(16, dtype=np.int32)] * 3, indexing="ij")
        gx = (key[0] * 16 + xx).astype(np.float32) * vs
        gy = (key[1] * 16 + yy).astype(np.float32) * vs
        gz = (key[2] * 16 + zz).astype(np.float32) * vs
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.float32(1.0) / gz
            u = np.float32(160.0) * gx * inv + np.float32(160.0)
            v = np.float32(160.0) * gy * inv + np.float32(160.0)
        inb = (u >= 0) & (v >= 0) & (u <= 319) & (v <= 319)
        upd = inb & (gz > 0) & (np.float32(z0) - gz >= -TRUNC)
        ui, vi = np.where(upd, u, 0).astype(np.int32), np.where(upd, v, 0).astype(np.int32)   # truncation, as (int)u
        uf = np.float32(fc) * ((ui.astype(np.float32) - np.float32(160.0)) / np.float32(160.0)) + np.float32(CW / 2.0)
        vf = np.float32(fc) * ((vi.astype(np.float32) - np.float32(160.0)) / np.float32(160.0)) + np.float32(CH / 2.0)
        cin = upd & (uf >= 0) & (uf <= CW - 1) & (vf >= 0) & (vf <= CH - 1)
        cu = np.clip(np.floor(uf + np.float32(0.5)), 0, CW - 1).astype(np.int64)          # uf, vf >= 0 where used
        cv = np.clip(np.floor(vf + np.float32(0.5)), 0, CH - 1).astype(np.int64)
        want = np.where(cin[..., None], img[cv, cu].astype(np.float32), np.float32(0))
        assert np.array_equal(c2[i], want), key
        assert np.array_equal(w2[i], upd.astype(np.float32))
        checked += int(cin.sum())
    assert checked > 10000
    # second frame with another image: colour is the weight-averaged running mean (w = 1 -> (c + c2) / 2)
    img2 = (255 - img).astype(np.uint8)
    g.integrate(keys, d, K, I4, 4.0, TRUNC_MULT, color=img2, Kc=Kc)
    _, _, w3, c3 = g.export()
    both = (w3 == 2)[..., None] & (c2 + (255 - c2) != 0)
    mean = (np.float32(1) * c2 + (np.float32(255) - c2)) * (np.float32(1) / np.float32(2))
    sel = (w3 == 2) & (w2 == 1)
    cin_any = c2.sum(-1) > 0
    assert np.array_equal(c3[sel & cin_any], mean[sel & cin_any]) and both.any()


def test_edge_cases_empty_ragged_extreme(oracle):
    """Empty grid, frames that touch nothing, a single valid pixel, block keys at the ends of the +-2^20 key
    range, and K1 on degenerate frames."""
    g = oracle.Grid(VS)
    assert g.num_blocks == 0
    v, n, t, vk = g.extract_mesh(3.0)
    assert v.shape == (0, 3) and t.shape == (0, 3) and vk.shape == (0, 4)
    assert g.extract_points(3.0)[0].shape == (0, 3)
    with pytest.raises(RuntimeError, match="No block is touched"):
        g.touch(np.zeros((320, 320), np.float32), K, I4, 4.0, TRUNC_MULT)      # all pixels invalid
    one = np.zeros((320, 320), np.float32)
    one[160, 160] = 1.0                                                          # a single stride-4 sample
    keys = g.touch(one, K, I4, 4.0, TRUNC_MULT)
    assert 1 <= len(keys) <= 8 and set(keys[:, 2].tolist()) <= {2, 3}
    assert g.integrate(keys, one, K, I4, 4.0, TRUNC_MULT) > 0
    off = np.zeros((320, 320), np.float32)
    off[161, 161] = 1.0                                                          # not on the stride-4 lattice
    with pytest.raises(RuntimeError, match="No block is touched"):
        oracle.Grid(VS).touch(off, K, I4, 4.0, TRUNC_MULT)
    # keys at the limits of the packed 3 x 21-bit range survive a load / export round trip and extraction
    lim = (1 << 20) - 1
    far_keys = np.array([[lim, lim, lim], [-lim - 1, -lim - 1, -lim - 1], [lim - 1, lim, lim], [0, 0, 0]], np.int32)
    tsdf = np.linspace(-1, 1, 4 * 4096, dtype=np.float32).reshape(4, 16, 16, 16)
    h = oracle.Grid(VS)
    h.load(far_keys, tsdf, np.full_like(tsdf, 5.0))
    k2, t2, w2, _ = h.export()
    order = {tuple(k): i for i, k in enumerate(k2.tolist())}
    assert set(order) == {tuple(k) for k in far_keys.tolist()}
    for i, k in enumerate(far_keys.tolist()):
        assert np.array_equal(t2[order[tuple(k)]], tsdf[i])
    v, n, t, vk = h.extract_mesh(3.0)                  # (lim-1, lim, lim) and (lim, lim, lim) are x-neighbours
    assert np.isfinite(v).all() and (len(t) == 0 or t.max() < len(v))
    # K1 on degenerate frames (depth_data_io.py:80-85): all zeros / all ones / NaN / negative are rejected
    ok = np.full((8, 8), 0.5, np.float32)
    assert oracle.depth_valid(ok)
    for bad in (np.zeros((8, 8), np.float32), np.ones((8, 8), np.float32)):
        assert not oracle.depth_valid(bad)
    for value in (np.nan, -0.25):
        frame = ok.copy()
        frame[3, 4] = value
        assert not oracle.depth_valid(frame)
    lin = oracle.depth_to_linear(np.array([[1.0, 0.5, 0.0]], np.float32), 0.1, np.inf)
    # d = 1 (far plane at infinity) maps to 0 = "no measurement", d = 0 to the near plane (depth_utils.py:42-46)
    assert lin[0, 0] == 0 and lin[0, 1] == np.float32(0.2) and lin[0, 2] == np.float32(0.1)


def test_odometry_information_known_answers(oracle):
    """compute_odometry_information_matrix (make_fragments.py:142-150).  Identical frames under the
    identity: every valid pixel matches itself, so the translation block is n * I and the rotation block is
    sum(|p|^2 I - p p^T); a distance threshold below the residual leaves nothing; the independent NumPy restatement
    agrees on real frames under a real relative pose."""
    from oracle import numpy_mirror as mirror
    H = W = 24
    K = np.array([[20.0, 0, 12.0], [0, 20.0, 12.0], [0, 0, 1.0]])
    rng = np.random.default_rng(5)
    d = (1.0 + rng.random((H, W))).astype(np.float32)
    d[3, 4] = 0.0            # invalid: clipped
    d[5, 6] = 9.0            # beyond depth_max: clipped
    info = oracle.odometry_information(d, d, K, np.eye(4), 0.07, 1.0, 4.0)
    n = H * W - 2
    assert np.allclose(info[3:, 3:], n * np.eye(3)) and np.allclose(info, info.T)
    u, v = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    ok = (d > 0) & (d < 4.0)
    p = np.stack([(u - 12.0) * d / 20.0, (v - 12.0) * d / 20.0, d], -1)[ok].astype(np.float64)
    want = (p * p).sum() * np.eye(3) - p.T @ p
    assert np.allclose(info[:3, :3], want, rtol=1e-6)
    # a 10 cm offset along z with a 7 cm threshold: no correspondence survives; with 20 cm all do
    T = np.eye(4)
    T[2, 3] = 0.10
    flat = np.full((H, W), 2.0, np.float32)
    assert not oracle.odometry_information(flat, flat, K, T, 0.07, 1.0, 4.0).any()
    far = oracle.odometry_information(flat, flat, K, T, 0.20, 1.0, 4.0)
    assert far[3, 3] > 0.8 * H * W and far[3, 3] == far[4, 4] == far[5, 5]
    # depth_scale divides before the clip
    assert np.allclose(oracle.odometry_information(d * 1000, d * 1000, K, np.eye(4), 0.07, 1000.0, 4.0), info, rtol=1e-6)
    a = mirror.odometry_information(d, d, K, np.eye(4), 0.07, 1.0, 4.0)      # float64 products there, float32 here
    assert np.allclose(a, info, rtol=1e-5, atol=1e-3) and a[5, 5] == info[5, 5]
