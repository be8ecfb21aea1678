"""GPU parity: K1 depth conversion, K2 touch, K3 integrate (per-frame and fused sequence) against
the CPU oracle.  Bit-exact: float32 results compared with ==."""
import numpy as np
import pytest
import torch

from helpers import capture, oracle_integrate_sequence, pipeline_cameras, sort_blocks

pytestmark = pytest.mark.gpu

DEPTH_MAX = 4.0
TRUNC = 10.0


def _linear(orc, cap):
    ds = cap.dataset
    return np.stack([orc.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(len(ds))])


def test_depth_prepare_matches_oracle(cuda_device, oracle):
    from mq3d_b200.vbg import depth_prepare
    cap = capture(6)
    raw = cap.raw.copy()
    raw[1] = 0.0            # all zero  -> invalid
    raw[2] = 1.0            # all one   -> invalid
    raw[3, 5, 7] = np.nan   # NaN       -> invalid
    raw[4, 9, 9] = -0.25    # negative  -> invalid
    nears = np.array([0.1, 0.1, 0.1, 0.1, 0.1, 0.2])
    fars = np.array([np.inf, np.inf, 5.0, np.inf, 0.05, 7.5])   # finite far, far < near
    out, valid = depth_prepare(torch.from_numpy(raw).to(cuda_device), nears, fars)
    want = np.stack([oracle.depth_to_linear(raw[i], nears[i], fars[i]) for i in range(6)])
    got = out.cpu().numpy()
    nan = np.isnan(want)
    assert nan.sum() == 1 and np.array_equal(np.isnan(got), nan)     # NaN payload bits are not compared
    for i in range(6):
        assert np.array_equal(got[i][~nan[i]].view(np.uint32), want[i][~nan[i]].view(np.uint32)), f"frame {i}"
    assert valid.cpu().tolist() == [int(oracle.depth_valid(raw[i])) for i in range(6)] == [1, 0, 0, 0, 0, 1]


def test_depth_prepare_mask(cuda_device, oracle):
    from mq3d_b200.vbg import depth_prepare
    cap = capture(3)
    rng = np.random.default_rng(0)
    conf = rng.random(cap.raw.shape)
    conf[0, :4] = 0.02   # exactly at threshold: strict '<' keeps it
    count = rng.integers(0, 5, cap.raw.shape).astype(np.int32)
    has = np.array([1, 0, 1], np.uint8)   # frame 1 has no confidence map -> unfiltered
    ds = cap.dataset
    out, _ = depth_prepare(torch.from_numpy(cap.raw).to(cuda_device), ds.nears, ds.fars,
                           torch.from_numpy(conf).to(cuda_device), torch.from_numpy(count).to(cuda_device),
                           torch.from_numpy(has).to(cuda_device), 0.02, 2)
    lin = _linear(oracle, cap)
    want = lin.copy()
    for i in (0, 2):
        want[i] = oracle.depth_mask(lin[i], conf[i], count[i], 0.02, 2)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("voxel", [0.02, 0.01])
def test_touch_and_integrate_per_frame(cuda_device, oracle, voxel):
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(8)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    lin = _linear(oracle, cap)
    og = oracle.Grid(voxel)
    vbg = VoxelBlockGrid(voxel_size=voxel, block_count=64, device=cuda_device)   # forces growth
    for i in range(len(lin)):
        okeys = og.touch(lin[i], K[i], Ewc[i], DEPTH_MAX, TRUNC)
        keys = vbg.compute_unique_block_coordinates(lin[i], K[i].astype(np.float64), Ewc[i].astype(np.float64),
                                                    1.0, DEPTH_MAX, TRUNC)
        got = {tuple(k) for k in keys.cpu().numpy().tolist()}
        assert len(got) == keys.shape[0], "touch returned duplicate keys"
        assert got == {tuple(k) for k in okeys.tolist()}
        og.integrate(okeys, lin[i], K[i], Ewc[i], DEPTH_MAX, TRUNC)
        vbg.integrate(keys, lin[i], K[i].astype(np.float64), Ewc[i].astype(np.float64), 1.0, DEPTH_MAX, TRUNC)
    assert vbg.num_blocks() == og.num_blocks
    k0, t0, w0 = sort_blocks(*oracle_export(og))
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1)
    assert np.array_equal(w0, w1)
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))


def oracle_export(og):
    k, t, w, _ = og.export()
    return k, t, w


@pytest.mark.parametrize("batch", [1, 5, 64])
def test_integrate_sequence_matches_frame_loop(cuda_device, oracle, batch):
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(12)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    lin = _linear(oracle, cap)
    og = oracle.Grid(0.02)
    visits, updated = oracle_integrate_sequence(oracle, og, lin, K, Ewc, DEPTH_MAX, TRUNC)
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=50, device=cuda_device)
    st = vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC, batch_frames=batch)
    assert st.frames_integrated == 12
    assert st.block_visits == visits
    assert st.voxel_updates == updated
    assert st.num_blocks == og.num_blocks
    k0, t0, w0 = sort_blocks(*oracle_export(og))
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1)
    assert np.array_equal(w0, w1)
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))


def test_sequence_skips_invalid_frames_and_raises_on_empty(cuda_device, oracle):
    from mq3d_b200 import _lib
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(4)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    lin = _linear(oracle, cap)
    valid = np.array([1, 0, 1, 1], np.int32)
    og = oracle.Grid(0.02)
    oracle_integrate_sequence(oracle, og, lin, K, Ewc, DEPTH_MAX, TRUNC, valid=valid)
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=1000, device=cuda_device)
    st = vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC,
                                frame_valid=torch.from_numpy(valid))
    assert st.frames_integrated == 3 and st.num_blocks == og.num_blocks
    # a valid frame that touches nothing aborts like Open3D's LogError -> RuntimeError
    lin2 = lin.copy()
    lin2[2] = 0.0
    vbg2 = VoxelBlockGrid(voxel_size=0.02, block_count=1000, device=cuda_device)
    with pytest.raises(RuntimeError, match="No block is touched"):
        vbg2.integrate_sequence(torch.from_numpy(lin2).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC)
    with pytest.raises(RuntimeError, match="No block is touched") as ei:
        vbg2.compute_unique_block_coordinates(lin2[2], K[2].astype(np.float64), Ewc[2].astype(np.float64),
                                              1.0, DEPTH_MAX, TRUNC)
    assert ei.value.code == _lib.MQ3D_ERR_NO_BLOCK_TOUCHED


def test_save_load_roundtrip(cuda_device, oracle, tmp_path):
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(4)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    lin = _linear(oracle, cap)
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=1000, device=cuda_device)
    vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC)
    p = tmp_path / "colorless_vbg.npz"
    vbg.save(str(p))
    z = np.load(p)
    assert z["voxel_size"].dtype == np.float32 and z["block_resolution"].dtype == np.int64
    assert z["key"].dtype == np.int32 and z["value_000"].shape[1:] == (16, 16, 16, 1)
    assert int(z["attr_name_tsdf"][0]) == 0 and int(z["attr_name_weight"][0]) == 1
    g2 = VoxelBlockGrid.load(str(p), device=cuda_device)
    a = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    b = sort_blocks(*[x.cpu().numpy() for x in g2.export_blocks()[:3]])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("color", [False, True])
def test_integrate_shapes_are_bit_identical(cuda_device, oracle, color, monkeypatch):
    """Every work-item shape of the fused kernel (MQ3D_INTEG_VARIANT: whole / half / quarter / eighth blocks,
    different CTA sizes, with and without the warp-level cull of all-rejected frames, 4 or 8 voxels per thread)
    produces the same bits and the same statistics."""
    import mq3d_b200  # noqa: F401
    from mq3d_b200 import synth
    from mq3d_b200.vbg import VoxelBlockGrid
    from helpers import capture, pipeline_cameras, sort_blocks
    n = 12
    cap = capture(n)
    K, Ewc, Ecw = pipeline_cameras(cap.dataset)
    ds = cap.dataset
    lin = torch.from_numpy(np.stack([oracle.depth_to_linear(cap.raw[i], ds.nears[i], ds.fars[i]) for i in range(n)])).to(cuda_device)
    kw = {}
    if color:
        f, cw, ch = 110.0, 160, 120
        cols = np.stack([synth.make_color_frame(Ecw[i], width=cw, height=ch, f=f) for i in range(n)])
        kw = dict(colors=torch.from_numpy(cols).to(cuda_device),
                  color_intrinsics=np.tile(np.array([[f, 0, cw / 2.0], [0, f, ch / 2.0], [0, 0, 1.0]]), (n, 1, 1)))
    attrs = ("tsdf", "weight", "color") if color else ("tsdf", "weight")
    results = {}
    for variant in ("0", "9", "8", "12", "20", "21", "22", "23", "24", "25", "26", "30", "31", "32", "33", "34", "40"):
        monkeypatch.setenv("MQ3D_INTEG_VARIANT", variant)
        g = VoxelBlockGrid(attr_names=attrs, voxel_size=0.02, block_count=3000, device=cuda_device)
        st = g.integrate_sequence(lin, K, Ewc, 4.0, 10.0, batch_frames=5, **kw)
        blocks = sort_blocks(*[x.cpu().numpy() if x is not None else None for x in g.export_blocks()])
        results[variant] = (st.block_visits, st.voxel_updates, st.num_blocks, blocks)
    ref = results["9"]
    for variant, got in results.items():
        assert got[:3] == ref[:3], variant
        for a, b in zip(got[3], ref[3]):
            if a is not None:
                assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                                      b.view(np.uint32) if b.dtype == np.float32 else b), variant


@pytest.mark.parametrize("pixel", [(8, 12), (9, 13)])      # on / off the stride-4 touch lattice
def test_tiny_depth_takes_the_guarded_division(cuda_device, oracle, pixel):
    """A depth in (0, 2^-75) makes |depth - z| < 2^-100 possible -- outside the validated range of the unguarded
    fast division -- so the batch that holds it is integrated by the guarded instantiation (k_touch scans every
    pixel of every valid frame).  Bits equal the oracle's either way; the other batches stay on the fast kernel."""
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(12)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    lin = _linear(oracle, cap)
    lin[7, pixel[0], pixel[1]] = 1e-30
    og = oracle.Grid(0.02)
    oracle_integrate_sequence(oracle, og, lin, K, Ewc, DEPTH_MAX, TRUNC)
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=1000, device=cuda_device)
    st = vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC, batch_frames=4)
    assert st.batches == 3 and st.slow_div_batches == 1
    k0, t0, w0 = sort_blocks(*oracle_export(og))
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1)
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))


@pytest.mark.parametrize("variant", ["0", "40"])
def test_rejected_depth_values_match_the_oracle(cuda_device, oracle, variant, monkeypatch):
    """Depth pixels the integrator rejects by value -- 0, -0.0, negative, above depth_max, +inf, -inf -- and the ones it
    keeps -- exactly depth_max, NaN (Open3D's compares let NaN through: tsdf and weight of the voxels that project
    there follow the oracle).  The default (packed) body sees them through the sanitised
    batch copy written by k_touch (-inf => `sdf < -trunc`); variant 40 is the scalar body with the four explicit tests."""
    from mq3d_b200.vbg import VoxelBlockGrid
    monkeypatch.setenv("MQ3D_INTEG_VARIANT", variant)
    cap = capture(8)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    lin = _linear(oracle, cap)
    rng = np.random.default_rng(11)
    specials = np.array([0.0, -0.0, -1.5, DEPTH_MAX, np.nextafter(np.float32(DEPTH_MAX), np.float32(10.0)), 17.0, np.inf,
                         -np.inf, np.nan], np.float32)
    for f in range(len(lin)):
        ys, xs = rng.integers(0, lin.shape[1], 600), rng.integers(0, lin.shape[2], 600)
        lin[f, ys, xs] = specials[rng.integers(0, len(specials), 600)]
    og = oracle.Grid(0.02)
    oracle_integrate_sequence(oracle, og, lin, K, Ewc, DEPTH_MAX, TRUNC)
    vbg = VoxelBlockGrid(voxel_size=0.02, block_count=2000, device=cuda_device)
    vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC, batch_frames=3)
    k0, t0, w0 = sort_blocks(*oracle_export(og))
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1)
    nan0, nan1 = np.isnan(t0), np.isnan(t1)
    assert np.array_equal(nan0, nan1)          # (Open3D's min(sdf, trunc) turns a NaN sdf into +trunc: usually no NaN survives)
    assert np.array_equal(t0[~nan0].view(np.uint32), t1[~nan1].view(np.uint32))


def test_sequence_resumes_after_mid_sequence_growth(cuda_device, oracle):
    """All batches of a call are enqueued without host synchronisation; a batch whose touch overflows the pool or
    the hash table stops the device-side pipeline, the host grows the grid and resumes from that batch.  Start so
    small that both the pool and the table overflow several times; a failed call (frame that touches nothing)
    leaves no stale frame bits behind for the next call on the same grid."""
    from mq3d_b200.vbg import VoxelBlockGrid
    cap = capture(12)
    K, Ewc, _ = pipeline_cameras(cap.dataset)
    lin = _linear(oracle, cap)
    og = oracle.Grid(0.01)
    visits, updated = oracle_integrate_sequence(oracle, og, lin, K, Ewc, DEPTH_MAX, TRUNC)
    vbg = VoxelBlockGrid(voxel_size=0.01, block_count=8, device=cuda_device)
    cap0 = vbg.capacity()
    st = vbg.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC, batch_frames=2)
    assert vbg.capacity() >= og.num_blocks > cap0
    assert (st.frames_integrated, st.block_visits, st.voxel_updates, st.num_blocks) == (12, visits, updated, og.num_blocks)
    k0, t0, w0 = sort_blocks(*oracle_export(og))
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in vbg.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1)
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    # error exit, then a clean call with another batch size on the same grid
    bad = lin[:5].copy()
    bad[3] = 0.0
    g2 = VoxelBlockGrid(voxel_size=0.02, block_count=1000, device=cuda_device)
    with pytest.raises(RuntimeError, match="No block is touched"):
        g2.integrate_sequence(torch.from_numpy(bad).to(cuda_device), K[:5], Ewc[:5], DEPTH_MAX, TRUNC, batch_frames=3)
    g2.reset()
    og2 = oracle.Grid(0.02)
    oracle_integrate_sequence(oracle, og2, lin, K, Ewc, DEPTH_MAX, TRUNC)
    g2.integrate_sequence(torch.from_numpy(lin).to(cuda_device), K, Ewc, DEPTH_MAX, TRUNC, batch_frames=40)
    k0, t0, w0 = sort_blocks(*oracle_export(og2))
    k1, t1, w1 = sort_blocks(*[x.cpu().numpy() for x in g2.export_blocks()[:3]])
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
