"""VoxelBlockGrid: the host-side mirror of ``o3d.t.geometry.VoxelBlockGrid`` for the hot path.

Same call shapes as the reference's call sites (processing/reconstruction/utils/o3d_utils.py:171-229,
processing/reconstruction/reconstruct_scene.py:87-108,186-189, dataio/reconstruction_data_io.py:42-55),
backed by the C ABI in include/mq3d.h.  Buffers are torch CUDA tensors (device memory + streams
only); all arithmetic happens in libmq3d.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

RES = 16
PEER_DESC_BYTES = 512      # MQ3D_PEER_DESC_BYTES


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_np(x, dtype, shape=None):
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        x = x.cpu().numpy() if hasattr(x, "cpu") else x.numpy()
    a = np.ascontiguousarray(np.asarray(x), dtype=dtype)
    if shape is not None and a.shape != shape:
        raise RuntimeError(f"expected shape {shape}, got {a.shape}")
    return a


def _device_index(device) -> int:
    if isinstance(device, int):
        return device
    if isinstance(device, torch.device):
        if device.type != "cuda":
            raise RuntimeError(f"device {device} is not a CUDA device; this build has no CPU fallback")
        return device.index if device.index is not None else torch.cuda.current_device()
    s = str(device).upper()
    if s.startswith("CUDA"):
        return int(s.split(":")[1]) if ":" in s else 0
    raise RuntimeError(f"device '{device}' is not a CUDA device; this build has no CPU fallback")


def as_depth_tensor(depth, device: torch.device) -> torch.Tensor:
    """float32 [H,W] (or [H,W,1]) depth on the grid's device."""
    if hasattr(depth, "as_tensor"):
        depth = depth.as_tensor()
    if hasattr(depth, "torch"):
        depth = depth.torch
    if isinstance(depth, np.ndarray):
        depth = torch.from_numpy(np.ascontiguousarray(depth))
    if not isinstance(depth, torch.Tensor):
        raise RuntimeError("depth must be an Image, a Tensor, a numpy array or a torch tensor")
    if depth.dim() == 3 and depth.shape[-1] == 1:
        depth = depth[..., 0]
    if depth.dim() != 2:
        raise RuntimeError(f"depth must be [H,W] or [H,W,1], got {tuple(depth.shape)}")
    if depth.dtype != torch.float32:
        raise RuntimeError(f"depth must be float32, got {depth.dtype}")
    return depth.to(device, non_blocking=True).contiguous()


def ghost_packed_len(m: int, has_color: bool) -> int:
    """float32 words of one packed ghost payload: keys (padded to 16 B) | tsdf | weight | colour."""
    return ((3 * m + 3) // 4) * 4 + m * RES ** 3 * (5 if has_color else 2)


def ghost_views(buf: torch.Tensor, m: int, has_color: bool):
    """(keys i32 [m,3], tsdf [m,4096], weight [m,4096], color [m,4096,3] | None) views of a packed payload;
    `keys.untyped_storage()` is the single contiguous buffer that travels between ranks."""
    v = RES ** 3
    k_len = ((3 * m + 3) // 4) * 4
    keys = buf[: 3 * m].view(torch.int32).view(m, 3)
    tsdf = buf[k_len: k_len + m * v].view(m, v)
    weight = buf[k_len + m * v: k_len + 2 * m * v].view(m, v)
    color = buf[k_len + 2 * m * v: k_len + 5 * m * v].view(m, v, 3) if has_color else None
    return keys, tsdf, weight, color


@dataclass
class SequenceStats:
    frames_integrated: int
    block_visits: int
    blocks_loaded: int
    num_blocks: int
    batches: int
    voxel_updates: int
    touch_ms: float = 0.0
    integrate_ms: float = 0.0
    slow_div_batches: int = 0

    @property
    def voxel_visits(self) -> int:
        return self.block_visits * RES ** 3


class VoxelBlockGrid:
    """GPU spatial hash + block pool with Open3D-0.19 VoxelBlockGrid semantics."""

    def __init__(self, attr_names: Sequence[str] = ("tsdf", "weight"), attr_dtypes=None, attr_channels=None,
                 voxel_size: float = 0.0058, block_resolution: int = 16, block_count: int = 10000,
                 device="CUDA:0"):
        names = tuple(attr_names)
        if names[:2] != ("tsdf", "weight") or any(n not in ("tsdf", "weight", "color") for n in names):
            raise RuntimeError(f"unsupported attributes {names}: expected ('tsdf','weight'[, 'color'])")
        if attr_channels is not None:
            ch = tuple(int(np.prod(c)) for c in attr_channels)
            want = (1, 1, 3)[: len(names)]
            if ch != want:
                raise RuntimeError(f"attr_channels {ch} do not match {want}")
        self.attr_names = names
        self.voxel_size = float(voxel_size)
        self.block_resolution = int(block_resolution)
        self.device_index = _device_index(device)
        self.device = torch.device("cuda", self.device_index)
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available; the B200 path has no CPU fallback")
        mask = _lib.ATTR_TSDF_WEIGHT | (_lib.ATTR_COLOR if "color" in names else 0)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_grid_create(C.c_float(voxel_size), self.block_resolution, int(block_count),
                                                   mask, self.device_index, C.byref(h)))
        self._h = h
        self._keys_scratch: Optional[torch.Tensor] = None
        self.partition = None      # (rank, world, tile_blocks, integrate_ghosts) once set_partition was called

    # -- life cycle ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().mq3d_grid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def has_color(self) -> bool:
        return "color" in self.attr_names

    def num_blocks(self) -> int:
        n = C.c_int64()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_grid_num_blocks(self._h, C.byref(n), _stream()))
        return int(n.value)

    def capacity(self) -> int:
        cap = C.c_int64()
        _lib.check(_lib.lib().mq3d_grid_info(self._h, None, None, C.byref(cap), None, None))
        return int(cap.value)

    def reserve(self, block_count: int):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_grid_reserve(self._h, int(block_count), _stream()))

    def reset(self):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_grid_reset(self._h, _stream()))

    def set_partition(self, rank: int, world: int, tile_blocks: int = 8, integrate_ghosts: bool = True):
        """Hash partition for multi-GPU runs.  integrate_ghosts=False integrates owned blocks only; call
        dist.exchange_ghosts(vbg) before extracting."""
        _lib.check(_lib.lib().mq3d_grid_set_partition(self._h, int(rank), int(world), int(tile_blocks)))
        _lib.check(_lib.lib().mq3d_grid_set_ghost_mode(self._h, int(bool(integrate_ghosts))))
        self.partition = (int(rank), int(world), int(tile_blocks), bool(integrate_ghosts))

    def ghost_select(self, dest_rank: int):
        """Owned blocks that `dest_rank` keeps as ghosts: (keys i32 [n,3], tsdf f32 [n,4096], weight f32
        [n,4096], color f32 [n,4096,3] | None) device tensors."""
        buf, m, has_color = self.ghost_select_packed(dest_rank)
        return ghost_views(buf, m, has_color)

    def ghost_counts(self) -> list:
        """Number of owned blocks every rank keeps as ghosts ([world] ints; one kernel, one sync)."""
        world = self.partition[1] if self.partition else 1
        out = (C.c_int64 * world)()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_grid_ghost_counts(self._h, out, _stream()))
        return [int(x) for x in out]

    def peer_descriptor(self) -> np.ndarray:
        """uint8 [512] description of this rank's pool (CUDA IPC handles + block count) for ghost_pull."""
        desc = np.zeros(PEER_DESC_BYTES, np.uint8)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_grid_peer_descriptor(self._h, desc.ctypes.data_as(C.c_void_p), _stream()))
        return desc

    def ghost_pull(self, descs: np.ndarray, want_count: bool = True) -> int:
        """Fetch the ghost shell straight from the owners' pools (peer memory over NVLink).  descs: uint8
        [world, 512], row r = rank r's peer_descriptor().  Everything is enqueued on the current stream without a
        host round trip; want_count=True additionally reads back the number of blocks fetched (synchronises),
        otherwise -1 is returned.  The caller fences (stream-ordered collective) before any grid changes again."""
        descs = np.ascontiguousarray(descs, np.uint8)
        world = self.partition[1] if self.partition else 1
        if descs.shape != (world, PEER_DESC_BYTES):
            raise RuntimeError(f"descs must be uint8 [{world}, {PEER_DESC_BYTES}]")
        n = C.c_int64(-1)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_grid_ghost_pull(self._h, descs.ctypes.data_as(C.c_void_p),
                                                       C.byref(n) if want_count else None, _stream()))
        return int(n.value)

    def ghost_select_packed(self, dest_rank: int, count: Optional[int] = None):
        """Same selection as one contiguous float32 payload (see ghost_packed_len / ghost_views):
        (buffer, block count, has_color).  With `count` (from ghost_counts) the call is asynchronous."""
        n = C.c_int64()
        with torch.cuda.device(self.device):
            if count is None:
                _lib.check(_lib.lib().mq3d_grid_ghost_select(self._h, int(dest_rank), C.byref(n), None, None, None,
                                                             None, _stream()))
                count = int(n.value)
            m = n.value = int(count)
            buf = torch.empty(ghost_packed_len(m, self.has_color), dtype=torch.float32, device=self.device)
            if m:
                keys, tsdf, weight, color = ghost_views(buf, m, self.has_color)
                buf[: ((3 * m + 3) // 4) * 4].zero_()      # key padding words travel too: keep them defined
                _lib.check(_lib.lib().mq3d_grid_ghost_select(self._h, int(dest_rank), C.byref(n), _lib.dptr(keys),
                                                             _lib.dptr(tsdf), _lib.dptr(weight), _lib.dptr(color),
                                                             _stream()))
                assert int(n.value) == m
        return buf, m, self.has_color

    # -- K2 -----------------------------------------------------------------------------------------
    def compute_unique_block_coordinates(self, depth, intrinsic, extrinsic, depth_scale: float = 1000.0,
                                         depth_max: float = 3.0, trunc_voxel_multiplier: float = 8.0) -> torch.Tensor:
        """int32 [N,3] block keys touched by the frame (order unspecified); does not allocate."""
        d = as_depth_tensor(depth, self.device)
        H, W = d.shape
        K = _as_np(intrinsic, np.float64, (3, 3))
        E = _as_np(extrinsic, np.float64, (4, 4))
        cap = (W // 4) * (H // 4) * 4
        if self._keys_scratch is None or self._keys_scratch.shape[0] < cap:
            self._keys_scratch = torch.empty((cap, 3), dtype=torch.int32, device=self.device)
        n = C.c_int64()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_touch(self._h, _lib.dptr(d), W, H, _lib.darr(K), _lib.darr(E),
                                             C.c_float(depth_scale), C.c_float(depth_max),
                                             C.c_float(trunc_voxel_multiplier), _lib.dptr(self._keys_scratch),
                                             C.byref(n), _stream()))
        return self._keys_scratch[: n.value].clone()

    # -- K3 -----------------------------------------------------------------------------------------
    def integrate(self, block_coords, depth, *args, **kwargs):
        """Open3D overloads:
        integrate(block_coords, depth, intrinsic, extrinsic, depth_scale, depth_max, trunc_voxel_multiplier)
        integrate(block_coords, depth, color, depth_intrinsic, color_intrinsic, extrinsic, depth_scale, ...)
        """
        names_d = ["intrinsic", "extrinsic", "depth_scale", "depth_max", "trunc_voxel_multiplier"]
        names_c = ["color", "depth_intrinsic", "color_intrinsic", "extrinsic", "depth_scale", "depth_max",
                   "trunc_voxel_multiplier"]
        with_color = "color" in kwargs or len(args) >= 4 and not np.isscalar(args[2])
        names = names_c if with_color else names_d
        p = dict(zip(names, args))
        p.update(kwargs)
        depth_scale = float(p.get("depth_scale", 1000.0))
        depth_max = float(p.get("depth_max", 3.0))
        trunc = float(p.get("trunc_voxel_multiplier", 8.0))
        Kd = _as_np(p["depth_intrinsic"] if with_color else p["intrinsic"], np.float64, (3, 3))
        E = _as_np(p["extrinsic"], np.float64, (4, 4))
        d = as_depth_tensor(depth, self.device)
        H, W = d.shape
        keys = block_coords.torch if hasattr(block_coords, "torch") else block_coords
        if isinstance(keys, np.ndarray):
            keys = torch.from_numpy(np.ascontiguousarray(keys, dtype=np.int32))
        keys = keys.to(self.device).to(torch.int32).contiguous()
        if keys.dim() != 2 or keys.shape[1] != 3:
            raise RuntimeError(f"block_coords must be [N,3], got {tuple(keys.shape)}")
        col, Kc, CW, CH = None, Kd, 0, 0
        if with_color and self.has_color:
            col = p["color"]
            col = col.as_tensor() if hasattr(col, "as_tensor") else col
            col = col.torch if hasattr(col, "torch") else col
            if isinstance(col, np.ndarray):
                col = torch.from_numpy(np.ascontiguousarray(col))
            if col.dtype != torch.uint8 or col.dim() != 3 or col.shape[2] != 3:
                raise RuntimeError("color must be uint8 [H,W,3]")
            col = col.to(self.device).contiguous()
            CH, CW = int(col.shape[0]), int(col.shape[1])
            Kc = _as_np(p["color_intrinsic"], np.float64, (3, 3))
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_integrate(self._h, _lib.dptr(keys), int(keys.shape[0]), _lib.dptr(d), W, H,
                                                 _lib.dptr(col), CW, CH, _lib.darr(Kd), _lib.darr(Kc), _lib.darr(E),
                                                 C.c_float(depth_scale), C.c_float(depth_max), C.c_float(trunc),
                                                 _stream()))

    def integrate_sequence(self, depths: torch.Tensor, intrinsics, extrinsics, depth_max: float,
                           trunc_voxel_multiplier: float, depth_scale: float = 1.0,
                           frame_valid: Optional[torch.Tensor] = None, colors: Optional[torch.Tensor] = None,
                           color_intrinsics=None, batch_frames: int = 64,
                           colors_rgbx: Optional[torch.Tensor] = None, batch_events=None) -> SequenceStats:
        """Fused frame loop of ``integrate()`` (o3d_utils.py:231-236) over [F,H,W] linear depth.

        batch_events: one torch.cuda.Event (or None) per batch of `batch_frames` frames; the call waits for event i
        on the device before it enqueues batch i, so the frames of later batches may still be in flight on another
        stream (upload + depth_prepare) when the call is made (mq3d_grid_set_batch_gates).

        colors: uint8 [F,CH,CW,3], a CUDA tensor or a *pinned* CPU tensor (read in place over PCIe: only the
        W x H sampled pixels per frame travel); other CPU tensors are copied to the device first.
        colors_rgbx: int32 [F,H,W] from color_resample() instead of colors (+ color_intrinsics)."""
        if depths.dim() != 3 or depths.dtype != torch.float32 or not depths.is_cuda:
            raise RuntimeError("depths must be a float32 CUDA tensor [F,H,W]")
        depths = depths.contiguous()
        F, H, W = depths.shape
        Kd = _as_np(intrinsics, np.float64, (F, 3, 3))
        E = _as_np(extrinsics, np.float64, (F, 4, 4))
        fv = None
        if frame_valid is not None:
            fv = frame_valid.to(self.device).to(torch.int32).contiguous()
        col, Kc, CW, CH = None, Kd, 0, 0
        st = _lib.SeqStats()
        if batch_events is not None:
            handles = (C.c_void_p * len(batch_events))(*[None if e is None else e.cuda_event for e in batch_events])
            _lib.check(_lib.lib().mq3d_grid_set_batch_gates(self._h, handles, len(batch_events)))
        if colors_rgbx is not None and self.has_color:
            if (colors_rgbx.dtype != torch.int32 or tuple(colors_rgbx.shape) != (F, H, W) or not colors_rgbx.is_cuda
                    or not colors_rgbx.is_contiguous()):
                raise RuntimeError("colors_rgbx must be a contiguous int32 CUDA tensor [F,H,W] (see color_resample)")
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().mq3d_integrate_sequence_rgbx(
                    self._h, _lib.dptr(depths), _lib.dptr(fv), F, W, H, _lib.dptr(colors_rgbx), _lib.darr(Kd),
                    _lib.darr(E), C.c_float(depth_scale), C.c_float(depth_max), C.c_float(trunc_voxel_multiplier),
                    int(batch_frames), C.byref(st), _stream()))
            return SequenceStats(st.frames_integrated, st.block_visits, st.blocks_loaded, st.num_blocks, st.batches,
                                 st.voxel_updates, st.touch_ms, st.integrate_ms, st.slow_div_batches)
        if colors is not None and self.has_color:
            if colors.dtype != torch.uint8 or colors.dim() != 4 or colors.shape[0] != F or colors.shape[3] != 3:
                raise RuntimeError("colors must be uint8 [F,H,W,3]")
            zero_copy = colors.device.type == "cpu" and colors.is_pinned() and colors.is_contiguous()
            col = colors if zero_copy else colors.to(self.device).contiguous()
            CH, CW = int(col.shape[1]), int(col.shape[2])
            Kc = _as_np(color_intrinsics, np.float64, (F, 3, 3))
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_integrate_sequence(
                self._h, _lib.dptr(depths), _lib.dptr(fv), F, W, H, _lib.dptr(col), CW, CH, _lib.darr(Kd),
                _lib.darr(Kc), _lib.darr(E), C.c_float(depth_scale), C.c_float(depth_max),
                C.c_float(trunc_voxel_multiplier), int(batch_frames), C.byref(st), _stream()))
        return SequenceStats(st.frames_integrated, st.block_visits, st.blocks_loaded, st.num_blocks, st.batches,
                             st.voxel_updates, st.touch_ms, st.integrate_ms, st.slow_div_batches)

    # -- pool views / persistence -------------------------------------------------------------------
    def _pool_ptrs(self):
        k, t, w, c = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(_lib.lib().mq3d_grid_pool(self._h, C.byref(k), C.byref(t), C.byref(w), C.byref(c)))
        return k.value, t.value, w.value, c.value

    def export_blocks(self):
        """(keys int32 [N,3], tsdf f32 [N,16,16,16], weight f32 [N,16,16,16], color f32 [N,16,16,16,3]|None)
        as torch CUDA tensors (copies of the active part of the pool)."""
        n = self.num_blocks()
        keys = torch.empty((n, 3), dtype=torch.int32, device=self.device)
        tsdf = torch.empty((n, RES, RES, RES), dtype=torch.float32, device=self.device)
        weight = torch.empty_like(tsdf)
        color = torch.empty((n, RES, RES, RES, 3), dtype=torch.float32, device=self.device) if self.has_color else None
        if n:
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().mq3d_grid_export(self._h, _lib.dptr(keys), _lib.dptr(tsdf), _lib.dptr(weight),
                                                       _lib.dptr(color), _stream()))
        return keys, tsdf, weight, color

    def import_blocks(self, keys, tsdf, weight, color=None):
        def dev(x, dt):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(np.ascontiguousarray(x))
            return x.to(self.device).to(dt).contiguous()
        keys, tsdf, weight = dev(keys, torch.int32), dev(tsdf, torch.float32), dev(weight, torch.float32)
        color = dev(color, torch.float32) if self.has_color else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_grid_import(self._h, _lib.dptr(keys), _lib.dptr(tsdf), _lib.dptr(weight),
                                                   _lib.dptr(color), int(keys.shape[0]), _stream()))
            torch.cuda.current_stream().synchronize()

    def save(self, path: str):
        """Open3D VoxelBlockGrid.save npz layout (reconstruction_data_io.py:51-55)."""
        keys, tsdf, weight, color = self.export_blocks()
        out = {
            "voxel_size": np.array([self.voxel_size], dtype=np.float32),
            "block_resolution": np.array([self.block_resolution], dtype=np.int64),
            f"CUDA:{self.device_index}": np.zeros((), dtype=np.uint8),
            "key": keys.cpu().numpy(),
        }
        vals = [tsdf.cpu().numpy()[..., None], weight.cpu().numpy()[..., None]]
        if color is not None:
            vals.append(color.cpu().numpy())
        for i, (name, v) in enumerate(zip(self.attr_names, vals)):
            out[f"attr_name_{name}"] = np.array([i], dtype=np.int32)
            out[f"value_{i:03d}"] = v
        path = str(path)
        np.savez(path if path.endswith(".npz") else path + ".npz", **out)

    @classmethod
    def load(cls, path: str, device="CUDA:0") -> "VoxelBlockGrid":
        z = np.load(str(path))
        names = sorted((k[len("attr_name_"):] for k in z.files if k.startswith("attr_name_")),
                       key=lambda n: int(z[f"attr_name_{n}"][0]))
        keys = z["key"]
        g = cls(attr_names=tuple(names), voxel_size=float(z["voxel_size"][0]),
                block_resolution=int(z["block_resolution"][0]), block_count=max(len(keys), 16), device=device)
        vals = {n: z[f"value_{int(z[f'attr_name_{n}'][0]):03d}"] for n in names}
        g.import_blocks(keys, vals["tsdf"][..., 0], vals["weight"][..., 0], vals.get("color"))
        return g

    # -- K5 -----------------------------------------------------------------------------------------
    def extract_triangle_mesh_arrays(self, weight_threshold: float = 3.0, with_keys: bool = False,
                                     with_colors: bool = False):
        """(vertices f32 [V,3], normals f32 [V,3], triangles i32 [T,3][, vertex_keys i32 [V,4]][, colors f32
        [V,3] in 0..1]) on device; colours need the colour attribute.

        The first extraction of a grid uses the count -> allocate -> fill pair; later ones allocate from the sizes
        seen before (with head room) and make ONE library call (mq3d_extract_mesh: classification, scan and emission
        enqueued back to back, one synchronisation), falling back to an exact-size fill if the mesh outgrew the
        guess.  The returned tensors are then leading-row views of slightly larger buffers."""
        V, T = C.c_int64(), C.c_int64()
        dev = self.device
        with torch.cuda.device(dev):
            def alloc(nv, nt):
                return (torch.empty((nv, 3), dtype=torch.float32, device=dev),
                        torch.empty((nv, 3), dtype=torch.float32, device=dev),
                        torch.empty((nt, 3), dtype=torch.int32, device=dev),
                        torch.empty((nv, 4), dtype=torch.int32, device=dev) if with_keys else None,
                        torch.empty((nv, 3), dtype=torch.float32, device=dev) if with_colors else None)

            def fill(bufs):
                _lib.check(_lib.lib().mq3d_extract_mesh_fill(self._h, _lib.dptr(bufs[0]), _lib.dptr(bufs[1]),
                                                             _lib.dptr(bufs[2]), _lib.dptr(bufs[3]), _stream()))
                if with_colors:
                    _lib.check(_lib.lib().mq3d_extract_mesh_colors(self._h, _lib.dptr(bufs[4]), _stream()))

            cap = getattr(self, "_mesh_cap", None)
            if cap is None:
                _lib.check(_lib.lib().mq3d_extract_mesh_count(self._h, C.c_float(weight_threshold), C.byref(V),
                                                              C.byref(T), _stream()))
                bufs = alloc(V.value, T.value)
                fill(bufs)
            else:
                bufs = alloc(*cap)
                _lib.check(_lib.lib().mq3d_extract_mesh(self._h, C.c_float(weight_threshold), _lib.dptr(bufs[0]),
                                                        _lib.dptr(bufs[1]), _lib.dptr(bufs[2]), _lib.dptr(bufs[3]),
                                                        _lib.dptr(bufs[4]), cap[0], cap[1], C.byref(V), C.byref(T),
                                                        _stream()))
                if V.value > cap[0] or T.value > cap[1]:
                    bufs = alloc(V.value, T.value)
                    fill(bufs)
            self._mesh_cap = (V.value + V.value // 8 + 1024, T.value + T.value // 8 + 1024)
            verts, normals, tris, vkeys, cols = (None if x is None else x[: (T.value if i == 2 else V.value)]
                                                 for i, x in enumerate(bufs))
            out = (verts, normals, tris) + ((vkeys,) if with_keys else ()) + ((cols,) if with_colors else ())
        return out

    def extract_point_cloud_arrays(self, weight_threshold: float = 3.0, with_keys: bool = False,
                                   with_colors: bool = False):
        """(points f32 [P,3], normals f32 [P,3][, point_keys i32 [P,4]][, colors f32 [P,3]]) on device."""
        P = C.c_int64()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mq3d_extract_points_count(self._h, C.c_float(weight_threshold), C.byref(P), _stream()))
            pts = torch.empty((P.value, 3), dtype=torch.float32, device=self.device)
            nrm = torch.empty((P.value, 3), dtype=torch.float32, device=self.device)
            pk = torch.empty((P.value, 4), dtype=torch.int32, device=self.device) if with_keys else None
            _lib.check(_lib.lib().mq3d_extract_points_fill(self._h, _lib.dptr(pts), _lib.dptr(nrm), _lib.dptr(pk),
                                                           _stream()))
            out = (pts, nrm) + ((pk,) if with_keys else ())
            if with_colors:
                cols = torch.empty((P.value, 3), dtype=torch.float32, device=self.device)
                _lib.check(_lib.lib().mq3d_extract_points_colors(self._h, _lib.dptr(cols), _stream()))
                out += (cols,)
        return out

    def extract_triangle_mesh(self, weight_threshold: float = 3.0, estimated_vertex_number: int = -1):
        """Open3D-shaped result; vertex colours are present when the grid has the colour attribute."""
        from .geometry import TriangleMesh
        out = self.extract_triangle_mesh_arrays(weight_threshold, with_colors=self.has_color)
        return TriangleMesh(out[0], out[2], out[1], out[3] if self.has_color else None)

    def extract_point_cloud(self, weight_threshold: float = 3.0, estimated_point_number: int = -1):
        from .geometry import PointCloud
        out = self.extract_point_cloud_arrays(weight_threshold, with_colors=self.has_color)
        return PointCloud(out[0], out[1], out[2] if self.has_color else None)


def color_resample(colors: torch.Tensor, depth_intrinsics, color_intrinsics, width: int, height: int,
                   device=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Colour frames resampled onto the depth pixel grid (mq3d_color_resample): int32 [F,height,width] CUDA
    tensor, R | G<<8 | B<<16, byte 3 = 0xFF where the depth pixel projects outside the colour image.
    colors: uint8 [F,CH,CW,3] on the GPU or in *pinned* host memory (read in place over PCIe: only the sampled
    pixels travel).  Asynchronous on the current stream."""
    if colors.dtype != torch.uint8 or colors.dim() != 4 or colors.shape[3] != 3 or not colors.is_contiguous():
        raise RuntimeError("colors must be a contiguous uint8 tensor [F,CH,CW,3]")
    if colors.is_cuda:
        device = colors.device
    elif not colors.is_pinned():
        raise RuntimeError("host colour frames must be pinned (torch.Tensor.pin_memory) to be read by the GPU")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    F, CH, CW = int(colors.shape[0]), int(colors.shape[1]), int(colors.shape[2])
    Kd = _as_np(depth_intrinsics, np.float64, (F, 3, 3))
    Kc = _as_np(color_intrinsics, np.float64, (F, 3, 3))
    if out is None:
        out = torch.empty((F, int(height), int(width)), dtype=torch.int32, device=device)
    elif (out.dtype != torch.int32 or tuple(out.shape) != (F, int(height), int(width)) or not out.is_cuda
          or not out.is_contiguous()):
        raise RuntimeError("out must be a contiguous int32 CUDA tensor [F,height,width]")
    if F == 0:
        return out
    with torch.cuda.device(device):
        _lib.check(_lib.lib().mq3d_color_resample(C.c_void_p(colors.data_ptr()), F, CW, CH, int(width), int(height),
                                                  _lib.darr(Kd), _lib.darr(Kc), _lib.dptr(out), int(device.index),
                                                  _stream()))
    return out


def depth_prepare(raw: torch.Tensor, nears, fars, conf: Optional[torch.Tensor] = None,
                  count: Optional[torch.Tensor] = None, has_conf: Optional[torch.Tensor] = None,
                  confidence_threshold: float = 0.0, valid_count_threshold: int = 0,
                  out: Optional[torch.Tensor] = None, valid_out: Optional[torch.Tensor] = None):
    """K1 over [F,H,W] raw NDC depth: returns (linear float32 [F,H,W], frame_valid int32 [F]).  out / valid_out:
    contiguous destination tensors of those shapes (e.g. slices of a whole-sequence buffer)."""
    if raw.dim() != 3 or raw.dtype != torch.float32 or not raw.is_cuda:
        raise RuntimeError("raw must be a float32 CUDA tensor [F,H,W]")
    raw = raw.contiguous()
    F, H, W = raw.shape
    near = _as_np(nears, np.float64, (F,))
    far = _as_np(fars, np.float64, (F,))
    if out is None:
        out = torch.empty_like(raw)
    elif (out.dtype != torch.float32 or tuple(out.shape) != (F, H, W) or out.device != raw.device
          or not out.is_contiguous()):
        raise RuntimeError("out must be a contiguous float32 tensor [F,H,W] on the device of raw")
    valid = valid_out if valid_out is not None else torch.empty((F,), dtype=torch.int32, device=raw.device)
    if (valid.dtype != torch.int32 or tuple(valid.shape) != (F,) or valid.device != raw.device
            or not valid.is_contiguous()):
        raise RuntimeError("valid_out must be a contiguous int32 tensor [F] on the device of raw")
    if conf is not None:
        conf = conf.to(raw.device).to(torch.float64).contiguous()
        count = count.to(raw.device).to(torch.int32).contiguous()
        if has_conf is not None:
            has_conf = has_conf.to(raw.device).to(torch.uint8).contiguous()
    with torch.cuda.device(raw.device):
        _lib.check(_lib.lib().mq3d_depth_prepare(_lib.dptr(raw), F, W, H, _lib.darr(near), _lib.darr(far),
                                                 _lib.dptr(conf), _lib.dptr(count), _lib.dptr(has_conf),
                                                 C.c_double(confidence_threshold), int(valid_count_threshold),
                                                 _lib.dptr(out), _lib.dptr(valid), _stream()))
    return out, valid


def estimate_confidence(depths: torch.Tensor, K, Ecw, Ecw_inv, target_frame_range: int, depth_max: float,
                        error_threshold: float, frame_valid: Optional[torch.Tensor] = None):
    """K4 over one side: returns (confidence float64 [F,H,W], valid_count int32 [F,H,W])."""
    if depths.dim() != 3 or depths.dtype != torch.float32 or not depths.is_cuda:
        raise RuntimeError("depths must be a float32 CUDA tensor [F,H,W]")
    depths = depths.contiguous()
    F, H, W = depths.shape
    K = _as_np(K, np.float32, (F, 3, 3))
    Ecw = _as_np(Ecw, np.float32, (F, 4, 4))
    Einv = _as_np(Ecw_inv, np.float32, (F, 4, 4))
    conf = torch.zeros((F, H, W), dtype=torch.float64, device=depths.device)
    cnt = torch.zeros((F, H, W), dtype=torch.int32, device=depths.device)
    fv = frame_valid.to(depths.device).to(torch.int32).contiguous() if frame_valid is not None else None
    with torch.cuda.device(depths.device):
        _lib.check(_lib.lib().mq3d_confidence(_lib.dptr(depths), _lib.dptr(fv), F, W, H, _lib.farr(K), _lib.farr(Ecw),
                                              _lib.farr(Einv), int(target_frame_range), C.c_double(depth_max),
                                              C.c_double(error_threshold), _lib.dptr(conf), _lib.dptr(cnt), _stream()))
    return conf, cnt
