"""Torch (device-side) renderer of the same analytic scene as ``synth.py`` -- used by bench.py to
produce large synthetic sequences (depth and 1280x960 colour) without minutes of NumPy time.
Plumbing only: input generation happens outside every timed region."""
from __future__ import annotations

import numpy as np
import torch

from . import synth


def _trace(origin: torch.Tensor, dirs: torch.Tensor):
    """origin [...,3] (broadcastable against dirs), dirs [...,3] (float64) -> (t [...], surface id [...])."""
    rmin, rmax, sc, sr, bmin, bmax = (torch.as_tensor(np.asarray(a, dtype=np.float64), device=dirs.device)
                                      for a in synth._scene_o3d())
    inv = 1.0 / dirs
    t_hi = torch.where(dirs > 0, (rmax - origin) * inv, (rmin - origin) * inv)
    t_hi = torch.where(dirs == 0, torch.full_like(t_hi, float("inf")), t_hi)
    t, ax = torch.min(t_hi, dim=-1)
    sid = 2 * ax + (torch.gather(dirs, -1, ax[..., None])[..., 0] > 0).long()
    oc = origin - sc
    a = (dirs * dirs).sum(-1)
    b = 2.0 * (dirs * oc).sum(-1)
    c = (oc * oc).sum(-1) - sr * sr
    disc = b * b - 4 * a * c
    ts = torch.where(disc > 0, (-b - torch.sqrt(disc.clamp_min(0))) / (2 * a), torch.full_like(a, float("inf")))
    ts = torch.where(ts > 0, ts, torch.full_like(ts, float("inf")))
    hit = ts < t
    t = torch.where(hit, ts, t)
    sid = torch.where(hit, torch.full_like(sid, 6), sid)
    t0 = (bmin - origin) * inv
    t1 = (bmax - origin) * inv
    tn = torch.minimum(t0, t1).amax(-1)
    tf = torch.maximum(t0, t1).amin(-1)
    tb = torch.where((tn <= tf) & (tn > 0), tn, torch.full_like(tn, float("inf")))
    hit = tb < t
    t = torch.where(hit, tb, t)
    sid = torch.where(hit, torch.full_like(sid, 7), sid)
    return t, sid


def render_depth(e_cw: np.ndarray, device, width=synth.DEPTH_W, height=synth.DEPTH_H, noise=0.005,
                 dropout=0.02, seed=1234, first_frame: int = 0, out: torch.Tensor = None, chunk: int = 32) -> torch.Tensor:
    """Raw NDC depth float32 [F,H,W] on `device` for camera->world poses e_cw [F,4,4] (Open3D frame).
    Frame i draws its noise from a generator seeded with seed + first_frame + i, so any slice of a sequence
    can be rendered on its own (ranks render disjoint slices of one capture)."""
    fx, fy, cx, cy = synth.depth_intrinsics(width, height)
    u = torch.arange(width, dtype=torch.float64, device=device)
    v = torch.arange(height, dtype=torch.float64, device=device)
    dx = ((u - (width - cx)) / fx)[None, :].expand(height, width)
    dy = ((v - cy) / fy)[:, None].expand(height, width)
    dirs_cam = torch.stack([dx, dy, torch.ones_like(dx)], dim=-1)
    gen = torch.Generator(device=device)
    n = len(e_cw)
    if out is None:
        out = torch.empty((n, height, width), dtype=torch.float32, device=device)
    e_all = torch.as_tensor(np.asarray(e_cw, dtype=np.float64), device=device)
    for i0 in range(0, n, chunk):
        e = e_all[i0:i0 + chunk]
        dirs = torch.einsum("hwk,bjk->bhwj", dirs_cam, e[:, :3, :3])
        z, _ = _trace(e[:, None, None, :3, 3], dirs)
        for j in range(z.shape[0]):
            gen.manual_seed(seed + first_frame + i0 + j)
            zj = z[j]
            if noise > 0:
                zj = zj * (1.0 + noise * torch.randn(zj.shape, generator=gen, device=device, dtype=torch.float64))
            d = 1.0 - synth.NEAR / zj
            if dropout > 0:
                d = torch.where(torch.rand(zj.shape, generator=gen, device=device) < dropout, torch.ones_like(d), d)
            out[i0 + j] = d.to(torch.float32)
    return out


def render_color(e_cw: np.ndarray, device, width=synth.COLOR_W, height=synth.COLOR_H, f=synth.COLOR_F) -> torch.Tensor:
    """Procedural RGB uint8 [F,H,W,3] on `device` (per-surface hue x 25 cm checker)."""
    u = torch.arange(width, dtype=torch.float64, device=device)
    v = torch.arange(height, dtype=torch.float64, device=device)
    dx = ((u - width / 2.0) / f)[None, :].expand(height, width)
    dy = ((v - height / 2.0) / f)[:, None].expand(height, width)
    dirs_cam = torch.stack([dx, dy, torch.ones_like(dx)], dim=-1)
    hues = torch.as_tensor(synth._HUES, device=device)
    out = torch.empty((len(e_cw), height, width, 3), dtype=torch.uint8, device=device)
    for i in range(len(e_cw)):
        e = torch.as_tensor(e_cw[i].astype(np.float64), device=device)
        dirs = dirs_cam @ e[:3, :3].T
        t, sid = _trace(e[:3, 3], dirs)
        p = e[:3, 3] + t[..., None] * dirs
        checker = torch.floor(p * 4).sum(-1) % 2
        rgb = hues[sid] * (0.55 + 0.45 * checker[..., None])
        out[i] = rgb.clamp(0, 255).to(torch.uint8)
    return out
