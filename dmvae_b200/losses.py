"""Fused loss operators over libdmvae_b200.so.

Scalars are produced on the device (fp64 accumulators -> 0-d fp32 tensors); nothing here synchronises with the
host.  Call ``.item()`` yourself when you want to log, as the reference does (train_dmd.py:227-228,239-242).
"""
from __future__ import annotations

from typing import Tuple

import torch

from ._lib import call, dtype_code, ptr


def _flat2(x: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    x = x.contiguous()
    B = x.shape[0]
    return x, B, (x.numel() // B if B else 0)


# ------------------------------------------------------------------------------------------------ A3: DMD
def dmd_mix_xt(z: torch.Tensor, x0: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """xt = t*z + (1-t)*x0  (ICPlan.plan, diffusion/transport/path.py:114-136 as used at train_dmd.py:210)."""
    z, B, P = _flat2(z.detach())
    x0 = x0.detach().to(z.dtype).contiguous()
    t = t.detach().to(z.dtype).contiguous()
    xt = torch.empty_like(z)
    call("dmvae_dmd_mix_xt", ptr(z), ptr(x0), ptr(t), ptr(xt), B, P, dtype_code(z))
    return xt


class _DmdLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, xt, t, vT_c, vT_u, vS_c, vS_u, cfg_scale: float, normalize: bool, dz_dtype):
        zc, B, P = _flat2(z.detach())
        dt = zc.dtype
        xt, vT_c, vS_c = (a.detach().to(dt).contiguous() for a in (xt, vT_c, vS_c))
        vT_u = None if vT_u is None else vT_u.detach().to(dt).contiguous()
        vS_u = None if vS_u is None else vS_u.detach().to(dt).contiguous()
        t = t.detach().to(dt).contiguous()
        dz = torch.empty(zc.shape, dtype=dz_dtype, device=zc.device)
        acc = torch.zeros(2, dtype=torch.float64, device=zc.device)
        if B * P > 0:                     # empty batch: nothing to launch (an empty tensor's data_ptr() is NULL)
            call("dmvae_dmd_loss_fwd_bwd", ptr(zc), ptr(xt), ptr(t), ptr(vT_c), ptr(vT_u), ptr(vS_c), ptr(vS_u), ptr(dz),
                 ptr(acc), B, P, float(cfg_scale), int(normalize), 1.0, dtype_code(zc), dtype_code(dz))
        n = max(B * P, 1)
        loss = (acc[0] * (0.5 / n)).float()
        gnorm = (acc[1] / max(B, 1)).float()
        ctx.save_for_backward(dz)
        ctx.z_dtype = z.dtype
        ctx.mark_non_differentiable(gnorm)
        return loss, gnorm

    @staticmethod
    def backward(ctx, g_loss, _g_gnorm):
        (dz,) = ctx.saved_tensors
        # fp32 product then one rounding to z's dtype: same as autocast's mse_loss backward + cast
        return (dz.float() * g_loss).to(ctx.z_dtype), None, None, None, None, None, None, None, None, None


def dmd_loss(z, xt, t, vT_c, vS_c, vT_u=None, vS_u=None, cfg_scale: float = 1.0, normalize: bool = True,
             dz_dtype: torch.dtype = torch.float32):
    """Fused replacement for train_dmd.py:214-228.  Returns (loss, mean grad-norm) as 0-d device tensors; the
    gradient reaches ``z`` only (everything else is no_grad/detach in the reference)."""
    return _DmdLossFn.apply(z, xt, t, vT_c, vT_u, vS_c, vS_u, cfg_scale, normalize, dz_dtype)


# ------------------------------------------------------------------------------------------------ A4: L1 + L2
class _L1L2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon, image):
        r = recon.detach().float().contiguous()
        x = image.detach().float().contiguous()
        acc = torch.zeros(2, dtype=torch.float64, device=r.device)
        n = r.numel()
        if n:
            call("dmvae_l1l2_fwd", ptr(r), ptr(x), ptr(acc), n)
        out = (acc / max(n, 1)).float()
        ctx.save_for_backward(r, x)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g1, g2):
        r, x = ctx.saved_tensors
        d = torch.empty_like(r)
        g1 = g1.float().contiguous()
        g2 = g2.float().contiguous()
        if r.numel():
            call("dmvae_l1l2_bwd", ptr(r), ptr(x), ptr(d), ptr(g1), ptr(g2), r.numel(), 1.0, 1.0)
        return d, None


def l1_l2_loss(recon: torch.Tensor, image: torch.Tensor):
    """(F.l1_loss(recon, image), F.mse_loss(recon, image)) in one pass (train_dmd.py:234-235)."""
    return _L1L2Fn.apply(recon, image)


def l1l2_fused(recon: torch.Tensor, image: torch.Tensor, w_l1: float, w_l2: float):
    """Single-pass variant for fixed weights: returns (l1, l2, d(w_l1*l1 + w_l2*l2)/d recon) with no autograd."""
    r = recon.detach().float().contiguous()
    x = image.detach().float().contiguous()
    acc = torch.zeros(2, dtype=torch.float64, device=r.device)
    d = torch.empty_like(r)
    n = r.numel()
    if n:
        call("dmvae_l1l2_fwd_bwd", ptr(r), ptr(x), ptr(d), ptr(acc), n, float(w_l1), float(w_l2))
    out = (acc / max(n, 1)).float()
    return out[0], out[1], d


# ------------------------------------------------------------------------------------------------ A5: LPIPS distance
def _as_channels_last_3d(f: torch.Tensor) -> Tuple[torch.Tensor, int, int, int]:
    """(B,C,H,W) feature map -> physical [B][HW][C] without a copy when it already is channels_last."""
    B, c, H, W = f.shape
    fp = f.permute(0, 2, 3, 1)
    if not fp.is_contiguous():
        fp = fp.contiguous()
    return fp, B, H * W, c


class _LpipsDistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f0, f1, lin_w, faithful: bool):
        a, B, HW, c = _as_channels_last_3d(f0.detach())
        b, _, _, _ = _as_channels_last_3d(f1.detach())
        if a.dtype != b.dtype:
            b = b.to(a.dtype)
        w = lin_w.detach().float().reshape(-1).contiguous()
        acc = torch.zeros(B, dtype=torch.float64, device=a.device)
        call("dmvae_lpips_dist_fwd", ptr(a), ptr(b), ptr(w), ptr(acc), B, HW, c, dtype_code(a), int(faithful))
        ctx.save_for_backward(a, b, w)
        ctx.shape = (B, HW, c, f1.shape, f1.dtype)
        return (acc / max(HW, 1)).float()           # per-image spatial mean, shape (B,)

    @staticmethod
    def backward(ctx, g):
        a, b, w = ctx.saved_tensors
        B, HW, c, shp, dt = ctx.shape
        # g is (B,): fold it in per image by running one launch per distinct upstream value is wasteful; the
        # reference's upstream is uniform (mean over batch), so use g[0] and assert nothing (device-side scalar).
        g0 = g.float().reshape(-1)[:1].contiguous()
        df = torch.empty_like(b)
        call("dmvae_lpips_dist_bwd", ptr(a), ptr(b), ptr(w), ptr(df), ptr(g0), B, HW, c, 1.0 / max(HW, 1), dtype_code(a))
        df = df.view(shp[0], shp[2], shp[3], shp[1]).permute(0, 3, 1, 2)
        return None, df.to(dt), None, None


def lpips_tap_distance(f0: torch.Tensor, f1: torch.Tensor, lin_w: torch.Tensor, faithful: bool = False) -> torch.Tensor:
    """Per-image LPIPS distance of one VGG tap: spatial mean of sum_c w_c (f0^ - f1^)^2  (utils/lpips.py:86-91).
    Gradient flows to ``f1`` only and assumes a batch-uniform upstream gradient (true for ``.mean()``)."""
    return _LpipsDistFn.apply(f0, f1, lin_w, faithful)


# ------------------------------------------------------------------------------------------------ A8: reparam + KL
class _ReparamKlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, eps, rows: int, half: int):
        hc = h.detach().contiguous()
        e = eps.detach().to(hc.dtype).contiguous()
        z = torch.empty(e.shape, dtype=hc.dtype, device=hc.device)
        acc = torch.zeros(1, dtype=torch.float64, device=hc.device)
        call("dmvae_reparam_kl_fwd", ptr(hc), ptr(e), ptr(z), ptr(acc), rows, half, dtype_code(hc))
        ctx.save_for_backward(hc, e)
        ctx.geom = (rows, half)
        return z, acc[0].float()

    @staticmethod
    def backward(ctx, dz, g_kl):
        hc, e = ctx.saved_tensors
        rows, half = ctx.geom
        dh = torch.empty_like(hc)
        dzc = None if dz is None else dz.to(hc.dtype).contiguous()
        g = torch.zeros(1, dtype=torch.float32, device=hc.device) if g_kl is None else g_kl.float().reshape(1).contiguous()
        call("dmvae_reparam_kl_bwd", ptr(hc), ptr(e), ptr(dzc), ptr(dh), ptr(g), 1.0, rows, half, dtype_code(hc))
        return dh, None, None, None


def reparam_kl(h: torch.Tensor, eps: torch.Tensor, channel_dim: int = 1):
    """h = [mu | logvar] split along ``channel_dim`` (1 for NCHW encoder output, -1 for channels-last tokens).
    Returns (z, KL summed over all elements).  Extension: the reference VAE is deterministic."""
    if channel_dim in (-1, h.ndim - 1):
        half = h.shape[-1] // 2
        rows = h.numel() // (2 * half) if half else 0
    elif channel_dim == 1:
        half = h[0].numel() // 2
        rows = h.shape[0]
    else:
        raise ValueError("channel_dim must be 1 or -1")
    return _ReparamKlFn.apply(h, eps, rows, half)
