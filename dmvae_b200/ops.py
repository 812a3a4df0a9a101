"""Host-side operators over libdmvae_b200.so: raw kernel wrappers and torch.autograd.Functions.

Activations inside the library are channels-last bf16 tensors of shape (B, H, W, C).  Parameters stay the
reference's fp32 ``nn.Parameter``s in state_dict layout; the bf16 tap-major GEMM operands are a derived cache
(``WeightPack``) refreshed when the parameter's version counter moves (optimizer.step / load_state_dict).

Every Function here honours: retain_graph=True with repeated partial ``autograd.grad`` calls (saved tensors are
never written in place -- the adaptive-GAN-weight code of train_dmd.py:248-251 relies on it), ``no_grad`` /
``inference_mode`` (nothing is saved), and DDP / clip_grad_norm_ / AdamW on the unmodified parameters.
"""
from __future__ import annotations

import contextlib

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import call, ptr, query

GN_EPS = 1e-6
EPI_RELU, EPI_MASK = 1, 2      # conv epilogue flags (csrc/conv_tc.cu): ReLU on the output / `residual` gates the output
# The conv epilogue reduces the next GroupNorm's statistics.  The line-coalesced epilogues of the pair / halo tiles and the transposed
# tile do it for group widths 4 / 8 / 16 (C = 128 / 256 / 512) at a few FADDs per stored vector; the per-thread-row epilogue of the
# remaining tiles handles any width that divides 32.  Below this group width the separate (HBM-bound) gn_stats pass is used.
import os as _os
FUSE_GN_STATS_MIN_CPG = int(_os.environ.get("DMVAE_FUSE_GN_STATS_MIN_CPG", "4"))


class ZeroPool:
    """Bump allocator over a buffer its owner zero-fills once per training step, for the many small accumulators a step needs
    (GroupNorm statistics, the group sums of its backward, affine / bias gradients: ~150 tensors of a few KB) -- instead of one
    ``torch.zeros`` fill kernel each.  GradArena owns one: the pool is the tail of the gradient buffer's allocation, so the
    arena's per-step ``zero()`` clears it in the same memset.  Allocations come from the pool only inside ``with pool:`` (a
    trainer's forward + backward); tensors handed out are valid until the owner's next ``zero()`` -- every user consumes them
    within the step (statistics saved for backward included).  A full pool falls back to ``torch.zeros``."""
    current: Optional["ZeroPool"] = None
    enabled = bool(int(_os.environ.get("DMVAE_ZERO_POOL", "1")))

    def __init__(self, buf: torch.Tensor):
        self.buf = buf.view(torch.uint8)
        self.off = 0
        self.high_water = 0
        self._prev: Optional["ZeroPool"] = None

    def reset(self) -> None:
        """The owner has just zero-filled the buffer."""
        self.off = 0

    def __enter__(self):
        self._prev, ZeroPool.current = ZeroPool.current, (self if ZeroPool.enabled else None)
        return self

    def __exit__(self, *exc):
        ZeroPool.current = self._prev
        return False


def small_zeros(shape, dtype, device) -> torch.Tensor:
    """Zero-filled tensor for a per-step accumulator: carved out of the active ZeroPool when there is one."""
    pool = ZeroPool.current
    if pool is not None and pool.buf.device == device:
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = (n * dtype.itemsize + 31) // 32 * 32
        if pool.off + nbytes <= pool.buf.numel():
            # a fresh tensor over the pool's storage, NOT a view of it: views share one autograd version counter, and any
            # in-place write to the arena (its memset, AccumulateGrad's +=) would invalidate every pool tensor saved for backward
            t = torch.empty((0,), dtype=dtype, device=device).set_(
                pool.buf.untyped_storage(), (pool.buf.storage_offset() + pool.off) // dtype.itemsize, tuple(int(d) for d in shape))
            pool.off += nbytes
            pool.high_water = max(pool.high_water, pool.off)
            return t
    return torch.zeros(shape, dtype=dtype, device=device)


def _chk_nhwc(x: torch.Tensor, name: str) -> torch.Tensor:
    if x.dtype != torch.bfloat16 or x.ndim != 4:
        raise _lib.DmvaeError(f"{name}: expected a (B,H,W,C) bfloat16 tensor, got {tuple(x.shape)} {x.dtype}")
    if not x.is_cuda:
        raise _lib.DmvaeError(f"{name}: tensor is on {x.device}; dmvae_b200 has no CPU path")
    return x if x.is_contiguous() else x.contiguous()


# ------------------------------------------------------------------------------------------------ weights
class WeightPack:
    """bf16 GEMM operands derived from one fp32 conv weight (Cout, Cin, KH, KW)."""

    __slots__ = ("version", "data_ptr", "w_fwd", "w_dgrad")

    def __init__(self):
        self.version = -1
        self.data_ptr = 0
        self.w_fwd = None
        self.w_dgrad = None

    def get(self, weight: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        v = weight._version
        if self.w_fwd is None or v != self.version or weight.data_ptr() != self.data_ptr or self.w_fwd.device != weight.device:
            w16 = getattr(weight, "_dmvae_w16", None)
            if w16 is not None and getattr(weight, "_dmvae_w16_version", None) == v and w16.device == weight.device:
                # the optimizer kernel keeps a bf16 copy of this (tap-major) parameter: that IS w_fwd; only the transposed
                # data-gradient operand is derived here, bf16 -> bf16 (optim.FlatAdamWEMA, csrc/optim.cu)
                taps, cout, cin = w16.shape
                wd = getattr(weight, "_dmvae_wd16", None)          # ... or not even that: the optimizer's batched transpose
                if wd is None:
                    wd = torch.empty((taps, cin, cout), dtype=torch.bfloat16, device=w16.device)
                    call("dmvae_pack_dgrad_bf16", ptr(w16), ptr(wd), cout, cin, taps)
                self.w_fwd, self.w_dgrad, self.version, self.data_ptr = w16, wd, v, weight.data_ptr()
                return self.w_fwd, self.w_dgrad
            w = weight.detach()
            if w.dtype != torch.float32:
                w = w.float()
            w = w.contiguous()
            cout, cin, kh, kw = w.shape
            # fresh tensors on every repack: graphs that saved the old operands keep valid data
            wf = torch.empty((kh * kw, cout, cin), dtype=torch.bfloat16, device=w.device)
            wd = torch.empty((kh * kw, cin, cout), dtype=torch.bfloat16, device=w.device)
            call("dmvae_pack_weights", ptr(w), ptr(wf), ptr(wd), cout, cin, kh, kw)
            self.w_fwd, self.w_dgrad, self.version, self.data_ptr = wf, wd, v, weight.data_ptr()
        return self.w_fwd, self.w_dgrad


# ------------------------------------------------------------------------------------------------ raw kernels
def conv_forward_raw(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor],
                     kh: int, kw: int, stride: int = 1, pad_tl: Tuple[int, int] = (1, 1),
                     out_hw: Optional[Tuple[int, int]] = None, force_direct: bool = False,
                     pad_br: Optional[Tuple[int, int]] = None, want_gn_stats: bool = False, flags: int = 0) -> torch.Tensor:
    """y = conv(x, w_packed[tap][Cout][Cin]) + bias (+ residual).  Picks the tcgen05 tile when the shape allows.
    pad_tl / pad_br: zero padding (top, left) / (bottom, right); pad_br defaults to pad_tl.
    flags: EPI_RELU -> y = max(y, 0); EPI_MASK -> ``residual`` is not added but gates the output (y where residual > 0, else 0)."""
    B, H, W, cin = x.shape
    taps, cout, cin_w = w_packed.shape
    assert taps == kh * kw and cin_w == cin, (w_packed.shape, x.shape, kh, kw)
    pt, pl = pad_tl
    if out_hw is None:
        pb, pr = pad_tl if pad_br is None else pad_br
        out_hw = ((H + pt + pb - kh) // stride + 1, (W + pl + pr - kw) // stride + 1)
    OH, OW = out_hw
    y = torch.empty((B, OH, OW, cout), dtype=torch.bfloat16, device=x.device)
    if bias is not None and bias.dtype != torch.float32:
        bias = bias.float()
    same = stride == 1 and OH == H and OW == W and pt == (kh - 1) // 2 and pl == (kw - 1) // 2
    if same and not force_direct and query("dmvae_conv_tc_supported", B, H, W, cin, cout, kh, kw):
        stats = None
        if want_gn_stats and cout % 32 == 0 and cout // 32 >= FUSE_GN_STATS_MIN_CPG and \
                (cout // 32 in (1, 2, 4, 8, 16) or (cout // 32) % 32 == 0):
            stats = small_zeros((B, 32, 2), torch.float64, x.device)
        call("dmvae_conv_tc_fwd", ptr(x), ptr(w_packed), ptr(bias), ptr(residual), ptr(y), ptr(stats), B, H, W, cin, cout, kh, kw, flags)
        if stats is not None:
            y._dmvae_gnstats = (stats, y.data_ptr(), tuple(y.shape), _ver(y))       # consumed by the next GroupNorm
    elif (stride == 2 and residual is None and not force_direct and not flags
          and query("dmvae_conv_tc_strided_supported", B, H, W, cin, OH, OW, cout, kh, kw, stride)):
        call("dmvae_conv_tc_fwd_strided", ptr(x), ptr(w_packed), ptr(bias), ptr(y), B, H, W, cin, OH, OW, cout, kh, kw, stride, pt, pl)
    else:
        call("dmvae_conv_direct_fwd", ptr(x), ptr(w_packed), ptr(bias), ptr(residual), ptr(y), B, H, W, cin, OH, OW, cout,
             kh, kw, stride, pt, pl, flags)
    return y


class _DirectGrads:
    """While a trainer's ``backward()`` runs inside ``direct_param_grads(arena)``, the custom Functions below accumulate parameter
    gradients straight into the parameters' slots of the flat gradient arena (the kernels already accumulate: ``red.add`` /
    ``+=``) and return None to autograd, instead of materialising a gradient tensor that AccumulateGrad then adds into ``.grad``
    with one more elementwise kernel per parameter (the reference's DDP path does that add too, train_dmd.py:348).  Outside the
    context -- ``torch.autograd.grad`` on ``conv_out.weight`` (train_dmd.py:249-250), plain ``.backward()`` -- nothing changes."""
    arena = None


@contextlib.contextmanager
def direct_param_grads(arena):
    prev, _DirectGrads.arena = _DirectGrads.arena, arena
    try:
        yield
    finally:
        _DirectGrads.arena = prev


def _grad_slot(p) -> Optional[torch.Tensor]:
    """The arena view to accumulate into, or None when direct mode is off / p is not an arena parameter / p.grad was re-pointed."""
    a = _DirectGrads.arena
    if a is None or p is None:
        return None
    slot = a.slot_of(p)
    if slot is None or p.grad is None or p.grad.data_ptr() != slot.data_ptr():
        return None
    return slot


def _grad_done(p) -> None:
    a = _DirectGrads.arena
    if a is not None:
        a.notify(p)


def _is_tap_major(t: torch.Tensor, cout: int, cin: int, kh: int, kw: int) -> bool:
    """t is a (Cout, Cin, KH, KW) view of a dense [KH*KW][Cout][Cin] block (train_arena.arena_view), 16-byte aligned."""
    return (t.dtype == torch.float32 and tuple(t.shape) == (cout, cin, kh, kw) and t.data_ptr() % 16 == 0
            and (kh * kw == 1 and t.is_contiguous() or t.stride() == (cin, 1, kw * cout * cin, cout * cin)))


def conv_wgrad_raw(x: torch.Tensor, dy: torch.Tensor, kh: int, kw: int, stride: int = 1, pad_tl: Tuple[int, int] = (1, 1),
                   force_direct: bool = False, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """dw (Cout, Cin, KH, KW) fp32 = sum_pixels dy (x) x.  With ``out`` (an fp32 tensor of that shape) the result is
    ACCUMULATED into it and None is returned; when ``out`` is a tap-major arena view the tensor-core kernels accumulate
    straight into its storage (no scratch, no un-packing pass)."""
    B, H, W, cin = x.shape
    _, OH, OW, cout = dy.shape
    pt, pl = pad_tl
    same = stride == 1 and OH == H and OW == W and pt == (kh - 1) // 2 and pl == (kw - 1) // 2
    direct_out = out is not None and _is_tap_major(out, cout, cin, kh, kw)
    if same and not force_direct and query("dmvae_conv_tc_wgrad_supported", B, H, W, cin, cout, kh, kw):
        if direct_out:
            call("dmvae_conv_tc_wgrad", ptr(x), ptr(dy), ptr(out), B, H, W, cin, cout, kh, kw)
            return None
        scratch = torch.zeros((kh * kw, cout, cin), dtype=torch.float32, device=x.device)
        call("dmvae_conv_tc_wgrad", ptr(x), ptr(dy), ptr(scratch), B, H, W, cin, cout, kh, kw)
        if out is not None and out.is_contiguous():
            call("dmvae_wgrad_unpack", ptr(scratch), ptr(out), cout, cin, kh * kw, 1)
            return None
        dw = torch.empty((cout, cin, kh, kw), dtype=torch.float32, device=x.device)
        call("dmvae_wgrad_unpack", ptr(scratch), ptr(dw), cout, cin, kh * kw, 0)
        if out is not None:
            out.add_(dw)
            return None
        return dw
    if stride == 2 and not force_direct and query("dmvae_conv_tc_strided_supported", B, H, W, cin, OH, OW, cout, kh, kw, stride):
        if direct_out:
            call("dmvae_conv_tc_wgrad_strided", ptr(x), ptr(dy), ptr(out), B, H, W, cin, OH, OW, cout, kh, kw, stride, pt, pl)
            return None
        scratch = torch.zeros((kh * kw, cout, cin), dtype=torch.float32, device=x.device)
        call("dmvae_conv_tc_wgrad_strided", ptr(x), ptr(dy), ptr(scratch), B, H, W, cin, OH, OW, cout, kh, kw, stride, pt, pl)
        if out is not None and out.is_contiguous():
            call("dmvae_wgrad_unpack", ptr(scratch), ptr(out), cout, cin, kh * kw, 1)
            return None
        dw = torch.empty((cout, cin, kh, kw), dtype=torch.float32, device=x.device)
        call("dmvae_wgrad_unpack", ptr(scratch), ptr(dw), cout, cin, kh * kw, 0)
        if out is not None:
            out.add_(dw)
            return None
        return dw
    dw = torch.zeros((cout, cin, kh, kw), dtype=torch.float32, device=x.device)
    call("dmvae_conv_direct_wgrad", ptr(x), ptr(dy), ptr(dw), B, H, W, cin, OH, OW, cout, kh, kw, stride, pt, pl)
    if out is not None:
        out.add_(dw)
        return None
    return dw


def conv_dgrad_raw(dy: torch.Tensor, w_fwd: torch.Tensor, w_dgrad: torch.Tensor, in_hw: Tuple[int, int], kh: int, kw: int,
                   stride: int = 1, pad_tl: Tuple[int, int] = (1, 1), force_direct: bool = False,
                   relu_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """relu_mask: the conv's own input when that input is a ReLU output -- the data gradient is zeroed where it is <= 0 in the
    epilogue (the upstream ReLU's backward, fused); stride-1 "same" convs only."""
    B, OH, OW, cout = dy.shape
    H, W = in_hw
    pt, pl = pad_tl
    cin = w_fwd.shape[2]
    if stride == 1 and OH == H and OW == W:
        # "same" conv: dX = conv(dY, flipped/transposed weights) with the mirrored padding
        return conv_forward_raw(dy, w_dgrad, None, relu_mask, kh, kw, 1, (kh - 1 - pt, kw - 1 - pl), (H, W), force_direct,
                                flags=EPI_MASK if relu_mask is not None else 0)
    if relu_mask is not None:
        raise _lib.DmvaeError("conv_dgrad_raw: relu_mask is only supported for stride-1 'same' convolutions")
    if (stride == 2 and kh == 3 and kw == 3 and pt == 0 and pl == 0 and H == 2 * OH and W == 2 * OW and cout % 8 == 0
            and not force_direct and query("dmvae_conv_tc_supported", B, H, W, cout, cin, 3, 3)):
        # Downsample (pad (0,1,0,1)): dx = conv3x3_same(zero-inserted dy, flipped weights) on the tensor cores
        dyz = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=dy.device)
        call("dmvae_zero_insert2x", ptr(dy), ptr(dyz), B, OH, OW, cout)
        return conv_forward_raw(dyz, w_dgrad, None, None, 3, 3, 1, (1, 1), (H, W))
    dx = torch.empty((B, H, W, cin), dtype=torch.bfloat16, device=dy.device)
    call("dmvae_conv_direct_dgrad_strided", ptr(dy), ptr(w_fwd), ptr(dx), B, H, W, cin, OH, OW, cout, kh, kw, stride, pt, pl)
    return dx


def bias_grad_raw(dy: torch.Tensor, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Column sums of dy; with ``out`` (fp32 [C]) they are accumulated into it (the kernel adds) and None is returned."""
    c = dy.shape[-1]
    if out is not None:
        call("dmvae_bias_grad", ptr(dy), ptr(out), dy.numel() // c, c)
        return None
    db = small_zeros((c,), torch.float32, dy.device)
    call("dmvae_bias_grad", ptr(dy), ptr(db), dy.numel() // c, c)
    return db


def _tagged_gn_stats(x: torch.Tensor) -> Optional[torch.Tensor]:
    tag = getattr(x, "_dmvae_gnstats", None)
    if tag is not None and tag[1] == x.data_ptr() and tag[2] == tuple(x.shape) and tag[3] == _ver(x):
        return tag[0]
    return None


def gn_stats_raw(x: torch.Tensor) -> torch.Tensor:
    B, H, W, c = x.shape
    stats = small_zeros((B, 32, 2), torch.float64, x.device)
    call("dmvae_gn_stats", ptr(x), ptr(stats), B, H * W, c)
    return stats


def gn_apply_raw(x, stats, gamma, beta, silu: bool, eps: float = GN_EPS) -> torch.Tensor:
    B, H, W, c = x.shape
    y = torch.empty_like(x)
    call("dmvae_gn_apply", ptr(x), ptr(stats), ptr(gamma), ptr(beta), ptr(y), B, H * W, c, eps, int(silu))
    return y


def gn_bwd_raw(da, x, stats, gamma, beta, silu: bool, dres=None, eps: float = GN_EPS, want_colsum: bool = False,
               out_dgamma: Optional[torch.Tensor] = None, out_dbeta: Optional[torch.Tensor] = None):
    """Returns (dx, dgamma, dbeta); with ``out_dgamma`` / ``out_dbeta`` (fp32 [C]) the affine gradients are accumulated into them
    (the kernel adds atomically) and None is returned in their place."""
    B, H, W, c = x.shape
    gsum = small_zeros((B, 32, 2), torch.float64, x.device)
    # one zero-filled fp32 buffer for dgamma | dbeta | column sums
    direct = out_dgamma is not None and out_dbeta is not None
    small = small_zeros((1 if direct else 3, c), torch.float32, x.device)
    dgamma, dbeta, colsum = (out_dgamma, out_dbeta, small[0]) if direct else (small[0], small[1], small[2])
    dx = torch.empty_like(x)
    call("dmvae_gn_bwd", ptr(da), ptr(x), ptr(stats), ptr(gamma), ptr(beta), ptr(gsum), ptr(dgamma), ptr(dbeta), ptr(dres),
         ptr(dx), ptr(colsum) if want_colsum else None, B, H * W, c, eps, int(silu))
    if want_colsum:
        _tag_colsum(dx, colsum)
    return (dx, None, None) if direct else (dx, dgamma, dbeta)


def _ver(t: torch.Tensor) -> int:
    """Version counter, or -1 for inference tensors (which cannot be modified in place outside inference mode)."""
    return -1 if t.is_inference() else t._version


def _tag_colsum(t: torch.Tensor, colsum: torch.Tensor) -> None:
    """Side channel from the kernel that produced a gradient tensor to the conv backward that consumes it: the
    per-channel sum over pixels (= that conv's bias gradient) was accumulated while the tensor was written."""
    t._dmvae_colsum = (colsum, t.data_ptr(), tuple(t.shape), _ver(t))


def _tagged_colsum(t: torch.Tensor) -> Optional[torch.Tensor]:
    tag = getattr(t, "_dmvae_colsum", None)
    if tag is not None and tag[1] == t.data_ptr() and tag[2] == tuple(t.shape) and tag[3] == _ver(t):
        return tag[0]
    return None


# ------------------------------------------------------------------------------------------------ autograd
def thin_output_grads(x: Optional[torch.Tensor], dy: torch.Tensor, w_fwd: torch.Tensor, kh: int, kw: int,
                      pad_tl: Tuple[int, int], need_dx: bool, need_dw: bool):
    """dx and dw of a stride-1 "same" conv with taps*Cout <= 32 (the 128 -> 3 head) as two 1x1 GEMMs on the tcgen05
    tiles over the gradient-patch matrix P (see dmvae_grad_patches)."""
    B, H, W, cout = dy.shape
    taps, _, cin = w_fwd.shape                      # packed forward operand [tap][Cout][Cin] (bf16)
    P = torch.empty((B, H, W, 32), dtype=torch.bfloat16, device=dy.device)
    call("dmvae_grad_patches", ptr(dy), ptr(P), B, H, W, cout, kh, kw, pad_tl[0], pad_tl[1])
    dx = dw = None
    if need_dx:
        # Wm[j = tap*Cout + co][ci] = w[co][ci][tap]  ->  packed 1x1 operand [1][Cout_gemm = cin][Cin_gemm = 32]
        wm = torch.zeros((1, cin, 32), dtype=torch.bfloat16, device=dy.device)
        wm[0, :, :taps * cout] = w_fwd.permute(2, 0, 1).reshape(cin, taps * cout)
        dx = conv_forward_raw(P, wm, None, None, 1, 1, 1, (0, 0))
    if need_dw:
        scratch = torch.zeros((1, 32, cin), dtype=torch.float32, device=dy.device)     # [1 tap][32 "couts"][cin]
        call("dmvae_conv_tc_wgrad", ptr(x), ptr(P), ptr(scratch), B, H, W, cin, 32, 1, 1)
        dw = scratch[0, :taps * cout].reshape(taps, cout, cin).permute(1, 2, 0).reshape(cout, cin, kh, kw).contiguous()
    return dx, dw


def _thin_output_ok(dy: torch.Tensor, w_fwd: torch.Tensor, kh: int, kw: int, stride: int, pad_tl, in_hw) -> bool:
    B, OH, OW, cout = dy.shape
    cin = w_fwd.shape[2]
    return (stride == 1 and (OH, OW) == tuple(in_hw) and kh * kw * cout <= 32 and pad_tl == ((kh - 1) // 2, (kw - 1) // 2)
            and cin % 8 == 0 and cin >= 32
            and bool(query("dmvae_conv_tc_supported", B, OH, OW, 32, cin, 1, 1))
            and bool(query("dmvae_conv_tc_wgrad_supported", B, OH, OW, cin, 32, 1, 1)))


class ConvFn(torch.autograd.Function):
    """nn.Conv2d on channels-last bf16 (models/flux_ae.py:32-35,63,65,67,89,101,133,158,210,237,274)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, pack: WeightPack, stride: int, pad_tl, pad_br=None, want_gn_stats=False):
        x = _chk_nhwc(x, "conv")
        cout, cin, kh, kw = weight.shape
        w_fwd, w_dgrad = pack.get(weight)
        if residual is not None:
            residual = _chk_nhwc(residual, "conv residual")
        y = conv_forward_raw(x, w_fwd, None if bias is None else bias.detach(), residual, kh, kw, stride, pad_tl, pad_br=pad_br,
                             want_gn_stats=want_gn_stats)
        ctx.geom = (kh, kw, stride, pad_tl, x.shape[1:3])
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.params = (weight, bias)                  # for direct accumulation into the gradient arena (see _DirectGrads)
        ctx.save_for_backward(x if ctx.needs_input_grad[1] else None, w_fwd, w_dgrad)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_fwd, w_dgrad = ctx.saved_tensors
        kh, kw, stride, pad_tl, in_hw = ctx.geom
        dy = _chk_nhwc(dy, "conv backward")
        dx = dw = db = dres = None
        if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and _thin_output_ok(dy, w_fwd, kh, kw, stride, tuple(pad_tl), in_hw):
            dx, dw = thin_output_grads(x, dy, w_fwd, kh, kw, tuple(pad_tl), ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        else:
            if ctx.needs_input_grad[0]:
                dx = conv_dgrad_raw(dy, w_fwd, w_dgrad, in_hw, kh, kw, stride, pad_tl)
            if ctx.needs_input_grad[1]:
                slot = _grad_slot(ctx.params[0])
                dw = conv_wgrad_raw(x, dy, kh, kw, stride, pad_tl, out=slot)
                if slot is not None:
                    _grad_done(ctx.params[0])
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _tagged_colsum(dy)                 # already reduced by the kernel that wrote dy (gn_bwd)
            if db is None:
                slot = _grad_slot(ctx.params[1])
                db = bias_grad_raw(dy, out=slot)
                if slot is not None:
                    _grad_done(ctx.params[1])
        if ctx.has_res and ctx.needs_input_grad[3]:
            dres = dy
        return dx, dw, db, dres, None, None, None, None, None


def _affine_slots(ctx):
    """Arena slots of a GroupNorm's (gamma, beta) when both can be accumulated directly, else (None, None)."""
    if not (ctx.needs_input_grad[1] and ctx.needs_input_grad[2]):
        return None, None
    sg, sb = _grad_slot(ctx.params[0]), _grad_slot(ctx.params[1])
    if sg is None or sb is None:
        return None, None
    return sg, sb


def _affine_done(ctx, sg) -> None:
    if sg is not None:                      # after the kernel that accumulates into the slots has been enqueued
        _grad_done(ctx.params[0])
        _grad_done(ctx.params[1])


class GroupNormSiluFn(torch.autograd.Function):
    """swish(GroupNorm(32, C, eps=1e-6)(x))  /  GroupNorm alone (silu=False)   (models/flux_ae.py:21-22,70-77,38)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, silu: bool):
        x = _chk_nhwc(x, "group_norm")
        g, b = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        stats = _tagged_gn_stats(x)                   # already reduced by the conv epilogue that wrote x
        if stats is None:
            stats = gn_stats_raw(x)
        y = gn_apply_raw(x, stats, g, b, silu)
        ctx.silu = silu
        ctx.params = (gamma, beta)
        ctx.save_for_backward(x, stats, g, b)
        return y

    @staticmethod
    def backward(ctx, da):
        x, stats, g, b = ctx.saved_tensors
        da = _chk_nhwc(da, "group_norm backward")
        sg, sb = _affine_slots(ctx)
        dx, dgamma, dbeta = gn_bwd_raw(da, x, stats, g, b, ctx.silu, want_colsum=True, out_dgamma=sg, out_dbeta=sb)
        _affine_done(ctx, sg)
        return dx, dgamma, dbeta, None


class GroupNormSiluSkipFn(torch.autograd.Function):
    """(swish(GroupNorm(x)), x): the second output is the residual branch of ResnetBlock / AttnBlock
    (models/flux_ae.py:52,82).  Returning it from the same node lets backward fold the gradient arriving over the
    skip connection into the GroupNorm backward kernel (no separate add pass)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, silu: bool):
        x = _chk_nhwc(x, "group_norm")
        g, b = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        stats = _tagged_gn_stats(x)
        if stats is None:
            stats = gn_stats_raw(x)
        y = gn_apply_raw(x, stats, g, b, silu)
        ctx.silu = silu
        ctx.params = (gamma, beta)
        ctx.save_for_backward(x, stats, g, b)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, da, dskip):
        x, stats, g, b = ctx.saved_tensors
        if da is None:                       # only the skip branch was used
            return dskip, None, None, None
        da = _chk_nhwc(da, "group_norm backward")
        dres = None if dskip is None else _chk_nhwc(dskip, "group_norm skip gradient")
        sg, sb = _affine_slots(ctx)
        dx, dgamma, dbeta = gn_bwd_raw(da, x, stats, g, b, ctx.silu, dres=dres, want_colsum=True, out_dgamma=sg, out_dbeta=sb)
        _affine_done(ctx, sg)
        return dx, dgamma, dbeta, None


class ConvReluFn(torch.autograd.Function):
    """relu(conv3x3(x) + b) with FROZEN weights: VGG16's conv -> ReLU pairs (utils/lpips.py:116-153, requires_grad=False at :121).
    The ReLU forward is the conv epilogue; its backward is never a pass of its own:
      * ``x_is_relu``: x is itself a ReLU output, so the data gradient is gated by x > 0 in the dgrad epilogue (that is the
        backward of the ReLU that produced x);
      * ``dy_premasked``: whoever consumes y already gates the gradient it sends back by y > 0 (a ConvReluFn with x_is_relu, or
        PoolTapFn); only when that is not the case (the last tap) is a small mask kernel run here."""

    @staticmethod
    def forward(ctx, x, weight, bias, pack: WeightPack, x_is_relu: bool, dy_premasked: bool):
        if weight.requires_grad or (bias is not None and bias.requires_grad):
            raise _lib.DmvaeError("ConvReluFn is for frozen convolutions (LPIPS' VGG16); use conv2d for trainable ones")
        x = _chk_nhwc(x, "conv_relu")
        cout, cin, kh, kw = weight.shape
        w_fwd, w_dgrad = pack.get(weight)
        y = conv_forward_raw(x, w_fwd, None if bias is None else bias.detach(), None, kh, kw, 1, ((kh - 1) // 2, (kw - 1) // 2),
                             flags=EPI_RELU)
        ctx.geom = (kh, kw, x.shape[1:3])
        ctx.save_for_backward(x if x_is_relu else None, None if dy_premasked else y, w_fwd, w_dgrad)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, w_fwd, w_dgrad = ctx.saved_tensors
        kh, kw, in_hw = ctx.geom
        dy = _chk_nhwc(dy, "conv_relu backward")
        if y is not None:
            masked = torch.empty_like(dy)
            call("dmvae_relu_mask", ptr(y), ptr(dy), ptr(masked), dy.numel())
            dy = masked
        dx = conv_dgrad_raw(dy, w_fwd, w_dgrad, in_hw, kh, kw, 1, ((kh - 1) // 2, (kw - 1) // 2), relu_mask=x)
        return dx, None, None, None, None, None


class PoolTapFn(torch.autograd.Function):
    """y -> (y, max_pool2d(y, 2, 2)) for a ReLU output y that is both an LPIPS tap and the input of the next VGG slice
    (utils/lpips.py:127-153).  Backward adds the tap gradient and the routed pool gradient and gates the sum by y > 0 in one
    kernel (no ATen max-pool backward, no gradient add, no ReLU backward)."""

    @staticmethod
    def forward(ctx, y):
        y = _chk_nhwc(y, "pool_tap")
        B, H, W, c = y.shape
        if H % 2 or W % 2 or c % 8:
            raise _lib.DmvaeError(f"pool_tap: need even H, W and C % 8 == 0, got {tuple(y.shape)}")
        p = torch.empty((B, H // 2, W // 2, c), dtype=y.dtype, device=y.device)
        call("dmvae_maxpool2x2_fwd", ptr(y), ptr(p), B, H // 2, W // 2, c)
        ctx.save_for_backward(y)
        return y.view_as(y), p

    @staticmethod
    def backward(ctx, d_tap, d_pooled):
        (y,) = ctx.saved_tensors
        B, H, W, c = y.shape
        if d_pooled is None:                             # only the tap was used
            if d_tap is None:
                return None
            d_tap = _chk_nhwc(d_tap, "pool_tap backward")
            out = torch.empty_like(y)
            call("dmvae_relu_mask", ptr(y), ptr(d_tap), ptr(out), y.numel())
            return out
        d_pooled = _chk_nhwc(d_pooled, "pool_tap backward")
        d_tap = None if d_tap is None else _chk_nhwc(d_tap, "pool_tap backward")
        dy = torch.empty_like(y)
        call("dmvae_pool_tap_bwd", ptr(y), ptr(d_pooled), ptr(d_tap), ptr(dy), B, H // 2, W // 2, c, 1)
        return dy


def conv_relu(x, weight, bias, pack: WeightPack, x_is_relu: bool, dy_premasked: bool = True):
    return ConvReluFn.apply(x, weight, bias, pack, x_is_relu, dy_premasked)


def pool_tap(y):
    """Returns (tap, pooled)."""
    return PoolTapFn.apply(y)


class Upsample2xFn(torch.autograd.Function):
    """F.interpolate(scale_factor=2, mode='nearest') (models/flux_ae.py:104)."""

    @staticmethod
    def forward(ctx, x):
        x = _chk_nhwc(x, "upsample2x")
        B, H, W, c = x.shape
        y = torch.empty((B, 2 * H, 2 * W, c), dtype=x.dtype, device=x.device)
        call("dmvae_upsample2x_fwd", ptr(x), ptr(y), B, H, W, c)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _chk_nhwc(dy, "upsample2x backward")
        B, H2, W2, c = dy.shape
        dx = torch.empty((B, H2 // 2, W2 // 2, c), dtype=dy.dtype, device=dy.device)
        call("dmvae_upsample2x_bwd", ptr(dy), ptr(dx), B, H2 // 2, W2 // 2, c)
        return dx


# flux_ae.Upsample (nearest 2x + 3x3) in sub-pixel form: four 2x2 phase convs on the low-res tensor (csrc/conv_tc.cu, dmvae_conv_up2x_*).
# 2.25x fewer FLOPs in all three passes and no 4x intermediate; the tap sums are rounded to bf16 once (the reference rounds every 3x3
# tap), so outputs agree with the two-kernel form to bf16 tolerance, not bit for bit -- DMVAE_SUBPIXEL_UPSAMPLE=0 restores that form.
SUBPIXEL_UPSAMPLE = bool(int(_os.environ.get("DMVAE_SUBPIXEL_UPSAMPLE", "1")))


class SubpixelPack:
    """bf16 16-tap operands of the sub-pixel Upsample conv derived from its fp32 (Cout, Cin, 3, 3) weight: (wp_fwd[16][Cout][Cin],
    wp_dgrad[16][Cin][Cout]); refreshed when the parameter's version counter moves, fresh tensors on every refresh."""

    __slots__ = ("version", "data_ptr", "wp_fwd", "wp_dgrad")

    def __init__(self):
        self.version, self.data_ptr, self.wp_fwd, self.wp_dgrad = -1, 0, None, None

    def get(self, weight: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        v = weight._version
        if self.wp_fwd is None or v != self.version or weight.data_ptr() != self.data_ptr or self.wp_fwd.device != weight.device:
            w = weight.detach()
            s_co, s_ci, s_kh, s_kw = w.stride()
            if w.dtype != torch.float32 or s_kh != 3 * s_kw:           # neither state_dict nor tap-major storage: make it contiguous
                w = w.float().contiguous()
                s_co, s_ci, s_kh, s_kw = w.stride()
            cout, cin = w.shape[:2]
            wf = torch.empty((16, cout, cin), dtype=torch.bfloat16, device=w.device)
            wd = torch.empty((16, cin, cout), dtype=torch.bfloat16, device=w.device)
            call("dmvae_subpixel_pack", ptr(w), s_co, s_ci, s_kw, ptr(wf), ptr(wd), cout, cin)
            self.wp_fwd, self.wp_dgrad, self.version, self.data_ptr = wf, wd, v, weight.data_ptr()
        return self.wp_fwd, self.wp_dgrad


def upsample_conv_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    B, H, W, cin = x.shape
    return (SUBPIXEL_UPSAMPLE and x.is_cuda and weight.shape[1] == cin and tuple(weight.shape[2:]) == (3, 3)
            and bool(query("dmvae_conv_up2x_supported", B, H, W, cin, weight.shape[0])))


class UpsampleConvFn(torch.autograd.Function):
    """conv3x3(F.interpolate(x, scale_factor=2, mode="nearest")) + bias (models/flux_ae.py:103-107) without materialising the
    upsampled tensor: x (B, H, W, Cin) -> y (B, 2H, 2W, Cout)."""

    @staticmethod
    def forward(ctx, x, weight, bias, pack: SubpixelPack, want_gn_stats: bool):
        x = _chk_nhwc(x, "upsample_conv")
        B, H, W, cin = x.shape
        cout = weight.shape[0]
        wpf, wpd = pack.get(weight)
        y = torch.empty((B, 2 * H, 2 * W, cout), dtype=torch.bfloat16, device=x.device)
        stats = None
        if want_gn_stats and cout // 32 in (4, 8, 16) and cout % 32 == 0 and cout // 32 >= FUSE_GN_STATS_MIN_CPG:
            stats = small_zeros((B, 32, 2), torch.float64, x.device)
        b = None if bias is None else bias.detach()
        if b is not None and b.dtype != torch.float32:
            b = b.float()
        call("dmvae_conv_up2x_fwd", ptr(x), ptr(wpf), ptr(b), ptr(y), ptr(stats), B, H, W, cin, cout)
        if stats is not None:
            y._dmvae_gnstats = (stats, y.data_ptr(), tuple(y.shape), _ver(y))
        ctx.has_bias = bias is not None
        ctx.params = (weight, bias)
        ctx.save_for_backward(x if ctx.needs_input_grad[1] else None, wpd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wpd = ctx.saved_tensors
        weight, bias = ctx.params
        dy = _chk_nhwc(dy, "upsample_conv backward")
        B, H2, W2, cout = dy.shape
        H, W = H2 // 2, W2 // 2
        cin = wpd.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((B, H, W, cin), dtype=torch.bfloat16, device=dy.device)
            call("dmvae_conv_up2x_dgrad", ptr(dy), ptr(wpd), ptr(dx), B, H, W, cin, cout)
        if ctx.needs_input_grad[1]:
            dwp = torch.zeros((16, cout, cin), dtype=torch.float32, device=dy.device)
            call("dmvae_conv_up2x_wgrad", ptr(x), ptr(dy), ptr(dwp), B, H, W, cin, cout)
            slot = _grad_slot(weight)
            tgt = slot
            if tgt is None or tgt.dtype != torch.float32 or tgt.stride(2) != 3 * tgt.stride(3):
                tgt = dw = torch.zeros((cout, cin, 3, 3), dtype=torch.float32, device=dy.device)
            s_co, s_ci, _, s_kw = tgt.stride()
            call("dmvae_subpixel_fold_wgrad", ptr(dwp), ptr(tgt), s_co, s_ci, s_kw, cout, cin)
            if tgt is slot:
                _grad_done(weight)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _tagged_colsum(dy)
            if db is None:
                slot = _grad_slot(bias)
                db = bias_grad_raw(dy, out=slot)
                if slot is not None:
                    _grad_done(bias)
        return dx, dw, db, None, None


def upsample_conv(x, weight, bias, pack: SubpixelPack, want_gn_stats: bool = False):
    return UpsampleConvFn.apply(x, weight, bias, pack, want_gn_stats)


class ToChannelsLastFn(torch.autograd.Function):
    """(B, C, H, W) fp32|bf16 -> (B, H, W, C) bf16."""

    @staticmethod
    def forward(ctx, x):
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        B, c, H, W = x.shape
        ctx.in_dtype = x.dtype
        y = torch.empty((B, H, W, c), dtype=torch.bfloat16, device=x.device)
        call("dmvae_nchw_to_nhwc", ptr(x), ptr(y), B, c, H * W, _lib.dtype_code(x))
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _chk_nhwc(dy, "to_channels_last backward")
        B, H, W, c = dy.shape
        dx = torch.empty((B, c, H, W), dtype=ctx.in_dtype, device=dy.device)
        call("dmvae_nhwc_to_nchw", ptr(dy), ptr(dx), B, c, H * W, _lib.dtype_code(dx))
        return dx


class ToNchwFn(torch.autograd.Function):
    """(B, H, W, C) bf16 -> (B, C, H, W) in out_dtype."""

    @staticmethod
    def forward(ctx, x, out_dtype):
        x = _chk_nhwc(x, "to_nchw")
        B, H, W, c = x.shape
        y = torch.empty((B, c, H, W), dtype=out_dtype, device=x.device)
        call("dmvae_nhwc_to_nchw", ptr(x), ptr(y), B, c, H * W, _lib.dtype_code(y))
        return y

    @staticmethod
    def backward(ctx, dy):
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        dy = dy.contiguous()
        B, c, H, W = dy.shape
        dx = torch.empty((B, H, W, c), dtype=torch.bfloat16, device=dy.device)
        call("dmvae_nchw_to_nhwc", ptr(dy), ptr(dx), B, c, H * W, _lib.dtype_code(dy))
        return dx, None


class SingleHeadAttentionFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(C)) v for one head of width C (models/flux_ae.py:43-47: SDPA with a single 512-wide head).

    Library GEMMs (cuBLAS bmm) with fp32 logits, fp32 softmax and bf16 probabilities -- the same rounding points as
    the fused SDPA kernels the reference dispatches to.  Written as explicit GEMMs because at head_dim 512 the
    stock SDPA backward falls back to a memory-efficient kernel that takes ~3 ms per step here (N4 in SURVEY 8(f)
    replaces this with a tcgen05 flash kernel)."""

    @staticmethod
    def forward(ctx, q, k, v):
        # q, k, v: (B, N, C) bf16 contiguous
        scale = q.shape[-1] ** -0.5
        s = torch.bmm(q, k.transpose(1, 2), out_dtype=torch.float32)
        p = torch.softmax(s * scale, dim=-1).to(q.dtype)
        o = torch.bmm(p, v)
        ctx.save_for_backward(q, k, v, p)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, p = ctx.saved_tensors
        do = do.contiguous()
        dv = torch.bmm(p.transpose(1, 2), do)
        dp = torch.bmm(do, v.transpose(1, 2), out_dtype=torch.float32)
        pf = p.float()
        ds = (pf * (dp - (dp * pf).sum(-1, keepdim=True)) * ctx.scale).to(q.dtype)
        dq = torch.bmm(ds, k)
        dk = torch.bmm(ds.transpose(1, 2), q)
        return dq, dk, dv


def single_head_attention(q, k, v):
    return SingleHeadAttentionFn.apply(q, k, v)


def conv2d(x, weight, bias, pack: WeightPack, stride: int = 1, pad_tl=(1, 1), residual=None, pad_br=None,
           want_gn_stats: bool = False):
    """want_gn_stats: the output feeds a GroupNorm(32): have the conv epilogue reduce its statistics."""
    return ConvFn.apply(x, weight, bias, residual, pack, stride, pad_tl, pad_br, want_gn_stats)


def group_norm_silu(x, gamma, beta, silu: bool = True):
    return GroupNormSiluFn.apply(x, gamma, beta, silu)


def group_norm_silu_skip(x, gamma, beta, silu: bool = True):
    """Returns (activation, x_for_the_residual_branch)."""
    return GroupNormSiluSkipFn.apply(x, gamma, beta, silu)


def upsample2x(x):
    return Upsample2xFn.apply(x)


def to_channels_last(x):
    return ToChannelsLastFn.apply(x)


def to_nchw(x, out_dtype=torch.float32):
    return ToNchwFn.apply(x, out_dtype)
