"""Drop-in replacements for the reference's ``models/flux_ae.py`` modules, computing on libdmvae_b200.so.

Same class names, constructor signatures, attribute tree and ``state_dict`` keys/shapes as the reference
(models/flux_ae.py:25-277), so checkpoints interchange and ``train_tokenizer.py`` / ``train_dmd.py`` can import
these instead.  Parameters live in ordinary ``nn.Conv2d`` / ``nn.GroupNorm`` holders (so ``init_weights``'
isinstance dispatch, DDP, AdamW, EMA deepcopy all behave as before); only ``forward`` differs: it runs the
sm_100a kernels on channels-last bf16 activations.

Tensors crossing module boundaries are NCHW-*logical* (as in the reference).  Between our own modules they carry
channels-last strides, so ``x.permute(0, 2, 3, 1)`` is a free view of the [B][H][W][C] buffer the kernels use.
Arithmetic follows the reference under ``torch.autocast(bfloat16)``: bf16 conv operands/outputs with fp32
accumulation, GroupNorm + swish evaluated in fp32 and rounded once, bf16 residual adds.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import ops


def _to_cl(x: Tensor) -> Tensor:
    """NCHW-logical tensor -> contiguous (B, H, W, C) bf16 (zero-copy when already channels-last bf16)."""
    if x.dtype == torch.bfloat16:
        v = x.permute(0, 2, 3, 1)
        if v.is_contiguous():
            return v
    return ops.to_channels_last(x)


def _from_cl(y: Tensor) -> Tensor:
    return y.permute(0, 3, 1, 2)


class _PackMixin:
    """Lazily created bf16 operand caches, one per conv; never part of state_dict, never deep-copied."""

    def _pack(self, name: str) -> ops.WeightPack:
        packs = self.__dict__.setdefault("_packs", {})
        p = packs.get(name)
        if p is None:
            p = packs[name] = ops.WeightPack()
        return p

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_packs" else copy.deepcopy(v, memo)
        return new


def _conv(mod: "_PackMixin", name: str, conv: nn.Conv2d, x_cl: Tensor, stride=1, pad_tl=(1, 1), residual=None, pad_br=None,
          gn_next=False) -> Tensor:
    """gn_next: the output goes into a GroupNorm next, so let the conv epilogue produce its statistics."""
    return ops.conv2d(x_cl, conv.weight, conv.bias, mod._pack(name), stride, pad_tl, residual, pad_br, gn_next)


def _gn(norm: nn.GroupNorm, x_cl: Tensor, silu: bool) -> Tensor:
    return ops.group_norm_silu(x_cl, norm.weight, norm.bias, silu)


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


class AttnBlock(_PackMixin, nn.Module):
    """models/flux_ae.py:25-52"""

    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = _norm(in_channels)
        for name in ("q", "k", "v", "proj_out"):
            setattr(self, name, nn.Conv2d(in_channels, in_channels, kernel_size=1))

    def _forward_cl(self, x: Tensor) -> Tensor:
        B, H, W, c = x.shape
        h, x = ops.group_norm_silu_skip(x, self.norm.weight, self.norm.bias, False)
        q = _conv(self, "q", self.q, h, 1, (0, 0)).view(B, H * W, c)        # "b c h w -> b 1 (h w) c" is free here
        k = _conv(self, "k", self.k, h, 1, (0, 0)).view(B, H * W, c)
        v = _conv(self, "v", self.v, h, 1, (0, 0)).view(B, H * W, c)
        o = ops.single_head_attention(q, k, v)                              # single head, d = C (library GEMMs, 0.3% of FLOPs)
        o = o.reshape(B, H, W, c)
        return _conv(self, "proj_out", self.proj_out, o, 1, (0, 0), residual=x, gn_next=True)

    def forward(self, x: Tensor) -> Tensor:
        return _from_cl(self._forward_cl(_to_cl(x)))


class ResnetBlock(_PackMixin, nn.Module):
    """models/flux_ae.py:55-82"""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.norm1, self.conv1 = _norm(in_channels), _conv3(in_channels, out_channels)
        self.norm2, self.conv2 = _norm(out_channels), _conv3(out_channels, out_channels)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1)

    def _forward_cl(self, x: Tensor) -> Tensor:
        a1, x = ops.group_norm_silu_skip(x, self.norm1.weight, self.norm1.bias, True)   # x: residual branch
        h = _conv(self, "conv1", self.conv1, a1, gn_next=True)
        h = _gn(self.norm2, h, True)
        if self.in_channels != self.out_channels:
            x = _conv(self, "nin_shortcut", self.nin_shortcut, x, 1, (0, 0))
        return _conv(self, "conv2", self.conv2, h, residual=x, gn_next=True)      # residual add fused into the conv epilogue

    def forward(self, x):
        return _from_cl(self._forward_cl(_to_cl(x)))


class Downsample(_PackMixin, nn.Module):
    """models/flux_ae.py:85-95: pad (0,1,0,1) + 3x3 stride 2 (the pad is the kernel's bounds check)."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def _forward_cl(self, x: Tensor) -> Tensor:
        return _conv(self, "conv", self.conv, x, 2, (0, 0), pad_br=(1, 1))

    def forward(self, x: Tensor):
        return _from_cl(self._forward_cl(_to_cl(x)))


class Upsample(_PackMixin, nn.Module):
    """models/flux_ae.py:98-107: nearest 2x + 3x3"""

    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def _forward_cl(self, x: Tensor) -> Tensor:
        if ops.upsample_conv_supported(x, self.conv.weight):
            # sub-pixel form: four 2x2 phase convs on the low-res tensor, the upsampled tensor is never materialised
            packs = self.__dict__.setdefault("_packs", {})
            pack = packs.get("conv.subpixel")
            if pack is None:
                pack = packs["conv.subpixel"] = ops.SubpixelPack()
            return ops.upsample_conv(x, self.conv.weight, self.conv.bias, pack, True)
        return _conv(self, "conv", self.conv, ops.upsample2x(x), gn_next=True)

    def forward(self, x: Tensor):
        return _from_cl(self._forward_cl(_to_cl(x)))


def _out_dtype(x: Tensor) -> torch.dtype:
    if x.dtype == torch.bfloat16 or torch.is_autocast_enabled():
        return torch.bfloat16
    return torch.float32


def _norm(c: int) -> nn.GroupNorm:
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


def _conv3(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1)


def _stage(c_in: int, c_out: int, n_blocks: int) -> nn.Module:
    """One resolution level: `.block` (ResnetBlocks, first one changes width) and an empty `.attn` list, the
    attribute names the reference's state_dict uses (down.N.block.M / up.N.block.M)."""
    stage = nn.Module()
    stage.block = nn.ModuleList(ResnetBlock(c_in if i == 0 else c_out, c_out) for i in range(n_blocks))
    stage.attn = nn.ModuleList()
    return stage


def _middle(c: int) -> nn.Module:
    mid = nn.Module()
    mid.block_1 = ResnetBlock(c, c)
    mid.attn_1 = AttnBlock(c)
    mid.block_2 = ResnetBlock(c, c)
    return mid


class Encoder(_PackMixin, nn.Module):
    """models/flux_ae.py:110-181 (unused by the reference's VAE; exercised by the 512x512 stress config)."""

    def __init__(self, resolution: int, in_channels: int, ch: int, ch_mult: list[int], num_res_blocks: int, z_channels: int):
        super().__init__()
        self.ch = ch
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.in_ch_mult = (1,) + tuple(ch_mult)
        self.conv_in = _conv3(in_channels, ch)
        widths = [ch * m for m in self.in_ch_mult]
        self.down = nn.ModuleList()
        for lvl in range(self.num_resolutions):
            last = lvl == self.num_resolutions - 1
            stage = _stage(widths[lvl], widths[lvl + 1], num_res_blocks)
            if not last:
                stage.downsample = Downsample(widths[lvl + 1])
            self.down.append(stage)
        top = widths[-1]
        self.mid = _middle(top)
        self.norm_out = _norm(top)
        self.conv_out = _conv3(top, z_channels * 2)

    def forward(self, x: Tensor) -> Tensor:
        out_dtype = _out_dtype(x)
        h = _conv(self, "conv_in", self.conv_in, _to_cl(x))
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block]._forward_cl(h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block]._forward_cl(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample._forward_cl(h)
        h = self.mid.block_1._forward_cl(h)
        h = self.mid.attn_1._forward_cl(h)
        h = self.mid.block_2._forward_cl(h)
        h = _conv(self, "conv_out", self.conv_out, _gn(self.norm_out, h, True))
        return ops.to_nchw(h, out_dtype)


class Decoder(_PackMixin, nn.Module):
    """models/flux_ae.py:184-277"""

    def __init__(self, ch: int, out_ch: int, ch_mult: list[int], num_res_blocks: int, in_channels: int, resolution: int,
                 z_channels: int):
        super().__init__()
        self.ch = ch
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.ffactor = 2 ** (self.num_resolutions - 1)
        widths = [ch * m for m in ch_mult]
        self.block_in = widths[-1]
        low_res = resolution // self.ffactor
        self.z_shape = (1, z_channels, low_res, low_res)
        self.conv_in = _conv3(z_channels, self.block_in)
        self.mid = _middle(self.block_in)
        # stages are built top-down (lowest resolution first) but stored so that up[0] is the full-resolution one
        stages, c_in = [], self.block_in
        for lvl in reversed(range(self.num_resolutions)):
            stage = _stage(c_in, widths[lvl], num_res_blocks + 1)
            c_in = widths[lvl]
            if lvl != 0:
                stage.upsample = Upsample(c_in)
            stages.append(stage)
        self.up = nn.ModuleList(reversed(stages))
        self.norm_out = _norm(c_in)
        self.conv_out = _conv3(c_in, out_ch)

    def forward(self, z: Tensor, grad_ckpt=False) -> Tensor:
        out_dtype = _out_dtype(z)
        if z.ndim == 3:
            # (B, h*w, C) tokens are already channels-last; the reference hard-codes h = w = 16 (:244-245)
            B, n, c = z.shape
            zc = z.reshape(B, 16, 16, c)
            zc = zc if zc.dtype == torch.bfloat16 else zc.to(torch.bfloat16)
            zc = zc.contiguous()
        else:
            zc = _to_cl(z)
        if isinstance(self.conv_in, nn.Sequential):         # post_init stem: Upsample(z) + 3x3
            h = self.conv_in[0]._forward_cl(zc)
            h = _conv(self, "conv_in.1", self.conv_in[1], h, gn_next=True)
        else:
            h = _conv(self, "conv_in", self.conv_in, zc, gn_next=True)
        h = self.mid.block_1._forward_cl(h)
        h = self.mid.attn_1._forward_cl(h)
        h = self.mid.block_2._forward_cl(h)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block]._forward_cl(h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block]._forward_cl(h)
            if i_level != 0:
                h = self.up[i_level].upsample._forward_cl(h)
        h = _conv(self, "conv_out", self.conv_out, _gn(self.norm_out, h, True))
        return ops.to_nchw(h, out_dtype)

    def post_init(self, z_channels):
        """Swap the stem for nearest-2x + 3x3 followed by the 3x3 widening conv (reference :271-275)."""
        self.conv_in = nn.Sequential(Upsample(z_channels), _conv3(z_channels, self.block_in))

    def get_last_layer(self):
        return self.conv_out.weight
