"""LightningDiT velocity network (s_real / s_fake of the DMD stage) -- SURVEY.md section 8(f) row N1, first step.

A stock-PyTorch restatement of ``diffusion/lightningdit/lightningdit.py`` (RMSNorm, QK-norm, 2-D RoPE, SwiGLU, adaLN)
under the reference's parameter names, so LightningDiT checkpoints (``model`` / ``ema`` entries) load with ``strict=True``.
On the hot path these networks are black-box producers of v_teacher / v_student (4 no-grad forwards per VAE turn,
train_dmd.py:212-217).  What this module adds over the reference:

* under ``no_grad`` + ``autocast(bf16)`` on a GPU (exactly those scoring passes) the elementwise chains around the GEMMs run as two
  library kernels -- RMSNorm + adaLN modulate -> bf16, and QK-norm + 2-D RoPE + head split -> bf16 (csrc/dit_ops.cu) -- in place
  of the reference's ``@torch.compile`` sites; the autograd path (the student's training step) is the stock-PyTorch restatement;

* ``forward_cond_uncond``: the conditional and the unconditional pass of classifier-free guidance as ONE batched forward
  (2B rows) instead of two, halving launches for the DMD loss (train_dmd.py:214-217 calls the model twice);
* no ``torch.compile`` sites (the reference decorates five functions; the north star rules Triton out) and no
  fairscale / timm imports.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ._lib import call, dtype_code, ptr


def _fused_ok(x: torch.Tensor) -> bool:
    """The no-grad scoring passes of the DMD loss on a GPU under autocast(bf16): the fused glue kernels apply."""
    return (x.is_cuda and not torch.is_grad_enabled() and torch.is_autocast_enabled()
            and torch.get_autocast_dtype("cuda") == torch.bfloat16)


def rmsnorm_modulate(norm: "RMSNorm", x: torch.Tensor, shift, scale) -> torch.Tensor:
    """modulate(norm(x), shift, scale) -> bf16 in one kernel; shift / scale: (B, D) bf16 views of the adaLN output (or None)."""
    B, N, D = x.shape
    x = x.contiguous()
    y = torch.empty((B, N, D), dtype=torch.bfloat16, device=x.device)
    ref = scale if scale is not None else shift
    stride = ref.stride(0) if ref is not None else 0
    call("dmvae_rmsnorm_modulate", ptr(x), dtype_code(x), ptr(norm.weight), ptr(shift), ptr(scale), stride, ptr(y), B * N, N, D,
         float(norm.eps))
    return y


def modulate(x, shift, scale):
    """lightningdit.py:27-31"""
    if shift is None:
        return x * (1 + scale.unsqueeze(1))
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


class RMSNorm(nn.Module):
    """rms_norm.py:34-77: fp32 normalisation, result cast back, then the learnable gain."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x) * self.weight


class Rope2D(nn.Module):
    """VisionRotaryEmbeddingFast (pos_embed.py:96-134), 'lang' frequencies: buffers ``freqs_cos`` / ``freqs_sin`` of shape
    (seq*seq, 2*dim); applied as t*cos + rotate_half(t)*sin on the head dimension."""

    def __init__(self, dim: int, pt_seq_len: int = 16, theta: float = 10000.0):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        t = torch.arange(pt_seq_len) / pt_seq_len * pt_seq_len
        f = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)          # (seq, dim)
        fh = f[:, None, :].expand(pt_seq_len, pt_seq_len, -1)
        fw = f[None, :, :].expand(pt_seq_len, pt_seq_len, -1)
        full = torch.cat([fh, fw], dim=-1).reshape(pt_seq_len * pt_seq_len, -1)      # (seq*seq, 2*dim)
        self.register_buffer("freqs_cos", full.cos())
        self.register_buffer("freqs_sin", full.sin())

    @staticmethod
    def _rotate_half(x):
        x1, x2 = x[..., 0::2], x[..., 1::2]
        return torch.stack((-x2, x1), dim=-1).flatten(-2)

    def forward(self, t):
        return t * self.freqs_cos + self._rotate_half(t) * self.freqs_sin


class Attention(nn.Module):
    """lightningdit.py:34-92 with qkv_bias=True, fused SDPA."""

    def __init__(self, dim: int, num_heads: int, qk_norm: bool, use_rmsnorm: bool):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        norm = RMSNorm if use_rmsnorm else (lambda d: nn.LayerNorm(d))
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = norm(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = norm(self.head_dim) if qk_norm else nn.Identity()
        self.proj = nn.Linear(dim, dim)

    def forward(self, x, rope=None):
        B, N, C = x.shape
        if (_fused_ok(x) and isinstance(self.q_norm, (RMSNorm, nn.Identity)) and type(self.q_norm) is type(self.k_norm)
                and self.head_dim % 2 == 0 and self.head_dim <= 192):
            qkv = self.qkv(x)                               # bf16 under autocast, (B, N, 3*C) = [B][N][3][H][hd]
            if qkv.dtype == torch.bfloat16 and qkv.is_contiguous():
                H, hd = self.num_heads, self.head_dim
                q, k, v = (torch.empty((B, H, N, hd), dtype=torch.bfloat16, device=x.device) for _ in range(3))
                qn = isinstance(self.q_norm, RMSNorm)
                call("dmvae_qk_norm_rope", ptr(qkv), ptr(self.q_norm.weight) if qn else None, ptr(self.k_norm.weight) if qn else None,
                     ptr(rope.freqs_cos) if rope is not None else None, ptr(rope.freqs_sin) if rope is not None else None,
                     ptr(q), ptr(k), ptr(v), B, N, H, hd, float(self.q_norm.eps) if qn else 0.0)
                o = F.scaled_dot_product_attention(q, k, v)
                return self.proj(o.transpose(1, 2).reshape(B, N, C))
        q, k, v = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4).unbind(0)
        q, k = self.q_norm(q), self.k_norm(k)
        if rope is not None:
            q, k = rope(q), rope(k)
        x = F.scaled_dot_product_attention(q, k, v)            # under autocast q/k (fp32 after the norm gain) are cast back
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class SwiGLUFFN(nn.Module):
    """swiglu_ffn.py:15-36"""

    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.w12 = nn.Linear(dim, 2 * hidden, bias=True)
        self.w3 = nn.Linear(hidden, dim, bias=True)

    def forward(self, x):
        x1, x2 = self.w12(x).chunk(2, dim=-1)
        return self.w3(F.silu(x1) * x2)


class _GeluMlp(nn.Module):
    """timm Mlp with tanh-GELU (the use_swiglu=False branch, lightningdit.py:212-218); keys fc1 / fc2."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x), approximate="tanh"))


class TimestepEmbedder(nn.Module):
    """lightningdit.py:95-139"""

    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.frequency_embedding_size = frequency_embedding_size
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size), nn.SiLU(), nn.Linear(hidden_size, hidden_size))

    def forward(self, t):
        half = self.frequency_embedding_size // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
        args = t[:, None].float() * freqs[None]
        return self.mlp(torch.cat([torch.cos(args), torch.sin(args)], dim=-1))


class LabelEmbedder(nn.Module):
    """lightningdit.py:142-170 (training-time label dropout included)."""

    def __init__(self, num_classes, hidden_size, dropout_prob):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), hidden_size)
        self.num_classes, self.dropout_prob = num_classes, dropout_prob

    def forward(self, labels, train, force_drop_ids=None):
        if (train and self.dropout_prob > 0) or force_drop_ids is not None:
            drop = (torch.rand(labels.shape[0], device=labels.device) < self.dropout_prob) if force_drop_ids is None else force_drop_ids == 1
            labels = torch.where(drop, self.num_classes, labels)
        return self.embedding_table(labels)


class LightningDiTBlock(nn.Module):
    """lightningdit.py:173-252"""

    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0, use_qknorm=False, use_swiglu=False, use_rmsnorm=False, wo_shift=False):
        super().__init__()
        mk = (lambda: RMSNorm(hidden_size)) if use_rmsnorm else (lambda: nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6))
        self.norm1, self.norm2 = mk(), mk()
        self.attn = Attention(hidden_size, num_heads, use_qknorm, use_rmsnorm)
        hid = int(hidden_size * mlp_ratio)
        self.mlp = SwiGLUFFN(hidden_size, int(2 / 3 * hid)) if use_swiglu else _GeluMlp(hidden_size, hid)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, (4 if wo_shift else 6) * hidden_size))
        self.wo_shift = wo_shift

    def forward(self, x, c, feat_rope=None):
        m = self.adaLN_modulation(c)
        if self.wo_shift:
            scale_msa, gate_msa, scale_mlp, gate_mlp = m.chunk(4, dim=1)
            shift_msa = shift_mlp = None
        else:
            shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = m.chunk(6, dim=1)
        if _fused_ok(x) and isinstance(self.norm1, RMSNorm) and x.shape[-1] in (256, 768, 1152) and scale_msa.dtype == torch.bfloat16:
            x = x + gate_msa.unsqueeze(1) * self.attn(rmsnorm_modulate(self.norm1, x, shift_msa, scale_msa), rope=feat_rope)
            return x + gate_mlp.unsqueeze(1) * self.mlp(rmsnorm_modulate(self.norm2, x, shift_mlp, scale_mlp))
        x = x + gate_msa.unsqueeze(1) * self.attn(modulate(self.norm1(x), shift_msa, scale_msa), rope=feat_rope)
        return x + gate_mlp.unsqueeze(1) * self.mlp(modulate(self.norm2(x), shift_mlp, scale_mlp))


class FinalLayer(nn.Module):
    """lightningdit.py:254-274"""

    def __init__(self, hidden_size, patch_size, out_channels, use_rmsnorm=False):
        super().__init__()
        self.norm_final = RMSNorm(hidden_size) if use_rmsnorm else nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size))

    def forward(self, x, c):
        shift, scale = self.adaLN_modulation(c).chunk(2, dim=1)
        if _fused_ok(x) and isinstance(self.norm_final, RMSNorm) and x.shape[-1] in (256, 768, 1152) and scale.dtype == torch.bfloat16:
            return self.linear(rmsnorm_modulate(self.norm_final, x, shift, scale))
        return self.linear(modulate(self.norm_final(x), shift, scale))


class _PatchEmbed(nn.Module):
    """timm PatchEmbed as the reference uses it: Conv2d(k = s = patch) then flatten(2).transpose(1, 2)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


def sincos_pos_embed_2d(embed_dim: int, grid_size: int) -> np.ndarray:
    """get_2d_sincos_pos_embed (lightningdit.py:468-514): w-first meshgrid, [sin | cos] per axis, h half then w half."""
    def one_d(dim, pos):
        omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    gw, gh = np.meshgrid(np.arange(grid_size, dtype=np.float32), np.arange(grid_size, dtype=np.float32))
    return np.concatenate([one_d(embed_dim // 2, gw), one_d(embed_dim // 2, gh)], axis=1)


class LightningDiT(nn.Module):
    """lightningdit.py:277-421"""

    def __init__(self, input_size=32, patch_size=2, in_channels=32, hidden_size=1152, depth=28, num_heads=16, mlp_ratio=4.0,
                 class_dropout_prob=0.1, num_classes=1000, learn_sigma=False, use_qknorm=True, use_swiglu=True, use_rope=True,
                 use_rmsnorm=True, wo_shift=False, use_checkpoint=False):
        super().__init__()
        self.learn_sigma, self.in_channels = learn_sigma, in_channels
        self.out_channels = in_channels * (2 if learn_sigma else 1)
        self.patch_size, self.num_heads, self.depth, self.hidden_size = patch_size, num_heads, depth, hidden_size
        self.num_classes = num_classes
        self.x_embedder = _PatchEmbed(input_size, patch_size, in_channels, hidden_size)
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.y_embedder = LabelEmbedder(num_classes, hidden_size, class_dropout_prob)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.x_embedder.num_patches, hidden_size), requires_grad=False)
        self.feat_rope = Rope2D(hidden_size // num_heads // 2, input_size // patch_size) if use_rope else None
        self.blocks = nn.ModuleList([LightningDiTBlock(hidden_size, num_heads, mlp_ratio, use_qknorm, use_swiglu, use_rmsnorm, wo_shift)
                                     for _ in range(depth)])
        self.final_layer = FinalLayer(hidden_size, patch_size, self.out_channels, use_rmsnorm)
        self.initialize_weights()

    def initialize_weights(self):
        """lightningdit.py:343-378 (note: the final layer is zero-initialised, so a fresh model outputs v = 0)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        g = int(self.x_embedder.num_patches ** 0.5)
        self.pos_embed.data.copy_(torch.from_numpy(sincos_pos_embed_2d(self.pos_embed.shape[-1], g)).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.zeros_(self.x_embedder.proj.bias)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for blk in self.blocks:
            nn.init.zeros_(blk.adaLN_modulation[-1].weight)
            nn.init.zeros_(blk.adaLN_modulation[-1].bias)
        for lin in (self.final_layer.adaLN_modulation[-1], self.final_layer.linear):
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)

    def unpatchify(self, x):
        c, p = self.out_channels, self.patch_size
        h = w = int(x.shape[1] ** 0.5)
        x = x.reshape(x.shape[0], h, w, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, h * p, w * p)

    def _trunk(self, x, c):
        for blk in self.blocks:
            x = blk(x, c, self.feat_rope)
        x = self.unpatchify(self.final_layer(x, c))
        return x.chunk(2, dim=1)[0] if self.learn_sigma else x

    def forward(self, x, t=None, y=None):
        x = self.x_embedder(x) + self.pos_embed
        c = self.t_embedder(t) + self.y_embedder(y, self.training)
        return self._trunk(x, c)

    def forward_cond_uncond(self, x, t, y) -> Tuple[torch.Tensor, torch.Tensor]:
        """(v(x, t, y), v(x, t, null-class)) from one batched pass: rows [0, B) conditional, [B, 2B) unconditional."""
        B = x.shape[0]
        tok = self.x_embedder(x) + self.pos_embed
        te = self.t_embedder(t)
        ye = self.y_embedder.embedding_table(torch.cat([y, torch.full_like(y, self.num_classes)]))
        out = self._trunk(torch.cat([tok, tok]), torch.cat([te, te]) + ye)
        return out[:B], out[B:]


def LightningDiT_Mini_1(**kw):
    return LightningDiT(depth=6, hidden_size=256, patch_size=1, num_heads=4, **kw)


def LightningDiT_XL_1(**kw):
    return LightningDiT(depth=28, hidden_size=1152, patch_size=1, num_heads=16, **kw)


def LightningDiT_B_1(**kw):
    return LightningDiT(depth=12, hidden_size=768, patch_size=1, num_heads=12, **kw)
