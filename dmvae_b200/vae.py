"""VAE wrapper with the reference's surface (models/vae.py:71-121): ``encoder`` (DINOv2 ViT, frozen in stage 1) ->
``bottle_neck`` MLP -> ``decoder`` (our sm_100a Decoder).

The encoder stays stock PyTorch (SDPA + cuBLAS): it is a black-box feature extractor on this path.
The reference builds it with ``timm.create_model('vit_{base,large}_patch14_dinov2.lvd142m', pretrained=True,
patch_size=16, img_size=256)`` (models/vae.py:47-50); timm and its weights are not available offline, so
``DinoViT`` below restates that architecture (pre-norm blocks, LayerScale, cls token + learned position
embedding, no register tokens) under timm's parameter names, so a reference checkpoint's ``encoder.model.*``
tensors load with ``strict=True``.  Weights are random-initialised unless a checkpoint is loaded.
"""
from __future__ import annotations

import os
from contextlib import nullcontext

import torch
from torch import nn

from ._lib import call, ptr
from .autoencoder import Decoder

_VIT = {"base": dict(dim=768, depth=12, heads=12), "large": dict(dim=1024, depth=24, heads=16)}


def init_weights(model: nn.Module, conv_std_or_gain: float = 0.02, other_std: float = 0.02):
    """models/init_param.py:4-33 semantics: trunc-normal(std) for conv / linear / embedding weights, zero biases,
    norm layers to (1, 0); |conv_std_or_gain| > 10 skips; negative selects xavier_normal with that gain."""
    if abs(conv_std_or_gain) > 10:
        return
    convs = (nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d)
    norms = (nn.LayerNorm, nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.SyncBatchNorm, nn.GroupNorm,
             nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)
    for m in model.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.trunc_normal_(m.weight.data, std=other_std)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
            if isinstance(m, nn.Embedding) and m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()
        elif isinstance(m, convs):
            if conv_std_or_gain > 0:
                nn.init.trunc_normal_(m.weight.data, std=conv_std_or_gain)
            else:
                nn.init.xavier_normal_(m.weight.data, gain=-conv_std_or_gain)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, norms):
            if m.bias is not None:
                m.bias.data.zero_()
            if m.weight is not None:
                m.weight.data.fill_(1.0)


def _frozen_cast(mod: nn.Module, name: str) -> torch.Tensor:
    """Under autocast every Linear re-casts its fp32 weight to bf16 on every forward.  For frozen parameters (the
    encoder in stage 1, train_tokenizer.py:295-297) the cast is cached on the module across steps; numerics unchanged."""
    p = getattr(mod, name)
    if p is None or p.requires_grad or not p.is_cuda or not torch.is_autocast_enabled():
        return p
    dt = torch.get_autocast_dtype("cuda")
    cache = mod.__dict__.setdefault("_cast_cache", {})
    hit = cache.get(name)
    if hit is None or hit[0] != p._version or hit[1] != p.data_ptr() or hit[2].dtype != dt:
        hit = (p._version, p.data_ptr(), p.detach().to(dt))
        cache[name] = hit
    return hit[2]


def _linear(mod: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    return nn.functional.linear(x, _frozen_cast(mod, "weight"), _frozen_cast(mod, "bias"))


class _Affine(nn.Module):
    """(x - a) / b or x * b + a with per-channel buffers named ``mean`` / ``std`` (models/vae.py:10-31)."""

    def __init__(self, mean, std, inverse: bool):
        super().__init__()
        self.register_buffer("mean", torch.tensor(mean).view(1, -1, 1, 1))
        self.register_buffer("std", torch.tensor(std).view(1, -1, 1, 1))
        self.inverse = inverse

    def forward(self, x):
        return x * self.std + self.mean if self.inverse else (x - self.mean) / self.std


class _PatchEmbed(nn.Module):
    def __init__(self, patch, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        p = self.proj.kernel_size[0]
        if (x.is_cuda and not torch.is_grad_enabled() and torch.is_autocast_enabled() and x.shape[-1] % p == 0 and x.shape[-2] % p == 0
                and torch.get_autocast_dtype("cuda") == torch.bfloat16):
            # frozen pass: the stride = kernel conv is a plain GEMM over non-overlapping patches (cuDNN runs it as an fp32
            # implicit-GEMM conv, 0.24 ms per step); same products, bf16 operands as autocast would give the conv
            B, C, H, W = x.shape
            cols = x.to(torch.bfloat16).view(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), C * p * p)
            w = _frozen_cast(self.proj, "weight").reshape(self.proj.out_channels, -1)
            return nn.functional.linear(cols, w, _frozen_cast(self.proj, "bias"))
        return self.proj(x).flatten(2).transpose(1, 2)


class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, D = x.shape
        q, k, v = _linear(self.qkv, x).view(B, N, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4)
        o = nn.functional.scaled_dot_product_attention(q, k, v)
        return _linear(self.proj, o.transpose(1, 2).reshape(B, N, D))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return _linear(self.fc2, nn.functional.gelu(_linear(self.fc1, x)))


class _Gamma(nn.Module):
    def __init__(self, dim, init=1e-5):
        super().__init__()
        self.gamma = nn.Parameter(init * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


def _fused_glue_ok(x: torch.Tensor, dim: int) -> bool:
    """The no-grad encoder pass of stage 1 (train_tokenizer.py:295-297, vae.py:92-93) under autocast(bf16) on a GPU: LayerNorm ->
    bf16 and LayerScale + residual run as one library kernel each instead of two / two ATen passes."""
    return (x.is_cuda and not torch.is_grad_enabled() and x.dtype == torch.float32 and x.is_contiguous()
            and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16 and dim in (384, 512, 768, 1024))


def _ln_bf16(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    call("dmvae_layernorm_bf16", ptr(x), ptr(norm.weight), ptr(norm.bias), ptr(y), x.numel() // x.shape[-1], x.shape[-1], float(norm.eps))
    return y


def _scale_residual_(x: torch.Tensor, y: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """x += y * gamma in place (x: fp32 residual stream owned by the encoder pass, y: bf16 branch output)."""
    if y.dtype != torch.bfloat16 or not y.is_contiguous():
        return x.add_(y * gamma)
    call("dmvae_scale_residual", ptr(x), ptr(y), ptr(gamma), x.numel() // x.shape[-1], x.shape[-1])
    return x


def _scale_residual_ln_(x: torch.Tensor, y: torch.Tensor, gamma: torch.Tensor, norm: nn.LayerNorm) -> torch.Tensor:
    """x += y * gamma in place, and returns bf16 LayerNorm(x) with ``norm``'s parameters (the next branch's pre-norm): one kernel,
    bit-identical to _scale_residual_ followed by _ln_bf16."""
    if y.dtype != torch.bfloat16 or not y.is_contiguous():
        return _ln_bf16(norm, _scale_residual_(x, y, gamma))
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    call("dmvae_scale_residual_layernorm", ptr(x), ptr(y), ptr(gamma), ptr(norm.weight), ptr(norm.bias), ptr(out),
         x.numel() // x.shape[-1], x.shape[-1], float(norm.eps))
    return out


class _Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, heads)
        self.ls1 = _Gamma(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, 4 * dim)
        self.ls2 = _Gamma(dim)

    def forward(self, x, own_buffer: bool = False):
        if _fused_glue_ok(x, x.shape[-1]):
            if not own_buffer:
                x = x.clone()                       # the residual stream is updated in place below
            x = _scale_residual_(x, self.attn(_ln_bf16(self.norm1, x)), self.ls1.gamma)
            return _scale_residual_(x, self.mlp(_ln_bf16(self.norm2, x)), self.ls2.gamma)
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class DinoViT(nn.Module):
    """DINOv2 ViT-B/L trunk with timm's VisionTransformer parameter names."""

    def __init__(self, dim, depth, heads, patch_size=16, img_size=256):
        super().__init__()
        self.num_prefix_tokens = 1
        n = (img_size // patch_size) ** 2
        self.patch_embed = _PatchEmbed(patch_size, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n + 1, dim) * 0.02)
        self.blocks = nn.Sequential(*[_Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embed
        if _fused_glue_ok(x, x.shape[-1]) and len(self.blocks) > 0:
            # frozen pass: every LayerScale + residual update is fused with the LayerNorm that follows it (the same block's norm2, or
            # the next block's norm1), so a block costs two glue launches instead of four; x is a fresh tensor owned by this pass
            h = _ln_bf16(self.blocks[0].norm1, x)
            for i, blk in enumerate(self.blocks):
                h = _scale_residual_ln_(x, blk.attn(h), blk.ls1.gamma, blk.norm2)
                y = blk.mlp(h)
                if i + 1 < len(self.blocks):
                    h = _scale_residual_ln_(x, y, blk.ls2.gamma, self.blocks[i + 1].norm1)
                else:
                    x = _scale_residual_(x, y, blk.ls2.gamma)
            return self.norm(x)
        for blk in self.blocks:                     # x is a fresh tensor owned by this pass
            x = blk(x, own_buffer=True)
        return self.norm(x)


class DINOEncoder(nn.Module):
    """models/vae.py:34-53"""

    def __init__(self, model_size="base", patch_size=16, image_size=256):
        super().__init__()
        cfg = _VIT[model_size]
        self.dim = cfg["dim"]
        self.de_scale = _Affine([0.5, 0.5, 0.5], [0.5, 0.5, 0.5], inverse=True)
        self.scale = _Affine([0.485, 0.456, 0.406], [0.229, 0.224, 0.225], inverse=False)
        self.model = DinoViT(cfg["dim"], cfg["depth"], cfg["heads"], patch_size, image_size)

    def forward(self, x):
        return self.model.forward_features(self.scale(self.de_scale(x)))[:, self.model.num_prefix_tokens:]


class MLP(nn.Module):
    """models/vae.py:56-68"""

    def __init__(self, in_dim, out_dim, hidden_dim=2048):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.SiLU(), nn.Linear(hidden_dim, out_dim))

    def forward(self, x):
        return self.mlp(x)

    def get_last_layer(self):
        return self.mlp[-1].weight


class VAE(nn.Module):
    """models/vae.py:71-121"""

    def __init__(self, z_channels: int = 16, image_size: int = 256, model_size: str = "base", patch_size: int = 16,
                 conv_std_or_gain: float = 0.02):
        super().__init__()
        self.encoder = DINOEncoder(model_size, patch_size=patch_size)
        self.decoder = Decoder(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, resolution=256,
                               z_channels=16)
        self.decoder.post_init(z_channels=z_channels)
        self.bottle_neck = MLP(in_dim=self.encoder.dim, out_dim=z_channels)
        init_weights(self.bottle_neck, conv_std_or_gain)
        init_weights(self.decoder, conv_std_or_gain)

    def forward(self, x, freeze_encoder=False, return_latent=False):
        with (torch.no_grad() if freeze_encoder else nullcontext()):
            latent_tokens = self.encoder(x)
        latent_tokens = self.bottle_neck(latent_tokens)
        x_rec = self.decoder(latent_tokens).float()
        return (x_rec, latent_tokens) if return_latent else x_rec

    @torch.inference_mode()
    def encode(self, x):
        return self.bottle_neck(self.encoder(x))

    @torch.inference_mode()
    def decode(self, latent_tokens):
        return self.decoder(latent_tokens)

    def load_pretrained(self, state_dict_path, ema=False):
        if not os.path.exists(state_dict_path):
            print(f"[WARNING] VAE state_dict_path {state_dict_path} not found, skip loading")
            return
        try:
            ckpt = torch.load(state_dict_path, map_location="cpu")
        except Exception:
            ckpt = torch.load(state_dict_path, map_location="cpu", weights_only=False)
        key = "vae_ema" if (ema and "vae_ema" in ckpt) else "vae_wo_ddp"
        self.load_state_dict(ckpt[key], strict=True)


def latents_to_spatial(tokens: torch.Tensor) -> torch.Tensor:
    """train_dmd.py:408-416 (p = 1): (B, h*w, C) -> (B, C, h, w), a pure index permutation."""
    B, n, c = tokens.shape
    s = int(n ** 0.5)
    assert s * s == n
    return tokens.transpose(1, 2).reshape(B, c, s, s)
