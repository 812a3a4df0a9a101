"""CPU oracle for the DMVAE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional PyTorch-on-CPU restatement of the reference's arithmetic for every row of SURVEY.md section 8(a).
Only tests/, __graft_entry__.smoke() and bench.py's baseline / checker legs may import this module -- in bench.py those are
cpu_baseline and --impl reference (this file timed on the host cores), and the two legs the round-1 review asked for, whose
reference-side arms are stock PyTorch steps assembled from these functions (scripts/stock_arms.py): gpu_baseline (the stock
cuDNN-eager step on the same GPU) and loss_parity (strict-fp32 anchor and cuDNN-autocast control the real trainer is compared
with).  It is never on the path that produces `value` / `e2e`, and nothing under dmvae_b200/ imports it (the product path fails
loudly without its CUDA library).

Pinning: tests/golden/make_golden.py runs the *real* reference modules (imported from /root/reference in the
build container) on seeded inputs and stores input/output vectors; tests/test_oracle_golden.py checks every
function here against those vectors (fp32, tolerance 1e-5) -- so parity is pinned for rows A1-A5.  Row A8
(reparameterize + KL) has no reference implementation: PARITY UNPINNED for that row only.

Functions take a flat ``state_dict``-style mapping (reference key names) instead of nn.Modules, so the file
shares no structure with the reference's classes.

``bf16=True`` inserts the roundings CUDA autocast(bf16) inserts in the reference run: conv/linear operands and
outputs are bf16, GroupNorm / swish / losses run in fp32 (torch autocast policy), elementwise ops on bf16 tensors
round after every op.  That is the mode the CUDA kernels are compared against.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Mapping[str, Tensor]


def r16(x: Tensor) -> Tensor:
    """round-trip through bf16 (value stays fp32)"""
    return x.to(torch.bfloat16).to(torch.float32)


# ----------------------------------------------------------------------------------------------------------------
# A1 / A2: flux_ae blocks                                  reference: models/flux_ae.py
# ----------------------------------------------------------------------------------------------------------------
def conv2d(sd: SD, name: str, x: Tensor, stride: int = 1, padding: int = 0, bf16: bool = False) -> Tensor:
    """nn.Conv2d forward (flux_ae.py:32-35,63,65,67,89,101,133,158,210,237). Under autocast: bf16 operands and output."""
    w, b = sd[name + ".weight"].float(), sd.get(name + ".bias")
    if bf16:
        y = F.conv2d(r16(x), r16(w), None, stride=stride, padding=padding)
        if b is not None:
            y = y + b.float().view(1, -1, 1, 1)       # cuDNN adds the (fp32-held) bias before the single rounding
        return r16(y)
    return F.conv2d(x, w, None if b is None else b.float(), stride=stride, padding=padding)


def group_norm(sd: SD, name: str, x: Tensor, swish: bool, eps: float = 1e-6) -> Tensor:
    """GroupNorm(32, C, eps=1e-6) [+ x*sigmoid(x)]  (flux_ae.py:21-22,30,62,64,157,236); fp32 under autocast."""
    y = F.group_norm(x.float(), 32, sd[name + ".weight"].float(), sd[name + ".bias"].float(), eps)
    return y * torch.sigmoid(y) if swish else y


def resnet_block(sd: SD, p: str, x: Tensor, bf16: bool = False) -> Tensor:
    """flux_ae.py:69-82"""
    h = conv2d(sd, p + ".conv1", group_norm(sd, p + ".norm1", x, True), 1, 1, bf16)
    h = conv2d(sd, p + ".conv2", group_norm(sd, p + ".norm2", h, True), 1, 1, bf16)
    if (p + ".nin_shortcut.weight") in sd:
        x = conv2d(sd, p + ".nin_shortcut", x, 1, 0, bf16)
    y = x + h
    return r16(y) if bf16 else y


def attn_block(sd: SD, p: str, x: Tensor, bf16: bool = False) -> Tensor:
    """flux_ae.py:37-52: GN -> q,k,v 1x1 -> single-head softmax(q k^T / sqrt(C)) v -> 1x1 -> +x"""
    h = group_norm(sd, p + ".norm", x, False)
    q, k, v = (conv2d(sd, p + "." + n, h, 1, 0, bf16) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2) for t in (q, k, v))      # (b, hw, c)
    att = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1)
    o = att @ v
    if bf16:
        o = r16(o)
    o = o.transpose(1, 2).reshape(b, c, hh, ww)
    y = x + conv2d(sd, p + ".proj_out", o, 1, 0, bf16)
    return r16(y) if bf16 else y


def upsample(sd: SD, p: str, x: Tensor, bf16: bool = False) -> Tensor:
    """flux_ae.py:103-107: nearest 2x then 3x3"""
    x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    return conv2d(sd, p + ".conv", x, 1, 1, bf16)


def downsample(sd: SD, p: str, x: Tensor, bf16: bool = False) -> Tensor:
    """flux_ae.py:91-95: zero-pad right/bottom by one, 3x3 stride 2"""
    return conv2d(sd, p + ".conv", F.pad(x, (0, 1, 0, 1)), 2, 0, bf16)


def _count(sd: SD, prefix: str) -> int:
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


def decoder_forward(sd: SD, z: Tensor, bf16: bool = False, token_hw: Optional[Tuple[int, int]] = (16, 16)) -> Tensor:
    """flux_ae.Decoder.forward (:239-269) incl. the post_init stem (:271-275).  z: (B, HW, C) tokens or (B, C, H, W)."""
    if z.ndim == 3:
        h_, w_ = token_hw
        z = z.transpose(1, 2).reshape(z.shape[0], z.shape[2], h_, w_)
    z = z.float()
    if bf16:
        z = r16(z)
    if "conv_in.0.conv.weight" in sd:       # post_init: Upsample(z) + 3x3
        h = conv2d(sd, "conv_in.1", upsample(sd, "conv_in.0", z, bf16), 1, 1, bf16)
    else:
        h = conv2d(sd, "conv_in", z, 1, 1, bf16)
    h = resnet_block(sd, "mid.block_1", h, bf16)
    h = attn_block(sd, "mid.attn_1", h, bf16)
    h = resnet_block(sd, "mid.block_2", h, bf16)
    levels = _count(sd, "up")
    for lvl in reversed(range(levels)):
        for blk in range(_count(sd, f"up.{lvl}.block")):
            h = resnet_block(sd, f"up.{lvl}.block.{blk}", h, bf16)
        if f"up.{lvl}.upsample.conv.weight" in sd:
            h = upsample(sd, f"up.{lvl}.upsample", h, bf16)
    h = group_norm(sd, "norm_out", h, True)
    return conv2d(sd, "conv_out", h, 1, 1, bf16)


def encoder_forward(sd: SD, x: Tensor, bf16: bool = False) -> Tensor:
    """flux_ae.Encoder.forward (:160-181)"""
    x = x.float()
    h = conv2d(sd, "conv_in", x, 1, 1, bf16)
    levels = _count(sd, "down")
    for lvl in range(levels):
        for blk in range(_count(sd, f"down.{lvl}.block")):
            h = resnet_block(sd, f"down.{lvl}.block.{blk}", h, bf16)
        if f"down.{lvl}.downsample.conv.weight" in sd:
            h = downsample(sd, f"down.{lvl}.downsample", h, bf16)
    h = resnet_block(sd, "mid.block_1", h, bf16)
    h = attn_block(sd, "mid.attn_1", h, bf16)
    h = resnet_block(sd, "mid.block_2", h, bf16)
    h = group_norm(sd, "norm_out", h, True)
    return conv2d(sd, "conv_out", h, 1, 1, bf16)


def latents_to_spatial(tokens: Tensor) -> Tensor:
    """train_dmd.py:408-416 with p=1: (B, h*w, C) -> (B, C, h, w); a pure transpose (bit-exact)."""
    b, n, c = tokens.shape
    s = int(round(math.sqrt(n)))
    assert s * s == n
    return tokens.transpose(1, 2).reshape(b, c, s, s)


# ----------------------------------------------------------------------------------------------------------------
# A3: DMD loss                                              reference: train_dmd.py:204-230
# ----------------------------------------------------------------------------------------------------------------
def _bt(t: Tensor, x: Tensor) -> Tensor:
    return t.view(t.shape[0], *([1] * (x.ndim - 1)))


def dmd_mix_xt(z: Tensor, x0: Tensor, t: Tensor) -> Tensor:
    """ICPlan.compute_mu_t (path.py:114-124): alpha_t * x1 + sigma_t * x0, alpha=t, sigma=1-t.  Tensor ops in the
    tensors' own dtype, so bf16 inputs round after every op exactly like the reference."""
    tb = _bt(t, z)
    return tb * z + (1 - tb) * x0


def dmd_loss(z: Tensor, xt: Tensor, t: Tensor, vT_c: Tensor, vS_c: Tensor, vT_u: Optional[Tensor] = None,
             vS_u: Optional[Tensor] = None, cfg_scale: float = 1.0, normalize: bool = True):
    """train_dmd.py:214-228 (normalize=True) / toy_example_2d/dmd.py:349-360 (normalize=False, no CFG).
    Returns (loss fp32 scalar, mean per-sample grad norm, dz = dloss/dz in fp32)."""
    vT, vS = vT_c, vS_c
    if cfg_scale > 1 and vT_u is not None:
        vT = vT + (cfg_scale - 1) * (vT - vT_u)
        vS = vS + (cfg_scale - 1) * (vS - vS_u)
    omt = _bt(1 - t, xt)
    pred_T = xt + vT * omt
    pred_S = xt + vS * omt
    p_real = z - pred_T
    p_student = z - pred_S
    if normalize:
        w = p_real.abs().mean(dim=list(range(1, z.ndim)), keepdim=True)
        grad = (p_real - p_student) / w
    else:
        grad = p_real - p_student
    grad = torch.nan_to_num(grad)
    target = z - grad
    zf = z.float().detach().requires_grad_(True)
    loss = 0.5 * F.mse_loss(zf, target.float(), reduction="mean")     # autocast runs mse_loss in fp32
    (dz,) = torch.autograd.grad(loss, zf)
    gnorm = torch.norm(grad.float().flatten(1), dim=1).mean()
    return loss.detach(), gnorm, dz


# ----------------------------------------------------------------------------------------------------------------
# A4: L1 + L2                                               reference: train_dmd.py:234-235
# ----------------------------------------------------------------------------------------------------------------
def l1l2(recon: Tensor, image: Tensor, w_l1: float = 1.0, w_l2: float = 0.0):
    r = recon.float().detach().requires_grad_(True)
    l1 = F.l1_loss(r, image.float())
    l2 = F.mse_loss(r, image.float())
    (g,) = torch.autograd.grad(w_l1 * l1 + w_l2 * l2, r)
    return l1.detach(), l2.detach(), g


# ----------------------------------------------------------------------------------------------------------------
# A5: LPIPS                                                 reference: utils/lpips.py
# ----------------------------------------------------------------------------------------------------------------
VGG_TAPS = {3: 0, 8: 1, 15: 2, 22: 3, 29: 4}      # torchvision vgg16.features index after which a tap is read (:140-149)
VGG_CONVS = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]
VGG_POOLS = [4, 9, 16, 23]
LPIPS_SHIFT = (-0.030, -0.088, -0.188)
LPIPS_SCALE = (0.458, 0.448, 0.450)


def vgg_features(sd: SD, x: Tensor, bf16: bool = False) -> List[Tensor]:
    """ScalingLayer (:97-104) + torchvision VGG16.features sliced at relu1_2..relu5_3 (:116-153).
    sd keys: net.slice{k}.{idx}.{weight,bias}."""
    shift = torch.tensor(LPIPS_SHIFT).view(1, 3, 1, 1)
    scale = torch.tensor(LPIPS_SCALE).view(1, 3, 1, 1)
    h = (x.float() - shift) / scale
    feats = []
    slice_of = lambda i: 1 + sum(i > t for t in (3, 8, 15, 22))
    for i in range(30):
        if i in VGG_CONVS:
            h = conv2d(sd, f"net.slice{slice_of(i)}.{i}", h, 1, 1, bf16)
            h = F.relu(h)
        elif i in VGG_POOLS:
            h = F.max_pool2d(h, 2, 2)
        if i in VGG_TAPS:
            feats.append(h)
    return feats


def lpips_distance(f0s: Sequence[Tensor], f1s: Sequence[Tensor], lin_ws: Sequence[Tensor], faithful: bool = False) -> Tensor:
    """lpips.py:86-94 from the five feature maps (NCHW) onward; returns the 0-d loss.
    faithful=True adds the bf16 roundings of the autocast'ed 1x1 `lin` conv and the bf16 tail (mean/sum/mean)."""
    val = None
    for f0, f1, w in zip(f0s, f1s, lin_ws):
        f0, f1 = f0.float(), f1.float()
        n0 = torch.sqrt((f0 ** 2).sum(1, keepdim=True))
        n1 = torch.sqrt((f1 ** 2).sum(1, keepdim=True))
        d = (f0 / (n0 + 1e-10) - f1 / (n1 + 1e-10)) ** 2
        wv = w.float().view(1, -1, 1, 1)
        if faithful:
            lin = r16((r16(d) * r16(wv)).sum(1, keepdim=True))
            res = r16(lin.mean([2, 3], keepdim=True))
        else:
            res = (d * wv).sum(1, keepdim=True).mean([2, 3], keepdim=True)
        val = res if val is None else (r16(val + res) if faithful else val + res)
    out = val.mean()
    return r16(out) if faithful else out


def lpips_forward(sd: SD, x: Tensor, y: Tensor, bf16: bool = False) -> Tensor:
    """LPIPS.forward(input, target) (:81-94)"""
    lin_ws = [sd[f"lin{k}.model.1.weight"].flatten() for k in range(5)]
    return lpips_distance(vgg_features(sd, x, bf16), vgg_features(sd, y, bf16), lin_ws, faithful=bf16)


# ----------------------------------------------------------------------------------------------------------------
# A8: reparameterize + KL (extension -- no reference implementation; parity unpinned)
# ----------------------------------------------------------------------------------------------------------------
def reparam_kl(mu: Tensor, logvar: Tensor, eps: Tensor):
    """z = mu + exp(logvar/2) * eps ;  KL(N(mu, e^lv) || N(0, I)) summed over all elements."""
    mu, logvar, eps = mu.float(), logvar.float(), eps.float()
    z = mu + torch.exp(0.5 * logvar) * eps
    kl = 0.5 * (mu ** 2 + torch.exp(logvar) - 1.0 - logvar).sum()
    return z, kl


# ----------------------------------------------------------------------------------------------------------------
# Weight construction (mirrors models/init_param.py: trunc-normal std 0.02 for conv/linear, GN weight 1 / bias 0)
# ----------------------------------------------------------------------------------------------------------------
def _conv_p(sd: Dict[str, Tensor], name: str, cin: int, cout: int, k: int, g: torch.Generator, std: float):
    w = torch.empty(cout, cin, k, k)
    torch.nn.init.trunc_normal_(w, std=std, generator=g)
    sd[name + ".weight"] = w
    sd[name + ".bias"] = torch.zeros(cout)


def _gn_p(sd, name, c):
    sd[name + ".weight"] = torch.ones(c)
    sd[name + ".bias"] = torch.zeros(c)


def _res_p(sd, p, cin, cout, g, std):
    _gn_p(sd, p + ".norm1", cin); _conv_p(sd, p + ".conv1", cin, cout, 3, g, std)
    _gn_p(sd, p + ".norm2", cout); _conv_p(sd, p + ".conv2", cout, cout, 3, g, std)
    if cin != cout:
        _conv_p(sd, p + ".nin_shortcut", cin, cout, 1, g, std)


def _attn_p(sd, p, c, g, std):
    _gn_p(sd, p + ".norm", c)
    for n in ("q", "k", "v", "proj_out"):
        _conv_p(sd, f"{p}.{n}", c, c, 1, g, std)


def make_decoder_state(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=32, seed=0, std=0.02,
                       post_init=True, randomize_affine=False) -> Dict[str, Tensor]:
    """A state_dict with the reference Decoder's keys/shapes (flux_ae.py:185-237,271-275)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    block_in = ch * ch_mult[-1]
    if post_init:
        _conv_p(sd, "conv_in.0.conv", z_channels, z_channels, 3, g, std)
        _conv_p(sd, "conv_in.1", z_channels, block_in, 3, g, std)
    else:
        _conv_p(sd, "conv_in", z_channels, block_in, 3, g, std)
    _res_p(sd, "mid.block_1", block_in, block_in, g, std)
    _attn_p(sd, "mid.attn_1", block_in, g, std)
    _res_p(sd, "mid.block_2", block_in, block_in, g, std)
    for lvl in reversed(range(len(ch_mult))):
        block_out = ch * ch_mult[lvl]
        for blk in range(num_res_blocks + 1):
            _res_p(sd, f"up.{lvl}.block.{blk}", block_in, block_out, g, std)
            block_in = block_out
        if lvl != 0:
            _conv_p(sd, f"up.{lvl}.upsample.conv", block_in, block_in, 3, g, std)
    _gn_p(sd, "norm_out", block_in)
    _conv_p(sd, "conv_out", block_in, out_ch, 3, g, std)
    if randomize_affine:
        for k in list(sd):
            if ".norm" in k or k.startswith("norm_out"):
                sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
            elif k.endswith(".bias"):
                sd[k] = 0.02 * torch.randn(sd[k].shape, generator=g)
    return sd


def make_encoder_state(in_channels=3, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=32, seed=0, std=0.02,
                       randomize_affine=False) -> Dict[str, Tensor]:
    """A state_dict with the reference Encoder's keys/shapes (flux_ae.py:111-158)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    _conv_p(sd, "conv_in", in_channels, ch, 3, g, std)
    in_mult = (1,) + tuple(ch_mult)
    block_in = ch
    for lvl in range(len(ch_mult)):
        block_in = ch * in_mult[lvl]
        block_out = ch * ch_mult[lvl]
        for blk in range(num_res_blocks):
            _res_p(sd, f"down.{lvl}.block.{blk}", block_in, block_out, g, std)
            block_in = block_out
        if lvl != len(ch_mult) - 1:
            _conv_p(sd, f"down.{lvl}.downsample.conv", block_in, block_in, 3, g, std)
    _res_p(sd, "mid.block_1", block_in, block_in, g, std)
    _attn_p(sd, "mid.attn_1", block_in, g, std)
    _res_p(sd, "mid.block_2", block_in, block_in, g, std)
    _gn_p(sd, "norm_out", block_in)
    _conv_p(sd, "conv_out", block_in, 2 * z_channels, 3, g, std)
    if randomize_affine:
        for k in list(sd):
            if ".norm" in k or k.startswith("norm_out"):
                sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
            elif k.endswith(".bias"):
                sd[k] = 0.02 * torch.randn(sd[k].shape, generator=g)
    return sd


VGG_CFG = [(0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256),
           (17, 256, 512), (19, 512, 512), (21, 512, 512), (24, 512, 512), (26, 512, 512), (28, 512, 512)]


def make_lpips_state(seed=0, lin_ckpt: Optional[str] = None) -> Dict[str, Tensor]:
    """LPIPS state_dict (utils/lpips.py keys).  VGG weights: Kaiming-normal like torchvision's vgg16(weights=None)
    (ImageNet weights are not available offline); lin weights from ckpt if given else |N(0,1)|/C."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    slice_of = lambda i: 1 + sum(i > t for t in (3, 8, 15, 22))
    for idx, cin, cout in VGG_CFG:
        std = math.sqrt(2.0 / (cout * 9))
        sd[f"net.slice{slice_of(idx)}.{idx}.weight"] = torch.randn(cout, cin, 3, 3, generator=g) * std
        sd[f"net.slice{slice_of(idx)}.{idx}.bias"] = torch.zeros(cout)
    chns = [64, 128, 256, 512, 512]
    lin = torch.load(lin_ckpt, map_location="cpu", weights_only=True) if lin_ckpt else {}
    for k, c in enumerate(chns):
        key = f"lin{k}.model.1.weight"
        sd[key] = lin[key].float() if key in lin else (torch.randn(1, c, 1, 1, generator=g).abs() / c)
    sd["scaling_layer.shift"] = torch.tensor(LPIPS_SHIFT).view(1, 3, 1, 1)
    sd["scaling_layer.scale"] = torch.tensor(LPIPS_SCALE).view(1, 3, 1, 1)
    return sd
