"""LPIPS with the reference's surface (utils/lpips.py:52-161) and a fused feature-distance reduction.

State-dict keys are the reference's: ``scaling_layer.{shift,scale}``, ``net.slice{1..5}.{idx}.{weight,bias}``,
``lin{0..4}.model.1.weight``.  The frozen VGG16 trunk runs on this library's conv tiles with ReLU / max-pool fused
(``vgg16.forward_b200``); everything after the taps -- channel-normalise, squared difference, 1x1 ``lin``
weighting, spatial mean (utils/lpips.py:86-91, ~50 ATen launches and six HBM passes in the reference) -- is one kernel
per tap reading each feature map exactly once.
"""
from __future__ import annotations

import os
import warnings

import torch
from torch import nn

from . import ops
from ._lib import DmvaeError
from .losses import lpips_tap_distance

_VGG_PLAN = [  # torchvision vgg16.features indices: (idx, kind, cin, cout)
    (0, "c", 3, 64), (1, "r"), (2, "c", 64, 64), (3, "r"),
    (4, "p"), (5, "c", 64, 128), (6, "r"), (7, "c", 128, 128), (8, "r"),
    (9, "p"), (10, "c", 128, 256), (11, "r"), (12, "c", 256, 256), (13, "r"), (14, "c", 256, 256), (15, "r"),
    (16, "p"), (17, "c", 256, 512), (18, "r"), (19, "c", 512, 512), (20, "r"), (21, "c", 512, 512), (22, "r"),
    (23, "p"), (24, "c", 512, 512), (25, "r"), (26, "c", 512, 512), (27, "r"), (28, "c", 512, 512), (29, "r"),
]
_SLICE_END = (4, 9, 16, 23, 30)


def _layer(spec):
    if spec[1] == "c":
        return nn.Conv2d(spec[2], spec[3], kernel_size=3, padding=1)
    if spec[1] == "r":
        return nn.ReLU(inplace=False)
    return nn.MaxPool2d(kernel_size=2, stride=2)


class vgg16(nn.Module):
    """The five VGG16 slices ending at relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 (utils/lpips.py:116-153)."""

    def __init__(self, requires_grad=False, pretrained=True):
        super().__init__()
        self.N_slices = 5
        start = 0
        for k, end in enumerate(_SLICE_END, start=1):
            seq = nn.Sequential()
            for spec in _VGG_PLAN[start:end]:
                seq.add_module(str(spec[0]), _layer(spec))
            setattr(self, f"slice{k}", seq)
            start = end
        if pretrained:
            self._try_load_imagenet()
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def _try_load_imagenet(self):
        try:
            from torchvision.models import vgg16 as tv_vgg16
            feats = tv_vgg16(weights="IMAGENET1K_V1").features.state_dict()   # needs a cached download
            mapped = {}
            for k, end in enumerate(_SLICE_END, start=1):
                lo = 0 if k == 1 else _SLICE_END[k - 2]
                for name, v in feats.items():
                    if lo <= int(name.split(".")[0]) < end:
                        mapped[f"slice{k}.{name}"] = v
            self.load_state_dict(mapped, strict=True)
        except Exception as e:  # offline: keep the default init, say so once
            warnings.warn(f"LPIPS: ImageNet VGG16 weights unavailable ({type(e).__name__}); using random init")

    def forward(self, X):
        outs, h = [], X
        for k in range(1, 6):
            h = getattr(self, f"slice{k}")(h)
            outs.append(h)
        return outs

    def forward_b200(self, X):
        """Frozen-weight VGG16 pass on the library's own kernels: channels-last bf16 activations, tcgen05
        implicit-GEMM convs with the ReLU in their epilogue (thin-input kernel for the 3->64 stem), own 2x2 max-pool; in backward
        every ReLU gate rides in a data-gradient epilogue or in the pool backward, which also adds the LPIPS tap gradient
        (ops.ConvReluFn / ops.PoolTapFn) -- no elementwise or ATen pooling launches.  Same bf16 / fp32 rounding points as the
        autocast'ed cuDNN path.  X: (B, 3, H, W) fp32 or bf16; returns the five taps as NCHW-logical views."""
        packs = self.__dict__.setdefault("_packs", {})
        h = ops.to_channels_last(X)                                   # (B, H, W, 3) bf16
        outs = []
        x_is_relu = False                                             # the image / a pooled map: no ReLU gate on the way back
        for k in range(1, 6):
            convs = [(name, m) for name, m in getattr(self, f"slice{k}").named_children() if isinstance(m, nn.Conv2d)]
            if k > 1:                                                 # the slice starts with the pool of the previous tap
                tap, h = ops.pool_tap(h)
                outs.append(tap.permute(0, 3, 1, 2))
                x_is_relu = False
            for i, (name, m) in enumerate(convs):
                key = f"slice{k}.{name}"
                pack = packs.get(key)
                if pack is None:
                    pack = packs[key] = ops.WeightPack()
                last = k == 5 and i == len(convs) - 1                 # relu5_3: consumed by the LPIPS tap only, nobody gates for it
                h = ops.conv_relu(h, m.weight, m.bias, pack, x_is_relu, dy_premasked=not last)
                x_is_relu = True
        outs.append(h.permute(0, 3, 1, 2))                            # NCHW-logical views of the channels-last buffers
        return outs


class ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.Tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class NetLinLayer(nn.Module):
    """Holds the 1x1 conv weight under the reference's key (``model.1.weight`` when dropout is configured)."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else []
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False)]
        self.model = nn.Sequential(*layers)

    @property
    def weight(self):
        return self.model[-1].weight


class LPIPS(nn.Module):
    def __init__(self, ckpt_path=None, use_dropout=True, pretrained_vgg=True, faithful=None):
        """faithful: reproduce the bf16 roundings autocast puts around the 1x1 lin conv and the bf16 tail
        (None = do so exactly when autocast(bf16) is active, as the reference run would).
        The VGG16 trunk always runs on this library's tcgen05 conv tiles (bf16 operands, fp32 accumulation -- the arithmetic of
        the reference's autocast run, train_dmd.py:516); there is no cuDNN or CPU variant of the forward pass."""
        super().__init__()
        self.scaling_layer = ScalingLayer()
        self.chns = [64, 128, 256, 512, 512]
        self.net = vgg16(pretrained=pretrained_vgg, requires_grad=False)
        for k, c in enumerate(self.chns):
            setattr(self, f"lin{k}", NetLinLayer(c, use_dropout=use_dropout))
        self.faithful = faithful
        if ckpt_path is not None:
            self.load_from_pretrained(ckpt_path)
        for p in self.parameters():
            p.requires_grad = False

    def load_from_pretrained(self, ckpt_path=None, name="vgg_lpips"):
        if not os.path.exists(ckpt_path):
            warnings.warn(f"LPIPS: {ckpt_path} not found; lin weights stay at their random init")
            return
        self.load_state_dict(torch.load(ckpt_path, map_location=torch.device("cpu"), weights_only=True), strict=False)

    def forward(self, input, target):
        """LPIPS(input, target) -> 0-d tensor; like the reference call LPIPS(images, recon) the gradient flows to
        ``target`` (utils/lpips.py:81-94)."""
        if self.training:
            raise RuntimeError("LPIPS is an eval-only module here (reference: LPIPS(...).eval(), train_dmd.py:189)")
        faithful = self.faithful
        if faithful is None:
            faithful = torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16
        if not input.is_cuda:
            raise DmvaeError(f"LPIPS: input is on {input.device}; dmvae_b200 has no CPU path")
        if any(p.requires_grad for p in self.net.parameters()):
            raise DmvaeError("LPIPS: the VGG16 trunk is frozen on this path (utils/lpips.py:119-121 requires_grad=False)")
        with torch.autocast("cuda", enabled=False):
            with torch.no_grad():
                f0 = self.net.forward_b200(self.scaling_layer(input.float()))
            f1 = self.net.forward_b200(self.scaling_layer(target.float()))
        val = None
        for k in range(5):
            w = getattr(self, f"lin{k}").weight
            d = lpips_tap_distance(f0[k], f1[k], w, faithful)            # (B,) spatial means
            if faithful:
                d = d.to(torch.bfloat16)
            val = d if val is None else val + d
        return val.mean()
