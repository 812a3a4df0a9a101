#!/usr/bin/env python
"""CHECKER / CONTEXT code, not the product: the train_tokenizer.py step (frozen DINOv2 ViT -> bottleneck MLP -> flux Decoder,
L1 + LPIPS(VGG16), clip 1.0, AdamW) executed with STOCK PyTorch ops on the GPU, built from the oracle's functional restatement
(oracle/dmvae_oracle.py) of the reference's modules.  Two precisions:

  mode="autocast"  the way the reference itself executes on a GPU: torch.autocast(bf16) around cuDNN convs, ATen GroupNorm /
                   SiLU (fp32 by autocast policy), SDPA-free explicit attention, the LPIPS tail written op by op like
                   utils/lpips.py:81-94 + :156-161 (so autocast puts its bf16 roundings around the 1x1 `lin` conv and the
                   bf16 mean / sum tail exactly where the reference has them), torch.optim.AdamW.
                   = the CONTROL arm of the loss-parity run and the `gpu_baseline` of bench.py (cudnn.benchmark on).
  mode="fp32"      the same graph in strict fp32 (TF32 off): the exact-arithmetic anchor both bf16 pipelines are measured against.

Used by scripts/loss_parity.py, bench.py (`loss_parity`, `gpu_baseline` legs) and tests/.  Nothing under dmvae_b200/ imports it."""
from __future__ import annotations

import copy
import os
import sys
from contextlib import nullcontext
from typing import Dict, List

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import dmvae_oracle as O  # noqa: E402  (checker only)


def _vgg(lp: Dict[str, torch.Tensor], x: torch.Tensor) -> List[torch.Tensor]:
    """ScalingLayer + VGG16 slices (utils/lpips.py:97-104,116-153) with plain F.conv2d / relu / max_pool2d."""
    shift = torch.tensor(O.LPIPS_SHIFT, device=x.device).view(1, 3, 1, 1)
    scale = torch.tensor(O.LPIPS_SCALE, device=x.device).view(1, 3, 1, 1)
    h = (x - shift) / scale
    feats = []
    slice_of = lambda i: 1 + sum(i > t for t in (3, 8, 15, 22))   # noqa: E731
    for i in range(30):
        if i in O.VGG_CONVS:
            k = f"net.slice{slice_of(i)}.{i}"
            h = F.relu(F.conv2d(h, lp[k + ".weight"], lp[k + ".bias"], padding=1))
        elif i in O.VGG_POOLS:
            h = F.max_pool2d(h, 2, 2)
        if i in O.VGG_TAPS:
            feats.append(h)
    return feats


def _lpips_tail(lp: Dict[str, torch.Tensor], f0s, f1s) -> torch.Tensor:
    """utils/lpips.py:86-94 op by op (normalize_tensor :156-158, 1x1 lin conv, spatial_average :161, running sum, batch mean), so
    that under autocast every op gets the dtype the reference run gives it."""
    val = None
    for k, (f0, f1) in enumerate(zip(f0s, f1s)):
        n0 = f0 / (torch.sqrt(torch.sum(f0 ** 2, dim=1, keepdim=True)) + 1e-10)
        n1 = f1 / (torch.sqrt(torch.sum(f1 ** 2, dim=1, keepdim=True)) + 1e-10)
        d = (n0 - n1) ** 2
        res = F.conv2d(d, lp[f"lin{k}.model.1.weight"]).mean([2, 3], keepdim=True)
        val = res if val is None else val + res
    return val.mean()


class StockStep:
    """One arm.  ``decoder_sd`` / ``lpips_sd``: reference-keyed state dicts; ``encoder`` / ``mlp``: nn.Modules (stock PyTorch)."""

    def __init__(self, encoder, mlp, decoder_sd, lpips_sd, mode: str, lr: float = 1e-4, micro: int = 0):
        assert mode in ("autocast", "fp32")
        self.mode = mode
        self.enc = copy.deepcopy(encoder).eval()
        for p in self.enc.parameters():
            p.requires_grad = False
        self.mlp = copy.deepcopy(mlp)
        self.sd = {k: v.detach().clone().float().requires_grad_(True) for k, v in decoder_sd.items()}
        self.lp = {k: v.detach().clone().float() for k, v in lpips_sd.items()}
        self.params = list(self.sd.values()) + list(self.mlp.parameters())
        self.opt = torch.optim.AdamW(self.params, lr=lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
        self.micro = micro

    def _ctx(self):
        return torch.autocast("cuda", dtype=torch.bfloat16) if self.mode == "autocast" else nullcontext()

    def _loss(self, x: torch.Tensor) -> torch.Tensor:
        with self._ctx():
            # grad mode stays ON for the frozen encoder pass so that dmvae_b200's fused no-grad glue kernels are NOT taken: this
            # arm must be stock ATen / cuBLAS / SDPA end to end
            tok = self.enc(x).detach()
            z = self.mlp(tok)
            rec = O.decoder_forward(self.sd, z, bf16=False).float()          # plain torch ops; autocast (if on) picks bf16 convs
            l1 = F.l1_loss(rec, x)
            with torch.no_grad():
                f0 = _vgg(self.lp, x)
            f1 = _vgg(self.lp, rec)
            return l1 + _lpips_tail(self.lp, f0, f1)

    def step(self, x: torch.Tensor) -> torch.Tensor:
        """x: the GLOBAL batch of the step; processed in micro-batches of ``micro`` images (gradient accumulation = the mean over
        equal shards that DDP computes).  Returns the mean loss (0-d, device, fp32)."""
        self.opt.zero_grad(set_to_none=True)
        n = x.shape[0]
        m = self.micro if self.micro and self.micro < n else n
        assert n % m == 0
        total = torch.zeros((), device=x.device, dtype=torch.float32)
        for i in range(0, n, m):
            loss = self._loss(x[i:i + m])
            (loss * (m / n)).backward()
            total += loss.detach().float() * (m / n)
        torch.nn.utils.clip_grad_norm_(self.params, 1.0)
        self.opt.step()
        return total


def arms_from_vae(vae, lpips_module, modes=("fp32", "autocast"), lr: float = 1e-4, micro: int = 0) -> Dict[str, StockStep]:
    """Stock arms holding copies of the weights of a dmvae_b200 ``VAE`` (+ ``LPIPS``) at this moment."""
    dec_sd = {k: v for k, v in vae.decoder.state_dict().items()}
    lp_sd = {k: v for k, v in lpips_module.state_dict().items()}
    return {m: StockStep(vae.encoder, vae.bottle_neck, dec_sd, lp_sd, m, lr=lr, micro=micro) for m in modes}
