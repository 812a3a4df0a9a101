"""Host-side training step for the hot path: the reference's ``VAELossFunction`` surface (train_dmd.py:169-262,
train_tokenizer.py:153-199) on the fused kernels, a flat gradient arena with one NCCL allreduce per step (the DDP
exchange of train_dmd.py:348 / train_tokenizer.py:302), and a small trainer used by bench.py and the tests.

Discriminator / GAN branches are out of scope (stock PyTorch in the reference, not on this path).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as tdist
from torch import nn

from . import losses
from .vae import latents_to_spatial


# ------------------------------------------------------------------------------------------------ transport glue
def sample_t_x0(x1: torch.Tensor, time_dist_shift: float = 1.0, t0: float = 0.0, t1: float = 1.0,
                cpu_generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Transport.sample (diffusion/transport/transport.py:105-116) for the linear / velocity plan:
    x0 ~ N(0, I) from the device generator, t ~ U(t0, t1) drawn on the CPU generator then moved to x1's device and
    dtype (so t is bf16 when the latents are), followed by the time-distribution shift."""
    x0 = torch.randn_like(x1)
    t = torch.rand((x1.shape[0],), generator=cpu_generator) * (t1 - t0) + t0
    t = t.to(x1)
    t = 1 - time_dist_shift * (1 - t) / (1 + (time_dist_shift - 1) * (1 - t))
    return t, x0


@dataclass
class LossConfig:
    """Field names follow the reference's Tap args (train_dmd.py:26-100)."""
    l1: float = 1.0
    l2: float = 0.0
    lpips: float = 1.0
    dmd_weight: float = 10.0
    dmd_cfg_scale: float = 5.0
    num_classes: int = 1000
    t0: float = 0.0
    t1: float = 1.0
    time_dist_shift: float = 1.0


class VAELossFunction:
    """Generator-side losses of the reference's VAELossFunction.  ``base_model`` (teacher, s_real) and ``sit``
    (student, s_fake) are black-box callables ``v = model(xt, t, labels)`` (LightningDiT in the reference)."""

    def __init__(self, args: LossConfig, lpips_loss: Optional[nn.Module] = None, sit: Optional[Callable] = None,
                 base_model: Optional[Callable] = None):
        self.args = args
        self.lpips_loss = lpips_loss
        self.l1, self.l2, self.lpips, self.dmd_weight = args.l1, args.l2, args.lpips, args.dmd_weight
        self.sit_wo_ddp = sit
        self.base_model = base_model

    def compute_distribution_matching_loss(self, latents_norm: torch.Tensor, labels: torch.Tensor, step: int = 0,
                                           cpu_generator: Optional[torch.Generator] = None,
                                           t: Optional[torch.Tensor] = None, x0: Optional[torch.Tensor] = None):
        """train_dmd.py:204-230.  Returns (loss, log) with device-side scalars in ``log`` (no host sync here).
        ``t`` / ``x0`` may be injected (what ``self.transport.sample`` returns in the reference) for reproducible tests."""
        a = self.args
        if t is None or x0 is None:
            t, x0 = sample_t_x0(latents_norm, a.time_dist_shift, cpu_generator=cpu_generator)
        t = t * (a.t1 - a.t0) + a.t0
        xt = losses.dmd_mix_xt(latents_norm, x0, t)
        with torch.no_grad():
            v_teacher = self.base_model(xt, t, labels)
            v_student = self.sit_wo_ddp(xt, t, labels)
            vT_u = vS_u = None
            if a.dmd_cfg_scale > 1:
                uncond = torch.ones_like(labels) * a.num_classes
                vT_u = self.base_model(xt, t, uncond)
                vS_u = self.sit_wo_ddp(xt, t, uncond)
        loss, gnorm = losses.dmd_loss(latents_norm, xt, t, v_teacher, v_student, vT_u, vS_u, a.dmd_cfg_scale, True)
        return loss, {"dmd_loss": loss.detach(), "dmd_gradient_norm": gnorm}

    def forward_generator(self, images_pm1, recon_image, latents=None, labels=None, compute_dmd=False, step=0):
        """train_dmd.py:233-262 without the discriminator branch."""
        l1, l2 = losses.l1_l2_loss(recon_image, images_pm1)
        rec_loss = l1 * self.l1 + l2 * self.l2
        log = {"L1": l1.detach(), "L2": l2.detach()}
        if self.lpips_loss is not None and self.lpips != 0:
            lp = self.lpips_loss(images_pm1, recon_image).mean()
            rec_loss = rec_loss + lp * self.lpips
            log["LPIPS"] = lp.detach()
        log["rec_loss"] = rec_loss.detach()
        if compute_dmd:
            dmd, dmd_log = self.compute_distribution_matching_loss(latents, labels, step)
            log.update(dmd_log)
            rec_loss = rec_loss + dmd * self.dmd_weight
        return rec_loss, log


# ------------------------------------------------------------------------------------------------ data parallel
class GradArena:
    """All trainable gradients as views of one flat fp32 buffer, exchanged with one NCCL all-reduce per chunk over
    NVLink/NVSwitch instead of DDP's 25 MB buckets.  With ``overlap=True`` a chunk's all-reduce is issued from an
    autograd hook as soon as every gradient inside it has been accumulated (backward runs output-to-input, so the
    tail of the arena completes first), i.e. the exchange overlaps the rest of the backward pass like DDP's bucketed
    reduction (train_dmd.py:348).  With world_size == 1 it is just a flat buffer."""

    def __init__(self, params: Iterable[nn.Parameter], chunks: int = 4, overlap: bool = True):
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.chunks = max(1, min(chunks, len(self.params) or 1))
        # chunk boundaries on parameter boundaries, roughly equal in bytes
        target = (n + self.chunks - 1) // self.chunks
        self.bounds: List[Tuple[int, int]] = []          # [start, end) element ranges
        self.chunk_of: Dict[int, int] = {}
        start, off = 0, 0
        for i, p in enumerate(self.params):
            self.chunk_of[id(p)] = len(self.bounds)
            off += p.numel()
            if off - start >= target or i == len(self.params) - 1:
                self.bounds.append((start, off))
                start = off
        self.members = [sum(1 for p in self.params if self.chunk_of[id(p)] == c) for c in range(len(self.bounds))]
        self._pending = list(self.members)
        self._handles: List = []
        self._launched = [False] * len(self.bounds)
        self._attach()
        self.overlap = overlap and self._distributed()
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad_ready)

    @staticmethod
    def _distributed() -> bool:
        return tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1

    def _attach(self):
        off = 0
        for p in self.params:      # (re-)attach in case an optimizer dropped the views (set_to_none)
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()
        self._attach()
        self._pending = list(self.members)
        self._launched = [False] * len(self.bounds)
        self._handles = []

    def _launch(self, c: int):
        s, e = self.bounds[c]
        self._launched[c] = True
        self._handles.append(tdist.all_reduce(self.flat[s:e], op=tdist.ReduceOp.SUM, async_op=True))

    def _on_grad_ready(self, p: nn.Parameter):
        c = self.chunk_of[id(p)]
        self._pending[c] -= 1
        if self._pending[c] == 0 and not self._launched[c]:
            self._launch(c)

    def allreduce(self):
        """Finish the exchange: launch whatever was not launched from hooks, wait, average."""
        if not self._distributed():
            return
        for c in range(len(self.bounds)):
            if not self._launched[c]:
                self._launch(c)
        for h in self._handles:
            h.wait()
        self._handles = []
        self.flat.div_(tdist.get_world_size())


@torch.no_grad()
def update_ema(ema_params: List[torch.Tensor], params: List[torch.Tensor], decay: float = 0.9999):
    """train_tokenizer.py:140-150 as two foreach passes."""
    torch._foreach_mul_(ema_params, decay)
    torch._foreach_add_(ema_params, params, alpha=1 - decay)


class TokenizerTrainer:
    """One VAE-pretrain step of train_tokenizer.py:403-437 (frozen encoder, recon losses, clip, AdamW, EMA)."""

    def __init__(self, vae: nn.Module, loss_fn: VAELossFunction, lr: float = 1e-4, wd: float = 0.0, ema: bool = True,
                 clip: float = 1.0):
        self.vae = vae
        self.loss_fn = loss_fn
        self.clip = clip
        self.arena = GradArena(vae.parameters())
        self.params = self.arena.params
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=wd, betas=(0.9, 0.95), eps=1e-8, fused=self.params[0].is_cuda)
        self.ema = [p.detach().clone() for p in self.params] if ema else None

    def step(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.arena.zero()
        with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16):
            recon = self.vae(images, freeze_encoder=True)
            loss, log = self.loss_fn.forward_generator(images, recon)
        loss.backward()
        self.arena.allreduce()
        # clip_grad_norm_(params, 1.0) (train_tokenizer.py:415) on the flat arena: one norm, one scale
        total = torch.linalg.vector_norm(self.arena.flat, 2)
        self.arena.flat.mul_(torch.clamp(self.clip / (total + 1e-6), max=1.0))
        log["vae_norm"] = total
        self.opt.step()
        if self.ema is not None:
            update_ema(self.ema, [p.data for p in self.params])
        log["loss"] = loss.detach()
        return log
