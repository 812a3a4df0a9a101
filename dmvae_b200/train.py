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
            vT_u = vS_u = None
            pair_T = getattr(self.base_model, "forward_cond_uncond", None)
            pair_S = getattr(self.sit_wo_ddp, "forward_cond_uncond", None)
            if a.dmd_cfg_scale > 1 and pair_T is not None and pair_S is not None:
                # conditional + unconditional rows in one batched pass per network (2 forwards instead of 4)
                v_teacher, vT_u = pair_T(xt, t, labels)
                v_student, vS_u = pair_S(xt, t, labels)
            else:
                v_teacher = self.base_model(xt, t, labels)
                v_student = self.sit_wo_ddp(xt, t, labels)
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


from .train_arena import GradArena  # noqa: E402,F401  (re-exported)


def dit_training_loss(model: Callable, latents: torch.Tensor, labels: torch.Tensor, time_dist_shift: float = 1.0,
                      cpu_generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Transport.training_losses for the linear path / velocity prediction (diffusion/transport/transport.py:119-142):
    xt = t*x1 + (1-t)*x0, target ut = x1 - x0, loss = mean over batch of the per-sample mean squared error."""
    t, x0 = sample_t_x0(latents, time_dist_shift, cpu_generator=cpu_generator)
    xt = losses.dmd_mix_xt(latents, x0, t)
    ut = latents - x0
    out = model(xt, t, labels)
    return ((out - ut) ** 2).flatten(1).mean(1).mean()


@torch.no_grad()
def update_ema(ema_params: List[torch.Tensor], params: List[torch.Tensor], decay: float = 0.9999):
    """train_tokenizer.py:140-150 as two foreach passes."""
    torch._foreach_mul_(ema_params, decay)
    torch._foreach_add_(ema_params, params, alpha=1 - decay)


class TokenizerTrainer:
    """One VAE-pretrain step of train_tokenizer.py:403-437 (frozen encoder, recon losses, clip, AdamW, EMA)."""

    def __init__(self, vae: nn.Module, loss_fn: VAELossFunction, lr: float = 1e-4, wd: float = 0.0, ema: bool = True,
                 clip: float = 1.0, fused_optimizer: bool = True):
        self.vae = vae
        self.loss_fn = loss_fn
        self.clip = clip
        self.arena = GradArena(vae.parameters())
        self.params = self.arena.params
        self.fused = None
        if fused_optimizer and self.params[0].is_cuda:
            from .optim import FlatAdamWEMA
            self.fused = FlatAdamWEMA(self.params, lr=lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd, max_norm=clip,
                                      ema_decay=0.9999 if ema else None, arena=self.arena)
            self.opt, self.ema = None, None
            return
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=wd, betas=(0.9, 0.95), eps=1e-8, fused=self.params[0].is_cuda)
        self.ema = [p.detach().clone() for p in self.params] if ema else None

    def _forward_backward(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.arena.zero()
        with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16):
            recon = self.vae(images, freeze_encoder=True)
            loss, log = self.loss_fn.forward_generator(images, recon)
        with self.arena.direct():                   # conv / GroupNorm gradients accumulate straight into the arena
            loss.backward()
        log["loss"] = loss.detach()
        return log

    def capture_cuda_graph(self, example_images: torch.Tensor, warmup: int = 2) -> bool:
        """Capture forward + backward of one step (≈1100 kernel launches at fixed shapes) into a CUDA graph; ``step`` then copies
        the batch into the graph's input buffer and replays it, leaving only the gradient exchange and the two optimizer
        kernels (whose step count / learning rate are host-side scalars) to be issued from Python.  With several ranks the NCCL
        exchange is issued after the replay (0.5 ms for 207 MB over NVLink, not overlapped).  Returns False -- and stays eager --
        if capture fails."""
        if self.fused is None or not example_images.is_cuda:
            return False
        self._graph = None
        self.arena.hooks_enabled = False            # nothing may enqueue a collective from inside the captured backward
        try:
            self._gx = example_images.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):           # warm up off the default stream: kernel attributes, tensor maps, cuDNN / cuBLAS plans
                for _ in range(max(1, warmup)):
                    self._forward_backward(self._gx)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            # the bf16 operand packs are refreshed only when a parameter's version counter moved: make sure the refresh is part
            # of what gets captured, so that every replay re-packs the weights the optimizer has just updated
            torch.autograd.graph.increment_version(self.params)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._glog = self._forward_backward(self._gx)
            self._graph = graph
            return True
        except Exception as e:                      # noqa: BLE001 -- any capture problem means: keep the eager path
            import warnings
            warnings.warn(f"TokenizerTrainer: CUDA graph capture failed ({type(e).__name__}: {e}); staying eager")
            self._graph = None
            self.arena.hooks_enabled = True
            torch.cuda.synchronize()
            return False

    def step(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        if getattr(self, "_graph", None) is not None and images.shape == self._gx.shape:
            self._gx.copy_(images, non_blocking=True)
            self._graph.replay()
            log = dict(self._glog)
        else:
            log = self._forward_backward(images)
        loss = log["loss"]
        self.arena.allreduce()
        if self.fused is not None:                  # clip + AdamW + EMA in two kernels over the flat arenas
            log["vae_norm"] = self.fused.step()
            return log
        # clip_grad_norm_(params, 1.0) (train_tokenizer.py:415) on the flat arena: one norm, one scale
        total = torch.linalg.vector_norm(self.arena.flat, 2)
        self.arena.flat.mul_(torch.clamp(self.clip / (total + 1e-6), max=1.0))
        log["vae_norm"] = total
        self.opt.step()
        if self.ema is not None:
            update_ema(self.ema, [p.data for p in self.params])
        log["loss"] = loss.detach()
        return log


class DmdTrainer:
    """One train_dmd.py iteration (:506-575) without the discriminator: the VAE turn (whole VAE trainable incl. the encoder,
    :519; recon + LPIPS + dmd_weight * DMD; clip; AdamW) followed by the student-DiT flow-matching step (:563-575)."""

    def __init__(self, vae: nn.Module, sit: nn.Module, base_model: nn.Module, lpips_loss: Optional[nn.Module], cfg: LossConfig,
                 lr_vae: float = 2e-5, lr_sit: float = 2e-5, latent_mean: float = 0.0, latent_scale: float = 1.0,
                 wd: float = 0.005, fused_optimizer: bool = True):
        self.vae, self.sit, self.base = vae, sit, base_model
        self.cfg, self.latent_mean, self.latent_scale = cfg, latent_mean, latent_scale
        self.loss_fn = VAELossFunction(cfg, lpips_loss=lpips_loss, sit=sit, base_model=base_model)
        for p in base_model.parameters():
            p.requires_grad = False
        self.arena_vae = GradArena(vae.parameters())
        self.arena_sit = GradArena(sit.parameters())
        self.fused = fused_optimizer and self.arena_vae.params[0].is_cuda
        if self.fused:          # clip + AdamW in two kernels per network over the flat arenas (train_dmd.py has no EMA)
            from .optim import FlatAdamWEMA
            kw = dict(betas=(0.9, 0.95), eps=1e-8, weight_decay=wd, max_norm=1.0, ema_decay=None)
            self.opt_vae = FlatAdamWEMA(self.arena_vae.params, lr=lr_vae, arena=self.arena_vae, **kw)
            self.opt_sit = FlatAdamWEMA(self.arena_sit.params, lr=lr_sit, arena=self.arena_sit, **kw)
        else:
            self.opt_vae = torch.optim.AdamW(self.arena_vae.params, lr=lr_vae, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd)
            self.opt_sit = torch.optim.AdamW(self.arena_sit.params, lr=lr_sit, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd)

    def _clip_step(self, arena: GradArena, opt, max_norm: float = 1.0) -> torch.Tensor:
        """clip_grad_norm_(params, 1.0); optimizer.step() (train_dmd.py:540-542, :568-570)."""
        if self.fused:
            return opt.step()
        total = torch.linalg.vector_norm(arena.flat, 2)
        arena.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        opt.step()
        return total

    def step(self, images: torch.Tensor, labels: torch.Tensor, vae_turn: bool = True) -> Dict[str, torch.Tensor]:
        dev = images.device.type
        log: Dict[str, torch.Tensor] = {}
        self.arena_vae.zero()
        with torch.autocast(device_type=dev, dtype=torch.bfloat16):
            if vae_turn:
                recon, z = self.vae(images, return_latent=True)
            else:
                with torch.no_grad():
                    z = self.vae.encode(images)
            latents = latents_to_spatial((z - self.latent_mean) * self.latent_scale)
            if vae_turn:
                self.sit.eval()
                for p in self.sit.parameters():
                    p.requires_grad = False
                loss, log = self.loss_fn.forward_generator(images, recon, latents, labels, compute_dmd=self.cfg.dmd_weight > 0)
        if vae_turn:
            with self.arena_vae.direct():
                loss.backward()
            self.arena_vae.allreduce()
            log["vae_norm"] = self._clip_step(self.arena_vae, self.opt_vae)
            log["loss"] = loss.detach()
        # 2. train the student DiT on the (detached) latents
        for p in self.sit.parameters():
            p.requires_grad = True
        self.sit.train()
        self.arena_sit.zero()
        with torch.autocast(device_type=dev, dtype=torch.bfloat16):
            dloss = dit_training_loss(self.sit, latents.detach(), labels, self.cfg.time_dist_shift)
        dloss.backward()
        self.arena_sit.allreduce()
        log["sit_norm"] = self._clip_step(self.arena_sit, self.opt_sit)
        log["diffusion_loss"] = dloss.detach()
        return log
