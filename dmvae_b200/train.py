"""Host-side training step for the hot path: the reference's ``VAELossFunction`` surface (train_dmd.py:169-262,
train_tokenizer.py:153-199) on the fused kernels, a flat gradient arena with one NCCL allreduce per step (the DDP
exchange of train_dmd.py:348 / train_tokenizer.py:302), and the two trainers used by bench.py and the tests.

Discriminator / GAN branches are out of scope (stock PyTorch in the reference, not on this path).
There is no PyTorch-optimizer or CPU variant of the trainers: they need CUDA parameters and the library's fused
clip + AdamW (+ EMA) kernels (``optim.FlatAdamWEMA``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import losses
from .vae import latents_to_spatial


# ------------------------------------------------------------------------------------------------ transport glue
def sample_t(batch: int, like: torch.Tensor, t0: float = 0.0, t1: float = 1.0,
             cpu_generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """t ~ U(t0, t1) drawn on the CPU generator, then moved to ``like``'s device and dtype -- so t is bf16 when the latents are
    (diffusion/transport/transport.py:111-114)."""
    return (torch.rand((batch,), generator=cpu_generator) * (t1 - t0) + t0).to(like)


def shift_t(t: torch.Tensor, time_dist_shift: float) -> torch.Tensor:
    """The time-distribution shift of Transport.sample (transport.py:115)."""
    return 1 - time_dist_shift * (1 - t) / (1 + (time_dist_shift - 1) * (1 - t))


def sample_t_x0(x1: torch.Tensor, time_dist_shift: float = 1.0, t0: float = 0.0, t1: float = 1.0,
                cpu_generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Transport.sample (diffusion/transport/transport.py:105-116) for the linear / velocity plan:
    x0 ~ N(0, I) from the device generator, t ~ U(t0, t1) drawn on the CPU generator then moved to x1's device and
    dtype, followed by the time-distribution shift."""
    x0 = torch.randn_like(x1)
    return shift_t(sample_t(x1.shape[0], x1, t0, t1, cpu_generator), time_dist_shift), x0


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
    disc_weight: float = 0.0
    disc_start_step: int = 0
    bcr: float = 0.0


class VAELossFunction:
    """The reference's VAELossFunction (train_dmd.py:169-285).  ``base_model`` (teacher, s_real) and ``sit`` (student, s_fake)
    are black-box callables ``v = model(xt, t, labels)`` (LightningDiT in the reference); ``disc`` is a black-box discriminator
    ``logits = disc(images)`` (DinoDisc / PatchGAN in the reference: stock PyTorch, out of scope for kernels -- SURVEY section 2),
    ``last_layer`` the tensor the adaptive GAN weight is measured on (``vae.decoder.get_last_layer()``, :247), ``aug`` /
    ``bcr_aug`` the differentiable augmentations applied to the discriminator input (DiffAug, :199-200; identity when None)."""

    def __init__(self, args: LossConfig, lpips_loss: Optional[nn.Module] = None, sit: Optional[Callable] = None,
                 base_model: Optional[Callable] = None, disc: Optional[nn.Module] = None, last_layer: Optional[torch.Tensor] = None,
                 aug: Optional[Callable] = None, bcr_aug: Optional[Callable] = None):
        self.args = args
        self.lpips_loss = lpips_loss
        self.l1, self.l2, self.lpips, self.dmd_weight = args.l1, args.l2, args.lpips, args.dmd_weight
        self.disc_weight, self.bcr_weight = args.disc_weight, args.bcr
        self.sit_wo_ddp = sit
        self.base_model = base_model
        self.disc, self.last_layer = disc, last_layer
        self.aug = aug if aug is not None else (lambda x, fade=0.0: x)
        self.bcr_aug = bcr_aug if bcr_aug is not None else (lambda x, fade=0.0: x)

    def compute_distribution_matching_loss(self, latents_norm: torch.Tensor, labels: torch.Tensor, step: int = 0,
                                           cpu_generator: Optional[torch.Generator] = None,
                                           t: Optional[torch.Tensor] = None, x0: Optional[torch.Tensor] = None):
        """train_dmd.py:204-230.  Returns (loss, log) with device-side scalars in ``log`` (no host sync here).
        ``t`` / ``x0`` may be injected (what ``self.transport.sample`` returns in the reference): tests use it for
        reproducibility, the CUDA-graph trainer to feed the CPU-drawn t through a static device buffer."""
        a = self.args
        if x0 is None:
            x0 = torch.randn_like(latents_norm)
        if t is None:
            t = shift_t(sample_t(latents_norm.shape[0], latents_norm, cpu_generator=cpu_generator), a.time_dist_shift)
        t = t * (a.t1 - a.t0) + a.t0
        if isinstance(self.base_model, nn.Module) and self.base_model.training:
            raise RuntimeError("DMD teacher must be in eval mode (train_dmd.py:372): a train-mode LightningDiT drops labels at random")
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

    def forward_generator(self, images_pm1, recon_image, latents=None, labels=None, compute_dmd=False, step=0, t=None):
        """train_dmd.py:233-262.  Log values are device scalars (the reference calls .item() on each: six host syncs per step)."""
        l1, l2 = losses.l1_l2_loss(recon_image, images_pm1)
        rec_loss = l1 * self.l1 + l2 * self.l2
        log = {"L1": l1.detach(), "L2": l2.detach()}
        if self.lpips_loss is not None and self.lpips != 0:
            lp = self.lpips_loss(images_pm1, recon_image).mean()
            rec_loss = rec_loss + lp * self.lpips
            log["LPIPS"] = lp.detach()
        log["rec_loss"] = rec_loss.detach()
        if self.disc is not None and self.disc_weight > 0 and step >= self.args.disc_start_step:
            # :244-257 -- adaptive GAN weight: ratio of the two losses' gradient norms at the decoder's last layer.  The two partial
            # autograd.grad calls run through the custom Functions with retain_graph=True (saved tensors are never written in
            # place) and outside GradArena.direct(), so they return tensors instead of accumulating into the arena.
            self.disc.eval()
            for p in self.disc.parameters():
                p.requires_grad = False
            d_loss = -self.disc(self.aug(recon_image, 0)).mean()
            g_rec = torch.autograd.grad(rec_loss, self.last_layer, retain_graph=True)[0]
            g_gan = torch.autograd.grad(d_loss, self.last_layer, retain_graph=True)[0]
            w = (g_rec.detach().norm() / g_gan.detach().norm().add_(1e-6)).clamp_(0.0, 1e4)
            d_weight = self.disc_weight * w
            rec_loss = rec_loss + d_loss * d_weight
            log["d_weight"] = d_weight.detach()
        if compute_dmd:
            dmd, dmd_log = self.compute_distribution_matching_loss(latents, labels, step, t=t)
            log.update(dmd_log)
            rec_loss = rec_loss + dmd * self.dmd_weight
        return rec_loss, log


    def forward_discriminator(self, images_pm1, recon_image, fade_blur_schedule=0.0):
        """train_dmd.py:265-285: hinge loss on real / reconstructed images + balanced consistency regularisation."""
        for p in self.disc.parameters():
            p.requires_grad = True
        self.disc.train()
        bs = images_pm1.size(0)
        both = torch.cat([images_pm1, recon_image], dim=0)
        logits = self.disc(self.aug(both, fade_blur_schedule)).float()
        logits_real, logits_fake = logits[:bs], logits[bs:]
        d_loss = 0.5 * (torch.relu(1.0 - logits_real).mean() + torch.relu(1.0 + logits_fake).mean())
        log = {"d_loss": d_loss.detach(), "acc_real": (logits_real.detach() > 0).float().mean() * 100,
               "acc_fake": (logits_fake.detach() < 0).float().mean() * 100}
        if self.bcr_weight > 0:
            logits2 = self.disc(self.bcr_aug(both, 0.0)).float()
            lbcr = nn.functional.mse_loss(logits2, logits) * self.bcr_weight
            log["bcr_loss"] = lbcr.detach()
            d_loss = d_loss + lbcr
        return d_loss, log


from .train_arena import GradArena  # noqa: E402,F401  (re-exported)


def dit_training_loss(model: Callable, latents: torch.Tensor, labels: torch.Tensor, time_dist_shift: float = 1.0,
                      cpu_generator: Optional[torch.Generator] = None, t: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Transport.training_losses for the linear path / velocity prediction (diffusion/transport/transport.py:119-142):
    xt = t*x1 + (1-t)*x0, target ut = x1 - x0, loss = mean over batch of the per-sample mean squared error.
    ``t`` (already shifted) may be injected, see compute_distribution_matching_loss."""
    x0 = torch.randn_like(latents)
    if t is None:
        t = shift_t(sample_t(latents.shape[0], latents, cpu_generator=cpu_generator), time_dist_shift)
    xt = losses.dmd_mix_xt(latents, x0, t)
    ut = latents - x0
    out = model(xt, t, labels)
    return ((out - ut) ** 2).flatten(1).mean(1).mean()


# ------------------------------------------------------------------------------------------------ CUDA-graph sections
class GraphCaptureError(RuntimeError):
    pass


class _GraphedSection:
    """``fn()`` (no arguments, fixed shapes, reads its inputs from static buffers) captured once into a CUDA graph.

    Capture protocol: warm up on a side stream (kernel attributes, tensor maps, cuBLAS / cuDNN plans, NCCL channels), bump
    the parameters' version counters so the bf16 operand re-packs are part of what gets captured, capture, then replay once so
    that every buffer the capture allocated in the graph's private pool holds real data (an eager pass right after the capture
    would otherwise see an unchanged version counter and read packs that were only *recorded*, never written)."""

    def __init__(self, fn: Callable[[], Dict[str, torch.Tensor]], params: List[nn.Parameter], warmup: int = 2,
                 optimizers: Tuple = ()):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if params:
            torch.autograd.graph.increment_version(params)
        for opt in optimizers:                          # ... with the optimizer-maintained bf16 operand copies marked current, so the
            opt.sync_w16()                              # capture records the cheap path (only the dgrad transposes are re-derived)
        self.graph = torch.cuda.CUDAGraph()
        # with a process group alive its watchdog thread polls CUDA events concurrently: only this thread's calls belong to the capture
        mode = "thread_local" if GradArena._distributed() else "global"
        with torch.cuda.graph(self.graph, capture_error_mode=mode):
            self.out = fn()
        self.graph.replay()
        torch.cuda.synchronize()

    def replay(self) -> Dict[str, torch.Tensor]:
        self.graph.replay()
        return dict(self.out)

    def release(self) -> None:
        """Destroy the graph (and with it the captured NCCL operations).  Must happen before the process group goes away."""
        self.graph = None
        self.out = {}


def shutdown_distributed(*trainers, timeout_s: float = 30.0) -> None:
    """Orderly end of a multi-rank run: release every captured CUDA graph FIRST (a graph that holds captured NCCL collectives
    keeps the communicator busy: destroying the process group while such graphs are alive dead-locks against ProcessGroupNCCL's
    watchdog), drain the device, then destroy the process group -- under a watchdog of our own, because a hang at exit would
    otherwise cost the NCCL timeout (10 minutes) on every rank."""
    import gc
    import os
    import threading
    import torch.distributed as tdist
    for tr in trainers:
        if tr is not None and hasattr(tr, "release_graphs"):
            tr.release_graphs()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if not (tdist.is_available() and tdist.is_initialized()):
        return

    def _bail():
        import sys
        sys.stderr.write("[dmvae_b200] process-group teardown did not finish in time; exiting without it\n")
        sys.stderr.flush()
        os._exit(0)
    timer = threading.Timer(timeout_s, _bail)
    timer.daemon = True
    timer.start()
    try:
        tdist.barrier()
        torch.cuda.synchronize()
        tdist.destroy_process_group()
    finally:
        timer.cancel()


def _capture_with_exchange(fn: Callable[[], Dict[str, torch.Tensor]], arenas: List[GradArena], params: List[nn.Parameter],
                           warmup: int, optimizers: Tuple = ()) -> Tuple[_GraphedSection, str]:
    """Capture ``fn`` (which ends with ``arena.allreduce()``) with the chunked NCCL all-reduces INSIDE the graph: the hooks fire
    during the captured backward, each collective lands on NCCL's stream as a forked branch of the graph and overlaps the
    rest of backward on every replay.  If the process group cannot be captured, fall back to a graph without the exchange
    (hooks off; the caller issues ``arena.allreduce()`` after each replay).  Returns (section, "in-graph" | "post-replay")."""
    distributed = GradArena._distributed()
    if not distributed:
        return _GraphedSection(fn, params, warmup, optimizers), "none"
    try:
        return _GraphedSection(fn, params, warmup, optimizers), "in-graph"
    except Exception as e:                              # noqa: BLE001
        import warnings
        warnings.warn(f"NCCL exchange could not be captured ({type(e).__name__}: {e}); capturing without it")
        torch.cuda.synchronize()
    for a in arenas:
        a.hooks_enabled = False
        a.begin_step()
    return _GraphedSection(fn, params, warmup, optimizers), "post-replay"


class _PinnedTimes:
    """The CPU-drawn diffusion times (transport.py:111-114: ``torch.rand`` on the CPU generator) on their way to a static device
    buffer without a host sync: a small ring of pinned fp32 buffers, each guarded by an event so that a buffer is only redrawn
    after the asynchronous copy that read it has completed (the host may run many replayed steps ahead of the device)."""

    def __init__(self, batch: int, depth: int = 8):
        self.bufs = [torch.empty(batch, dtype=torch.float32, pin_memory=True) for _ in range(depth)]
        self.events: List[Optional[torch.cuda.Event]] = [None] * depth
        self.i = 0

    def send(self, dst: torch.Tensor, cpu_generator: Optional[torch.Generator], t0: float = 0.0, t1: float = 1.0) -> None:
        k = self.i % len(self.bufs)
        self.i += 1
        if self.events[k] is not None:
            self.events[k].synchronize()
        buf = self.bufs[k]
        torch.rand(buf.shape, generator=cpu_generator, out=buf)
        if t0 != 0.0 or t1 != 1.0:
            buf.mul_(t1 - t0).add_(t0)
        dst.copy_(buf, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev


class TokenizerTrainer:
    """One VAE-pretrain step of train_tokenizer.py:403-437 (frozen encoder, recon losses, clip, AdamW, EMA).  When the loss
    function carries a discriminator (``loss_fn.disc``, BASELINE configs[3]) the step is the full GAN iteration: generator branch
    with the adaptive weight (:179-199), then the discriminator's hinge step on (images, recon.detach()) (:420-428) with its own
    arena / fused optimizer -- the discriminator itself is whatever stock-PyTorch module the caller hands in.

    **Pipelined exchange** (``pipelined=True``; the default with several ranks and no discriminator).  The encoder is frozen in
    this stage (train_tokenizer.py:295-297), so its forward pass does not depend on the optimizer update.  ``step(k)`` therefore
    runs  [all-reduce of step k-1's gradients on NCCL's stream  ||  encoder forward of batch k]  ->  clip + AdamW + EMA with
    those gradients  ->  bottleneck + decoder forward, losses, backward of batch k -- the same arithmetic in the same order as the
    sequential loop (the update still precedes the first use of the updated weights), but the exchange hides under the ViT's
    library GEMMs, which tolerate the SMs NCCL takes, instead of under the persistent one-CTA-per-SM conv kernels, which do not
    (a 148-CTA persistent grid that finds a few SMs busy finishes late by up to one NCCL kernel: measured +1.0 ms per step at 4
    GPUs with the exchange overlapped with backward).  The update of the LAST step is applied by ``flush()`` (called by
    ``weights_changed`` / ``release_graphs`` and to be called before reading weights, e.g. for a checkpoint)."""

    def __init__(self, vae: nn.Module, loss_fn: VAELossFunction, lr: float = 1e-4, wd: float = 0.0, ema: bool = True,
                 clip: float = 1.0, lr_disc: float = 1e-4, pipelined: Optional[bool] = None):
        from .optim import FlatAdamWEMA
        self.vae = vae
        self.loss_fn = loss_fn
        self.clip = clip
        self.arena = GradArena(vae.parameters())
        self.params = self.arena.params
        self.fused = FlatAdamWEMA(self.params, lr=lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd, max_norm=clip,
                                  ema_decay=0.9999 if ema else None, arena=self.arena)
        self.arena_disc = self.fused_disc = None
        self.global_step = 0
        if loss_fn.disc is not None and loss_fn.disc_weight > 0:
            for p in loss_fn.disc.parameters():
                p.requires_grad = True
            self.arena_disc = GradArena(loss_fn.disc.parameters())
            self.fused_disc = FlatAdamWEMA(self.arena_disc.params, lr=lr_disc, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd,
                                           max_norm=clip, ema_decay=None, arena=self.arena_disc)
            if loss_fn.last_layer is None:
                loss_fn.last_layer = vae.decoder.get_last_layer()
        self._section: Optional[_GraphedSection] = None
        self._exchange_outside = True       # eager: _forward_backward ends with the exchange itself
        self.exchange_mode = "eager"
        if pipelined is None:
            pipelined = GradArena._distributed() and self.arena_disc is None
        self.pipelined = bool(pipelined) and self.arena_disc is None
        self._pending = False               # pipelined mode: the previous step's gradients are in the arena, not yet applied
        self._sec_enc: Optional[_GraphedSection] = None
        self._sec_dec: Optional[_GraphedSection] = None
        if self.pipelined:
            self.arena.hooks_enabled = False            # the exchange is issued at the start of the NEXT step, never from backward
            self.exchange_mode = "pipelined (eager)"

    # ---- pipelined mode: the step in two pieces
    def _encode(self, images: torch.Tensor) -> torch.Tensor:
        with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16), torch.no_grad():
            return self.vae.encoder(images)             # models/vae.py:92-93 with freeze_encoder=True

    def _decode_backward(self, images: torch.Tensor, tokens: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.arena.zero()
        with self.arena.scratch():                  # small per-step accumulators come out of the arena's zero pool
            with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16):
                recon = self.vae.decoder(self.vae.bottle_neck(tokens)).float()      # models/vae.py:94-97
                loss, log = self.loss_fn.forward_generator(images, recon, step=self.global_step)
            with self.arena.direct():
                loss.backward()
        log["loss"] = loss.detach()
        return log

    def flush(self) -> Optional[torch.Tensor]:
        """Pipelined mode: exchange and apply the gradients of the last ``step()`` now.  Returns their (pre-clip) norm."""
        if not self._pending:
            return None
        self.arena.allreduce()
        self._pending = False
        return self.fused.step()

    def _forward_backward(self, images: torch.Tensor, exchange: bool = True) -> Dict[str, torch.Tensor]:
        self.arena.zero()
        with self.arena.scratch():                  # small per-step accumulators come out of the arena's zero pool
            with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16):
                recon = self.vae(images, freeze_encoder=True)
                loss, log = self.loss_fn.forward_generator(images, recon, step=self.global_step)
            with self.arena.direct():               # conv / GroupNorm gradients accumulate straight into the arena
                loss.backward()
        if exchange:
            self.arena.allreduce()                  # chunks not yet launched from the hooks, wait, (average inside NCCL)
        log["loss"] = loss.detach()
        if self.arena_disc is not None and self.global_step >= self.loss_fn.args.disc_start_step:
            self.arena_disc.zero()
            with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16):
                d_loss, d_log = self.loss_fn.forward_discriminator(images, recon.detach())
            d_loss.backward()
            if exchange:
                self.arena_disc.allreduce()
            log.update(d_log)
        return log

    def capture_cuda_graph(self, example_images: torch.Tensor, warmup: int = 2, strict: bool = False) -> bool:
        """Capture forward + backward + gradient exchange of one step (≈1100 kernel launches at fixed shapes) into a CUDA
        graph; ``step`` then copies the batch into the graph's input buffer and replays it, leaving only the two optimizer
        kernels (whose step count / learning rate are host-side scalars) to be issued from Python.  With several ranks the
        chunked NCCL all-reduces are captured too (forked onto NCCL's stream from the hooks, so they overlap the rest of
        backward in every replay); ``exchange_mode`` says whether that worked ("in-graph") or the exchange runs after each
        replay ("post-replay").  ``strict``: raise GraphCaptureError instead of returning False when capture fails."""
        self._section = None
        if self.pipelined:
            try:                                    # two graphs; the NCCL exchange stays an eager launch between them
                self.flush()
                self._gx = example_images.clone()
                self._gtok = self._encode(self._gx).clone()

                def enc():
                    self._gtok.copy_(self._encode(self._gx))
                    return {}
                self._sec_enc = _GraphedSection(enc, [], warmup)
                self._sec_dec = _GraphedSection(lambda: self._decode_backward(self._gx, self._gtok), self.params, warmup, (self.fused,))
                self.exchange_mode = "pipelined under the next step's encoder forward"
                return True
            except Exception as e:                  # noqa: BLE001
                self._sec_enc = self._sec_dec = None
                torch.cuda.synchronize()
                if strict:
                    raise GraphCaptureError(f"TokenizerTrainer: CUDA graph capture failed ({type(e).__name__}: {e})") from e
                import warnings
                warnings.warn(f"TokenizerTrainer: CUDA graph capture failed ({type(e).__name__}: {e}); staying eager")
                return False
        self.arena.hooks_enabled = True
        try:
            self._gx = example_images.clone()

            def body():                             # hooks off (fallback capture) = the exchange is issued after each replay
                return self._forward_backward(self._gx, exchange=self.arena.hooks_enabled)
            arenas = [self.arena] + ([self.arena_disc] if self.arena_disc is not None else [])
            self._section, self.exchange_mode = _capture_with_exchange(body, arenas, self.params, warmup, (self.fused,))
            self._exchange_outside = self.exchange_mode == "post-replay"
            return True
        except Exception as e:                      # noqa: BLE001
            self._section = None
            self.arena.hooks_enabled = True
            self.exchange_mode = "eager"
            torch.cuda.synchronize()
            if strict:
                raise GraphCaptureError(f"TokenizerTrainer: CUDA graph capture failed ({type(e).__name__}: {e})") from e
            import warnings
            warnings.warn(f"TokenizerTrainer: CUDA graph capture failed ({type(e).__name__}: {e}); staying eager")
            return False

    @property
    def graphed(self) -> bool:
        return self._section is not None or self._sec_dec is not None

    def release_graphs(self) -> None:
        self.flush()
        for sec in (self._section, self._sec_enc, self._sec_dec):
            if sec is not None:
                sec.release()
        self._section = self._sec_enc = self._sec_dec = None
        self._exchange_outside = True
        self.exchange_mode = "pipelined (eager)" if self.pipelined else "eager"
        self.arena.hooks_enabled = not self.pipelined

    def weights_changed(self) -> None:
        """Call after writing the trainable weights out of band (``load_state_dict``, ``load_pretrained``, manual edits) once the
        trainer exists: bumps the version counters and rebuilds the optimizer-maintained bf16 operand copies that a captured CUDA
        graph reads directly (an eager step would notice the stale version stamp by itself; a replayed graph cannot).  In pipelined
        mode call ``flush()`` BEFORE overwriting the weights: a still-pending update is dropped here, not applied on top of them."""
        if self._pending:
            self.arena.begin_step()
            self._pending = False
        torch.autograd.graph.increment_version(self.params)
        self.fused.sync_w16()

    def _step_pipelined(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        graphed = self._sec_dec is not None and images.shape == self._gx.shape
        if self._pending:
            self.arena.launch_all()                 # step k-1's gradients: on NCCL's stream, behind the backward already queued
        if graphed:
            self._gx.copy_(images, non_blocking=True)
            self._sec_enc.replay()                  # ... concurrently with this batch's (frozen) encoder forward
        else:
            tokens = self._encode(images)
        norm = None
        if self._pending:
            self.arena.finish()
            norm = self.fused.step()                # clip + AdamW + EMA with step k-1's exchanged gradients
        log = self._sec_dec.replay() if graphed else self._decode_backward(images, tokens)
        if norm is not None:
            log["vae_norm"] = norm                  # of the update applied at the start of this call (step k-1's gradients)
        self._pending = True
        self.global_step += 1
        return log

    def step(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        if self.pipelined:
            return self._step_pipelined(images)
        if self._section is not None and images.shape == self._gx.shape:
            self._gx.copy_(images, non_blocking=True)
            log = self._section.replay()
            if self._exchange_outside:
                self.arena.allreduce()
                if self.arena_disc is not None:
                    self.arena_disc.allreduce()
        else:
            log = self._forward_backward(images)
        log["vae_norm"] = self.fused.step()         # clip + AdamW + EMA in two kernels over the flat arenas
        if self.fused_disc is not None and "d_loss" in log:
            log["disc_norm"] = self.fused_disc.step()
        self.global_step += 1
        return log


class DmdTrainer:
    """One train_dmd.py iteration (:506-575) without the discriminator: the VAE turn (whole VAE trainable incl. the encoder,
    :519; recon + LPIPS + dmd_weight * DMD; clip; AdamW) followed by the student-DiT flow-matching step (:563-575).

    ``capture_cuda_graphs`` turns the iteration into three replayed graphs -- VAE turn (forward, 4 DiT forwards batched as 2,
    fused DMD loss, backward, exchange), encode-only (the other iterations, :522-524) and the student step -- that hand the
    latents to each other through a static buffer; the CPU-drawn diffusion times (transport.py:111-114) travel through two
    small static device buffers, and only the fused optimizer kernels are issued from Python."""

    def __init__(self, vae: nn.Module, sit: nn.Module, base_model: nn.Module, lpips_loss: Optional[nn.Module], cfg: LossConfig,
                 lr_vae: float = 2e-5, lr_sit: float = 2e-5, latent_mean: float = 0.0, latent_scale: float = 1.0,
                 wd: float = 0.005, cpu_generator: Optional[torch.Generator] = None):
        from .optim import FlatAdamWEMA
        self.vae, self.sit, self.base = vae, sit, base_model
        self.cfg, self.latent_mean, self.latent_scale = cfg, latent_mean, latent_scale
        self.loss_fn = VAELossFunction(cfg, lpips_loss=lpips_loss, sit=sit, base_model=base_model)
        self.cpu_generator = cpu_generator
        base_model.eval()                               # train_dmd.py:372
        for p in base_model.parameters():
            p.requires_grad = False
        self.arena_vae = GradArena(vae.parameters())
        self.arena_sit = GradArena(sit.parameters())
        # clip + AdamW in two kernels per network over the flat arenas (train_dmd.py has no EMA)
        kw = dict(betas=(0.9, 0.95), eps=1e-8, weight_decay=wd, max_norm=1.0, ema_decay=None)
        self.opt_vae = FlatAdamWEMA(self.arena_vae.params, lr=lr_vae, arena=self.arena_vae, **kw)
        self.opt_sit = FlatAdamWEMA(self.arena_sit.params, lr=lr_sit, arena=self.arena_sit, **kw)
        self._sections: Optional[Dict[str, _GraphedSection]] = None
        self.exchange_mode = "eager"

    # ---- the three pieces of an iteration (each runs eagerly or under capture)
    def _latents(self, z: torch.Tensor) -> torch.Tensor:
        return latents_to_spatial((z - self.latent_mean) * self.latent_scale)       # train_dmd.py:525-526

    def _vae_turn(self, images, labels, t_dmd, exchange: bool = True) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
        self.sit.eval()                                 # train_dmd.py:529-531: the student is a frozen scorer in the VAE turn
        for p in self.sit.parameters():
            p.requires_grad = False
        self.arena_vae.zero()
        with self.arena_vae.scratch():
            with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16):
                recon, z = self.vae(images, return_latent=True)
                latents = self._latents(z)
                if t_dmd is None:
                    t_dmd = self._draw_t(latents)
                loss, log = self.loss_fn.forward_generator(images, recon, latents, labels, compute_dmd=self.cfg.dmd_weight > 0, t=t_dmd)
            with self.arena_vae.direct():
                loss.backward()
        if exchange:
            self.arena_vae.allreduce()
        log["loss"] = loss.detach()
        return log, latents.detach()

    def _encode_only(self, images) -> torch.Tensor:
        with torch.autocast(device_type=images.device.type, dtype=torch.bfloat16), torch.no_grad():
            return self._latents(self.vae.encode(images))

    def _sit_step(self, latents, labels, t_sit, exchange: bool = True) -> Dict[str, torch.Tensor]:
        for p in self.sit.parameters():
            p.requires_grad = True
        self.sit.train()
        self.arena_sit.zero()
        if t_sit is None:
            t_sit = self._draw_t(latents)
        with torch.autocast(device_type=latents.device.type, dtype=torch.bfloat16):
            dloss = dit_training_loss(self.sit, latents, labels, self.cfg.time_dist_shift, t=t_sit)
        dloss.backward()
        if exchange:
            self.arena_sit.allreduce()
        return {"diffusion_loss": dloss.detach()}

    def _draw_t(self, like: torch.Tensor) -> torch.Tensor:
        return shift_t(sample_t(like.shape[0], like, cpu_generator=self.cpu_generator), self.cfg.time_dist_shift)

    # ---- CUDA graphs
    def capture_cuda_graphs(self, images: torch.Tensor, labels: torch.Tensor, warmup: int = 2, strict: bool = False) -> bool:
        self._sections = None
        try:
            self._gx, self._gy = images.clone(), labels.clone()
            with torch.no_grad():
                lat0 = self._encode_only(self._gx)
            self._glat = lat0.clone()
            B = lat0.shape[0]
            # raw U(0,1) draws in fp32; cast to the latents' dtype and time-shifted INSIDE the graphs (the reference does both on the device)
            self._gt_dmd = torch.rand(B, device=lat0.device, dtype=torch.float32)
            self._gt_sit = torch.rand(B, device=lat0.device, dtype=torch.float32)
            self._pin_dmd, self._pin_sit = _PinnedTimes(B), _PinnedTimes(B)
            shift = self.cfg.time_dist_shift
            for a in (self.arena_vae, self.arena_sit):
                a.hooks_enabled = True

            def turn():
                log, lat = self._vae_turn(self._gx, self._gy, shift_t(self._gt_dmd.to(self._glat.dtype), shift),
                                          exchange=self.arena_vae.hooks_enabled)
                self._glat.copy_(lat)
                return log

            def enc():
                self._glat.copy_(self._encode_only(self._gx))
                return {}

            def sit():
                return self._sit_step(self._glat, self._gy, shift_t(self._gt_sit.to(self._glat.dtype), shift),
                                      exchange=self.arena_sit.hooks_enabled)
            s_turn, m1 = _capture_with_exchange(turn, [self.arena_vae], self.arena_vae.params, warmup, (self.opt_vae,))
            s_enc = _GraphedSection(enc, [], warmup)
            s_sit, m2 = _capture_with_exchange(sit, [self.arena_sit], self.arena_sit.params, warmup, (self.opt_sit,))
            self._sections = {"turn": s_turn, "enc": s_enc, "sit": s_sit}
            self._post = {"vae": m1 == "post-replay", "sit": m2 == "post-replay"}
            self.exchange_mode = m1 if m1 == m2 else f"{m1}/{m2}"
            return True
        except Exception as e:                          # noqa: BLE001
            self._sections = None
            for a in (self.arena_vae, self.arena_sit):
                a.hooks_enabled = True
            self.exchange_mode = "eager"
            torch.cuda.synchronize()
            if strict:
                raise GraphCaptureError(f"DmdTrainer: CUDA graph capture failed ({type(e).__name__}: {e})") from e
            import warnings
            warnings.warn(f"DmdTrainer: CUDA graph capture failed ({type(e).__name__}: {e}); staying eager")
            return False

    @property
    def graphed(self) -> bool:
        return self._sections is not None

    def release_graphs(self) -> None:
        for sec in (self._sections or {}).values():
            sec.release()
        self._sections = None
        self.exchange_mode = "eager"
        for a in (self.arena_vae, self.arena_sit):
            a.hooks_enabled = True

    def step(self, images: torch.Tensor, labels: torch.Tensor, vae_turn: bool = True) -> Dict[str, torch.Tensor]:
        log: Dict[str, torch.Tensor] = {}
        if self._sections is not None and images.shape == self._gx.shape:
            self._gx.copy_(images, non_blocking=True)
            self._gy.copy_(labels, non_blocking=True)
            if vae_turn:
                self._pin_dmd.send(self._gt_dmd, self.cpu_generator)
                log = self._sections["turn"].replay()
                if self._post["vae"]:
                    self.arena_vae.allreduce()
                log["vae_norm"] = self.opt_vae.step()
            else:
                self._sections["enc"].replay()
            self._pin_sit.send(self._gt_sit, self.cpu_generator)
            log.update(self._sections["sit"].replay())
            if self._post["sit"]:
                self.arena_sit.allreduce()
            log["sit_norm"] = self.opt_sit.step()
            return log
        if vae_turn:
            log, latents = self._vae_turn(images, labels, None)
            log["vae_norm"] = self.opt_vae.step()       # clip_grad_norm_(1.0); AdamW (train_dmd.py:540-542)
        else:
            latents = self._encode_only(images)
        # 2. train the student DiT on the (detached) latents (train_dmd.py:563-575)
        log.update(self._sit_step(latents, labels, None))
        log["sit_norm"] = self.opt_sit.step()
        return log
