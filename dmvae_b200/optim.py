"""Fused optimizer step on flat arenas: global-norm clip + AdamW + EMA in two kernels.

Replaces, for the trainable VAE parameters, the reference's per-step sequence
    clip_grad_norm_(params, 1.0); optimizer.step(); update_ema(ema_model, model)
(train_tokenizer.py:415-417,437 with update_ema :140-150; train_dmd.py:540-544).  Parameters, gradients, both Adam moments
and the EMA copy each live in one flat fp32 buffer; the ``nn.Parameter``s stay ordinary parameters whose ``.data`` / ``.grad``
are views into those buffers, so ``state_dict()``, ``load_state_dict()`` (in-place copy), DDP-free all-reduce (GradArena)
and the packed-weight caches (version counters are bumped after every step) keep working.

Checkpointing follows the reference's entries (train_tokenizer.py:439-453: ``opt_vae`` and a full ``vae_ema`` state_dict):
``state_dict()`` / ``load_state_dict()`` carry the moments, the step count and the EMA arena; ``ema_state_dict(module)`` merges
the EMA views with the module's frozen parameters and buffers into a dict that ``VAE.load_pretrained(ema=True)`` loads strictly.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional

import torch
from torch import nn

from ._lib import call, ptr
from .train_arena import GradArena, arena_view


class FlatAdamWEMA:
    def __init__(self, params: Iterable[nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.95), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_norm: float = 1.0, ema_decay: Optional[float] = 0.9999,
                 arena: Optional[GradArena] = None):
        self.arena = arena if arena is not None else GradArena(params)
        self.params: List[nn.Parameter] = self.arena.params
        if not self.params or not self.params[0].is_cuda:
            raise RuntimeError("FlatAdamWEMA needs CUDA parameters (dmvae_b200 has no CPU path)")
        n = self.arena.flat.numel()
        dev = self.arena.flat.device
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(self.params, self.arena.offsets):     # move every parameter's storage into the flat buffer
                view = arena_view(self.flat_p, off, p)              # (3x3 conv weights: tap-major storage, strided view)
                view.copy_(p.data)
                p.data = view
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.ema = self.flat_p.clone() if ema_decay is not None else None
        # bf16 copy of the parameters, rewritten by the optimizer kernel: for conv weights (tap-major, or 1x1) the view
        # [taps][Cout][Cin] of it IS the conv tiles' packed forward operand (ops.WeightPack picks it up through p._dmvae_w16)
        self.w16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        # ... and its per-tap transpose [taps-1-tap][Cin][Cout], the data-gradient operand, at the same offsets of a second
        # arena: rebuilt for ALL conv weights by one batched launch after the update (DMVAE_BATCHED_DGRAD_PACK=0: ops.WeightPack
        # derives it per layer, one small launch each, the first time a step uses the weight)
        batched = bool(int(os.environ.get("DMVAE_BATCHED_DGRAD_PACK", "1")))
        self.wd16 = torch.zeros(n if batched else 0, dtype=torch.bfloat16, device=dev)
        desc, tiles = [], 0
        for p, off in zip(self.params, self.arena.offsets):
            if p.ndim == 4 and (p.shape[2] * p.shape[3] == 1 or (p.shape[2] == 3 and p.shape[3] == 3)):
                co, ci, kh, kw = p.shape
                p._dmvae_w16 = self.w16[off:off + p.numel()].view(kh * kw, co, ci)
                p._dmvae_w16_version = -1
                if batched:
                    p._dmvae_wd16 = self.wd16[off:off + p.numel()].view(kh * kw, ci, co)
                    desc.append((off, co, ci, kh * kw, tiles))
                    tiles += kh * kw * ((co + 31) // 32) * ((ci + 31) // 32)
                elif hasattr(p, "_dmvae_wd16"):
                    del p._dmvae_wd16
        self._pack_desc = torch.tensor(desc, dtype=torch.int64, device=dev).reshape(-1, 5) if desc else None
        self._pack_tiles = tiles
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.ema_decay = 0.0 if ema_decay is None else ema_decay
        self.t = 0
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ema_synced_at = tuple(p._version for p in self.params)
        self.sync_w16()

    def sync_w16(self) -> None:
        """Rebuild the bf16 copy from the fp32 parameters as they are now and mark it current (construction, after
        ``load_state_dict``, before a CUDA-graph capture).  Until this or the next ``step()`` runs, ops.WeightPack sees a stale
        version stamp and packs from the fp32 parameter itself."""
        call("dmvae_cast_bf16", ptr(self.flat_p), ptr(self.w16), self.flat_p.numel())
        self._pack_dgrad()
        for p in self.params:
            if hasattr(p, "_dmvae_w16"):
                p._dmvae_w16_version = p._version

    def _pack_dgrad(self) -> None:
        if self._pack_desc is not None:
            call("dmvae_pack_dgrad_batched", ptr(self.w16), ptr(self.wd16), ptr(self._pack_desc), self._pack_desc.shape[0],
                 self._pack_tiles)

    def sync_ema(self) -> None:
        """EMA := current weights.  The reference deep-copies the model into ``vae_ema`` after the weights are in place
        (train_tokenizer.py:397); call this after ``load_state_dict`` / ``load_pretrained`` on a model whose trainer already
        exists.  ``step()`` does it by itself when it sees, before the first update, that the weights changed since construction."""
        if self.ema is not None:
            self.ema.copy_(self.flat_p)
        self._ema_synced_at = tuple(p._version for p in self.params)

    @torch.no_grad()
    def step(self, lr: Optional[float] = None) -> torch.Tensor:
        """One optimizer step on the gradients currently in the arena.  Returns the pre-clip gradient norm (0-d, device)."""
        if self.t == 0 and self.ema is not None and tuple(p._version for p in self.params) != self._ema_synced_at:
            self.sync_ema()                             # weights were loaded after the trainer was built
        self.t += 1
        n = self.flat_p.numel()
        self._sumsq.zero_()
        call("dmvae_grad_sumsq", ptr(self.arena.flat), ptr(self._sumsq), n)
        call("dmvae_adamw_ema_step", ptr(self.flat_p), ptr(self.arena.flat), ptr(self.m), ptr(self.v), ptr(self.ema), ptr(self.w16),
             ptr(self._sumsq), ptr(self._norm), n, float(self.lr if lr is None else lr), float(self.betas[0]),
             float(self.betas[1]), float(self.eps), float(self.wd), int(self.t), float(self.max_norm), float(self.ema_decay))
        self._pack_dgrad()
        torch.autograd.graph.increment_version(self.params)      # the kernel wrote through raw pointers
        for p in self.params:
            if hasattr(p, "_dmvae_w16"):
                p._dmvae_w16_version = p._version                # ... and refreshed the bf16 operand copy with them
        return self._norm[0]

    # ------------------------------------------------------------------------------------------------ EMA views / checkpoints
    def ema_state(self, named_params: Dict[str, nn.Parameter]) -> Dict[str, torch.Tensor]:
        """EMA tensors keyed like ``named_parameters()`` (the trainable part of the reference's ``vae_ema`` checkpoint entry)."""
        out = {}
        by_id = {id(p): k for k, p in named_params.items()}
        for p, off in zip(self.params, self.arena.offsets):
            if id(p) in by_id and self.ema is not None:
                out[by_id[id(p)]] = arena_view(self.ema, off, p)
        return out

    def ema_state_dict(self, module: nn.Module) -> Dict[str, torch.Tensor]:
        """A full reference-format ``vae_ema`` entry (train_tokenizer.py:441): ``module.state_dict()`` with every trainable
        tensor replaced by its EMA value; frozen parameters (the stage-1 encoder) and buffers are the live ones, exactly what
        ``update_ema`` leaves in the reference's deep-copied model (train_tokenizer.py:140-150 only touches requires_grad
        parameters).  Loads with ``strict=True``."""
        sd = {k: v.detach().clone(memory_format=torch.contiguous_format) for k, v in module.state_dict().items()}
        for k, v in self.ema_state(dict(module.named_parameters())).items():
            sd[k] = v.detach().clone(memory_format=torch.contiguous_format)
        return sd

    def state_dict(self) -> Dict[str, object]:
        """The ``opt_vae`` checkpoint entry: moments, step count, hyper-parameters and the EMA arena (flat, in arena order and
        layout -- 3x3 conv weights tap-major, parameters padded to 8 elements; ``shapes`` guards against loading into a different
        parameter set)."""
        return {"step": self.t, "exp_avg": self.m.detach().clone(), "exp_avg_sq": self.v.detach().clone(),
                "ema": None if self.ema is None else self.ema.detach().clone(),
                "lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd, "max_norm": self.max_norm,
                "ema_decay": self.ema_decay, "shapes": [tuple(p.shape) for p in self.params]}

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        shapes = [tuple(s) for s in sd["shapes"]]
        if shapes != [tuple(p.shape) for p in self.params]:
            raise ValueError("FlatAdamWEMA.load_state_dict: parameter shapes differ from the checkpoint's")
        with torch.no_grad():
            self.m.copy_(sd["exp_avg"])
            self.v.copy_(sd["exp_avg_sq"])
            if self.ema is not None:
                if sd.get("ema") is not None:
                    self.ema.copy_(sd["ema"])
                else:
                    self.ema.copy_(self.flat_p)
        self.t = int(sd["step"])
        self.lr, self.betas, self.eps, self.wd = float(sd["lr"]), tuple(sd["betas"]), float(sd["eps"]), float(sd["weight_decay"])
        self.max_norm = float(sd["max_norm"])
        self._ema_synced_at = tuple(p._version for p in self.params)
