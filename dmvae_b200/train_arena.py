"""Flat gradient arena + data-parallel exchange (row A7): the DDP gradient all-reduce of train_dmd.py:348 /
train_tokenizer.py:302 as chunked NCCL all-reduces over one contiguous fp32 buffer."""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import torch
import torch.distributed as tdist
from torch import nn


ALIGN = 8          # elements: every parameter starts on a 32-byte (fp32) / 16-byte (bf16 copy) boundary (vector red.add, TMA)


def tap_major(p: torch.Tensor) -> bool:
    """3x3 conv weights are kept TAP-MAJOR in the arenas -- physical order [kh*kw][Cout][Cin], the layout of the tcgen05 weight
    gradient and of the packed bf16 forward operand -- and exposed to PyTorch as a strided (Cout, Cin, 3, 3) view.  So the weight
    gradient is accumulated by the kernel straight into ``p.grad``'s storage (no un-packing pass) and the optimizer's bf16 copy
    of the updated weights is the conv tiles' operand as it stands (no packing pass).  1x1 conv weights are tap-major as they are."""
    return p.ndim == 4 and p.shape[2] == 3 and p.shape[3] == 3


def arena_view(flat: torch.Tensor, off: int, p: torch.Tensor) -> torch.Tensor:
    """The view of ``flat[off : off + p.numel()]`` that has p's shape (strided for tap-major parameters)."""
    seg = flat[off:off + p.numel()]
    if tap_major(p):
        co, ci, kh, kw = p.shape
        return seg.view(kh * kw, co, ci).permute(1, 2, 0).unflatten(2, (kh, kw))
    return seg.view(p.shape)


class GradArena:
    """All trainable gradients as views of one flat fp32 buffer, exchanged with one NCCL all-reduce per chunk over
    NVLink/NVSwitch instead of DDP's 25 MB buckets.  With ``overlap=True`` a chunk's all-reduce is issued from an
    autograd hook as soon as every gradient inside it has been accumulated (backward runs output-to-input, so the
    tail of the arena completes first), i.e. the exchange overlaps the rest of the backward pass like DDP's bucketed
    reduction (train_dmd.py:348).  With world_size == 1 it is just a flat buffer.

    Step protocol (what a trainer does every step, eagerly or while capturing a CUDA graph):
        arena.zero()            # clears the buffer AND the per-step exchange bookkeeping
        loss.backward()         # hooks / notify() launch the chunk all-reduces as chunks complete
        arena.allreduce()       # launches what is left, waits, averages, resets the bookkeeping
    ``allreduce()`` leaves the bookkeeping reset, so a loop that never calls ``zero()`` from Python between steps (a
    replayed CUDA graph only re-runs the captured ``flat.zero_()`` kernel) still exchanges every chunk on every step.
    On NCCL the average is taken inside the collective (ReduceOp.AVG): no separate division pass."""

    SCRATCH_BYTES = 4 << 20     # zero pool behind the gradients (ops.ZeroPool): cleared by the same memset as the arena

    def __init__(self, params: Iterable[nn.Parameter], chunks: int = 4, overlap: bool = True):
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        self.offsets: List[int] = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        dev = self.params[0].device if self.params else torch.device("cpu")
        tail = self.SCRATCH_BYTES // 4 if dev.type == "cuda" else 0
        self._storage = torch.zeros(n + tail, dtype=torch.float32, device=dev)
        self.flat = self._storage[:n]                                    # padding elements stay zero (and inert in the optimizer)
        from .ops import ZeroPool
        self.pool = ZeroPool(self._storage[n:])
        self.chunks = max(1, min(chunks, len(self.params) or 1))
        # chunk boundaries on parameter boundaries, roughly equal in bytes
        target = (n + self.chunks - 1) // self.chunks
        self.bounds: List[Tuple[int, int]] = []          # [start, end) element ranges
        self.chunk_of: Dict[int, int] = {}
        start = 0
        for i, p in enumerate(self.params):
            self.chunk_of[id(p)] = len(self.bounds)
            off = self.offsets[i + 1] if i + 1 < len(self.params) else n
            if off - start >= target or i == len(self.params) - 1:
                self.bounds.append((start, off))
                start = off
        self.members = [sum(1 for p in self.params if self.chunk_of[id(p)] == c) for c in range(len(self.bounds))]
        self._slots: Dict[int, torch.Tensor] = {}
        self.hooks_enabled = True          # False: no exchange is issued from backward; allreduce() does it all
        self.exchanges = 0                 # chunk all-reduces launched so far (tests / bench accounting)
        self.begin_step()
        self.overlap = False
        self._attach()
        self.overlap = overlap and self._distributed()
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad_ready)

    @staticmethod
    def _distributed() -> bool:
        return tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1

    def _attach(self):
        for p, off in zip(self.params, self.offsets):      # (re-)attach in case an optimizer dropped the views (set_to_none)
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = arena_view(self.flat, off, p)
            self._slots[id(p)] = p.grad

    def flatten(self, tensors) -> torch.Tensor:
        """Per-parameter tensors (in ``self.params`` order, parameter shapes) laid out like the arena: padded to ALIGN elements,
        3x3 conv weights tap-major.  For comparisons against ``self.flat`` in tests and consistency checks."""
        out = torch.zeros_like(self.flat)
        for p, off, t in zip(self.params, self.offsets, tensors):
            arena_view(out, off, p).copy_(t)
        return out

    def slot_of(self, p):
        """The arena view that is ``p.grad`` (None for tensors this arena does not own)."""
        return self._slots.get(id(p))

    def notify(self, p: nn.Parameter):
        """A custom Function accumulated p's gradient straight into its slot (ops.direct_param_grads): what the
        post-accumulate-grad hook would have reported."""
        if self.overlap:
            self._on_grad_ready(p)

    def direct(self):
        """Context manager for ``loss.backward()``: dmvae_b200 ops write parameter gradients directly into this arena."""
        from . import ops
        return ops.direct_param_grads(self)

    def begin_step(self):
        """Reset the per-step exchange bookkeeping (host side only, no device work)."""
        self._pending = list(self.members)
        self._launched = [False] * len(self.bounds)
        self._handles: List = []
        self._ready = set()

    def zero(self):
        self._storage.zero_()               # gradients + the zero pool behind them
        self.pool.reset()
        self._attach()
        self.begin_step()

    def scratch(self):
        """Context manager for one forward + backward that started with ``zero()``: the pass's small zero-initialised
        accumulators (GroupNorm statistics, ...) are carved out of the pool this arena clears, not filled one by one."""
        return self.pool

    def _launch(self, c: int):
        s, e = self.bounds[c]
        self._launched[c] = True
        self.exchanges += 1
        # NCCL averages inside the collective; other backends (gloo in the CPU tests) sum here and divide in allreduce()
        op = tdist.ReduceOp.AVG if tdist.get_backend() == "nccl" else tdist.ReduceOp.SUM
        self._handles.append(tdist.all_reduce(self.flat[s:e], op=op, async_op=True))

    def _on_grad_ready(self, p: nn.Parameter):
        # once per parameter and step: a Function that accumulated directly reports through notify(), and autograd still
        # runs the parameter's (empty) AccumulateGrad node with its post hooks afterwards
        if not self.hooks_enabled or id(p) in self._ready:
            return
        self._ready.add(id(p))
        c = self.chunk_of[id(p)]
        self._pending[c] -= 1
        if self._pending[c] == 0 and not self._launched[c]:
            self._launch(c)

    def launch_all(self):
        """Issue (asynchronously, on the backend's own stream, ordered after everything already queued on the current stream)
        every chunk's all-reduce that has not been launched from a hook.  ``finish()`` completes the exchange."""
        if not self._distributed():
            return
        for c in range(len(self.bounds)):
            if not self._launched[c]:
                self._launch(c)

    def finish(self):
        """Make the current stream wait for the launched all-reduces, average (non-NCCL backends), reset for the next step."""
        if not self._distributed():
            return
        for h in self._handles:
            h.wait()
        if tdist.get_backend() != "nccl":
            self.flat.div_(tdist.get_world_size())
        self.begin_step()

    def allreduce(self):
        """Finish the exchange: launch whatever was not launched from hooks, wait, average, reset for the next step."""
        self.launch_all()
        self.finish()
