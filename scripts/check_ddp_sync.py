#!/usr/bin/env python
"""Multi-GPU consistency check (torchrun --nproc-per-node N scripts/check_ddp_sync.py): after a few training steps on
different per-rank batches, every rank must hold bit-identical parameters (the gradient exchange covered every parameter
exactly once), and the averaged gradient of step 1 must equal the mean of the per-rank local gradients."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    tr = bench.build_trainer(dev, model_size="base")
    g = torch.Generator().manual_seed(1000 + rank)
    xs = [(torch.rand(4, 3, 256, 256, generator=g) * 2 - 1).to(dev) for _ in range(3)]
    # reference for step 1: local gradients without any exchange, averaged explicitly
    tr.arena.zero()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, _ = tr.loss_fn.forward_generator(xs[0], tr.vae(xs[0], freeze_encoder=True))
    tr.arena._launched = [True] * len(tr.arena.bounds)          # suppress the exchange for this pass: purely local gradients
    with tr.arena.direct():
        loss.backward()
    local_grad = tr.arena.flat.clone()
    dist.all_reduce(local_grad)
    local_grad /= world
    # the trainer's own path (overlapped chunked all-reduce from the readiness notifications)
    tr.arena.zero()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, _ = tr.loss_fn.forward_generator(xs[0], tr.vae(xs[0], freeze_encoder=True))
    with tr.arena.direct():
        loss.backward()
    tr.arena.allreduce()
    err = ((tr.arena.flat - local_grad).norm() / local_grad.norm()).item()
    for x in xs:
        tr.step(x)
    flat_p = torch.cat([p.detach().reshape(-1) for p in tr.params])
    lo, hi = flat_p.clone(), flat_p.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = (hi - lo).abs().max().item()
    if rank == 0:
        print(f"world {world}: exchanged-vs-explicit mean gradient rel err {err:.3e}; parameter spread across ranks after 3 steps {spread:.3e}")
        assert err < 5e-3, "gradient exchange disagrees with the explicit average (beyond run-to-run bf16 noise)"
        assert spread == 0.0, "ranks diverged: some parameter was reduced before its gradient was complete"
        print("DDP SYNC OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
