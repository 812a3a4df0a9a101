#!/usr/bin/env python
"""Multi-GPU consistency check (torchrun --nproc-per-node N scripts/check_ddp_sync.py): after a few training steps on
different per-rank batches, every rank must hold bit-identical parameters (the gradient exchange covered every parameter
exactly once), and the averaged gradient must equal the mean of the per-rank local gradients -- on the eager path
(hook-driven overlapped exchange) AND under CUDA-graph replay (exchange captured inside the graph, or issued after the
replay if the process group could not be captured), for several replayed steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def spread(tr):
    flat_p = tr.fused.flat_p
    lo, hi = flat_p.clone(), flat_p.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return (hi - lo).abs().max().item()


def local_mean_gradient(tr, x, world):
    """Purely local gradients (no exchange), averaged explicitly with a plain all-reduce."""
    tr.arena.zero()
    tr.arena._launched = [True] * len(tr.arena.bounds)          # suppress the exchange for this pass
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, _ = tr.loss_fn.forward_generator(x, tr.vae(x, freeze_encoder=True))
    with tr.arena.direct():
        loss.backward()
    g = tr.arena.flat.clone()
    tr.arena.begin_step()
    dist.all_reduce(g)
    return g / world


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    tr = bench.build_trainer(dev, model_size="base")
    g = torch.Generator().manual_seed(1000 + rank)
    xs = [(torch.rand(4, 3, 256, 256, generator=g) * 2 - 1).to(dev) for _ in range(4)]

    # ---- eager path: the exchange finished by allreduce() equals the explicit mean
    ref = local_mean_gradient(tr, xs[0], world)
    n0 = tr.arena.exchanges
    tr._forward_backward(xs[0])
    err_eager = ((tr.arena.flat - ref).norm() / ref.norm()).item()
    assert tr.arena.exchanges - n0 == len(tr.arena.bounds)
    for x in xs[:3]:
        tr.step(x)
    tr.flush()
    sp_eager = spread(tr)

    # ---- CUDA-graph replay
    assert tr.capture_cuda_graph(xs[0], strict=True)
    mode = tr.exchange_mode
    errs = []
    for it in range(4):
        ref = local_mean_gradient(tr, xs[it], world)             # same weights on every rank (spread is 0), this step's batch
        if tr.pipelined:
            tr.step(xs[it])                                      # local gradients of this batch are in the arena, update pending
            tr.arena.launch_all()                                # what the next step() / flush() does first
            tr.arena.finish()
            errs.append(((tr.arena.flat - ref).norm() / ref.norm()).item())
            tr._pending = False
            tr.fused.step()                                      # steps on the exchanged gradient
        else:
            tr._gx.copy_(xs[it])
            tr._section.replay()
            if tr._exchange_outside:
                tr.arena.allreduce()
            errs.append(((tr.arena.flat - ref).norm() / ref.norm()).item())
            tr.fused.step()                                      # steps on the replayed (exchanged) gradient
    for x in xs:                                                 # and the public step()
        tr.step(x)
    tr.flush()
    sp_graph = spread(tr)
    if rank == 0:
        print(f"world {world}: eager exchanged-vs-explicit mean gradient rel err {err_eager:.3e}; parameter spread after 3 eager steps {sp_eager:.3e}")
        print(f"world {world}: graph replay (exchange {mode}) rel err per replayed step {['%.3e' % e for e in errs]}; "
              f"parameter spread after 8 replayed steps {sp_graph:.3e}")
        assert err_eager < 5e-3, "gradient exchange disagrees with the explicit average (beyond run-to-run bf16 noise)"
        assert sp_eager == 0.0, "ranks diverged: some parameter was reduced before its gradient was complete"
        assert max(errs) < 5e-3, "a replayed step stepped on a gradient that is not the cross-rank mean"
        assert sp_graph == 0.0, "ranks diverged under CUDA-graph replay"
        print("DDP SYNC OK", flush=True)
    from dmvae_b200.train import shutdown_distributed
    shutdown_distributed(tr)


if __name__ == "__main__":
    main()
