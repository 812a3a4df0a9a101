#!/usr/bin/env python
"""Per-step loss parity of the REAL trainer (north star: "per-step loss matching the reference within 1e-3 relative over 100
steps", BASELINE.md section 4: 1 GPU and 8 GPUs).

Three arms train from identical weights on the identical synthetic stream (global batch G images per step, seeded per step):

  ours      dmvae_b200.train.TokenizerTrainer -- the production step: sm_100a kernels, gradient arena, direct gradient
            accumulation, CUDA-graph replay, in-graph NCCL exchange over `world` ranks (each rank takes G/world images),
            fused clip + AdamW + EMA.
  control   scripts/stock_arms.StockStep(mode="autocast"): the same step the way the reference executes it on a GPU
            (torch.autocast(bf16), cuDNN / ATen, torch.optim.AdamW).  Rank 0 only, on the whole global batch.
  anchor    StockStep(mode="fp32"): the same graph in strict fp32 (TF32 off) = exact arithmetic.  Rank 0 only.

Reported per arm pair: max / mean relative loss difference over the steps, and over steps 0..10.  `control vs anchor` is the
noise floor of the reference's own bf16 execution (its loss carries the bf16 rounding of the LPIPS tail, 2^-8 relative on that
term); `ours vs anchor` must not exceed it by more than run-to-run noise, and `ours vs control` is the two bf16 pipelines against
each other.

    python scripts/loss_parity.py [--steps 100 --global-batch 16 --size large]
    torchrun --nproc-per-node 8 scripts/loss_parity.py ...                       # 8 x 2 images against the same oracle arms
Prints one JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import torch  # noqa: E402


def _rel_stats(a, b, head=11):
    rel = [abs(x - y) / max(abs(y), 1e-30) for x, y in zip(a, b)]
    return {"max": max(rel), "mean": sum(rel) / len(rel), "max_first_steps": max(rel[:head]), "argmax": rel.index(max(rel))}


def run_parity(dev, steps=100, global_batch=16, size="large", cuda_graph=True, anchor=True, control=True, micro=4, lr=1e-4,
               seed=1234):
    """Run the arms; returns the result dict on rank 0 and None elsewhere.  Must be called by every rank of the process group."""
    import torch.distributed as dist
    from dmvae_b200.lpips import LPIPS
    from dmvae_b200.train import LossConfig, TokenizerTrainer, VAELossFunction
    from dmvae_b200.vae import VAE
    import stock_arms
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    assert global_batch % world == 0
    B = global_batch // world
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import warnings
    torch.manual_seed(seed)                        # identical initial weights on every rank
    vae = VAE(z_channels=32, model_size=size).to(dev)
    vae.encoder.eval()
    for p in vae.encoder.parameters():
        p.requires_grad = False
    with torch.no_grad():                          # LayerScale at its 1e-5 init would hide the encoder from the loss entirely
        for blk in vae.encoder.model.blocks:
            blk.ls1.gamma.fill_(0.1)
            blk.ls2.gamma.fill_(0.1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lp = LPIPS(ckpt_path=None, pretrained_vgg=False).eval().to(dev)
    arms = {}
    if rank == 0:
        modes = (["fp32"] if anchor else []) + (["autocast"] if control else [])
        arms = stock_arms.arms_from_vae(vae, lp, modes=modes, lr=lr, micro=micro)
    tr = TokenizerTrainer(vae, VAELossFunction(LossConfig(l1=1.0, l2=0.0, lpips=1.0, dmd_weight=0.0), lpips_loss=lp), lr=lr)

    def batch(step):
        g = torch.Generator().manual_seed(seed * 100003 + step)
        return torch.rand(global_batch, 3, 256, 256, generator=g) * 2 - 1

    if cuda_graph:
        tr.capture_cuda_graph(batch(0)[rank * B:(rank + 1) * B].to(dev), strict=True)
    ours, ctrl, anch = [], [], []
    t_arm = {"ours": 0.0, "control": 0.0, "anchor": 0.0}
    for s in range(steps):
        xg = batch(s).to(dev)
        t0 = time.perf_counter()
        loss = tr.step(xg[rank * B:(rank + 1) * B])["loss"].float().clone()
        if world > 1:                              # the global-batch loss is the mean of the equal-sized shards' losses
            dist.all_reduce(loss, op=dist.ReduceOp.AVG)
        ours.append(loss.item())
        t_arm["ours"] += time.perf_counter() - t0
        if rank == 0:
            if "autocast" in arms:
                t0 = time.perf_counter()
                ctrl.append(arms["autocast"].step(xg).item())
                t_arm["control"] += time.perf_counter() - t0
            if "fp32" in arms:
                t0 = time.perf_counter()
                anch.append(arms["fp32"].step(xg).item())
                t_arm["anchor"] += time.perf_counter() - t0
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    if world > 1:
        dist.barrier()
    if rank != 0:
        return None
    out = {"steps": steps, "global_batch": global_batch, "world": world, "per_rank_batch": B, "encoder": f"ViT-{size} (frozen)",
           "trainer": f"TokenizerTrainer, cuda_graph={tr.graphed}, exchange={tr.exchange_mode}, fused clip+AdamW+EMA, lr={lr}",
           "tolerance_north_star": 1e-3,
           "loss_first_last": {"ours": [ours[0], ours[-1]]},
           "seconds": {k: round(v, 1) for k, v in t_arm.items()}}
    if anch:
        out["ours_vs_anchor_fp32"] = _rel_stats(ours, anch)
        out["loss_first_last"]["anchor_fp32"] = [anch[0], anch[-1]]
    if ctrl:
        out["ours_vs_control_cudnn_autocast"] = _rel_stats(ours, ctrl)
        out["loss_first_last"]["control"] = [ctrl[0], ctrl[-1]]
    if anch and ctrl:
        out["control_vs_anchor_fp32"] = _rel_stats(ctrl, anch)
        floor = out["control_vs_anchor_fp32"]["max"]
        out["verdict"] = {"ours_max": out["ours_vs_anchor_fp32"]["max"], "reference_path_noise_floor_max": floor,
                          "within_1e-3": out["ours_vs_anchor_fp32"]["max"] <= 1e-3,
                          "within_reference_noise_floor": out["ours_vs_anchor_fp32"]["max"] <= 1.5 * floor}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--global-batch", type=int, default=16)
    ap.add_argument("--size", default="large", choices=["base", "large"])
    ap.add_argument("--micro", type=int, default=4, help="micro-batch of the stock arms (memory)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-anchor", action="store_true")
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = run_parity(dev, a.steps, a.global_batch, a.size, not a.no_graph, not a.no_anchor, True, a.micro)
    if out is not None:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
