#!/usr/bin/env python
"""Per-step loss parity of the REAL trainer (north star: "per-step loss matching the reference within 1e-3 relative over 100
steps", BASELINE.md section 4: 1 GPU and 8 GPUs).

Arms, all on the identical synthetic stream (global batch G images per step, seeded per step), from identical weights:

  ours      dmvae_b200.train.TokenizerTrainer -- the production step: sm_100a kernels, gradient arena, direct gradient
            accumulation, CUDA-graph replay, in-graph NCCL exchange over `world` ranks (each rank takes G/world images),
            fused clip + AdamW + EMA.
  control   scripts/stock_arms.StockStep(mode="autocast"): the same step the way the reference executes it on a GPU
            (torch.autocast(bf16), cuDNN / ATen, torch.optim.AdamW).  Rank 0 only, on the whole global batch.
  anchor    StockStep(mode="fp32"): the same graph in strict fp32 (TF32 off) = exact arithmetic.  Rank 0 only.

Two comparisons (see run_parity): SAME WEIGHTS PER STEP -- ours and the control are handed the anchor's weights before every
step, so the per-step loss difference is the pipelines' arithmetic and nothing else -- and FREE-RUNNING trajectories, where
`control vs anchor` shows how far the reference's own bf16 execution drifts from exact arithmetic.

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


def _make_trainer(dev, size, lr, seed):
    import warnings
    from dmvae_b200.lpips import LPIPS
    from dmvae_b200.train import LossConfig, TokenizerTrainer, VAELossFunction
    from dmvae_b200.vae import VAE
    torch.manual_seed(seed)                        # identical initial weights on every rank and in every arm
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
    tr = TokenizerTrainer(vae, VAELossFunction(LossConfig(l1=1.0, l2=0.0, lpips=1.0, dmd_weight=0.0), lpips_loss=lp), lr=lr)
    return vae, lp, tr


def _force_weights(tr, vae, arm, world):
    """Teacher forcing: overwrite the trainer's trainable weights with the stock arm's current ones (rank 0 holds the arm; the
    other ranks receive them by broadcast)."""
    import torch.distributed as dist
    tr.flush()                                     # pipelined trainer: apply the pending update BEFORE the weights are overwritten
    if arm is not None:
        src = dict(("decoder." + k, v) for k, v in arm.sd.items())
        src.update(("bottle_neck." + k, v) for k, v in arm.mlp.state_dict().items())
        with torch.no_grad():
            for name, p in vae.named_parameters():
                if p.requires_grad:
                    p.copy_(src[name])
    if world > 1:
        dist.broadcast(tr.fused.flat_p, src=0)
    tr.weights_changed()                           # out-of-band weight write: refresh the optimizer-maintained bf16 operands


def run_parity(dev, steps=100, global_batch=16, size="large", cuda_graph=True, micro=4, lr=1e-4, seed=1234, free_running=True):
    """Run the arms; returns the result dict on rank 0 and None elsewhere.  Must be called by every rank of the process group.

    anchor           StockStep(fp32), free running: the exact-arithmetic trajectory.
    *_forced         ours / control take the anchor's weights before every step, so each step compares the LOSS OF THE SAME
                     WEIGHTS ON THE SAME BATCH through the three pipelines (and still runs its own backward + optimizer step,
                     whose result the next forcing overwrites).  This is the per-step parity figure.
    *_free           ours / control train on their own from the common start: the trajectories.  AdamW's first steps move every
                     weight by ~lr * sign(g), so bf16-level gradient noise on near-zero gradients separates any two pipelines'
                     trajectories early on -- the control arm (the reference's own execution path) shows by how much."""
    import torch.distributed as dist
    import stock_arms
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    assert global_batch % world == 0
    B = global_batch // world
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    vae_f, lp, tr_forced = _make_trainer(dev, size, lr, seed)
    arms = {}
    if rank == 0:
        arms = stock_arms.arms_from_vae(vae_f, lp, modes=["fp32", "autocast"], lr=lr, micro=micro)
        arms["autocast_forced"] = arms.pop("autocast")
        if free_running:
            arms["autocast_free"] = stock_arms.arms_from_vae(vae_f, lp, modes=["autocast"], lr=lr, micro=micro)["autocast"]
    trainers = {"forced": (vae_f, tr_forced)}
    if free_running:
        vae_r, _, tr_free = _make_trainer(dev, size, lr, seed)
        trainers["free"] = (vae_r, tr_free)

    def batch(step):
        g = torch.Generator().manual_seed(seed * 100003 + step)
        return torch.rand(global_batch, 3, 256, 256, generator=g) * 2 - 1

    if cuda_graph:
        for _, tr in trainers.values():
            tr.capture_cuda_graph(batch(0)[rank * B:(rank + 1) * B].to(dev), strict=True)
    losses = {k: [] for k in ("anchor", "ours_forced", "ours_free", "control_forced", "control_free")}
    secs = {k: 0.0 for k in losses}

    def ours_step(tr, xg):
        loss = tr.step(xg[rank * B:(rank + 1) * B])["loss"].float().clone()
        if world > 1:                              # the global-batch loss is the mean of the equal-sized shards' losses
            dist.all_reduce(loss, op=dist.ReduceOp.AVG)
        return loss.item()

    for s in range(steps):
        xg = batch(s).to(dev)
        anchor = arms.get("fp32")
        # same weights in all three pipelines for this step
        _force_weights(tr_forced, vae_f, anchor, world)
        t0 = time.perf_counter()
        losses["ours_forced"].append(ours_step(tr_forced, xg))
        secs["ours_forced"] += time.perf_counter() - t0
        if free_running:
            t0 = time.perf_counter()
            losses["ours_free"].append(ours_step(trainers["free"][1], xg))
            secs["ours_free"] += time.perf_counter() - t0
        if rank == 0:
            cf = arms["autocast_forced"]
            with torch.no_grad():
                for k, v in anchor.sd.items():
                    cf.sd[k].copy_(v)
                cf.mlp.load_state_dict(anchor.mlp.state_dict())
            for name, arm_key in (("control_forced", "autocast_forced"), ("control_free", "autocast_free"), ("anchor", "fp32")):
                if arm_key in arms:
                    t0 = time.perf_counter()
                    losses[name].append(arms[arm_key].step(xg).item())
                    secs[name] += time.perf_counter() - t0
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    graph_state = (tr_forced.graphed, tr_forced.exchange_mode)
    for _, tr in trainers.values():
        tr.release_graphs()                        # captured NCCL operations must not outlive this run
    if world > 1:
        dist.barrier()
    if rank != 0:
        return None
    a = losses["anchor"]
    out = {"steps": steps, "global_batch": global_batch, "world": world, "per_rank_batch": B, "encoder": f"ViT-{size} (frozen)",
           "trainer": f"TokenizerTrainer, cuda_graph={graph_state[0]}, exchange={graph_state[1]}, fused clip+AdamW+EMA, lr={lr}",
           "tolerance_north_star": 1e-3,
           "same_weights_per_step": {
               "what": "every step, ours and the cuDNN-autocast control start from the fp32 anchor's weights: loss of identical weights "
                       "on the identical batch through three pipelines",
               "ours_vs_anchor_fp32": _rel_stats(losses["ours_forced"], a),
               "control_vs_anchor_fp32": _rel_stats(losses["control_forced"], a),
               "ours_vs_control": _rel_stats(losses["ours_forced"], losses["control_forced"])},
           "loss_first_last": {"anchor_fp32": [a[0], a[-1]], "ours_same_weights": [losses["ours_forced"][0], losses["ours_forced"][-1]]},
           "seconds": {k: round(v, 1) for k, v in secs.items() if v}}
    if free_running:
        out["free_running_trajectories"] = {
            "what": "each arm trains on its own from the common initial weights",
            "ours_vs_anchor_fp32": _rel_stats(losses["ours_free"], a),
            "control_vs_anchor_fp32": _rel_stats(losses["control_free"], a),
            "ours_vs_control": _rel_stats(losses["ours_free"], losses["control_free"])}
        out["loss_first_last"]["ours_free"] = [losses["ours_free"][0], losses["ours_free"][-1]]
        out["loss_first_last"]["control_free"] = [losses["control_free"][0], losses["control_free"][-1]]
    sw = out["same_weights_per_step"]
    out["verdict"] = {"ours_same_weights_max": sw["ours_vs_anchor_fp32"]["max"],
                      "reference_path_same_weights_max": sw["control_vs_anchor_fp32"]["max"],
                      "within_1e-3": sw["ours_vs_anchor_fp32"]["max"] <= 1e-3,
                      "within_reference_path_noise": sw["ours_vs_anchor_fp32"]["max"] <= 1.5 * sw["control_vs_anchor_fp32"]["max"] + 2e-4}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--global-batch", type=int, default=16)
    ap.add_argument("--size", default="large", choices=["base", "large"])
    ap.add_argument("--micro", type=int, default=4, help="micro-batch of the stock arms (memory)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-free", action="store_true", help="skip the free-running trajectories (same-weights comparison only)")
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = run_parity(dev, a.steps, a.global_batch, a.size, not a.no_graph, a.micro, free_running=not a.no_free)
    if out is not None:
        print(json.dumps(out), flush=True)
    from dmvae_b200.train import shutdown_distributed
    shutdown_distributed()


if __name__ == "__main__":
    main()
