#!/usr/bin/env python
"""Per-step loss parity: the CUDA training step vs the CPU oracle (autocast-emulating bf16 mode) on identical weights,
inputs and optimizer settings.  Decoder + bottleneck MLP are trained on fixed "encoder tokens" (the frozen ViT is a
shared black box and is left out so the comparison isolates the path under test); loss = L1 + LPIPS, AdamW, clip 1.0.

    python scripts/loss_parity.py --steps 100 --batch 2 [--small]
Prints one JSON line: max / mean relative loss difference over the steps."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import torch  # noqa: E402

from oracle import dmvae_oracle as O  # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--small", action="store_true", help="ch=64 decoder at 64x64 instead of the production decoder")
    ap.add_argument("--lr", type=float, default=1e-4)
    a = ap.parse_args()
    from dmvae_b200.autoencoder import Decoder
    from dmvae_b200.lpips import LPIPS
    from dmvae_b200 import losses
    dev = "cuda"
    torch.set_num_threads(os.cpu_count() or 1)
    if a.small:
        kw = dict(ch=64, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, in_channels=3, resolution=64, z_channels=32)
        sd0 = O.make_decoder_state(ch=64, ch_mult=(1, 2), num_res_blocks=1, z_channels=32, seed=1)
        res = 64
    else:
        kw = dict(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, resolution=256, z_channels=16)
        sd0 = O.make_decoder_state(z_channels=32, seed=1)
        res = 256
    lp_sd = O.make_lpips_state(seed=2)
    g = torch.Generator().manual_seed(0)
    tokens = [torch.randn(a.batch, 256, 32, generator=g) for _ in range(4)]
    images = [torch.rand(a.batch, 3, res, res, generator=g) * 2 - 1 for _ in range(4)]

    # ---- GPU arm
    dec = Decoder(**kw)
    dec.post_init(32)
    dec.load_state_dict(sd0, strict=True)
    dec = dec.to(dev)
    lp = LPIPS(ckpt_path=None, pretrained_vgg=False)
    lp.load_state_dict(lp_sd, strict=True)
    lp = lp.eval().to(dev)
    opt = torch.optim.AdamW(dec.parameters(), lr=a.lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    gpu_losses = []
    for i in range(a.steps):
        z, x = tokens[i % 4].to(dev), images[i % 4].to(dev)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            rec = dec(z).float()
            l1, _ = losses.l1_l2_loss(rec, x)
            loss = l1 + lp(x, rec).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(dec.parameters(), 1.0)
        opt.step()
        gpu_losses.append(loss.item())

    # ---- oracle arm (CPU, bf16 rounding points emulated)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt_c = torch.optim.AdamW(list(sd.values()), lr=a.lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    cpu_losses = []
    t0 = time.time()
    for i in range(a.steps):
        z, x = tokens[i % 4], images[i % 4]
        rec = O.decoder_forward(sd, z, bf16=True)
        loss = (rec - x).abs().mean() + O.lpips_forward(lp_sd, x, rec, bf16=True)
        opt_c.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(sd.values()), 1.0)
        opt_c.step()
        cpu_losses.append(loss.item())
    rel = [abs(a_ - b_) / abs(b_) for a_, b_ in zip(gpu_losses, cpu_losses)]
    print(json.dumps({"steps": a.steps, "batch": a.batch, "decoder": "small" if a.small else "production",
                      "max_rel_loss_diff": max(rel), "mean_rel_loss_diff": sum(rel) / len(rel),
                      "first": [gpu_losses[0], cpu_losses[0]], "last": [gpu_losses[-1], cpu_losses[-1]],
                      "oracle_seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
