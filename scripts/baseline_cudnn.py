#!/usr/bin/env python
"""Context baseline (not the product, not the CPU reference arm): the SAME training step executed the way the
reference executes it on a GPU -- stock PyTorch ops (cuDNN convs, ATen GroupNorm/SiLU, SDPA, ATen losses) under
torch.autocast(bf16), cudnn.benchmark=True -- using the oracle's functional restatement of the modules on cuda:0.
Prints images/s so profiles/ can show what the hand-written kernels are measured against."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from oracle import dmvae_oracle as O
from dmvae_b200.vae import DINOEncoder, MLP


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--channels-last", action="store_true")
    a = ap.parse_args()
    dev = "cuda"
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(42)
    enc = DINOEncoder("large").eval().to(dev)
    mlp = MLP(1024, 32).to(dev)
    sd = {k: v.to(dev).requires_grad_(True) for k, v in O.make_decoder_state(z_channels=32, seed=42).items()}
    lp = {k: v.to(dev) for k, v in O.make_lpips_state(seed=42).items()}

    params = list(sd.values()) + list(mlp.parameters())
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, fused=True)
    ema = [p.detach().clone() for p in params]
    x = torch.rand(a.batch, 3, 256, 256, device=dev) * 2 - 1
    shift = torch.tensor(O.LPIPS_SHIFT, device=dev).view(1, 3, 1, 1)
    scale = torch.tensor(O.LPIPS_SCALE, device=dev).view(1, 3, 1, 1)

    def vgg(inp):
        h = (inp - shift) / scale
        if a.channels_last:
            h = h.contiguous(memory_format=torch.channels_last)
        feats = []
        slice_of = lambda i: 1 + sum(i > t for t in (3, 8, 15, 22))
        for i in range(30):
            if i in O.VGG_CONVS:
                h = F.relu(F.conv2d(h, lp[f"net.slice{slice_of(i)}.{i}.weight"], lp[f"net.slice{slice_of(i)}.{i}.bias"], padding=1))
            elif i in O.VGG_POOLS:
                h = F.max_pool2d(h, 2, 2)
            if i in O.VGG_TAPS:
                feats.append(h)
        return feats

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            with torch.no_grad():
                tok = enc(x)
            z = mlp(tok)
            zz = z.transpose(1, 2).reshape(a.batch, 32, 16, 16)
            if a.channels_last:
                zz = zz.contiguous(memory_format=torch.channels_last)
            rec = O.decoder_forward(sd, zz, bf16=False).float()          # plain torch ops; autocast picks bf16 convs
            l1 = F.l1_loss(rec, x)
            with torch.no_grad():
                f0 = vgg(x)
            f1 = vgg(rec)
            lin_ws = [lp[f"lin{k}.model.1.weight"].flatten() for k in range(5)]
            loss = l1 + O.lpips_distance(f0, f1, lin_ws)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        torch._foreach_mul_(ema, 0.9999)
        torch._foreach_add_(ema, [p.data for p in params], alpha=1e-4)
        return loss

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"impl": "stock PyTorch/cuDNN eager (reference execution path) on the same B200", "channels_last": a.channels_last,
                      "batch": a.batch, "ms_per_step": round(ms, 2), "images_per_s": round(a.batch / ms * 1e3, 1)}))


if __name__ == "__main__":
    main()
