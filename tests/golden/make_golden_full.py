"""Production-size pin: the REAL reference Decoder (49.6 M parameters, models/flux_ae.py:184-277, post_init(32)) run here on weights
that oracle/dmvae_oracle.py:make_decoder_state regenerates from a seed, so only inputs / outputs / a few gradients are stored
(tests/golden/decoder_full.pt, ~2 MB) -- plus reference ResnetBlocks at the decoder's remaining channel configurations.

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_full.py        (build container only; needs /root/reference)
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import load_flux_ae  # noqa: E402
from oracle import dmvae_oracle as O  # noqa: E402

SEED = 11


def main():
    R = load_flux_ae()
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    dec = R.Decoder(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, resolution=256, z_channels=16)
    dec.post_init(32)
    sd = O.make_decoder_state(z_channels=32, seed=SEED, randomize_affine=True)
    dec.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(SEED + 1)
    z = torch.randn(1, 256, 32, generator=g, requires_grad=True)            # token input, the path VAE.forward uses (vae.py:96)
    y = dec(z)
    dy = torch.randn(y.shape, generator=g) / y.numel()
    names = ["conv_out.weight", "norm_out.weight", "up.0.block.0.conv1.weight", "up.0.block.0.nin_shortcut.weight",
             "up.2.upsample.conv.bias", "mid.attn_1.q.weight", "mid.attn_1.norm.bias", "conv_in.0.conv.weight"]
    params = dict(dec.named_parameters())
    grads = torch.autograd.grad(y, [z] + [params[n] for n in names], dy)
    out["decoder_full"] = dict(seed=SEED, z=z.detach(), y=y.detach().half(), dy_seed=SEED + 1, dz=grads[0],
                               dparams={n: gr for n, gr in zip(names, grads[1:])})
    # ResnetBlock at the decoder's channel configurations not covered by flux_ae.pt (64->64, 64->32 are there)
    for cin, cout, hw in [(512, 512, 8), (512, 256, 8), (256, 256, 8), (256, 128, 8), (128, 128, 8)]:
        blk = R.ResnetBlock(cin, cout)
        gb = torch.Generator().manual_seed(cin + cout)
        with torch.no_grad():
            for n, p in blk.named_parameters():
                if p.ndim > 1:
                    p.copy_(torch.randn(p.shape, generator=gb) * 0.02)
                elif "norm" in n and n.endswith("weight"):
                    p.copy_(1 + 0.1 * torch.randn(p.shape, generator=gb))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=gb))
        x = torch.randn(1, cin, hw, hw, generator=gb, requires_grad=True)
        yb = blk(x)
        dyb = torch.randn(yb.shape, generator=gb)
        (dx,) = torch.autograd.grad(yb, x, dyb)
        out[f"resnet_{cin}_{cout}"] = dict(seed=cin + cout, hw=hw, y=yb.detach().half(), dx=dx.half())
    # A2 Encoder at production size (config 5's module, models/flux_ae.py:110-181), weights regenerated from the seed
    enc = R.Encoder(resolution=256, in_channels=3, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16)
    esd = O.make_encoder_state(z_channels=16, seed=SEED + 2, randomize_affine=True)
    enc.load_state_dict(esd, strict=True)
    ge = torch.Generator().manual_seed(SEED + 3)
    x = (torch.rand(1, 3, 256, 256, generator=ge) * 2 - 1).requires_grad_(True)
    h = enc(x)
    dh = torch.randn(h.shape, generator=ge)
    eparams = dict(enc.named_parameters())
    enames = ["conv_in.weight", "down.0.block.0.conv1.weight", "down.1.downsample.conv.bias", "mid.attn_1.norm.weight", "conv_out.weight"]
    eg = torch.autograd.grad(h, [x] + [eparams[n] for n in enames], dh)
    out["encoder_full"] = dict(seed=SEED + 2, x_seed=SEED + 3, y=h.detach(), dx=eg[0].half(), dparams={n: gr for n, gr in zip(enames, eg[1:])})
    torch.save(out, os.path.join(HERE, "decoder_full.pt"))
    print({k: (list(v.keys())) for k, v in out.items()})


if __name__ == "__main__":
    main()
