"""tests/golden/gan.pt: the GAN branch of the reference's VAELossFunction (BASELINE configs[3]) executed from its source.

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_gan.py        (build container only: reads /root/reference)

`forward_generator` (train_dmd.py:232-262: L1 + L2 + LPIPS, hinge generator loss, the adaptive weight from two
`torch.autograd.grad(..., last_layer, retain_graph=True)` calls, clamp, `disc_weight`) and `forward_discriminator` (:265-285: hinge
loss on real / reconstructed logits, accuracies, balanced consistency regularisation) are cut out of the reference file and run,
fp32 on CPU, on a stub `self`: the networks around them are black boxes to this path, so the decoder is one
3x3 conv (its weight is `get_last_layer()`), the discriminator a three-layer conv net, LPIPS a differentiable stand-in and the two
augmentations identity / horizontal flip.  The fixture holds every input, the stand-ins' weights and the reference's outputs and
gradients; tests/test_modules_gpu.py rebuilds the same stand-ins and runs dmvae_b200.train.VAELossFunction on them."""
import os
import re
import sys
import textwrap
import types

import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


class TinyDisc(torch.nn.Module):
    """Stand-in discriminator (the reference's DinoDisc / PatchGAN are stock modules outside the path)."""

    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(3, 8, 3, stride=2, padding=1)
        self.c2 = torch.nn.Conv2d(8, 16, 3, stride=2, padding=1)
        self.c3 = torch.nn.Conv2d(16, 1, 3, padding=1)

    def forward(self, x, grad_ckpt=False):
        h = F.leaky_relu(self.c1(x), 0.2)
        h = F.leaky_relu(self.c2(h), 0.2)
        return self.c3(h).flatten(1)


def lpips_standin(a, b):
    return (a - b).abs().add(1e-3).pow(1.5).mean((1, 2, 3))


def requires_grad(model, flag=True):          # train_dmd.py's helper of the same name
    for p in model.parameters():
        p.requires_grad = flag


def cut(src, name):
    m = re.search(r"^    def %s\(self.*?(?=^    def |^\S)" % name, src, re.S | re.M)       # up to the next method / top-level statement
    return textwrap.dedent(m.group(0))


def main():
    src = open(f"{REF}/train_dmd.py").read()
    ns = {"torch": torch, "F": F, "requires_grad": requires_grad}
    exec(cut(src, "forward_generator"), ns)
    exec(cut(src, "forward_discriminator"), ns)
    gen_fn, disc_fn = ns["forward_generator"], ns["forward_discriminator"]

    cases = {}
    for tag, disc_weight, bcr, l2, start in [("w0.5_bcr4", 0.5, 4.0, 0.0, 0), ("w0.1_l2_bcr0", 0.1, 0.0, 0.5, 0), ("not_started", 0.5, 4.0, 0.0, 10)]:
        g = torch.Generator().manual_seed(17)
        torch.manual_seed(23)
        last = torch.nn.Conv2d(6, 3, 3, padding=1)
        disc = TinyDisc()
        h = torch.randn(3, 6, 32, 32, generator=g)
        images = torch.rand(3, 3, 32, 32, generator=g) * 2 - 1
        s = types.SimpleNamespace()
        s.l1, s.l2, s.lpips, s.disc_weight, s.bcr_weight = 1.0, l2, 1.0, disc_weight, bcr
        s.args = types.SimpleNamespace(disc_start_step=start, dmd_weight=10.0)
        s.lpips_loss = lpips_standin
        s.disc_wo_ddp = s.disc_ddp = disc
        s.daug = types.SimpleNamespace(aug=lambda x, p: x)
        s.bcr_strong_aug = types.SimpleNamespace(aug=lambda x, p: x.flip(-1))
        s.vae_wo_ddp = types.SimpleNamespace(decoder=types.SimpleNamespace(get_last_layer=lambda: last.weight))

        hh = h.clone().requires_grad_(True)
        recon = last(hh)
        loss, log = gen_fn(s, images, recon, None, None, compute_dmd=False, step=5)
        d_last, d_h = torch.autograd.grad(loss, [last.weight, hh])
        d_loss, d_log = disc_fn(s, images, recon.detach())
        d_params = torch.autograd.grad(d_loss, list(disc.parameters()))
        cases[tag] = dict(disc_weight=disc_weight, bcr=bcr, l2=l2, disc_start_step=start, step=5, h=h, images=images,
                          last_sd={k: v.detach().clone() for k, v in last.state_dict().items()},
                          disc_sd={k: v.detach().clone() for k, v in disc.state_dict().items()},
                          loss=loss.detach(), log=log, d_last=d_last, d_h=d_h, d_loss=d_loss.detach(), d_log=d_log,
                          d_params={n: gr for (n, _), gr in zip(disc.named_parameters(), d_params)})
        print(tag, float(loss), log, float(d_loss), d_log)
    torch.save(cases, os.path.join(OUT, "gan.pt"))
    print("gan.pt", os.path.getsize(os.path.join(OUT, "gan.pt")) // 1024, "KiB")


if __name__ == "__main__":
    sys.exit(main())
