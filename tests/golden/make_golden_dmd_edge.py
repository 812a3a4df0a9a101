"""Edge case of the DMD loss pinned on the REAL reference method (train_dmd.py:204-230, executed from its source exactly as in
make_golden.py): a sample whose normaliser w_b = mean|p_real| is exactly 0 AND whose numerator is 0, i.e. grad = 0/0 = NaN, which
the reference maps to 0 with torch.nan_to_num (:224).  Values of that sample are small even integers and t = 0.5, so every
intermediate is exact in fp32 and in bf16.  Output: tests/golden/dmd_edge.pt.

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_dmd_edge.py        (build container only; needs /root/reference)
"""
import os
import re
import sys
import textwrap
import types

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    for modname, attrs in {"torchdiffeq": {"odeint": None}, "pytz": {"timezone": lambda *a, **k: None}, "tap": {"Tap": object},
                           "wandb": {}, "matplotlib": {}, "matplotlib.pyplot": {}}.items():
        if modname not in sys.modules:
            try:
                __import__(modname)
            except Exception:
                m = types.ModuleType(modname)
                for k, v in attrs.items():
                    setattr(m, k, v)
                sys.modules[modname] = m
    from diffusion.transport import path as RP
    src = open(f"{REF}/train_dmd.py").read()
    m = re.search(r"    def compute_distribution_matching_loss\(self.*?\n(?=\n\n    def )", src, re.S)
    ns = {"torch": torch, "expand_t_like_x": lambda t, x: t.view(t.size(0), *([1] * (x.dim() - 1)))}
    exec(textwrap.dedent(m.group(0)), ns)
    ref_fn = ns["compute_distribution_matching_loss"]

    class Args:
        t0, t1, dmd_cfg_scale, num_classes = 0.0, 1.0, 1.0, 1000

    cases = {}
    for tag, dtype in [("fp32", torch.float32), ("bf16", torch.bfloat16)]:
        g = torch.Generator().manual_seed(31)
        z = torch.randn(4, 32, 16, 16, generator=g)
        x0 = torch.randn(z.shape, generator=g)
        t = torch.rand(4, generator=g)
        Tc = torch.randn(z.shape, generator=g)
        Sc = torch.randn(z.shape, generator=g)
        # sample 0: exact arithmetic, teacher = student = z - x0  =>  pred = z for both  =>  p_real = p_student = 0, w_0 = 0
        z[0] = torch.randint(-3, 4, z[0].shape, generator=g).float() * 2
        x0[0] = torch.randint(-3, 4, z[0].shape, generator=g).float() * 2
        t[0] = 0.5
        Tc[0] = z[0] - x0[0]
        Sc[0] = Tc[0]
        z, x0, t, Tc, Sc = (v.to(dtype) for v in (z, x0, t, Tc, Sc))
        s = types.SimpleNamespace()
        s.args = Args()
        s.transport = types.SimpleNamespace(sample=lambda x1: (t, x0, x1), path_sampler=RP.ICPlan())
        s.base_model = lambda xt, tt, y: Tc
        s.sit_wo_ddp = lambda xt, tt, y: Sc
        zz = z.clone().requires_grad_(True)
        loss, log = ref_fn(s, zz, torch.zeros(4, dtype=torch.long))
        (dz,) = torch.autograd.grad(loss, zz)
        assert torch.isfinite(loss) and torch.isfinite(dz).all() and float(dz[0].abs().max()) == 0.0
        cases[tag] = dict(z=z, x0=x0, t=t, Tc=Tc, Sc=Sc, cfg=1.0, loss=loss.detach(), dz=dz, gnorm=log["dmd_gradient_norm"])
        print(tag, float(loss), log["dmd_gradient_norm"])
    torch.save(cases, os.path.join(OUT, "dmd_edge.pt"))


if __name__ == "__main__":
    main()
