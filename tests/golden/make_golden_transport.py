"""Transport.sample (diffusion/transport/transport.py:105-116) pinned: under a fixed torch seed the real reference draws x0, then t,
then applies the time-distribution shift; tests/golden/transport.pt holds (t, checksum of x0) for fp32 / bf16 latents and two shifts.
dmvae_b200.train.sample_t_x0 must reproduce them bit for bit on the CPU (same RNG consumption order, same arithmetic).

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_transport.py        (build container only; needs /root/reference)
"""
import os
import sys
import types

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    for modname, attrs in {"torchdiffeq": {"odeint": None}}.items():
        if modname not in sys.modules:
            try:
                __import__(modname)
            except Exception:
                m = types.ModuleType(modname)
                for k, v in attrs.items():
                    setattr(m, k, v)
                sys.modules[modname] = m
    from diffusion.transport import create_transport
    out = {}
    for shift in (1.0, 3.0):
        tr = create_transport("Linear", "velocity", None, None, None, time_dist_shift=shift) if "time_dist_shift" in create_transport.__code__.co_varnames \
            else create_transport("Linear", "velocity", None, None, None)
        tr.time_dist_shift = shift
        for dname, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
            torch.manual_seed(123)
            x1 = torch.randn(6, 32, 16, 16).to(dtype)
            torch.manual_seed(77)
            t, x0, _ = tr.sample(x1)
            out[f"{dname}_shift{shift}"] = dict(shift=shift, dtype=dname, t=t, x0_sum=float(x0.double().sum()), x0_head=x0.reshape(-1)[:8].clone())
    torch.save(out, os.path.join(HERE, "transport.pt"))
    print({k: v["t"].tolist()[:3] for k, v in out.items()})


if __name__ == "__main__":
    main()
