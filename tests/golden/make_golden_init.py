"""models/init_param.py:4-33 pinned: the REAL reference Decoder / Encoder initialised by the reference's init_weights under a fixed
torch seed; per-tensor checksums are stored (tests/golden/init.pt).  dmvae_b200's modules + dmvae_b200.vae.init_weights must
reproduce them bit for bit (same module order => same RNG consumption, same trunc_normal_/xavier rules).

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_init.py        (build container only; needs /root/reference)
"""
import contextlib
import importlib.util
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_flux_ae  # noqa: E402

CASES = {   # name: (kind, ctor kwargs, post_init z, seed, conv_std_or_gain)
    "decoder_std": ("Decoder", dict(ch=32, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, in_channels=3, resolution=16, z_channels=4), 4, 7, 0.02),
    "decoder_xavier": ("Decoder", dict(ch=32, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, in_channels=3, resolution=16, z_channels=4), 4, 9, -0.5),
    "encoder_std": ("Encoder", dict(resolution=16, in_channels=3, ch=32, ch_mult=(1, 2), num_res_blocks=1, z_channels=4), None, 8, 0.02),
}


def checksums(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum()), v.reshape(-1)[:4].clone()) for k, v in sd.items()}


def main():
    R = load_flux_ae()
    spec = importlib.util.spec_from_file_location("ref_init", "/root/reference/models/init_param.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    out = {}
    for name, (kind, kw, post, seed, arg) in CASES.items():
        mod = getattr(R, kind)(**kw)
        if post is not None:
            mod.post_init(post)
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            m.init_weights(mod, arg)
        out[name] = dict(kind=kind, kw=kw, post=post, seed=seed, arg=arg, sums=checksums(mod.state_dict()))
    torch.save(out, os.path.join(HERE, "init.pt"))
    print({k: len(v["sums"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
