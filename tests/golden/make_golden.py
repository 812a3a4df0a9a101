"""Generates tests/golden/*.pt by executing the REAL reference modules from /root/reference on seeded inputs.

Run in the build container only (the reference tree does not travel to the GPU box):
    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden.py
Everything is fp32 on CPU.  Fixtures hold the inputs, the module state_dict and the reference outputs / gradients,
so tests/test_oracle_golden.py can pin oracle/dmvae_oracle.py without the reference being present.
"""
import importlib.util
import os
import sys
import types

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(0)


def load_flux_ae():
    spec = importlib.util.spec_from_file_location("ref_flux_ae", f"{REF}/models/flux_ae.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.ndim > 1:
                torch.nn.init.trunc_normal_(p, std=0.05, generator=g)
            elif "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))


def sd_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def main():
    R = load_flux_ae()
    g = torch.Generator().manual_seed(123)
    fx = {}

    # --- A1a ResnetBlock (with and without nin_shortcut), A1b AttnBlock, A1c Upsample / Downsample
    for name, mod, shape in [
        ("resnet_64_64", R.ResnetBlock(64, 64), (2, 64, 8, 8)),
        ("resnet_64_32", R.ResnetBlock(64, 32), (2, 64, 6, 10)),
        ("attn_64", R.AttnBlock(64), (2, 64, 8, 8)),
        ("upsample_32", R.Upsample(32), (2, 32, 5, 7)),
        ("downsample_32", R.Downsample(32), (2, 32, 8, 8)),
    ]:
        randomize(mod, hash(name) % 1000)
        x = torch.randn(shape, generator=g, requires_grad=True)
        y = mod(x)
        dy = torch.randn(y.shape, generator=g)
        grads = torch.autograd.grad(y, [x] + list(mod.parameters()), dy)
        fx[name] = dict(sd=sd_of(mod), x=x.detach(), y=y.detach(), dy=dy, dx=grads[0],
                        dparams={n: gr for (n, _), gr in zip(mod.named_parameters(), grads[1:])})

    # --- A1 Decoder (tiny, post_init stem, 4-D latent input) and A2 Encoder (tiny)
    dec = R.Decoder(ch=32, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, in_channels=3, resolution=16, z_channels=4)
    dec.post_init(4)
    randomize(dec, 7)
    z = torch.randn(2, 4, 4, 4, generator=g, requires_grad=True)
    y = dec(z)
    dy = torch.randn(y.shape, generator=g)
    grads = torch.autograd.grad(y, [z, dec.conv_out.weight, dec.mid.block_1.conv1.weight, dec.conv_in[0].conv.weight], dy)
    fx["decoder_tiny"] = dict(sd=sd_of(dec), z=z.detach(), y=y.detach(), dy=dy, dz=grads[0], d_conv_out=grads[1],
                              d_mid_conv1=grads[2], d_stem=grads[3])
    # token input path (hard-coded 16x16 grid, models/flux_ae.py:244-245)
    zt = torch.randn(1, 256, 4, generator=g)
    fx["decoder_tiny_tokens"] = dict(z=zt, y=dec(zt).detach())

    enc = R.Encoder(resolution=16, in_channels=3, ch=32, ch_mult=(1, 2), num_res_blocks=1, z_channels=4)
    randomize(enc, 9)
    x = torch.randn(2, 3, 16, 16, generator=g)
    fx["encoder_tiny"] = dict(sd=sd_of(enc), x=x, y=enc(x).detach())

    # --- state_dict manifests of the production-size modules (names + shapes only)
    dec_full = R.Decoder(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, resolution=256, z_channels=16)
    dec_full.post_init(32)
    enc_full = R.Encoder(resolution=256, in_channels=3, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16)
    fx["manifest"] = dict(decoder={k: tuple(v.shape) for k, v in dec_full.state_dict().items()},
                          encoder={k: tuple(v.shape) for k, v in enc_full.state_dict().items()})
    torch.save(fx, os.path.join(OUT, "flux_ae.pt"))

    # --- A5 LPIPS: the reference module with torchvision's vgg16(weights=None) patched in (no network)
    sys.path.insert(0, REF)
    import torchvision
    import utils.lpips as RL
    _tv_vgg16 = torchvision.models.vgg16
    RL.models.vgg16 = lambda pretrained=True: _tv_vgg16(weights=None)
    torch.manual_seed(5)
    lp = RL.LPIPS(ckpt_path=f"{REF}/ckpt_vae/vgg.pth").eval()
    a = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    b = (a + 0.3 * torch.randn(a.shape, generator=g)).clamp(-1, 1).requires_grad_(True)
    val = lp(a, b)
    (db,) = torch.autograd.grad(val, b)
    # Feature-level fixture (the VGG weights, 14.7 M params, are too big to commit): the five tap feature maps of
    # both branches, the lin weights, the loss and d loss / d feats1.  Pins normalize/diff/lin/spatial-mean/sum/mean.
    with torch.no_grad():
        f0 = [t.detach().half().float() for t in lp.net(lp.scaling_layer(a))]     # stored as fp16: round first
    f1 = [t.detach().half().float().requires_grad_(True) for t in lp.net(lp.scaling_layer(b.detach()))]
    lins = [lp.lin0, lp.lin1, lp.lin2, lp.lin3, lp.lin4]
    val_f = None
    for kk in range(5):
        d = (RL.normalize_tensor(f0[kk]) - RL.normalize_tensor(f1[kk])) ** 2
        r = RL.spatial_average(lins[kk].model(d), keepdim=True)
        val_f = r if val_f is None else val_f + r
    val_f = val_f.mean()
    df1 = torch.autograd.grad(val_f, f1)
    assert torch.allclose(val_f, val, rtol=1e-2), (val_f, val)
    lin_sd = {k: v.detach().clone() for k, v in lp.state_dict().items() if k.startswith("lin") or k.startswith("scaling")}
    manifest = {k: tuple(v.shape) for k, v in lp.state_dict().items()}
    w0 = lp.state_dict()["net.slice1.0.weight"]
    torch.save(dict(lin_sd=lin_sd, a=a, b=b.detach(), val=val.detach(), db=db, manifest=manifest,
                    f0=[t.half() for t in f0], f1=[t.detach().half() for t in f1],
                    val_feats=val_f.detach(), df1=[t.clone() for t in df1],
                    vgg_seed=5, vgg_w0_sum=w0.double().sum().item()), os.path.join(OUT, "lpips.pt"))

    # --- A3 DMD loss: the reference method itself, with stub teacher / student networks (the DiTs are black boxes)
    for modname, attrs in {
        "torchdiffeq": {"odeint": None}, "pytz": {"timezone": lambda *a, **k: None}, "tap": {"Tap": object},
        "wandb": {}, "matplotlib": {}, "matplotlib.pyplot": {},
    }.items():
        if modname not in sys.modules:
            try:
                __import__(modname)
            except Exception:
                m = types.ModuleType(modname)
                for k, v in attrs.items():
                    setattr(m, k, v)
                sys.modules[modname] = m
    from diffusion.transport import path as RP

    class Args:
        t0, t1, dmd_cfg_scale, num_classes = 0.0, 1.0, 5.0, 1000

    # the body of VAELossFunction.compute_distribution_matching_loss (train_dmd.py:204-230), executed from its source
    import inspect, re, textwrap
    src = open(f"{REF}/train_dmd.py").read()
    m = re.search(r"    def compute_distribution_matching_loss\(self.*?\n(?=\n\n    def )", src, re.S)
    fn_src = textwrap.dedent(m.group(0))
    ns = {"torch": torch, "expand_t_like_x": lambda t, x: t.view(t.size(0), *([1] * (x.dim() - 1)))}
    exec(fn_src, ns)
    ref_fn = ns["compute_distribution_matching_loss"]

    cases = {}
    for tag, dtype, cfg in [("fp32_cfg5", torch.float32, 5.0), ("bf16_cfg5", torch.bfloat16, 5.0), ("fp32_cfg1", torch.float32, 1.0)]:
        gg = torch.Generator().manual_seed(11)
        z = torch.randn(4, 32, 16, 16, generator=gg).to(dtype)
        x0 = torch.randn(z.shape, generator=gg).to(dtype)
        t = torch.rand(4, generator=gg).to(dtype)
        outs = {k: torch.randn(z.shape, generator=gg).to(dtype) for k in ("Tc", "Tu", "Sc", "Su")}

        class Self:
            pass
        s = Self()
        s.args = Args()
        s.args.dmd_cfg_scale = cfg
        s.transport = types.SimpleNamespace(sample=lambda x1: (t, x0, x1), path_sampler=RP.ICPlan())
        s.base_model = lambda xt, tt, y: outs["Tu"] if int(y[0]) == 1000 else outs["Tc"]
        s.sit_wo_ddp = lambda xt, tt, y: outs["Su"] if int(y[0]) == 1000 else outs["Sc"]
        zz = z.clone().requires_grad_(True)
        labels = torch.zeros(4, dtype=torch.long)
        loss, log = ref_fn(s, zz, labels)
        (dz,) = torch.autograd.grad(loss, zz)
        _, xt, _ = RP.ICPlan().plan(t, x0, z)
        cases[tag] = dict(z=z, x0=x0, t=t, xt=xt, cfg=cfg, loss=loss.detach(), dz=dz, gnorm=log["dmd_gradient_norm"], **outs)
    # --- config 1 (toy_example_2d/dmd.py): the "dmd" branch of DMDLossFunction.compute_distribution_matching_loss
    # (:320-360: no CFG, no |p_real| normaliser), learnable points (1536, 2) from create_learnable_points (:139-145)
    tsrc = open(f"{REF}/toy_example_2d/dmd.py").read()
    m = re.search(r"    def compute_distribution_matching_loss\(self.*?\n(?=        elif loss_type == \"score_teacher\")", tsrc, re.S)
    toy_src = textwrap.dedent(m.group(0)) + "    return loss, grad\n"
    ns2 = {"torch": torch, "expand_t_like_x": lambda t, x: t.view(t.size(0), *([1] * (x.dim() - 1)))}
    exec(toy_src, ns2)
    toy_fn = ns2["compute_distribution_matching_loss"]
    torch.manual_seed(42)
    points = torch.rand(1536, 2) * 3.0 - 1.5
    gg = torch.Generator().manual_seed(21)
    t = torch.rand(1536, generator=gg)
    x0 = torch.randn(1536, 2, 1, 1, generator=gg)
    vT = torch.randn(1536, 2, 1, 1, generator=gg)
    vS = torch.randn(1536, 2, 1, 1, generator=gg)

    class ToyArgs:
        t0, t1, dmd_loss_type = 0.0, 1.0, "dmd"

    class Self2:
        pass
    s2 = Self2()
    s2.args = ToyArgs()
    s2.transport = types.SimpleNamespace(sample=lambda x1: (t, x0, x1), path_sampler=RP.ICPlan())
    s2.base_model = lambda xt, tt, y: vT
    s2.sit_wo_ddp = lambda xt, tt, y: vS
    pts = points.clone().requires_grad_(True)
    loss, grad = toy_fn(s2, pts, torch.zeros(1536, dtype=torch.long))
    (dpts,) = torch.autograd.grad(loss, pts)
    cases["toy_fp32"] = dict(points=points, x0=x0, t=t, vT=vT, vS=vS, loss=loss.detach(), dpoints=dpts, grad=grad.detach())
    torch.save(cases, os.path.join(OUT, "dmd.pt"))
    # --- N1: LightningDiT (tiny config + manifest of Mini/1) from the reference, timm / fairscale replaced by import shims
    class _PE(torch.nn.Module):          # timm.models.vision_transformer.PatchEmbed as the reference uses it
        def __init__(self, img_size, patch_size, in_chans, embed_dim, bias=True):
            super().__init__()
            self.patch_size = (patch_size, patch_size)
            self.num_patches = (img_size // patch_size) ** 2
            self.proj = torch.nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

        def forward(self, x):
            return self.proj(x).flatten(2).transpose(1, 2)

    class _Mlp(torch.nn.Module):
        def __init__(self, in_features, hidden_features, act_layer, drop=0):
            super().__init__()
            self.fc1, self.act, self.fc2 = torch.nn.Linear(in_features, hidden_features), act_layer(), torch.nn.Linear(hidden_features, in_features)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))
    for name, attrs in {"timm": {}, "timm.models": {}, "timm.models.vision_transformer": {"PatchEmbed": _PE, "Mlp": _Mlp},
                        "fairscale": {}, "fairscale.nn": {}, "fairscale.nn.model_parallel": {},
                        "fairscale.nn.model_parallel.initialize": {"get_model_parallel_world_size": lambda: 1},
                        "fairscale.nn.model_parallel.layers": {"ColumnParallelLinear": object, "ParallelEmbedding": object,
                                                               "RowParallelLinear": object}}.items():
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(mod, k, v)
            sys.modules[name] = mod
    os.environ["TORCHDYNAMO_DISABLE"] = "1"
    from diffusion.lightningdit import lightningdit as RD
    torch.manual_seed(3)
    dit = RD.LightningDiT(input_size=4, patch_size=1, in_channels=8, hidden_size=64, depth=2, num_heads=2, num_classes=10).eval()
    gg = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for n, p_ in dit.named_parameters():          # the zero-initialised layers would make v == 0
            if "adaLN_modulation" in n or "final_layer" in n or "norm" in n:
                p_.copy_((1.0 if n.endswith("norm1.weight") or n.endswith("norm2.weight") or "norm_final" in n or "q_norm" in n or "k_norm" in n else 0.0)
                         + 0.05 * torch.randn(p_.shape, generator=gg))
    x = torch.randn(3, 8, 4, 4, generator=gg)
    t = torch.rand(3, generator=gg)
    y = torch.tensor([1, 7, 10])
    with torch.no_grad():
        v = dit(x, t, y)
        v_unc = dit(x, t, torch.full_like(y, 10))
    mini = RD.LightningDiT_Mini_1(input_size=16, in_channels=32, num_classes=1000)
    torch.save(dict(sd={k: v_.clone() for k, v_ in dit.state_dict().items()}, x=x, t=t, y=y, v=v, v_unc=v_unc,
                    manifest_mini={k: tuple(v_.shape) for k, v_ in mini.state_dict().items()}), os.path.join(OUT, "dit.pt"))
    for f in ("flux_ae.pt", "lpips.pt", "dmd.pt", "dit.pt"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
