"""Pins oracle/dmvae_oracle.py to outputs of the real reference (tests/golden/*.pt, made by make_golden.py from
/root/reference).  fp32 on CPU; tolerance 1e-5 relative (different op order only), bf16 DMD case 1e-2 (normaliser
rounding, see test_kernels_gpu.py)."""
import os

import pytest
import torch

from oracle import dmvae_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(G, "flux_ae.pt"), weights_only=True)


@pytest.mark.parametrize("name,fn", [
    ("resnet_64_64", lambda sd, x: O.resnet_block({"b." + k: v for k, v in sd.items()}, "b", x)),
    ("resnet_64_32", lambda sd, x: O.resnet_block({"b." + k: v for k, v in sd.items()}, "b", x)),
    ("attn_64", lambda sd, x: O.attn_block({"b." + k: v for k, v in sd.items()}, "b", x)),
    ("upsample_32", lambda sd, x: O.upsample({"b." + k: v for k, v in sd.items()}, "b", x)),
    ("downsample_32", lambda sd, x: O.downsample({"b." + k: v for k, v in sd.items()}, "b", x)),
])
def test_blocks(fx, name, fn):
    c = fx[name]
    sd = {k: v.clone().requires_grad_(True) for k, v in c["sd"].items()}
    x = c["x"].clone().requires_grad_(True)
    y = fn(sd, x)
    assert rel(y, c["y"]) < 1e-5
    y.backward(c["dy"])
    assert rel(x.grad, c["dx"]) < 1e-5
    for k, g in c["dparams"].items():
        # absolute floor: d/d(k.bias) of softmax attention is identically zero, both sides hold 1e-8 noise there
        assert (sd[k].grad - g).norm() <= 1e-4 * g.norm() + 1e-6, k


def test_decoder_encoder(fx):
    c = fx["decoder_tiny"]
    sd = {k: v.clone().requires_grad_(True) for k, v in c["sd"].items()}
    z = c["z"].clone().requires_grad_(True)
    y = O.decoder_forward(sd, z)
    assert rel(y, c["y"]) < 1e-5
    y.backward(c["dy"])
    assert rel(z.grad, c["dz"]) < 1e-5
    assert rel(sd["conv_out.weight"].grad, c["d_conv_out"]) < 1e-5
    assert rel(sd["mid.block_1.conv1.weight"].grad, c["d_mid_conv1"]) < 1e-5
    assert rel(sd["conv_in.0.conv.weight"].grad, c["d_stem"]) < 1e-5
    t = fx["decoder_tiny_tokens"]
    assert rel(O.decoder_forward(c["sd"], t["z"]), t["y"]) < 1e-5
    e = fx["encoder_tiny"]
    assert rel(O.encoder_forward(e["sd"], e["x"]), e["y"]) < 1e-5


def test_state_factories_match_reference_manifest(fx):
    d = O.make_decoder_state(z_channels=32)
    assert {k: tuple(v.shape) for k, v in d.items()} == fx["manifest"]["decoder"]
    e = O.make_encoder_state(z_channels=16)
    assert {k: tuple(v.shape) for k, v in e.items()} == fx["manifest"]["encoder"]


def test_lpips_distance_and_manifest():
    c = torch.load(os.path.join(G, "lpips.pt"), weights_only=True)
    lin_ws = [c["lin_sd"][f"lin{k}.model.1.weight"].flatten() for k in range(5)]
    f0 = [t.float() for t in c["f0"]]
    f1 = [t.float().requires_grad_(True) for t in c["f1"]]
    val = O.lpips_distance(f0, f1, lin_ws)
    assert abs(val.item() - c["val_feats"].item()) < 1e-5 * abs(c["val_feats"].item())
    val.backward()
    for g, ref in zip(f1, c["df1"]):
        assert rel(g.grad, ref.float()) < 1e-5
    sd = O.make_lpips_state()
    assert {k: tuple(v.shape) for k, v in sd.items()} == c["manifest"]


def test_lpips_end_to_end_with_regenerated_vgg():
    """End-to-end (ScalingLayer + VGG16 + distance) against the reference value, regenerating torchvision's seeded
    vgg16(weights=None) instead of storing 14.7 M weights; skipped if this torch/torchvision seeds differently."""
    tv = pytest.importorskip("torchvision")
    c = torch.load(os.path.join(G, "lpips.pt"), weights_only=True)
    torch.manual_seed(c["vgg_seed"])
    feats = tv.models.vgg16(weights=None).features.state_dict()
    if abs(feats["0.weight"].double().sum().item() - c["vgg_w0_sum"]) > 1e-6:
        pytest.skip("torchvision init stream differs from the one the fixture was generated with")
    sd = dict(c["lin_sd"])
    slice_of = lambda i: 1 + sum(i > t for t in (3, 8, 15, 22))
    for k, v in feats.items():
        i = int(k.split(".")[0])
        sd[f"net.slice{slice_of(i)}.{k}"] = v
    b = c["b"].clone().requires_grad_(True)
    val = O.lpips_forward(sd, c["a"], b)
    assert abs(val.item() - c["val"].item()) < 1e-4 * abs(c["val"].item())
    val.backward()
    assert rel(b.grad, c["db"]) < 1e-3


@pytest.mark.parametrize("tag,tol", [("fp32_cfg5", 1e-5), ("fp32_cfg1", 1e-5), ("bf16_cfg5", 1e-2)])
def test_dmd(tag, tol):
    c = torch.load(os.path.join(G, "dmd.pt"), weights_only=True)[tag]
    xt = O.dmd_mix_xt(c["z"], c["x0"], c["t"])
    assert torch.equal(xt, c["xt"])
    loss, gnorm, dz = O.dmd_loss(c["z"], xt, c["t"], c["Tc"], c["Sc"], c["Tu"], c["Su"], c["cfg"], True)
    assert abs(loss.item() - c["loss"].item()) <= tol * abs(c["loss"].item())
    assert abs(gnorm.item() - c["gnorm"]) <= tol * abs(c["gnorm"])
    assert rel(dz, c["dz"].float()) <= tol


def test_toy_2d_dmd_config():
    """BASELINE configs[0]: toy_example_2d/dmd.py "dmd" branch (no CFG, no normaliser) on the (1536, 2) learnable points."""
    c = torch.load(os.path.join(G, "dmd.pt"), weights_only=True)["toy_fp32"]
    z = c["points"].view(1536, 2, 1, 1)
    xt = O.dmd_mix_xt(z, c["x0"], c["t"])
    loss, gnorm, dz = O.dmd_loss(z, xt, c["t"], c["vT"], c["vS"], None, None, 1.0, normalize=False)
    assert abs(loss.item() - c["loss"].item()) <= 1e-6 * abs(c["loss"].item())
    assert rel(dz.view(1536, 2), c["dpoints"]) < 1e-6


def test_latents_to_spatial_bit_exact():
    x = torch.randn(2, 16, 5)
    y = O.latents_to_spatial(x)
    assert y.shape == (2, 5, 4, 4) and torch.equal(y[1, 3, 2, 1], x[1, 9, 3])


# ---------------------------------------------------------------- production-size pins (tests/golden/make_golden_full.py)
FULL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decoder_full.pt")


def test_decoder_production_size_against_reference():
    """The oracle's Decoder at production size (49.6 M parameters regenerated from the fixture's seed) against outputs and
    gradients the REAL reference module produced for the same weights and tokens."""
    c = torch.load(FULL, weights_only=True)["decoder_full"]
    sd = {k: v.requires_grad_(True) for k, v in O.make_decoder_state(z_channels=32, seed=c["seed"], randomize_affine=True).items()}
    g = torch.Generator().manual_seed(c["dy_seed"])
    z = torch.randn(1, 256, 32, generator=g)
    assert torch.equal(z, c["z"])
    z.requires_grad_(True)
    y = O.decoder_forward(sd, z)
    assert y.shape == (1, 3, 256, 256)
    assert rel(y, c["y"].float()) < 1e-3                      # stored as fp16
    dy = torch.randn(y.shape, generator=g) / y.numel()
    y.backward(dy)
    assert rel(z.grad, c["dz"]) < 1e-4
    for name, ref in c["dparams"].items():
        assert rel(sd[name].grad, ref) < 1e-4, name


@pytest.mark.parametrize("cin,cout", [(512, 512), (512, 256), (256, 256), (256, 128), (128, 128)])
def test_resnet_block_decoder_channel_configs(cin, cout):
    """ResnetBlock (flux_ae.py:55-82) at the decoder's remaining channel configurations, weights regenerated from the seed."""
    c = torch.load(FULL, weights_only=True)[f"resnet_{cin}_{cout}"]
    gb = torch.Generator().manual_seed(c["seed"])
    shapes = [("norm1.weight", (cin,)), ("norm1.bias", (cin,)), ("conv1.weight", (cout, cin, 3, 3)), ("conv1.bias", (cout,)),
              ("norm2.weight", (cout,)), ("norm2.bias", (cout,)), ("conv2.weight", (cout, cout, 3, 3)), ("conv2.bias", (cout,))]
    if cin != cout:
        shapes += [("nin_shortcut.weight", (cout, cin, 1, 1)), ("nin_shortcut.bias", (cout,))]
    sd = {}
    for n, s in shapes:                                       # the generation rule of make_golden_full.py, same order
        if len(s) > 1:
            sd["b." + n] = torch.randn(s, generator=gb) * 0.02
        elif "norm" in n and n.endswith("weight"):
            sd["b." + n] = 1 + 0.1 * torch.randn(s, generator=gb)
        else:
            sd["b." + n] = 0.05 * torch.randn(s, generator=gb)
    x = torch.randn(1, cin, c["hw"], c["hw"], generator=gb, requires_grad=True)
    y = O.resnet_block(sd, "b", x)
    dy = torch.randn(y.shape, generator=gb)
    (dx,) = torch.autograd.grad(y, x, dy)
    assert rel(y, c["y"].float()) < 1e-3 and rel(dx, c["dx"].float()) < 1e-3      # fixtures stored as fp16


@pytest.mark.parametrize("tag,tol", [("fp32", 1e-5), ("bf16", 1e-2)])
def test_dmd_zero_normaliser_nan_to_num(tag, tol):
    """A sample with w_b = mean|p_real| = 0 and a zero numerator: grad = 0/0 = NaN -> 0 (train_dmd.py:222-224), pinned on the real
    reference method (tests/golden/make_golden_dmd_edge.py).  The sample contributes nothing; the others are unaffected."""
    c = torch.load(os.path.join(G, "dmd_edge.pt"), weights_only=True)[tag]
    xt = O.dmd_mix_xt(c["z"], c["x0"], c["t"])
    loss, gnorm, dz = O.dmd_loss(c["z"], xt, c["t"], c["Tc"], c["Sc"], None, None, 1.0, True)
    assert torch.isfinite(loss) and torch.isfinite(dz).all()
    assert float(dz[0].abs().max()) == 0.0 and float(c["dz"][0].abs().max()) == 0.0
    assert abs(loss.item() - c["loss"].item()) <= tol * abs(c["loss"].item())
    assert abs(gnorm.item() - c["gnorm"]) <= tol * abs(c["gnorm"])
    assert rel(dz, c["dz"].float()) <= tol


def test_encoder_production_size_against_reference():
    """The oracle's Encoder at production size (models/flux_ae.py:110-181; BASELINE configs[4]'s module) against the real reference
    module's output and gradients for weights regenerated from the fixture's seed."""
    c = torch.load(FULL, weights_only=True)["encoder_full"]
    sd = {k: v.requires_grad_(True) for k, v in O.make_encoder_state(z_channels=16, seed=c["seed"], randomize_affine=True).items()}
    g = torch.Generator().manual_seed(c["x_seed"])
    x = (torch.rand(1, 3, 256, 256, generator=g) * 2 - 1).requires_grad_(True)
    h = O.encoder_forward(sd, x)
    assert h.shape == c["y"].shape == (1, 32, 32, 32)
    assert rel(h, c["y"]) < 1e-4
    dh = torch.randn(h.shape, generator=g)
    h.backward(dh)
    assert rel(x.grad, c["dx"].float()) < 2e-3                # stored as fp16
    for name, ref in c["dparams"].items():
        assert rel(sd[name].grad, ref) < 1e-4, name
