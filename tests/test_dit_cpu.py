"""LightningDiT restatement vs the reference's own module: outputs on a tiny config (fp32, 1e-5), the
state_dict surface of LightningDiT-Mini/1, and the batched cond+uncond forward."""
import os

import torch

from dmvae_b200.dit import LightningDiT, LightningDiT_Mini_1

G = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def test_tiny_dit_matches_reference_outputs():
    c = torch.load(os.path.join(G, "dit.pt"), weights_only=True)
    dit = LightningDiT(input_size=4, patch_size=1, in_channels=8, hidden_size=64, depth=2, num_heads=2, num_classes=10).eval()
    dit.load_state_dict(c["sd"], strict=True)                 # reference-keyed state_dict loads as is
    with torch.no_grad():
        v = dit(c["x"], c["t"], c["y"])
        vc, vu = dit.forward_cond_uncond(c["x"], c["t"], c["y"])
    assert rel(v, c["v"]) < 1e-5
    assert rel(vc, c["v"]) < 1e-5 and rel(vu, c["v_unc"]) < 1e-5


def test_mini_state_dict_surface():
    c = torch.load(os.path.join(G, "dit.pt"), weights_only=True)
    mini = LightningDiT_Mini_1(input_size=16, in_channels=32, num_classes=1000)
    assert {k: tuple(v.shape) for k, v in mini.state_dict().items()} == c["manifest_mini"]
    x = torch.randn(2, 32, 16, 16)
    with torch.no_grad():
        assert mini(x, torch.rand(2), torch.tensor([3, 1000])).abs().max() == 0          # zero-initialised head
