"""CPU-only checks: the C-ABI library loads and exports everything include/dmvae_b200.h declares, the drop-in modules
expose the reference's state_dict surface, the product path refuses to run without CUDA, and the data-parallel
gradient exchange is correct under gloo with world_size 2."""
import copy
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    from dmvae_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "dmvae_b200.h")).read()
    declared = set(re.findall(r"\b(dmvae_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared - {"dmvae_last_error"} == set(_lib.SIGNATURES), "ctypes table and header disagree"
    ver = int(re.search(r"#define\s+DMVAE_ABI_VERSION\s+(\d+)", hdr).group(1))
    assert lib.dmvae_abi_version() == ver == _lib.ABI_VERSION, "header, library and ctypes table must agree on the ABI version"


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dmvae_b200 import DmvaeError, losses
    with pytest.raises(DmvaeError):
        losses.l1_l2_loss(torch.zeros(4), torch.zeros(4))
    from dmvae_b200.autoencoder import ResnetBlock
    with pytest.raises(DmvaeError):
        ResnetBlock(32, 32)(torch.zeros(1, 32, 4, 4))


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dmvae_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_state_dict_surface_matches_reference_manifest():
    from dmvae_b200.autoencoder import Decoder, Encoder
    from dmvae_b200.lpips import LPIPS
    man = torch.load(os.path.join(G, "flux_ae.pt"), weights_only=True)["manifest"]
    dec = Decoder(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, resolution=256, z_channels=16)
    dec.post_init(32)
    assert {k: tuple(v.shape) for k, v in dec.state_dict().items()} == man["decoder"]
    assert list(dec.state_dict().keys()) == list(man["decoder"].keys())          # same order too
    enc = Encoder(resolution=256, in_channels=3, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16)
    assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == man["encoder"]
    assert all(v.dtype == torch.float32 for v in dec.state_dict().values())
    lman = torch.load(os.path.join(G, "lpips.pt"), weights_only=True)["manifest"]
    lp = LPIPS(ckpt_path=None, pretrained_vgg=False)
    assert {k: tuple(v.shape) for k, v in lp.state_dict().items()} == lman
    assert dec.get_last_layer() is dec.conv_out.weight
    c = copy.deepcopy(dec)
    assert c.conv_out.weight is not dec.conv_out.weight


def test_vae_surface():
    from dmvae_b200.vae import VAE
    vae = VAE(z_channels=32, model_size="base")
    keys = set(vae.state_dict().keys())
    for k in ("encoder.model.cls_token", "encoder.model.pos_embed", "encoder.model.patch_embed.proj.weight",
              "encoder.model.blocks.0.attn.qkv.weight", "encoder.model.blocks.11.ls2.gamma", "encoder.model.norm.bias",
              "encoder.scale.mean", "encoder.de_scale.std", "bottle_neck.mlp.0.weight", "bottle_neck.mlp.2.bias",
              "decoder.conv_in.0.conv.weight", "decoder.conv_in.1.weight", "decoder.conv_out.weight"):
        assert k in keys, k
    assert vae.bottle_neck.get_last_layer().shape == (32, 2048)
    assert vae.encoder.model.pos_embed.shape == (1, 257, 768)
    # encoder + bottleneck are plain PyTorch and run anywhere
    with torch.no_grad():
        t = vae.bottle_neck(vae.encoder(torch.zeros(1, 3, 256, 256)))
    assert t.shape == (1, 256, 32)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dmvae_b200.train import GradArena
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
arena = GradArena(net.parameters(), chunks=3)
g = torch.Generator().manual_seed(100)
x_all = torch.randn(8, 8, generator=g)
arena.zero()
net(x_all[rank * 4:(rank + 1) * 4]).square().mean().backward()      # each rank: its shard of the images
assert arena.overlap and any(arena._launched), "chunks must have been launched from the autograd hooks during backward"
arena.allreduce()
sharded = arena.flat.clone()
# the same 8 images on one rank, without touching .grad (no hooks, no exchange)
ref = torch.autograd.grad(net(x_all).square().mean(), list(net.parameters()))
ref = arena.flatten(ref)                      # arena layout: every parameter padded to 8 elements
assert torch.allclose(sharded, ref, rtol=1e-5, atol=1e-7), (sharded - ref).abs().max()
assert all(p.grad.data_ptr() >= arena.flat.data_ptr() for p in net.parameters())
# a second step reuses the arena: zero(), backward, allreduce
arena.zero()
net(x_all[rank * 4:(rank + 1) * 4]).square().mean().backward()
arena.allreduce()
assert torch.allclose(arena.flat, ref, rtol=1e-5, atol=1e-7)
# third step: a Function that accumulates its parameter gradients straight into the arena (what the CUDA conv / GroupNorm
# Functions do inside GradArena.direct()) -- the chunk bookkeeping must see those parameters exactly once, via notify()
from dmvae_b200 import ops


class DirectLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.params = (w, b)
        ctx.save_for_backward(x, w)
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dw, db = dy.t() @ x, dy.sum(0)
        sw, sb = ops._grad_slot(ctx.params[0]), ops._grad_slot(ctx.params[1])
        if sw is not None and sb is not None:
            sw.add_(dw); sb.add_(db)
            ops._grad_done(ctx.params[0]); ops._grad_done(ctx.params[1])
            dw = db = None
        return dy @ w, dw, db


def fwd(xs):
    h = torch.tanh(DirectLinear.apply(xs, net[0].weight, net[0].bias))
    return DirectLinear.apply(h, net[2].weight, net[2].bias)


assert ops._grad_slot(net[0].weight) is None, "direct mode must be off outside the context"
arena.zero()
with arena.direct():
    assert ops._grad_slot(net[0].weight).data_ptr() == net[0].weight.grad.data_ptr()
    fwd(x_all[rank * 4:(rank + 1) * 4]).square().mean().backward()
assert all(c == 0 for c in arena._pending) and all(arena._launched), (arena._pending, arena._launched)
arena.allreduce()
assert torch.allclose(arena.flat, ref, rtol=1e-5, atol=1e-7), (arena.flat - ref).abs().max()
# outside the context the same Function hands its gradients to autograd (AccumulateGrad + hooks), same result
arena.zero()
fwd(x_all[rank * 4:(rank + 1) * 4]).square().mean().backward()
arena.allreduce()
assert torch.allclose(arena.flat, ref, rtol=1e-5, atol=1e-7)
# CUDA-graph replay with the exchange issued after the replay: between replays only the captured flat.zero_() kernel runs,
# zero() is never called from Python.  Every step must still launch every chunk (round-1 bug: only the first replay exchanged).
arena.hooks_enabled = False
for it in range(3):
    arena.flat.zero_()
    n0 = arena.exchanges
    net(x_all[rank * 4:(rank + 1) * 4]).square().mean().backward()
    assert arena.exchanges == n0, "hooks are off: nothing may be launched from backward"
    arena.allreduce()
    assert arena.exchanges - n0 == len(arena.bounds), (it, arena.exchanges - n0)
    assert torch.allclose(arena.flat, ref, rtol=1e-5, atol=1e-7), it
# pipelined exchange (TokenizerTrainer with several ranks): step k's local gradients stay in the arena; step k+1 begins with
# launch_all(), does other work (the frozen-encoder forward) while the all-reduces run, then finish() + the weight update, and only
# then zero()s the arena for its own backward.  Two ranks on half batches must train exactly like one process on the full batch.
import copy
solo = copy.deepcopy(net)
lr = 0.1
pending = False
for it in range(3):
    xs = torch.randn(8, 8, generator=g)
    if pending:                                   # ---- start of step `it`: finish step it-1
        n0 = arena.exchanges
        arena.launch_all()
        assert arena.exchanges - n0 == len(arena.bounds)
        busy = torch.tanh(xs).sum()               # stands for the encoder forward that overlaps the exchange
        arena.finish()
        assert not any(arena._launched) and arena._handles == [], "finish() must leave the bookkeeping reset"
        with torch.no_grad():
            for p_ in net.parameters():
                p_.sub_(lr * p_.grad)
    arena.zero()
    net(xs[rank * 4:(rank + 1) * 4]).square().mean().backward()
    assert not any(arena._launched), "pipelined mode: nothing is exchanged from backward"
    pending = True
    gs = torch.autograd.grad(solo(xs).square().mean(), list(solo.parameters()))       # the sequential single-process run
    with torch.no_grad():
        for p_, g_ in zip(solo.parameters(), gs):
            p_.sub_(lr * g_)
arena.allreduce()                                 # flush(): the last step's update
with torch.no_grad():
    for p_ in net.parameters():
        p_.sub_(lr * p_.grad)
for p_, q_ in zip(net.parameters(), solo.parameters()):
    assert torch.allclose(p_, q_, rtol=1e-5, atol=1e-7), (p_ - q_).abs().max()
dist.destroy_process_group()
print("OK", rank)
'''


def test_grad_arena_allreduce_matches_single_process_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and "OK" in out, out


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): exactly one line on stdout, valid JSON, with the
    contract's keys; everything else (banners, warnings) must go to stderr."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, TORCHDYNAMO_DISABLE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


@pytest.mark.parametrize("case", ["decoder_std", "decoder_xavier", "encoder_std"])
def test_initial_weights_reproduce_reference_bit_for_bit(case):
    """Same seed => the same initial weights as the reference's modules + models/init_param.py (tests/golden/init.pt holds per-tensor
    checksums produced by the real reference): module registration order, RNG consumption and the init rules all match."""
    from dmvae_b200 import autoencoder
    from dmvae_b200.vae import init_weights
    c = torch.load(os.path.join(G, "init.pt"), weights_only=True)[case]
    mod = getattr(autoencoder, c["kind"])(**c["kw"])
    if c["post"] is not None:
        mod.post_init(c["post"])
    torch.manual_seed(c["seed"])
    init_weights(mod, c["arg"])
    sd = mod.state_dict()
    assert list(sd.keys()) == list(c["sums"].keys())
    for k, (s, a, head) in c["sums"].items():
        v = sd[k]
        assert float(v.double().sum()) == s and float(v.double().abs().sum()) == a and torch.equal(v.reshape(-1)[:4], head), k


@pytest.mark.parametrize("tag", ["fp32_shift1.0", "bf16_shift1.0", "fp32_shift3.0", "bf16_shift3.0"])
def test_sample_t_x0_reproduces_reference_transport_sample(tag):
    """train.sample_t_x0 vs the real Transport.sample (transport.py:105-116, tests/golden/transport.pt): same draws in the same order
    (x0 from the default generator, then t on the CPU generator, cast to the latents' dtype, then the time shift)."""
    from dmvae_b200.train import sample_t_x0
    c = torch.load(os.path.join(G, "transport.pt"), weights_only=True)[tag]
    dtype = torch.float32 if c["dtype"] == "fp32" else torch.bfloat16
    torch.manual_seed(123)
    x1 = torch.randn(6, 32, 16, 16).to(dtype)
    torch.manual_seed(77)
    t, x0 = sample_t_x0(x1, c["shift"])
    assert t.dtype == dtype and torch.equal(t, c["t"])
    assert float(x0.double().sum()) == c["x0_sum"] and torch.equal(x0.reshape(-1)[:8], c["x0_head"])


def test_abi_rejects_null_pointers_without_touching_the_device():
    """Error convention of the C ABI: every compute entry point validates its arguments before any CUDA call, returns a
    negative status and leaves a message naming itself in dmvae_last_error() -- checkable on a box without a GPU."""
    import ctypes
    from dmvae_b200 import _lib
    lib = _lib.load()
    lib.dmvae_last_error.restype = ctypes.c_char_p
    checked = 0
    for name, sig in _lib.SIGNATURES.items():
        if ctypes.c_void_p not in sig or name in ("dmvae_check_device",):
            continue
        args = []
        for ty in sig[:-1]:                                   # the last slot is the stream
            if ty is ctypes.c_void_p:
                args.append(None)
            elif ty is ctypes.c_float or ty is ctypes.c_double:
                args.append(ty(0.5))
            else:
                args.append(ty(1))
        rc = getattr(lib, name)(*args, None)
        msg = lib.dmvae_last_error().decode()
        assert rc < 0, f"{name} accepted null pointers (rc={rc})"
        assert name.replace("dmvae_", "").split("_fwd")[0].split("_bwd")[0][:6] in msg or "null" in msg or "pointer" in msg, (name, msg)
        checked += 1
    assert checked >= 25


def test_grad_arena_layout_tap_major_views_and_padding():
    """Arena layout contract (dmvae_b200/train_arena.py): every parameter starts on an 8-element boundary, 3x3 conv weights are stored
    tap-major ([tap][Cout][Cin]) behind a strided (Cout, Cin, 3, 3) view, everything else keeps its natural order; flatten() lays
    per-parameter tensors out the same way.  Pure host logic: runs without a GPU."""
    from dmvae_b200.train_arena import GradArena, arena_view, tap_major
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(5, 7, 3), torch.nn.Conv2d(7, 3, 1), torch.nn.GroupNorm(1, 3), torch.nn.Linear(3, 2))
    arena = GradArena(net.parameters())
    assert all(off % 8 == 0 for off in arena.offsets) and arena.flat.numel() % 8 == 0
    w3, w1 = net[0].weight, net[1].weight
    assert tap_major(w3) and not tap_major(w1) and not tap_major(net[3].weight)
    assert w3.grad.shape == w3.shape and w3.grad.stride() == (5, 1, 3 * 7 * 5, 7 * 5)
    assert w1.grad.is_contiguous() and net[0].bias.grad.is_contiguous()
    vals = [torch.randn(p.shape) for p in arena.params]
    for p, v in zip(arena.params, vals):
        p.grad.copy_(v)
    flat = arena.flatten(vals)
    assert torch.equal(flat, arena.flat)
    off = arena.offsets[0]
    assert torch.equal(arena.flat[off:off + w3.numel()].view(9, 7, 5), vals[0].permute(2, 3, 0, 1).reshape(9, 7, 5))
    used = torch.zeros_like(arena.flat, dtype=torch.bool)
    for p, off in zip(arena.params, arena.offsets):
        used[off:off + p.numel()] = True
    assert float(arena.flat[~used].abs().sum()) == 0.0               # padding stays zero
    assert arena_view(arena.flat, arena.offsets[1], net[0].bias).data_ptr() == net[0].bias.grad.data_ptr()
    # backward accumulates through the strided view like through any .grad
    arena.zero()
    x = torch.randn(2, 5, 6, 6)
    ref = torch.autograd.grad(net[:3](x).square().mean(), [w3])[0]
    net[:3](x).square().mean().backward()
    assert torch.allclose(w3.grad, ref) and w3.grad.data_ptr() == arena.flat.data_ptr() + 4 * arena.offsets[0]


def test_zero_pool_tensors_have_their_own_version_counters():
    """ops.ZeroPool hands out accumulators that alias the tail of the gradient arena's allocation.  They must NOT be views of it:
    views share one autograd version counter, so the arena's memset or an AccumulateGrad ``+=`` into any gradient slot would
    invalidate every pool tensor saved for backward (host logic only, exercised on a CPU buffer)."""
    from dmvae_b200 import ops
    storage = torch.zeros(1000 + 256, dtype=torch.float32)
    flat, pool = storage[:1000], ops.ZeroPool(storage[1000:])

    class Scale(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            acc = ops.small_zeros((3, 2), torch.float64, x.device)
            acc += x.sum().double()
            ctx.save_for_backward(acc)
            return x * 2

        @staticmethod
        def backward(ctx, g):
            (acc,) = ctx.saved_tensors
            return g * 2 + acc[0, 0].float()

    x = torch.ones(4, requires_grad=True)
    outside = ops.small_zeros((5,), torch.float32, x.device)
    with pool:
        y = Scale.apply(x)
        b = ops.small_zeros((5,), torch.float32, x.device)
        assert b.data_ptr() == storage.data_ptr() + 4000 + 64 and pool.off == 96          # 32-byte granules
        b.fill_(3.0)
        flat.add_(1.0)                     # what AccumulateGrad / the memset do to the shared allocation
        storage[1000:].add_(0.0)
        y.sum().backward()                 # would raise "modified by an inplace operation" for a view
        too_big = ops.small_zeros((4096,), torch.float32, x.device)
    assert torch.equal(x.grad, torch.full((4,), 6.0))
    assert not storage.data_ptr() <= outside.data_ptr() < storage.data_ptr() + 4 * storage.numel()
    assert not storage.data_ptr() <= too_big.data_ptr() < storage.data_ptr() + 4 * storage.numel()
    assert ops.ZeroPool.current is None
    assert torch.equal(storage[1016:1021], torch.full((5,), 3.0))
    storage.zero_(); pool.reset()
    with pool:
        again = ops.small_zeros((3, 2), torch.float64, x.device)
    assert again.data_ptr() == storage.data_ptr() + 4000 and float(again.abs().sum()) == 0.0


@pytest.mark.parametrize("tag", ["w0.5_bcr4", "w0.1_l2_bcr0"])
def test_discriminator_turn_against_reference_golden(tag):
    """VAELossFunction.forward_discriminator is host-side composition of stock ops (no library kernel), so it is pinned here,
    on the CPU, to the reference method's outputs (train_dmd.py:265-285, tests/golden/make_golden_gan.py); the generator turn,
    whose L1 / L2 come from the fused kernel, is pinned by the GPU suite against the same fixture."""
    import importlib.util
    from dmvae_b200.train import LossConfig, VAELossFunction
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_gan", os.path.join(here, "golden", "make_golden_gan.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    fx = torch.load(os.path.join(here, "golden", "gan.pt"), weights_only=False)[tag]
    last = torch.nn.Conv2d(6, 3, 3, padding=1)
    last.load_state_dict(fx["last_sd"])
    disc = G.TinyDisc()
    disc.load_state_dict(fx["disc_sd"])
    lf = VAELossFunction(LossConfig(disc_weight=fx["disc_weight"], bcr=fx["bcr"]), disc=disc, last_layer=last.weight,
                         aug=lambda x, fade=0.0: x, bcr_aug=lambda x, fade=0.0: x.flip(-1))
    with torch.no_grad():
        recon = last(fx["h"])
    d_loss, d_log = lf.forward_discriminator(fx["images"], recon)
    grads = torch.autograd.grad(d_loss, list(disc.parameters()))
    assert disc.training and all(p.requires_grad for p in disc.parameters())
    assert abs(d_loss.item() - fx["d_loss"].item()) < 1e-6
    for k in ("d_loss", "acc_real", "acc_fake") + (("bcr_loss",) if fx["bcr"] > 0 else ()):
        assert abs(float(d_log[k]) - fx["d_log"][k]) < 1e-5 * max(abs(fx["d_log"][k]), 1.0), k
    for (n, _), g in zip(disc.named_parameters(), grads):
        assert torch.allclose(g, fx["d_params"][n], rtol=1e-5, atol=1e-7), n


def test_shutdown_distributed_without_process_group_is_a_noop():
    from dmvae_b200.train import shutdown_distributed

    class T:
        released = False

        def release_graphs(self):
            self.released = True
    t = T()
    shutdown_distributed(t, None)
    assert t.released
