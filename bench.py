#!/usr/bin/env python
"""bench.py -- VAE(+DMD) training images/sec at 256x256, bf16, on N B200s.

A "step" is one pass of the hot path over one synthetic batch: the VAE-pretrain step of train_tokenizer.py:403-437
(BASELINE.json configs[1]): frozen DINOv2 ViT-L/16 encoder -> bottleneck MLP -> flux Decoder forward, L1 + LPIPS(VGG16)
reconstruction loss, backward through the decoder, gradient allreduce, clip, AdamW, EMA -- per-GPU batch 16
(scripts/train_tokenizer.sh:30).

    python bench.py [--gpus N --steps K --warmup W]            # our arm
    python bench.py --impl reference [...]                     # CPU arm: the oracle port on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # one rank per GPU

One JSON line on stdout (rank 0).  `value` = device-resident inputs, no host syncs in the timed region;
`e2e` = the same K steps through the public API with pinned HOST image batches copied in and the loss read back
every step.  `roofline` is the tcgen05 conv tile (forward + data-gradient launches) timed with CUDA events around
every launch of one extra profiled step; `roofline_hbm` is the fused DMD loss kernel at a bandwidth-bound size (64 Mi
latents); `dmd_stage` is the train_dmd.py iteration (BASELINE configs[2]) replayed from CUDA graphs; `gpu_baseline` is the
same tokenizer step through stock PyTorch/cuDNN eager on the same GPU (the reference's real execution path);
`loss_parity` is the 100-step loss comparison of the real trainer against a strict-fp32 anchor and a cuDNN-autocast
control arm (scripts/loss_parity.py); `cpu_baseline` is the oracle step timed on a bounded sample (1 image).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import torch  # noqa: E402

METRIC = "VAE+DMD training images/sec (256^2, bf16)"
UNIT = "images/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML in a thread (20 ms period), nvidia-smi fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.stop_flag, self.th, self.max_mhz = index, [], 0, False, None, None
        self.smi = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")

            def loop():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mask |= int(reasons_fn(h))
                    except Exception:
                        pass
                    time.sleep(0.02)
            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
        except Exception:
            self._start_smi()

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.smi = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                         "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.rows = []
            self.th = threading.Thread(target=lambda: [self.rows.append([c.strip() for c in l.split(",")]) for l in self.smi.stdout], daemon=True)
            self.th.start()
        except Exception:
            self.smi = None

    def stop(self):
        if self.smi is not None:
            self.smi.terminate()
            self.th.join(timeout=2)
            rows = [r for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": statistics.median(float(r[0]) for r in rows) if rows else None,
                    "sm_max_mhz": max(float(r[1]) for r in rows) if rows else None,
                    "reasons": [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)],
                    "samples": len(rows), "source": "nvidia-smi"}
        if self.th is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        self.stop_flag = True
        self.th.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for bit, n in self.REASONS.items() if self.mask & bit], "samples": len(self.sm), "source": "nvml"}


# ------------------------------------------------------------------------------------------------ algorithmic work per call
def _work(name, a):
    """FLOPs for the conv entry points (2*M*N*K, bias/GN excluded -- SURVEY.md section 8(d)); 0 otherwise."""
    if name == "dmvae_conv_tc_fwd":        # x, w, bias, res, y, gn_stats, B, H, W, Cin, Cout, KH, KW
        B, H, W, cin, cout, kh, kw = a[6:13]
        return 2.0 * B * H * W * cin * cout * kh * kw
    if name == "dmvae_conv_tc_wgrad":      # x, dy, dw, B, H, W, Cin, Cout, KH, KW
        B, H, W, cin, cout, kh, kw = a[3:10]
        return 2.0 * B * H * W * cin * cout * kh * kw
    if name in ("dmvae_conv_up2x_fwd", "dmvae_conv_up2x_dgrad", "dmvae_conv_up2x_wgrad"):
        # sub-pixel Upsample: EXECUTED flops (16 tap-GEMMs per low-res pixel; the reference op it replaces costs 36)
        B, H, W, cin, cout = a[-5:]
        return 2.0 * B * H * W * cin * cout * 16
    if name == "dmvae_conv_direct_fwd":    # x, w, bias, res, y, B,H,W,Cin,OH,OW,Cout,KH,KW,...
        B, H, W, cin, OH, OW, cout, kh, kw = a[5:14]
        return 2.0 * B * OH * OW * cin * cout * kh * kw
    if name == "dmvae_conv_direct_wgrad":
        B, H, W, cin, OH, OW, cout, kh, kw = a[3:12]
        return 2.0 * B * OH * OW * cin * cout * kh * kw
    return 0.0


# ------------------------------------------------------------------------------------------------ our arm
def patchgan_standin():
    """A PatchGAN-shaped discriminator in stock PyTorch (4x4 stride-2 conv stack, 64-128-256-512 channels, BatchNorm, LeakyReLU)
    for the GAN workload: the reference's discriminators (models/patchgan.py, models/dinodisc.py) are black boxes outside the
    accelerated path; what the workload exercises is the adaptive-weight generator branch through the decoder."""
    nn = torch.nn
    layers, c = [nn.Conv2d(3, 64, 4, 2, 1), nn.LeakyReLU(0.2)], 64
    for cout, stride in ((128, 2), (256, 2), (512, 1)):
        layers += [nn.Conv2d(c, cout, 4, stride, 1, bias=False), nn.BatchNorm2d(cout), nn.LeakyReLU(0.2)]
        c = cout
    layers += [nn.Conv2d(c, 1, 4, 1, 1)]
    return nn.Sequential(*layers)


def build_trainer(device, model_size="large", z_channels=32, gan=False):
    from dmvae_b200.lpips import LPIPS
    from dmvae_b200.train import LossConfig, TokenizerTrainer, VAELossFunction
    from dmvae_b200.vae import VAE
    torch.manual_seed(42)
    vae = VAE(z_channels=z_channels, model_size=model_size).to(device)
    vae.encoder.eval()
    for p in vae.encoder.parameters():           # train_tokenizer.py:295-297
        p.requires_grad = False
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lp = LPIPS(ckpt_path=os.path.join(ROOT, "ckpt_vae", "vgg.pth") if os.path.exists(os.path.join(ROOT, "ckpt_vae", "vgg.pth")) else None,
                   pretrained_vgg=False).eval().to(device)
    cfg = LossConfig(l1=1.0, l2=0.0, lpips=1.0, dmd_weight=0.0, disc_weight=0.5 if gan else 0.0)
    disc = patchgan_standin().to(device) if gan else None
    return TokenizerTrainer(vae, VAELossFunction(cfg, lpips_loss=lp, disc=disc), lr=1e-4)


class Stress512Trainer:
    """BASELINE configs[4]: flux_ae Encoder -> reparameterize + KL -> Decoder at 512x512 (a synthetic stress harness,
    the reference VAE hard-codes 256).  L1 + 1e-6*KL, fused clip/AdamW/EMA, gradient arena all-reduce."""

    def __init__(self, dev, res=512):
        from dmvae_b200 import losses
        from dmvae_b200.autoencoder import Decoder, Encoder
        from dmvae_b200.optim import FlatAdamWEMA
        from dmvae_b200.train_arena import GradArena
        from dmvae_b200.vae import init_weights
        torch.manual_seed(42)
        self.losses = losses
        self.enc = Encoder(resolution=res, in_channels=3, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16).to(dev)
        self.dec = Decoder(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, resolution=res, z_channels=16).to(dev)
        init_weights(self.enc, 0.02); init_weights(self.dec, 0.02)
        self.arena = GradArena(list(self.enc.parameters()) + list(self.dec.parameters()))
        self.opt = FlatAdamWEMA(self.arena.params, lr=1e-4, arena=self.arena)

        self.section = None
        self.exchange_mode = "eager"

    def _forward_backward(self, images):
        self.arena.zero()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            h = self.enc(images)
            eps = torch.randn(h.shape[0], h.shape[1] // 2, h.shape[2], h.shape[3], device=h.device, dtype=h.dtype)
            z, kl = self.losses.reparam_kl(h, eps, channel_dim=1)
            rec = self.dec(z).float()
            l1, _ = self.losses.l1_l2_loss(rec, images)
            loss = l1 + 1e-6 * kl
        with self.arena.direct():
            loss.backward()
        self.arena.allreduce()
        return {"loss": loss.detach()}

    def capture(self, images):
        from dmvae_b200.train import _capture_with_exchange
        self._gx = images.clone()
        self.section, self.exchange_mode = _capture_with_exchange(lambda: self._forward_backward(self._gx), [self.arena],
                                                                  self.arena.params, 2, (self.opt,))
        if self.exchange_mode == "post-replay":
            raise RuntimeError("stress512: NCCL exchange could not be captured")
        return True

    def release_graphs(self):
        if self.section is not None:
            self.section.release()
        self.section = None

    def step(self, images):
        if self.section is not None:
            self._gx.copy_(images, non_blocking=True)
            log = self.section.replay()
        else:
            log = self._forward_backward(images)
        log["vae_norm"] = self.opt.step()
        return log


class DmdStageTrainer:
    """BASELINE configs[2]: a train_dmd.py iteration with LightningDiT-Mini/1 teacher (frozen) and student -- VAE turn every
    ``vae_train_every`` = 5 iterations (scripts/train_dmd.sh), the student flow-matching step every iteration.  ViT-B encoder
    trainable (train_dmd.py:519), recon L1 + LPIPS + 10 * DMD(cfg 5)."""

    def __init__(self, dev, vae_train_every=5):
        import warnings
        from dmvae_b200.dit import LightningDiT_Mini_1
        from dmvae_b200.lpips import LPIPS
        from dmvae_b200.train import DmdTrainer, LossConfig
        from dmvae_b200.vae import VAE
        torch.manual_seed(42)
        vae = VAE(z_channels=32, model_size="large").to(dev)

        def mk():
            m = LightningDiT_Mini_1(input_size=16, in_channels=32, num_classes=1000)
            for lin in (m.final_layer.linear, m.final_layer.adaLN_modulation[-1]):   # zero-init head would give v = 0 everywhere
                torch.nn.init.normal_(lin.weight, std=0.02)
            return m.to(dev)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lp = LPIPS(pretrained_vgg=False).eval().to(dev)
        self.tr = DmdTrainer(vae, mk(), mk(), lp, LossConfig(l1=1.0, lpips=1.0, dmd_weight=10.0, dmd_cfg_scale=5.0))
        self.every, self.it = vae_train_every, 0
        self.labels = None

    def capture(self, images):
        self.labels = torch.randint(0, 1000, (images.shape[0],), device=images.device)
        return self.tr.capture_cuda_graphs(images, self.labels, strict=True)

    def step(self, images):
        if self.labels is None or self.labels.shape[0] != images.shape[0]:
            self.labels = torch.randint(0, 1000, (images.shape[0],), device=images.device)
        turn = self.it % self.every == 0
        self.it += 1
        log = self.tr.step(images, self.labels, vae_turn=turn)
        log.setdefault("loss", log["diffusion_loss"])
        return log


WORKLOADS = {
    "tokenizer": "train_tokenizer.py VAE pretrain step (BASELINE configs[1]): frozen ViT-L/16 encoder, flux Decoder "
                 "fwd+bwd, L1+LPIPS(VGG16), allreduce, clip, AdamW, EMA",
    "stress512": "BASELINE configs[4] stress: flux_ae Encoder -> reparam+KL -> Decoder fwd+bwd at 512x512, L1 + KL, "
                 "allreduce, clip, AdamW, EMA",
    "gan": "BASELINE configs[3]: the tokenizer step with the GAN generator branch (adaptive weight: two partial autograd.grad calls "
           "on decoder.conv_out.weight, train_dmd.py:244-257) and the discriminator's hinge step; PatchGAN-shaped stock-PyTorch "
           "discriminator stand-in",
    "dmd": "train_dmd.py iteration (BASELINE configs[2]): LightningDiT-Mini/1 teacher+student; every 5th iteration is a VAE "
           "turn (trainable ViT-L/16 encoder + flux Decoder fwd+bwd, L1+LPIPS+10*DMD cfg 5), every iteration a student "
           "flow-matching step; allreduce, clip, AdamW per network; steps rounded up to whole 5-iteration cycles",
}


def workload_config(workload, B, world, res):
    """`config` of the JSON line: names the workload; identical for our arm and the reference arm."""
    return {"workload": WORKLOADS[workload], "per_gpu_batch": B, "global_batch": B * world, "image": f"3x{res}x{res}",
            "z_channels": 16 if workload == "stress512" else 32, "parallelism": f"dp{world}", "weights": "random init (seed 42)"}


def _timed(fn, steps, world, dev):
    """K calls of fn bracketed by barrier + synchronize; device time by CUDA events, MAX over ranks.  Returns (ms, host issue ms/step)."""
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    host_ms = (time.perf_counter() - t0) * 1e3 / steps          # host time to ISSUE a step (no device wait)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), host_ms


def measure_dmd_stage(dev, world, rank, B, cycles=2):
    """BASELINE configs[2] as a secondary measured workload: the train_dmd.py iteration (VAE turn every 5th iteration + student
    step every iteration) replayed from three CUDA graphs, NCCL exchange inside the graphs.  Whole 5-iteration cycles are timed."""
    from dmvae_b200 import _lib
    tr = DmdStageTrainer(dev)
    g = torch.Generator().manual_seed(4242 * world + rank)
    xs = [(torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev) for _ in range(2)]
    for i in range(tr.every):                                    # eager warm-up: one full cycle
        tr.step(xs[i % 2])
    tr.it = 0
    _lib.Stats.reset()
    for i in range(tr.every):                                    # launches of one cycle (library kernels only), counted eagerly
        tr.step(xs[i % 2])
    lib_launches_per_cycle = _lib.Stats.launches
    tr.it = 0
    tr.capture(xs[0])
    for i in range(tr.every):
        tr.step(xs[i % 2])
    steps = cycles * tr.every
    ms, host_ms = _timed(lambda i: tr.step(xs[i % 2]), steps, world, dev)
    out = {"workload": WORKLOADS["dmd"], "value": round(world * B * steps / (ms * 1e-3), 2), "unit": UNIT,
           "ms_per_iteration": round(ms / steps, 3), "iterations": steps, "host_issue_ms_per_iteration": round(host_ms, 3),
           "cuda_graph": tr.tr.graphed, "exchange": tr.tr.exchange_mode, "per_gpu_batch": B,
           "library_launches_per_5_iterations": lib_launches_per_cycle,
           "exchange_bytes_per_vae_turn": 4 * tr.tr.arena_vae.flat.numel(), "exchange_bytes_per_student_step": 4 * tr.tr.arena_sit.flat.numel()}
    tr.tr.release_graphs()                         # captured NCCL operations must not outlive this leg (see shutdown_distributed)
    del tr
    gc.collect()
    torch.cuda.empty_cache()
    return out


def measure_dmd_kernel_roofline(dev, peaks, latents=64 << 20, reps=10):
    """The fused DMD loss + gradient kernel (train_dmd.py:214-228) at a bandwidth-bound size (at the real B=16 size it
    moves 1.8 MB and is launch-bound).  Algorithmic bytes: 6 bf16 reads + 1 bf16 write = 14 B per latent element.  Timed with
    CUDA events on the launching stream; 940 MB of operands per launch exceeds the 126 MB L2, so every launch streams from HBM."""
    from dmvae_b200 import losses
    P = 32 * 16 * 16
    Bn = latents // P
    g = torch.Generator(device=dev).manual_seed(7)
    z, xt, vTc, vTu, vSc, vSu = (torch.randn(Bn, 32, 16, 16, device=dev, generator=g).bfloat16() for _ in range(6))
    t = torch.rand(Bn, device=dev, generator=g).bfloat16()

    def once():
        return losses.dmd_loss(z, xt, t, vTc, vSc, vTu, vSu, 5.0, True, dz_dtype=torch.bfloat16)
    from dmvae_b200 import _lib
    for _ in range(3):
        once()
    _lib.Stats.reset()
    _lib.Stats.timing = True
    _lib.Stats.work_fn = lambda name, a: 14.0 * Bn * P if name == "dmvae_dmd_loss_fwd_bwd" else 0.0
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    _lib.Stats.timing = False
    ts = [a.elapsed_time(b) for n, a, b, w in _lib.Stats.events if n == "dmvae_dmd_loss_fwd_bwd"]
    ms = sum(ts) / len(ts)
    byts = 14.0 * Bn * P
    ach = byts / (ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_dmd_kernel_ncu.json")     # dram bytes of one launch at this size from an ncu --set full capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    return {"bound": "hbm", "kernel": "dmd_loss_bf16x2_kernel (fused DMD loss + dz, losses.cu)", "achieved": round(ach, 1),
            "peak": peaks["hbm"], "unit": "GB/s", "frac": round(ach / peaks["hbm"], 4), "peak_source": peaks["src"],
            "traffic": traffic, "algorithmic_bytes_per_launch": byts, "latent_elements": Bn * P, "avg_launch_ms": round(ms, 4),
            "launches_timed": len(ts), "real_size_note": "B=16 moves 1.8 MB: launch-latency bound (SURVEY D7)"}


def measure_gpu_baseline(dev, B, steps=5, warmup=3):
    """The real bar: the same tokenizer step through stock PyTorch / cuDNN eager under torch.autocast(bf16) with cudnn.benchmark on
    (the reference's own execution path on a GPU), same box, same weights shape, same batch.  scripts/stock_arms.py."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import stock_arms
    from oracle import dmvae_oracle as O
    from dmvae_b200.vae import DINOEncoder, MLP
    torch.manual_seed(42)
    enc = DINOEncoder("large").eval().to(dev)
    mlp = MLP(1024, 32).to(dev)
    sd = {k: v.to(dev) for k, v in O.make_decoder_state(z_channels=32, seed=42).items()}
    lp = {k: v.to(dev) for k, v in O.make_lpips_state(seed=42).items()}
    arm = stock_arms.StockStep(enc, mlp, sd, lp, "autocast", lr=1e-4)
    x = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
    for _ in range(warmup):
        arm.step(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        arm.step(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del arm
    torch.cuda.empty_cache()
    return {"value": round(B / ms * 1e3, 2), "unit": UNIT, "ms_per_step": round(ms, 2), "steps": steps, "warmup": warmup,
            "kind": "stock PyTorch + cuDNN eager, torch.autocast(bf16), cudnn.benchmark=True, torch.optim.AdamW (no EMA pass); "
                    "device-resident input; oracle's functional restatement of the reference modules"}


def run_ours(args):
    import torch.distributed as dist
    from dmvae_b200 import _lib
    from dmvae_b200.train import shutdown_distributed
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    B = args.batch
    res = 256
    if args.workload == "stress512":
        res = 512
        B = args.batch if args.batch != 16 else 4
        tr = Stress512Trainer(dev, res)
    elif args.workload == "dmd":
        tr = DmdStageTrainer(dev)
        if args.steps % tr.every:
            args.steps += tr.every - args.steps % tr.every        # whole vae_train_every cycles inside the timed region
    else:
        tr = build_trainer(dev, gan=args.workload == "gan")
    g = torch.Generator().manual_seed(42 * world + rank)
    graphed = False
    exchange_mode = "eager"
    n_pool = 4
    host = [(torch.rand(B, 3, res, res, generator=g) * 2 - 1).pin_memory() for _ in range(n_pool)]
    resident = [h.to(dev) for h in host]

    for i in range(args.warmup if args.workload != "dmd" else 2 * tr.every):
        tr.step(resident[i % n_pool])
    graph_launches = None
    if args.cuda_graph and args.workload in ("tokenizer", "gan"):
        # launches a replay stands for: an eager pass with the weight re-packs included (version counters bumped)
        torch.autograd.graph.increment_version(tr.params)
        _lib.Stats.reset()
        tr._forward_backward(resident[0])
        graph_launches = _lib.Stats.launches + 2                # + the two optimizer kernels issued eagerly
        tr.capture_cuda_graph(resident[0], strict=True)         # bench mode: a capture failure is an error, not a silent eager run
        graphed, exchange_mode = tr.graphed, tr.exchange_mode
        for i in range(2):
            tr.step(resident[i % n_pool])
    elif args.cuda_graph and args.workload == "stress512":
        torch.autograd.graph.increment_version(tr.arena.params)
        _lib.Stats.reset()
        tr._forward_backward(resident[0])
        graph_launches = _lib.Stats.launches + 2
        tr.capture(resident[0])
        graphed, exchange_mode = True, tr.exchange_mode
        for i in range(2):
            tr.step(resident[i % n_pool])
    elif args.cuda_graph and args.workload == "dmd":
        tr.it = 0
        _lib.Stats.reset()
        for i in range(tr.every):                                # library launches of one 5-iteration cycle, counted eagerly
            tr.step(resident[i % n_pool])
        graph_launches = _lib.Stats.launches / tr.every          # per iteration
        tr.it = 0
        tr.capture(resident[0])
        graphed, exchange_mode = tr.tr.graphed, tr.tr.exchange_mode
        for i in range(tr.every):
            tr.step(resident[i % n_pool])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.Stats.reset()
    ms, host_issue_ms = _timed(lambda i: tr.step(resident[i % n_pool]), args.steps, world, dev)
    launches = _lib.Stats.launches if graph_launches is None else int(round(graph_launches * args.steps))
    clocks = sampler.stop() if rank == 0 else None

    # end to end: pinned host batch in, loss out, every step
    last = {}

    def e2e_step(i):
        x = host[i % n_pool].to(dev, non_blocking=True)
        last["loss"] = tr.step(x)["loss"].item()
    for i in range(1 if args.workload != "dmd" else tr.every):
        e2e_step(i)
    ms_e2e, _ = _timed(e2e_step, args.steps, world, dev)

    # one profiled step: CUDA events around every library launch (issued eagerly so that every launch can be bracketed)
    per, dump, prof_total_ms = {}, [], 0.0
    per_kernel = {}
    if args.workload != "dmd" or not graphed:
        _lib.Stats.reset()
        _lib.Stats.work_fn = _work
        _lib.Stats.timing = True
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sec_attrs = [a for a in ("_section", "_sec_enc", "_sec_dec", "section") if hasattr(tr, a)]
        saved = {a: getattr(tr, a) for a in sec_attrs}
        for a in sec_attrs:
            setattr(tr, a, None)
        pe0.record()
        tr.step(resident[0])
        pe1.record()
        torch.cuda.synchronize()
        for a in sec_attrs:
            setattr(tr, a, saved[a])
        _lib.Stats.timing = False
        prof_total_ms = pe0.elapsed_time(pe1)
        for i, (name, a, b, work) in enumerate(_lib.Stats.events):
            d = per.setdefault(name, [0, 0.0, 0.0])
            t = a.elapsed_time(b)
            d[0] += 1; d[1] += t; d[2] += work
            dump.append((name, t, work))
            kid = _lib.Stats.kernel_ids[i] if i < len(_lib.Stats.kernel_ids) else 0
            if kid and "wgrad" not in name:                     # forward / data-gradient launches by the tile kernel that ran them
                k = per_kernel.setdefault(kid, [0, 0.0, 0.0])
                k[0] += 1; k[1] += t; k[2] += work
        if os.environ.get("BENCH_DUMP_LAUNCHES") and rank == 0:     # per-launch (entry point, ms, algorithmic flops) of the profiled step
            with open(os.environ["BENCH_DUMP_LAUNCHES"], "w") as f:
                json.dump([{"name": n, "ms": round(t, 4), "work": w, "args": list(_lib.Stats.args_log[i]) if i < len(_lib.Stats.args_log) else None}
                           for i, (n, t, w) in enumerate(dump)], f)
    step_ms_prof = prof_total_ms          # whole profiled step on the device (library kernels + the PyTorch ones)
    lib_ms = sum(v[1] for v in per.values())
    release = getattr(tr, "release_graphs", None) or getattr(getattr(tr, "tr", None), "release_graphs", None)
    if release is not None:
        release()                         # captured NCCL operations must not outlive the trainer (see shutdown_distributed)
    del tr
    gc.collect()
    torch.cuda.empty_cache()

    # secondary legs (every rank takes part in the ones that exchange)
    extra = {}
    if args.workload == "tokenizer" and not args.no_dmd_stage:
        extra["dmd_stage"] = measure_dmd_stage(dev, world, rank, B)
    if args.workload == "tokenizer" and args.parity_steps > 0:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import loss_parity
        lp_out = loss_parity.run_parity(dev, steps=args.parity_steps, global_batch=16, size="large", cuda_graph=True)
        torch.cuda.empty_cache()
        if rank == 0:
            extra["loss_parity"] = lp_out

    if rank != 0:
        shutdown_distributed()
        return
    peaks = load_peaks()
    tc = [0, 0.0, 0.0]                         # forward + data-gradient launches of the tcgen05 conv tiles (plain and sub-pixel entry points)
    for name in ("dmvae_conv_tc_fwd", "dmvae_conv_up2x_fwd", "dmvae_conv_up2x_dgrad"):
        v = per.get(name, [0, 0.0, 0.0])
        tc = [tc[0] + v[0], tc[1] + v[1], tc[2] + v[2]]
    step_ms = ms / args.steps
    knames = {1: "conv_tc_kernel (single-CTA per-tap tile)", 2: "conv_tc2_kernel (per-tap CTA pair: 1x1 / stride-2 / small images)",
              3: "conv_tc2h_kernel (halo-resident CTA-pair tile, cta_group::2 UMMA 256x256x16; plain 3x3 and sub-pixel Upsample launches)",
              4: "conv_tcT_kernel (transposed tile for 64 / 128 / thin outputs)"}

    def roof_entry(v, kernel):
        ach = v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0
        return {"bound": "tensor", "kernel": kernel, "achieved": round(ach, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / peaks["tf_sustained"], 4),
                "peak_source": f"{peaks['src']} (sustained: kernel timed inside a long step)",
                "frac_of_burst_peak": round(ach / peaks["tf_burst"], 4), "burst_peak": peaks["tf_burst"],
                "launches_per_step": v[0], "avg_launch_ms": round(v[1] / max(v[0], 1), 4),
                "flops_per_launch_avg": v[2] / max(v[0], 1), "traffic": None,
                "share_of_step": round(v[1] / max(step_ms, 1e-9), 4),
                "share_note": "sum of these launches' CUDA-event durations (taken in one eagerly issued step) / the timed step; "
                              "sub-pixel Upsample launches counted at their EXECUTED flops"}
    # the dominant kernel of the step = the tile kernel with the largest summed duration (attributed per launch by the library's
    # dmvae_conv_tc_last_kernel hook); the aggregate over every forward + data-gradient conv launch is kept beside it
    roof_all = roof_entry(tc, "all tcgen05 conv tiles: every forward + data-gradient launch of a step (conv_tc2h / conv_tcT / conv_tc2 / conv_tc)")
    if per_kernel:
        dom = max(per_kernel, key=lambda k: per_kernel[k][1])
        roof = roof_entry(per_kernel[dom], knames.get(dom, f"kernel id {dom}") + ": all its forward + data-gradient launches of a step")
        roof["by_kernel"] = {knames.get(k, str(k)).split(" ")[0]: {"n": v[0], "ms": round(v[1], 3), "tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1) if v[1] else None}
                             for k, v in sorted(per_kernel.items())}
    else:
        roof = roof_all
    wg = [0, 0.0, 0.0]
    for name in ("dmvae_conv_tc_wgrad", "dmvae_conv_up2x_wgrad"):
        v = per.get(name, [0, 0.0, 0.0])
        wg = [wg[0] + v[0], wg[1] + v[1], wg[2] + v[2]]
    kernels = {k.replace("dmvae_", ""): {"n": v[0], "ms": round(v[1], 3), **({"tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1)} if v[2] and v[1] else {})}
               for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}
    imgs = world * B * args.steps
    out = {
        "metric": METRIC, "value": round(imgs / (ms * 1e-3), 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args.workload, B, world, res),
        "run": {"cuda_graph": graphed, "gradient_exchange": exchange_mode,
                "l2": "per-step working set (several GB of activations) exceeds the 126 MB L2; 4 rotating input batches"},
        "e2e": {"value": round(imgs / (ms_e2e * 1e-3), 2), "unit": UNIT, "h2d_bytes_per_step": B * 3 * res * res * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3), "last_loss": last.get("loss")},
        "gpu_launches": launches, "host_issue_ms_per_step": round(host_issue_ms, 3), "clocks": clocks, "roofline": roof,
        "roofline_all_conv_tiles": roof_all,
        "wgrad": {"achieved_tflops": round(wg[2] / (wg[1] * 1e-3) / 1e12, 1) if wg[1] else None, "launches_per_step": wg[0],
                  "share_of_step": round(wg[1] / max(step_ms, 1e-9), 4)},
        "profiled_step_ms": {"total": round(prof_total_ms, 3), "library_kernels": round(lib_ms, 3)},
        "kernels_ms_per_step": kernels,
    }
    out.update(extra)
    if args.workload == "tokenizer":
        out["roofline_hbm"] = measure_dmd_kernel_roofline(dev, peaks)
        torch.cuda.empty_cache()
        if world == 1 and not args.no_gpu_baseline:
            out["gpu_baseline"] = measure_gpu_baseline(dev, B)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_step_baseline(sample_images=1, reps=1)
    emit(out)
    shutdown_distributed()


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
class OracleStep:
    """The same training step on the host cores: oracle/dmvae_oracle.py (decoder + losses, fp32, autograd) with the
    stock-PyTorch encoder module run on CPU, AdamW on the decoder/bottleneck tensors."""

    def __init__(self):
        from oracle import dmvae_oracle as O
        from dmvae_b200.vae import DINOEncoder, MLP
        self.O = O
        torch.manual_seed(42)
        self.enc = DINOEncoder("large").eval()
        self.mlp = MLP(1024, 32)
        self.sd = {k: v.requires_grad_(True) for k, v in O.make_decoder_state(z_channels=32, seed=42).items()}
        self.lp = O.make_lpips_state(seed=42)
        self.params = list(self.sd.values()) + list(self.mlp.parameters())
        self.opt = torch.optim.AdamW(self.params, lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)

    def step(self, images):
        O = self.O
        with torch.no_grad():
            tok = self.enc(images)
        z = self.mlp(tok)
        recon = O.decoder_forward(self.sd, z)
        loss = (recon - images).abs().mean() + O.lpips_forward(self.lp, images, recon)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 1.0)
        self.opt.step()
        return loss.item()


def cpu_step_baseline(sample_images=1, reps=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = OracleStep()
    x = torch.rand(sample_images, 3, 256, 256, generator=torch.Generator().manual_seed(0)) * 2 - 1
    st.step(x)                                   # warm-up (thread pool, oneDNN primitive caches)
    t0 = time.perf_counter()
    for _ in range(reps):
        st.step(x)
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(sample_images / dt, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{reps} oracle step(s) of {sample_images} image(s) after 1 warm-up, fp32, torch.set_num_threads({cores})"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = OracleStep()
    b = args.ref_batch
    g = torch.Generator().manual_seed(42)
    xs = [torch.rand(b, 3, 256, 256, generator=g) * 2 - 1 for _ in range(2)]
    for i in range(args.warmup):
        st.step(xs[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss = st.step(xs[i % 2])
    dt = time.perf_counter() - t0
    v = round(b * args.steps / dt, 4)
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config("tokenizer", args.batch, int(os.environ.get("WORLD_SIZE", 1)), 256),
        "run": {"arm": "the workload's step on the host cores: oracle port of the reference's PyTorch modules, fp32 (the Python "
                       "reference itself does not travel to the GPU box)", "images_per_sampled_step": b},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps after {args.warmup} warm-up, each a bounded sample of {b} image(s) of the "
                                   f"{args.batch}-image batch, fp32, torch.set_num_threads({cores})"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "last_loss": loss})


_JSON_FD = None


def _claim_stdout():
    """stdout carries exactly ONE line, the result JSON: NCCL (version banner), cuDNN and anything else that writes to fd 1 is
    sent to stderr for the rest of the process; emit() writes the line to the real stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (scripts/train_tokenizer.sh: local_bs 16)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock PyTorch/cuDNN eager step (N=1 only)")
    ap.add_argument("--no-dmd-stage", action="store_true", help="skip the secondary train_dmd.py-iteration measurement")
    ap.add_argument("--parity-steps", type=int, default=100,
                    help="steps of the loss-parity leg (real trainer vs fp32 anchor vs cuDNN-autocast control); 0 = skip")
    ap.add_argument("--quick", action="store_true", help="main timing only: no dmd_stage / loss_parity / gpu_baseline / cpu_baseline")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false",
                    help="tokenizer workload: issue every kernel from Python instead of replaying forward+backward from a CUDA graph")
    ap.add_argument("--workload", default="tokenizer", choices=["tokenizer", "stress512", "dmd", "gan"],
                    help="tokenizer = BASELINE configs[1] (the headline workload); dmd = configs[2]; stress512 = configs[4] "
                         "(no CPU baseline for the last two)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    if args.quick:
        args.no_cpu_baseline = args.no_gpu_baseline = args.no_dmd_stage = True
        args.parity_steps = 0
    run_ours(args)


if __name__ == "__main__":
    main()
