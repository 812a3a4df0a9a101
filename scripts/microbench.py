#!/usr/bin/env python
"""Per-kernel timing of the HBM-bound kernels at production / roofline sizes (CUDA events; buffers are rotated so the
footprint between two uses of the same buffer exceeds the 126 MB L2).  One JSON line per kernel:
   {"kernel":..., "shape":..., "us":..., "algo_bytes":..., "gbs":..., "frac_of_hbm_peak":...}
Usage: python scripts/microbench.py [--only substr] [--iters N]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dmvae_b200 import ops  # noqa: E402
from dmvae_b200._lib import call, ptr  # noqa: E402

DEV = "cuda"
PEAK = 6553.9
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = json.load(open(p))["hbm_gbs"]


def timeit(fn, iters, nbuf):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def report(name, shape, us, algo_bytes):
    gbs = algo_bytes / us / 1e3
    print(json.dumps({"kernel": name, "shape": list(shape), "us": round(us, 2), "algo_bytes": algo_bytes, "gbs": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / PEAK, 3)}), flush=True)


def nbuf_for(bytes_per_set):
    return max(2, int(400e6 // bytes_per_set) + 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    want = lambda n: a.only in n or any(tok and tok in n for tok in a.only.split(","))

    act_shapes = [(16, 256, 256, 128), (16, 128, 128, 256), (16, 64, 64, 512), (16, 32, 32, 512)]
    for shp in act_shapes:
        if not any(want(k) for k in ("gn_stats", "gn_apply", "gn_bwd", "bias_grad", "upsample")):
            break
        B, H, W, C = shp
        n = B * H * W * C
        nb = nbuf_for(n * 2 * 3)
        xs = [torch.randn(shp, device=DEV).bfloat16() for _ in range(nb)]
        das = [torch.randn(shp, device=DEV).bfloat16() for _ in range(nb)]
        gamma = torch.rand(C, device=DEV) + 0.5
        beta = torch.randn(C, device=DEV) * 0.1
        stats = [ops.gn_stats_raw(x) for x in xs]
        if want("gn_stats"):
            report("gn_stats", shp, timeit(lambda i: ops.gn_stats_raw(xs[i]), a.iters, nb), n * 2)
        if want("gn_apply"):
            report("gn_apply", shp, timeit(lambda i: ops.gn_apply_raw(xs[i], stats[i], gamma, beta, True), a.iters, nb), n * 4)
        if want("gn_bwd"):
            report("gn_bwd", shp, timeit(lambda i: ops.gn_bwd_raw(das[i], xs[i], stats[i], gamma, beta, True), a.iters, nb), n * 10)
        if want("bias_grad"):
            report("bias_grad", shp, timeit(lambda i: ops.bias_grad_raw(das[i]), a.iters, nb), n * 2)
        if want("upsample") and H <= 128:
            report("upsample2x_fwd", shp, timeit(lambda i: ops.upsample2x(xs[i]), a.iters, nb), n * 2 * 5)
        del xs, das, stats
        torch.cuda.empty_cache()

    if want("convtc"):
        from dmvae_b200 import _lib
        shapes = [(64, 64, 256, 3), (64, 128, 128, 3), (128, 128, 128, 3), (128, 256, 64, 3), (512, 512, 32, 3), (512, 512, 64, 3), (512, 512, 128, 3), (512, 256, 128, 3), (256, 256, 128, 3), (256, 256, 256, 3),
                  (256, 128, 256, 3), (128, 128, 256, 3), (512, 512, 32, 1), (512, 256, 128, 1), (256, 128, 256, 1), (32, 512, 32, 3)]
        for cin, cout, hw, k in shapes:
            B = 16
            x = torch.randn(B, hw, hw, cin, device=DEV).bfloat16()
            dy = torch.randn(B, hw, hw, cout, device=DEV).bfloat16()
            w = torch.randn(cout, cin, k, k, device=DEV) * 0.02
            bias = torch.zeros(cout, device=DEV)
            wf, wd = ops.WeightPack().get(w)
            flops = 2.0 * B * hw * hw * cin * cout * k * k
            res = {}
            for mode in (1, 2, 3):
                _lib.query("dmvae_conv_tc_set_tile_mode", mode)
                us = timeit(lambda i: ops.conv_forward_raw(x, wf, bias, None, k, k, 1, ((k - 1) // 2, (k - 1) // 2)), a.iters, 1)
                res[{1: "fwd_mt1", 2: "fwd_mt2", 3: "fwd_pair"}[mode] + "_tflops"] = round(flops / us / 1e6, 1)
            _lib.query("dmvae_conv_tc_set_tile_mode", 0)
            if k == 3:
                _lib.query("dmvae_conv_tc_set_tile_mode", 8)
                us = timeit(lambda i: ops.conv_forward_raw(x, wf, bias, None, k, k, 1, (1, 1)), a.iters, 1)
                res["fwd_halo_tflops"] = round(flops / us / 1e6, 1)
                rsd = torch.randn(B, hw, hw, cout, device=DEV).bfloat16()
                us = timeit(lambda i: ops.conv_forward_raw(x, wf, bias, rsd, k, k, 1, (1, 1)), a.iters, 1)
                res["fwd_halo_res_tflops"] = round(flops / us / 1e6, 1)
                del rsd
                _lib.query("dmvae_conv_tc_set_tile_mode", 7)
            # the tile the dispatcher picks by itself: plain, with the fused GroupNorm statistics, with residual + statistics (what
            # the decoder's conv1 / conv2 launch), with the ReLU epilogue and the ReLU-gated data gradient (the VGG16 launches)
            pad = ((k - 1) // 2, (k - 1) // 2)
            rsd = torch.randn(B, hw, hw, cout, device=DEV).bfloat16()
            msk = torch.relu(torch.randn(B, hw, hw, cout, device=DEV)).bfloat16()
            variants = {"auto": dict(), "auto_stats": dict(want_gn_stats=True), "auto_res_stats": dict(residual=rsd, want_gn_stats=True),
                        "auto_relu": dict(flags=ops.EPI_RELU), "auto_mask": dict(residual=msk, flags=ops.EPI_MASK)}
            for name, kw in variants.items():
                r = kw.pop("residual", None)
                us = timeit(lambda i: ops.conv_forward_raw(x, wf, bias, r, k, k, 1, pad, **kw), a.iters, 1)
                res[f"fwd_{name}_tflops"] = round(flops / us / 1e6, 1)
            del rsd, msk
            us = timeit(lambda i: ops.conv_wgrad_raw(x, dy, k, k, 1, ((k - 1) // 2, (k - 1) // 2)), a.iters, 1)
            res["wgrad_tflops"] = round(flops / us / 1e6, 1)
            if cin <= 128 and cout < 256:
                for m in (12, 14):
                    _lib.query("dmvae_conv_tc_set_tile_mode", m)
                    us = timeit(lambda i: ops.conv_wgrad_raw(x, dy, k, k, 1, ((k - 1) // 2, (k - 1) // 2)), a.iters, 1)
                    res[f"wgrad_tpc{m - 10}_tflops"] = round(flops / us / 1e6, 1)
                _lib.query("dmvae_conv_tc_set_tile_mode", 13)
            print(json.dumps({"kernel": "conv_tc", "cin": cin, "cout": cout, "hw": hw, "k": k, "gflop": round(flops / 1e9, 1), **res}), flush=True)
            del x, dy
            torch.cuda.empty_cache()

    if want("upconv"):
        # flux_ae.Upsample: nearest 2x + 3x3.  Two-kernel form (upsample2x + hi-res conv) vs the sub-pixel form, all three passes;
        # "tflops" are algorithmic FLOPs of the REFERENCE op (36 tap-GEMMs per low-res pixel) / time, so the two forms compare directly
        for c, hw in ((512, 32), (512, 64), (256, 128)):
            B = 16
            x = torch.randn(B, hw, hw, c, device=DEV).bfloat16().requires_grad_(True)
            w = (torch.randn(c, c, 3, 3, device=DEV) * 0.02).requires_grad_(True)
            bias = torch.zeros(c, device=DEV, requires_grad=True)
            dy = torch.randn(B, 2 * hw, 2 * hw, c, device=DEV).bfloat16()
            flops = 2.0 * B * 4 * hw * hw * c * c * 9
            pk, sp = ops.WeightPack(), ops.SubpixelPack()
            res = {}

            def plain_f(i):
                return ops.conv2d(ops.upsample2x(x), w, bias, pk, want_gn_stats=True)

            def sub_f(i):
                return ops.upsample_conv(x, w, bias, sp, True)
            for name, f in (("two_kernel", plain_f), ("subpixel", sub_f)):
                us_f = timeit(f, a.iters, 1)
                y = f(0)
                us_b = timeit(lambda i: torch.autograd.grad(y, [x, w, bias], dy, retain_graph=True), a.iters, 1)
                res[f"{name}_fwd_us"] = round(us_f, 1)
                res[f"{name}_bwd_us"] = round(us_b, 1)
                res[f"{name}_fwd_ref_tflops"] = round(flops / us_f / 1e6, 1)
                res[f"{name}_bwd_ref_tflops"] = round(2 * flops / us_b / 1e6, 1)
                del y
            print(json.dumps({"kernel": "upsample_conv", "c": c, "hw_in": hw, "ref_gflop_fwd": round(flops / 1e9, 1), **res}), flush=True)
            del x, dy
            torch.cuda.empty_cache()

    if want("conv_out"):
        shp = (16, 256, 256, 128)
        x = torch.randn(shp, device=DEV).bfloat16()
        w = torch.randn(3, 128, 3, 3, device=DEV) * 0.02
        wf, wd = ops.WeightPack().get(w)
        dy = torch.randn(16, 256, 256, 3, device=DEV).bfloat16()
        n = x.numel()
        report("conv_out_fwd(128->3)", shp, timeit(lambda i: ops.conv_forward_raw(x, wf, None, None, 3, 3), a.iters, 1), n * 2)
        report("conv_out_dgrad(3->128)", shp, timeit(lambda i: ops.conv_dgrad_raw(dy, wf, wd, (256, 256), 3, 3), a.iters, 1), n * 2)
        report("conv_out_wgrad", shp, timeit(lambda i: ops.conv_wgrad_raw(x, dy, 3, 3), a.iters, 1), n * 2)

    # A3: fused DMD loss at a bandwidth-bound size (>= 64 Mi latent elements) and at the real size
    for B in (8192, 16):
        if not want("dmd"):
            break
        shp = (B, 32, 16, 16)
        n = B * 8192
        nb = 2 if B > 16 else 64
        sets = [[torch.randn(shp, device=DEV).bfloat16() for _ in range(6)] + [torch.rand(B, device=DEV).bfloat16()] for _ in range(nb)]
        dz = torch.empty(shp, device=DEV, dtype=torch.bfloat16)
        acc = torch.zeros(2, device=DEV, dtype=torch.float64)

        def f(i):
            z, xt, a1, a2, a3, a4, t = sets[i]
            call("dmvae_dmd_loss_fwd_bwd", ptr(z), ptr(xt), ptr(t), ptr(a1), ptr(a2), ptr(a3), ptr(a4), ptr(dz), ptr(acc), B, 8192,
                 5.0, 1, 1.0, 1, 1)
        report("dmd_loss_fwd_bwd(bf16,cfg)", shp, timeit(f, a.iters, nb), n * 14)

        def g(i):
            z, x0, *_, t = sets[i]
            call("dmvae_dmd_mix_xt", ptr(z), ptr(x0), ptr(t), ptr(dz), B, 8192, 1)
        report("dmd_mix_xt(bf16)", shp, timeit(g, a.iters, nb), n * 6)
        del sets
        torch.cuda.empty_cache()

    if want("l1l2"):
        for B in (16, 512):
            n = B * 3 * 256 * 256
            nb = nbuf_for(n * 12)
            rs = [torch.randn(n, device=DEV) for _ in range(nb)]
            xs = [torch.randn(n, device=DEV) for _ in range(nb)]
            d = torch.empty(n, device=DEV)
            acc = torch.zeros(2, device=DEV, dtype=torch.float64)
            report("l1l2_fwd_bwd", (B, 3, 256, 256), timeit(lambda i: call("dmvae_l1l2_fwd_bwd", ptr(rs[i]), ptr(xs[i]), ptr(d), ptr(acc), n, 1.0, 0.5), a.iters, nb), n * 12)
            del rs, xs
            torch.cuda.empty_cache()

    if want("lpips"):
        for (C, S) in ((64, 256), (128, 128), (256, 64), (512, 32), (512, 16)):
            B = 16
            n = B * C * S * S
            nb = nbuf_for(n * 6)
            f0 = [torch.relu(torch.randn(B, S * S, C, device=DEV)).bfloat16() for _ in range(nb)]
            f1 = [torch.relu(torch.randn(B, S * S, C, device=DEV)).bfloat16() for _ in range(nb)]
            w = torch.rand(C, device=DEV)
            acc = torch.zeros(B, device=DEV, dtype=torch.float64)
            df = torch.empty_like(f0[0])
            g0 = torch.ones(1, device=DEV)
            report("lpips_dist_fwd", (B, C, S, S), timeit(lambda i: call("dmvae_lpips_dist_fwd", ptr(f0[i]), ptr(f1[i]), ptr(w), ptr(acc), B, S * S, C, 1, 0), a.iters, nb), n * 4)
            report("lpips_dist_bwd", (B, C, S, S), timeit(lambda i: call("dmvae_lpips_dist_bwd", ptr(f0[i]), ptr(f1[i]), ptr(w), ptr(df), ptr(g0), B, S * S, C, 1.0, 1), a.iters, nb), n * 6)
            del f0, f1
            torch.cuda.empty_cache()

    if want("reparam"):
        rows, half = 16 * 1024, 32 * 64
        n = rows * half
        h = torch.randn(rows, 2 * half, device=DEV).bfloat16()
        eps = torch.randn(rows, half, device=DEV).bfloat16()
        z = torch.empty_like(eps)
        acc = torch.zeros(1, device=DEV, dtype=torch.float64)
        report("reparam_kl_fwd", (rows, 2 * half), timeit(lambda i: call("dmvae_reparam_kl_fwd", ptr(h), ptr(eps), ptr(z), ptr(acc), rows, half, 1), a.iters, 1), n * 8)


if __name__ == "__main__":
    main()
