"""Kernel-level parity: every C-ABI entry point against the CPU oracle (oracle/dmvae_oracle.py) on seeded inputs.

Tolerances (stated per test): fp32 elementwise chains are compared at 1e-6 relative (only reduction order differs);
bf16 tensor-core outputs at 2^-8 relative per element (one bf16 ulp) plus an absolute floor scaled by the
accumulation length; reduced scalars at 1e-3 relative.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import dmvae_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from dmvae_b200 import ops, losses
    return ops, losses


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc(x):  # (B,C,H,W) fp32 cpu -> (B,H,W,C) bf16 cuda
    return x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)


def nchw(y):  # (B,H,W,C) cuda -> (B,C,H,W) fp32 cpu
    return y.float().permute(0, 3, 1, 2).contiguous().cpu()


# ------------------------------------------------------------------------------------------------ A3
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cfg,normalize", [(5.0, True), (3.7, True), (1.0, True), (1.0, False)])
@pytest.mark.parametrize("shape", [(4, 32, 16, 16), (3, 2, 1, 1), (5, 7, 3, 3)])
def test_dmd_loss(dtype, cfg, normalize, shape):
    _, L = _ops()
    g = torch.Generator().manual_seed(1)
    z, x0, vTc, vTu, vSc, vSu = (torch.randn(shape, generator=g).to(dtype) for _ in range(6))
    t = torch.rand(shape[0], generator=g).to(dtype)
    xt_ref = O.dmd_mix_xt(z, x0, t)
    xt = L.dmd_mix_xt(z.to(DEV), x0.to(DEV), t.to(DEV))
    assert torch.equal(xt.cpu(), xt_ref), "xt mix must be bit-exact (elementwise, per-op rounding)"
    loss_ref, gn_ref, dz_ref = O.dmd_loss(z, xt_ref, t, vTc, vSc, vTu, vSu, cfg, normalize)
    zc = z.to(DEV).requires_grad_(True)
    loss, gn = L.dmd_loss(zc, xt, t.to(DEV), vTc.to(DEV), vSc.to(DEV), vTu.to(DEV), vSu.to(DEV), cfg, normalize)
    loss.backward()
    # the per-sample mean|p_real| is a reduction: order differs, and in bf16 it is then rounded to 8 bits, so a
    # 1-ulp flip of the normaliser moves that sample's grad by 2^-8.  fp32: 1e-5.  bf16: 1e-2.
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert abs(loss.item() - loss_ref.item()) <= tol * abs(loss_ref.item()) + 1e-12
    assert abs(gn.item() - gn_ref.item()) <= tol * abs(gn_ref.item()) + 1e-12
    assert rel_err(zc.grad.float(), dz_ref) <= tol


def test_dmd_nan_to_num_and_empty():
    _, L = _ops()
    # sample 0: p_real == 0 everywhere -> normaliser 0 -> 0/0 = NaN -> 0 ; x/0 = inf -> dtype max (train_dmd.py:224)
    z = torch.zeros(2, 4, 2, 2); xt = torch.zeros_like(z); t = torch.tensor([0.5, 0.5])
    vT = torch.zeros_like(z); vS = torch.ones_like(z); vS[0, 0] = 0
    vT[1] = torch.randn(4, 2, 2)
    loss_ref, gn_ref, dz_ref = O.dmd_loss(z, xt, t, vT, vS, None, None, 1.0, True)
    zc = z.to(DEV).requires_grad_(True)
    loss, gn = L.dmd_loss(zc, xt.to(DEV), t.to(DEV), vT.to(DEV), vS.to(DEV), None, None, 1.0, True)
    loss.backward()
    assert torch.isfinite(zc.grad).all()
    assert rel_err(zc.grad, dz_ref) < 1e-5
    # +-inf -> +-dtype max makes the squared error overflow to inf in the reference as well
    assert loss.item() == loss_ref.item() or abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    # empty batch
    e = torch.zeros(0, 4, 2, 2, device=DEV)
    loss, gn = L.dmd_loss(e, e, torch.zeros(0, device=DEV), e, e)
    assert loss.item() == 0.0


def test_dmd_large_per_sample_uncached():
    _, L = _ops()
    g = torch.Generator().manual_seed(3)
    shape = (2, 16, 32, 32)   # 16384 elements per sample > shared-memory cache
    z, xt, vT, vS = (torch.randn(shape, generator=g) for _ in range(4))
    t = torch.rand(2, generator=g)
    loss_ref, gn_ref, dz_ref = O.dmd_loss(z, xt, t, vT, vS)
    zc = z.to(DEV).requires_grad_(True)
    loss, gn = L.dmd_loss(zc, xt.to(DEV), t.to(DEV), vT.to(DEV), vS.to(DEV))
    loss.backward()
    assert rel_err(zc.grad, dz_ref) < 1e-5 and abs(loss.item() - loss_ref.item()) < 1e-5 * loss_ref.item()


# ------------------------------------------------------------------------------------------------ A4
@pytest.mark.parametrize("n", [(2, 3, 64, 64), (1, 3, 5, 7), (0, 3, 4, 4)])
def test_l1l2(n):
    _, L = _ops()
    g = torch.Generator().manual_seed(2)
    r, x = torch.randn(n, generator=g), torch.randn(n, generator=g)
    if r.numel():
        r.view(-1)[0] = x.view(-1)[0]            # exact zero difference: sign(0) = 0
    rc = r.to(DEV).requires_grad_(True)
    l1, l2 = L.l1_l2_loss(rc, x.to(DEV))
    if r.numel() == 0:
        return
    l1_ref, l2_ref, g_ref = O.l1l2(r, x, 1.0, 0.7)
    (l1 * 1.0 + l2 * 0.7).backward()
    assert abs(l1.item() - l1_ref.item()) < 1e-6 * l1_ref.item()
    assert abs(l2.item() - l2_ref.item()) < 1e-6 * l2_ref.item()
    assert rel_err(rc.grad, g_ref) < 1e-6
    a, b, d = L.l1l2_fused(r.to(DEV), x.to(DEV), 1.0, 0.7)
    assert rel_err(d, g_ref) < 1e-6 and abs(a.item() - l1_ref.item()) < 1e-6 * l1_ref.item()


# ------------------------------------------------------------------------------------------------ A5
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,hw", [(64, 24), (128, 12), (256, 6), (512, 3)])
def test_lpips_distance(dtype, c, hw):
    _, L = _ops()
    g = torch.Generator().manual_seed(c)
    f0 = torch.relu(torch.randn(2, c, hw, hw, generator=g)).to(dtype)
    f1 = torch.relu(torch.randn(2, c, hw, hw, generator=g)).to(dtype)
    w = torch.rand(c, generator=g)
    f1r = f1.float().clone().requires_grad_(True)
    ref = O.lpips_distance([f0], [f1r], [w])
    ref.backward()
    f1c = f1.to(DEV).requires_grad_(True)
    d = L.lpips_tap_distance(f0.to(DEV), f1c, w.to(DEV)).mean()
    d.backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2      # bf16: the gradient is rounded to bf16 at the store
    assert abs(d.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_err(f1c.grad.float(), f1r.grad) < tol


def test_lpips_distance_faithful_bf16():
    _, L = _ops()
    g = torch.Generator().manual_seed(5)
    f0 = torch.relu(torch.randn(2, 64, 8, 8, generator=g)).bfloat16()
    f1 = torch.relu(torch.randn(2, 64, 8, 8, generator=g)).bfloat16()
    w = torch.rand(64, generator=g)
    ref = O.lpips_distance([f0], [f1], [w], faithful=True)
    d = L.lpips_tap_distance(f0.to(DEV), f1.to(DEV), w.to(DEV), faithful=True)
    d = O.r16(O.r16(d.cpu()).mean())
    assert abs(d.item() - ref.item()) <= 2 ** -7 * abs(ref.item())   # two bf16 ulps on the bf16 tail


# ------------------------------------------------------------------------------------------------ A8
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reparam_kl(dtype):
    _, L = _ops()
    g = torch.Generator().manual_seed(4)
    h = (torch.randn(3, 8, 4, 4, generator=g) * 0.5).to(dtype)
    eps = torch.randn(3, 4, 4, 4, generator=g).to(dtype)
    hr = h.float().clone().requires_grad_(True)
    mu, lv = hr.chunk(2, dim=1)
    z_ref, kl_ref = O.reparam_kl(mu, lv, eps)
    (z_ref.sum() * 0.3 + kl_ref * 0.01).backward()
    hc = h.to(DEV).requires_grad_(True)
    z, kl = L.reparam_kl(hc, eps.to(DEV), channel_dim=1)
    (z.float().sum() * 0.3 + kl * 0.01).backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(z.float(), z_ref) < tol
    assert abs(kl.item() - kl_ref.item()) < 1e-4 * abs(kl_ref.item())
    assert rel_err(hc.grad.float(), hr.grad) < tol


# ------------------------------------------------------------------------------------------------ GroupNorm
@pytest.mark.parametrize("c,h,w", [(32, 8, 8), (128, 16, 8), (256, 4, 12), (512, 8, 8)])
@pytest.mark.parametrize("silu", [True, False])
def test_group_norm_silu(c, h, w, silu):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(c + h)
    x = O.r16(torch.randn(2, c, h, w, generator=g) * 2 + 0.5)
    gamma = 1 + 0.2 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    da = O.r16(torch.randn(2, c, h, w, generator=g))
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y_ref = O.group_norm({"n.weight": gr, "n.bias": br}, "n", xr, silu)
    y_ref.backward(da)
    xc = nhwc(x).requires_grad_(True)
    gc, bc = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    y = ops.group_norm_silu(xc, gc, bc, silu)
    y.backward(nhwc(da))
    # outputs are bf16: one ulp = 2^-8 relative; norm-wise error is well below that
    assert rel_err(nchw(y), y_ref.detach()) < 4e-3
    assert rel_err(nchw(xc.grad), xr.grad) < 6e-3
    assert rel_err(gc.grad, gr.grad) < 2e-3 and rel_err(bc.grad, br.grad) < 2e-3


# ------------------------------------------------------------------------------------------------ layout
def test_layout_roundtrip_and_upsample():
    ops, _ = _ops()
    g = torch.Generator().manual_seed(7)
    x = O.r16(torch.randn(2, 40, 6, 10, generator=g))
    xl = ops.to_channels_last(x.to(DEV))
    assert torch.equal(nchw(xl), x)                                   # pure transpose: bit-exact
    back = ops.to_nchw(xl, torch.float32)
    assert torch.equal(back.cpu(), x)
    up = ops.upsample2x(xl)
    assert torch.equal(nchw(up), x.repeat_interleave(2, 2).repeat_interleave(2, 3))
    # adjoint
    xg = xl.clone().requires_grad_(True)
    dy = torch.randn(2, 12, 20, 40, generator=g).to(DEV, torch.bfloat16)
    ops.upsample2x(xg).backward(dy)
    ref = F.avg_pool2d(dy.float().permute(0, 3, 1, 2), 2) * 4
    assert rel_err(nchw(xg.grad), ref.cpu()) < 4e-3


# ------------------------------------------------------------------------------------------------ convolutions
def _conv_case(B, cin, cout, H, W, k, stride, pad_tl, residual, force_direct, seed=0, with_bias=True):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(seed)
    x = O.r16(torch.randn(B, cin, H, W, generator=g))
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = 0.1 * torch.randn(cout, generator=g) if with_bias else None
    pt, pl = pad_tl
    if stride == 2:   # Downsample: pad right/bottom by one (models/flux_ae.py:91-95)
        xin = F.pad(x, (0, 1, 0, 1))
        pad = 0
    else:
        xin, pad = x, pt
    xr, wr = xin.clone().requires_grad_(True), O.r16(w).requires_grad_(True)
    y_ref = F.conv2d(xr, wr, b, stride=stride, padding=pad)
    res = None
    if residual:
        res = O.r16(torch.randn(y_ref.shape, generator=g))
        y_full = O.r16(O.r16(y_ref) + res)
    else:
        y_full = O.r16(y_ref)
    dy = O.r16(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    dx_ref = xr.grad[..., :H, :W] if stride == 2 else xr.grad

    pack = ops.WeightPack()
    wf, wd = pack.get(w.to(DEV))
    xc = nhwc(x)
    y = ops.conv_forward_raw(xc, wf, None if b is None else b.to(DEV), None if res is None else nhwc(res), k, k, stride, pad_tl,
                             force_direct=force_direct, pad_br=(1, 1) if stride == 2 else None)
    dyc = nhwc(dy)
    dx = ops.conv_dgrad_raw(dyc, wf, wd, (H, W), k, k, stride, pad_tl, force_direct=force_direct)
    dw = ops.conv_wgrad_raw(xc, dyc, k, k, stride, pad_tl, force_direct=force_direct)
    db = ops.bias_grad_raw(dyc)
    torch.cuda.synchronize()
    # bf16 outputs (1 ulp = 2^-8); fp32 accumulation over K = cin*k*k terms
    assert rel_err(nchw(y), y_full) < 4e-3, "forward"
    assert rel_err(nchw(dx), dx_ref) < 4e-3, "dgrad"
    assert rel_err(dw, wr.grad) < 1e-3, "wgrad"
    assert rel_err(db, dy.sum((0, 2, 3))) < 1e-3, "bias grad"
    # element-wise bound as well.  One bf16 ulp is between 2^-8 and 2^-7 relative; a rounding flip costs one ulp of
    # the conv term and, with the residual add, one more of the sum: <= 2 * 2^-7 of the larger of the two (the add can
    # cancel), plus fp32 accumulation-order noise scaled by the typical magnitude.
    yy, rr = nchw(y), y_full
    mag = torch.maximum(rr.abs(), y_ref.detach().abs())
    nulp = 2 if residual else 1
    assert ((yy - rr).abs() <= nulp * 2 ** -7 * mag + 2 ** -8 * rr.abs().mean()).all()


@pytest.mark.parametrize("case", [
    # B, cin, cout, H, W, k, stride, pad, residual
    (2, 32, 32, 8, 8, 3, 1, (1, 1), False),      # decoder stem conv_in.0.conv
    (1, 32, 64, 6, 10, 3, 1, (1, 1), True),
    (2, 128, 3, 16, 16, 3, 1, (1, 1), False),    # conv_out head (Cout=3 kernel)
    (2, 3, 64, 12, 12, 3, 1, (1, 1), False),     # encoder stem (Cin=3, scalar path)
    (2, 64, 64, 8, 8, 3, 2, (0, 0), False),      # Downsample
    (1, 64, 128, 9, 7, 1, 1, (0, 0), False),     # 1x1 on a ragged image
    (1, 40, 24, 5, 5, 3, 1, (1, 1), False),
])
def test_conv_direct(case):
    _conv_case(*case, force_direct=True)


@pytest.mark.parametrize("case", [
    (1, 64, 64, 8, 16, 3, 1, (1, 1), False),      # one tile, one k-chunk per tap, masked N
    (2, 128, 128, 16, 16, 3, 1, (1, 1), True),    # BN=128, residual epilogue
    (1, 64, 256, 16, 8, 1, 1, (0, 0), False),     # 1x1, BN=256
    (2, 256, 512, 8, 32, 3, 1, (1, 1), False),    # 2 N tiles, 4 k-chunks
    (1, 512, 512, 32, 32, 3, 1, (1, 1), True),    # decoder mid-block shape
    (1, 128, 128, 4, 256, 3, 1, (1, 1), False),   # W=256 row tiles
    (3, 32, 512, 32, 32, 3, 1, (1, 1), False),    # Cin=32: half-filled K chunk (TMA zero fill)
    (1, 256, 128, 64, 64, 1, 1, (0, 0), False),   # nin_shortcut shape
    (5, 128, 256, 16, 16, 3, 1, (1, 1), False),   # more tiles than one wave at small grid
])
def test_conv_tcgen05(case):
    _conv_case(*case, force_direct=False)


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("case", [
    (2, 128, 128, 16, 16, 3, 1, (1, 1), True),    # BN=128: one 256-pixel tile per image
    (1, 512, 512, 32, 32, 3, 1, (1, 1), True),    # BN=256, single TMEM accumulator set when MT=2
    (3, 64, 256, 8, 64, 1, 1, (0, 0), False),
    (2, 128, 128, 2, 256, 3, 1, (1, 1), False),   # W=256: a 256-pixel tile is one image row
])
def test_conv_tcgen05_tile_modes(case, mode):
    """All CTA tile shapes (128 or 256 pixels per CTA, and the cta_group::2 CTA pair) must give the same answers."""
    from dmvae_b200 import _lib
    _lib.query("dmvae_conv_tc_set_tile_mode", mode)
    try:
        _conv_case(*case, force_direct=False, seed=mode)
    finally:
        _lib.query("dmvae_conv_tc_set_tile_mode", 0)


@pytest.mark.parametrize("case", [
    (2, 128, 128, 16, 16, 3, 1, (1, 1), True),    # BN=128, 2 chunks, every tile touches the image border (TMA zero fill = padding)
    (1, 512, 512, 32, 32, 3, 1, (1, 1), True),    # BN=256, 2 N tiles, 8 chunks: halo and weight rings wrap several times
    (3, 256, 128, 32, 16, 3, 1, (1, 1), False),   # rectangular image, odd batch
    (1, 64, 256, 16, 64, 3, 1, (1, 1), False),    # one chunk per tile
    (2, 128, 256, 48, 24, 3, 1, (1, 1), True),    # H, W not powers of two
    (2, 128, 128, 32, 16, 3, 1, (1, 1), True),    # Cout = 128, H % 32 == 0: transposed tile (channels in TMEM lanes), residual
    (1, 256, 128, 64, 24, 3, 1, (1, 1), False),   # transposed tile, 4 chunks, 6 tiles
    (3, 64, 128, 32, 8, 3, 1, (1, 1), True),      # transposed tile, image exactly one tile wide
    (2, 64, 64, 32, 16, 3, 1, (1, 1), True),      # Cout = 64 (VGG conv1_2): half of the transposed tile is zero fill
    (1, 128, 192, 32, 16, 3, 1, (1, 1), False),   # Cout = 192: second channel tile half empty
    (2, 128, 3, 32, 16, 3, 1, (1, 1), False),     # the 128 -> 3 head on the transposed tile (3 live channel lanes)
    (1, 64, 5, 64, 8, 3, 1, (1, 1), True),        # thin output with a residual
])
def test_conv_tcgen05_halo_tiles(case):
    """Halo-resident CTA-pair tiles (one (16+2) x (8+2) pixel tile serves all nine taps) against the fp32 reference, and
    bit-for-bit agreement is NOT required against the per-tap kernels (different fp32 accumulation order)."""
    from dmvae_b200 import _lib
    _lib.query("dmvae_conv_tc_set_tile_mode", 8)
    try:
        _conv_case(*case, force_direct=False, seed=11)
    finally:
        _lib.query("dmvae_conv_tc_set_tile_mode", 7)


def test_conv_halo_matches_per_tap_kernel():
    """Same inputs through the halo tiles and through the per-tap tiles: fp32 accumulation order differs, outputs agree to 1 bf16 ulp."""
    from dmvae_b200 import _lib
    ops, _ = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(4, 64, 64, 256, generator=g, device=DEV).bfloat16()
    w = torch.randn(256, 256, 3, 3, generator=g, device=DEV) / 48
    b = torch.randn(256, generator=g, device=DEV)
    wf, _wd = ops.WeightPack().get(w)
    try:
        _lib.query("dmvae_conv_tc_set_tile_mode", 8)
        y_h = ops.conv_forward_raw(x, wf, b, None, 3, 3)
        _lib.query("dmvae_conv_tc_set_tile_mode", 6)
        y_t = ops.conv_forward_raw(x, wf, b, None, 3, 3)
    finally:
        _lib.query("dmvae_conv_tc_set_tile_mode", 7)
    torch.cuda.synchronize()
    d = (y_h.float() - y_t.float()).abs()
    # one bf16 ulp of y, plus the fp32 reassociation error itself where y is a near-cancellation of O(1) partial sums
    assert (d <= y_t.float().abs() * 2 ** -7 + 3e-5).all(), d.max()
    assert (d > 0).float().mean() < 0.02          # roundings flip only where the fp32 sums straddle a bf16 boundary


@pytest.mark.parametrize("mode", [1, 2, 3, 8])
@pytest.mark.parametrize("cin,cout,hw,res", [(64, 128, 16, True), (128, 256, 16, False), (256, 512, 16, True), (64, 32, 16, False),
                                             (64, 128, 32, True), (128, 256, 32, True), (128, 512, 32, False)])
def test_conv_epilogue_group_norm_stats(cin, cout, hw, res, mode):
    """The statistics the conv epilogue reduces must equal a separate gn_stats pass over the stored bf16 output."""
    from dmvae_b200 import _lib
    ops, _ = _ops()
    g = torch.Generator(device=DEV).manual_seed(cout)
    x = torch.randn(2, hw, hw, cin, generator=g, device=DEV).bfloat16()
    w = torch.randn(cout, cin, 3, 3, generator=g, device=DEV) / math.sqrt(cin * 9)
    b = torch.randn(cout, generator=g, device=DEV)
    r = torch.randn(2, hw, hw, cout, generator=g, device=DEV).bfloat16() if res else None
    wf, _wd = ops.WeightPack().get(w)
    _lib.query("dmvae_conv_tc_set_tile_mode", mode)
    old = ops.FUSE_GN_STATS_MIN_CPG
    ops.FUSE_GN_STATS_MIN_CPG = 1           # exercise every group width
    try:
        y = ops.conv_forward_raw(x, wf, b, r, 3, 3, want_gn_stats=True)
    finally:
        ops.FUSE_GN_STATS_MIN_CPG = old
        _lib.query("dmvae_conv_tc_set_tile_mode", 0)
        _lib.query("dmvae_conv_tc_set_tile_mode", 7)
    fused = ops._tagged_gn_stats(y)
    assert fused is not None
    ref = ops.gn_stats_raw(y)
    torch.cuda.synchronize()
    # same bf16 values, different fp32 partial-sum order: 1e-5 relative on sums of ~10^3..10^4 terms
    assert torch.allclose(fused, ref, rtol=1e-5, atol=1e-3), (fused - ref).abs().max()


def test_thin_head_gradients_as_gemms():
    """conv_out (128 -> 3): dx and dw through the gradient-patch matrix + tcgen05 1x1 GEMMs vs the CUDA-core kernels."""
    ops, _ = _ops()
    g = torch.Generator(device=DEV).manual_seed(9)
    B, H, W, cin, cout = 2, 32, 64, 128, 3
    x = torch.randn(B, H, W, cin, generator=g, device=DEV).bfloat16()
    dy = torch.randn(B, H, W, cout, generator=g, device=DEV).bfloat16()
    w = torch.randn(cout, cin, 3, 3, generator=g, device=DEV) / 34
    wf, wd = ops.WeightPack().get(w)
    assert ops._thin_output_ok(dy, wf, 3, 3, 1, (1, 1), (H, W))
    dx, dw = ops.thin_output_grads(x, dy, wf, 3, 3, (1, 1), True, True)
    dx_ref = ops.conv_dgrad_raw(dy, wf, wd, (H, W), 3, 3, force_direct=True)
    dw_ref = ops.conv_wgrad_raw(x, dy, 3, 3, force_direct=True)
    assert rel_err(dx.float(), dx_ref.float()) < 4e-3
    assert rel_err(dw, dw_ref) < 1e-3
    # and against the fp32 reference math
    xr = x.float().permute(0, 3, 1, 2).cpu().requires_grad_(True)
    wr = O.r16(w.cpu()).requires_grad_(True)
    F.conv2d(xr, wr, None, padding=1).backward(dy.float().permute(0, 3, 1, 2).cpu())
    assert rel_err(nchw(dx), xr.grad) < 4e-3 and rel_err(dw, wr.grad) < 1e-3


@pytest.mark.parametrize("case", [
    (2, 64, 64, 32, 32, 3, 2, (0, 0), False),       # Downsample at small width
    (1, 128, 128, 64, 64, 3, 2, (0, 0), False),     # encoder level 0 -> 1
    (2, 256, 256, 32, 64, 3, 2, (0, 0), False),     # pair kernel (Cout = 256), rectangular
    (1, 512, 512, 32, 32, 3, 2, (0, 0), False),
])
def test_conv_tcgen05_stride2(case):
    """flux_ae.Downsample on the tensor cores: TMA element strides (fwd, wgrad) and zero insertion (dgrad)."""
    _conv_case(*case, force_direct=False)


def test_conv_tc_many_tiles_matches_direct():
    """Full-size layer (512->512 @64x64, B=2: 128 pixel tiles x 2 N tiles): tensor-core path vs CUDA-core path on device."""
    ops, _ = _ops()
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(2, 64, 64, 512, generator=g, device=DEV).bfloat16()
    w = (torch.randn(512, 512, 3, 3, generator=g, device=DEV) / 68).float()
    b = torch.randn(512, generator=g, device=DEV)
    wf, wd = ops.WeightPack().get(w)
    y_tc = ops.conv_forward_raw(x, wf, b, None, 3, 3)
    y_d = ops.conv_forward_raw(x, wf, b, None, 3, 3, force_direct=True)
    assert rel_err(y_tc.float(), y_d.float()) < 3e-3
    dy = torch.randn(2, 64, 64, 512, generator=g, device=DEV).bfloat16()
    dw_tc = ops.conv_wgrad_raw(x, dy, 3, 3)
    dw_d = ops.conv_wgrad_raw(x, dy, 3, 3, force_direct=True)
    assert rel_err(dw_tc, dw_d) < 1e-3


@pytest.mark.parametrize("cin,cout,hw,res", [(512, 512, 128, True), (256, 256, 256, True), (128, 128, 256, False), (256, 128, 256, False)])
def test_conv_production_shapes_at_bench_batch(cin, cout, hw, res):
    """The decoder's largest layers at the BENCH batch (B = 16): forward with the fused residual, dgrad, and the split-K weight
    gradient whose contraction runs over K = 16 * hw^2 pixels (up to 1 048 576) with red.global.add fan-in.  Reference: fp32 cuDNN
    on the same GPU with TF32 off (the CPU oracle needs minutes at this size; torch fp32 is the stated reference for a
    floating-point kernel).  Tolerances as in _conv_case: bf16 outputs 4e-3, fp32-accumulated weight gradient 1e-3."""
    ops, _ = _ops()
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        B = 16
        g = torch.Generator(device=DEV).manual_seed(cin + cout + hw)
        x = torch.randn(B, hw, hw, cin, generator=g, device=DEV).bfloat16()
        w = (torch.randn(cout, cin, 3, 3, generator=g, device=DEV) / math.sqrt(9 * cin)).bfloat16().float()
        b = torch.randn(cout, generator=g, device=DEV)
        r = torch.randn(B, hw, hw, cout, generator=g, device=DEV).bfloat16() if res else None
        dy = torch.randn(B, hw, hw, cout, generator=g, device=DEV).bfloat16()
        wf, wd = ops.WeightPack().get(w)
        y = ops.conv_forward_raw(x, wf, b, r, 3, 3)
        dx = ops.conv_dgrad_raw(dy, wf, wd, (hw, hw), 3, 3)
        dw = ops.conv_wgrad_raw(x, dy, 3, 3)
        torch.cuda.synchronize()
        # reference in image chunks (fp32 activations of the whole batch would not be needed at once)
        dw_ref = torch.zeros_like(w, dtype=torch.float64)
        for i in range(0, B, 4):
            xs = x[i:i + 4].float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            wr = w.clone().requires_grad_(True)
            yr = F.conv2d(xs, wr, b, padding=1)
            y_full = yr.detach().bfloat16().float()
            if res:
                y_full = (y_full + r[i:i + 4].float().permute(0, 3, 1, 2)).bfloat16().float()
            yr.backward(dy[i:i + 4].float().permute(0, 3, 1, 2).contiguous())
            assert rel_err(y[i:i + 4].float().permute(0, 3, 1, 2), y_full) < 4e-3, "forward"
            assert rel_err(dx[i:i + 4].float().permute(0, 3, 1, 2), xs.grad) < 4e-3, "dgrad"
            dw_ref += wr.grad.double()
        assert rel_err(dw, dw_ref) < 1e-3, "wgrad (split-K over the whole batch)"
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


@pytest.mark.parametrize("B,cin,cout,H,W,mode", [
    (2, 256, 256, 32, 16, 8),      # halo-resident CTA pair (line-coalesced epilogue)
    (2, 128, 128, 32, 16, 8),      # transposed tile
    (2, 64, 64, 32, 16, 8),        # transposed tile, half of the channel lanes live (VGG conv1_2)
    (1, 128, 256, 16, 16, 1),      # single-CTA per-tap tile (per-thread-row epilogue)
    (1, 64, 128, 6, 10, 0),        # ragged image: CUDA-core kernel
])
def test_conv_relu_epilogue_and_relu_gated_dgrad(B, cin, cout, H, W, mode):
    """N3: y = relu(conv(x) + b) in the conv epilogue (DMVAE_CONV_RELU) and dx = dgrad(dy) * [x > 0] in the data-gradient epilogue
    (DMVAE_CONV_MASK: the backward of the ReLU that produced x), on every tile family, vs torch fp32 on the bf16-rounded operands."""
    from dmvae_b200 import _lib
    ops, _ = _ops()
    g = torch.Generator().manual_seed(B * 1000 + cin + cout)
    x = O.r16(torch.relu(torch.randn(B, cin, H, W, generator=g)))            # a ReLU output: about half zeros
    w = O.r16(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    b = torch.randn(cout, generator=g) * 0.1
    dy = O.r16(torch.randn(B, cout, H, W, generator=g))
    xr = x.clone().requires_grad_(True)
    y_ref = O.r16(torch.relu(F.conv2d(xr, w, b, padding=1)))
    F.conv2d(xr, w, None, padding=1).backward(dy)
    dx_ref = O.r16(xr.grad) * (x > 0)
    wf, wd = ops.WeightPack().get(w.to(DEV))
    _lib.query("dmvae_conv_tc_set_tile_mode", mode)
    try:
        y = ops.conv_forward_raw(nhwc(x), wf, b.to(DEV), None, 3, 3, flags=ops.EPI_RELU)
        dx = ops.conv_dgrad_raw(nhwc(dy), wf, wd, (H, W), 3, 3, relu_mask=nhwc(x))
    finally:
        _lib.query("dmvae_conv_tc_set_tile_mode", 0)
        _lib.query("dmvae_conv_tc_set_tile_mode", 7)
    torch.cuda.synchronize()
    assert (nchw(y) >= 0).all()
    assert rel_err(nchw(y), y_ref) < 4e-3
    assert rel_err(nchw(dx), dx_ref) < 4e-3
    assert (nchw(dx)[x == 0] == 0).all(), "gradient must be exactly zero where the ReLU output was zero"


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 16, 24), (1, 512, 4, 4), (3, 8, 2, 2)])
def test_maxpool_and_fused_pool_tap_backward(B, C, H, W):
    """dmvae_maxpool2x2_fwd / dmvae_pool_tap_bwd vs ATen max_pool2d (+ its backward, + the tap-gradient add, + ReLU's
    threshold_backward) on a ReLU output with exact ties (zeros) in most windows.  Bit-exact: selection and one bf16 add."""
    ops, _ = _ops()
    g = torch.Generator(device=DEV).manual_seed(C + H)
    y = torch.relu(torch.randn(B, H, W, C, generator=g, device=DEV)).bfloat16()
    y[:, ::2, ::2, : C // 2] = y[:, 1::2, 1::2, : C // 2]                  # non-zero ties inside a window: first maximum wins
    d_tap = torch.randn(B, H, W, C, generator=g, device=DEV).bfloat16()
    d_pool = torch.randn(B, H // 2, W // 2, C, generator=g, device=DEV).bfloat16()
    yl = y.clone().requires_grad_(True)
    tap, pooled = ops.pool_tap(yl)
    yr = y.permute(0, 3, 1, 2).clone().requires_grad_(True)
    pr = F.max_pool2d(yr, 2, 2)
    assert torch.equal(pooled.permute(0, 3, 1, 2), pr)
    (tap.float() * d_tap.float()).sum().backward(retain_graph=True)        # tap only: relu gate on the tap gradient
    assert torch.equal(yl.grad, torch.where(y > 0, d_tap, torch.zeros_like(d_tap)))
    yl.grad = None
    torch.autograd.backward([tap, pooled], [d_tap, d_pool])
    pr.backward(d_pool.permute(0, 3, 1, 2))
    ref = (yr.grad.permute(0, 2, 3, 1).float() + d_tap.float()).bfloat16()
    ref = torch.where(y > 0, ref, torch.zeros_like(ref))
    assert torch.equal(yl.grad, ref)


@pytest.mark.parametrize("B,cin,cout,H,W", [(2, 256, 256, 16, 16), (1, 512, 512, 32, 16), (2, 256, 512, 16, 8), (4, 512, 256, 16, 24)])
def test_subpixel_upsample_conv(B, cin, cout, H, W):
    """flux_ae.Upsample (nearest 2x + 3x3, models/flux_ae.py:103-107) in sub-pixel form -- forward, data gradient, weight gradient,
    bias gradient and the fused GroupNorm statistics -- against the fp32 oracle on bf16-rounded operands.  The tap sums are rounded
    to bf16 once (the reference rounds each 3x3 tap): forward / dgrad 6e-3 (vs 4e-3 for the plain conv), wgrad 1e-3."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(cin + cout + H)
    x = O.r16(torch.randn(B, cin, H, W, generator=g))
    w = O.r16(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    b = torch.randn(cout, generator=g) * 0.1
    dy = O.r16(torch.randn(B, cout, 2 * H, 2 * W, generator=g))
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y_ref = O.upsample({"u.conv.weight": wr, "u.conv.bias": b}, "u", xr, bf16=False)
    y_ref.backward(dy)
    xc = nhwc(x).requires_grad_(True)
    wc = w.to(DEV).requires_grad_(True)
    bc = b.to(DEV).requires_grad_(True)
    assert ops.upsample_conv_supported(xc, wc)
    y = ops.upsample_conv(xc, wc, bc, ops.SubpixelPack(), True)
    y.backward(nhwc(dy))
    torch.cuda.synchronize()
    assert y.shape == (B, 2 * H, 2 * W, cout)
    assert rel_err(nchw(y), y_ref.detach()) < 6e-3, "forward"
    assert rel_err(nchw(xc.grad), xr.grad) < 6e-3, "dgrad"
    assert rel_err(wc.grad, wr.grad) < 1e-3, "wgrad"
    assert rel_err(bc.grad, dy.sum((0, 2, 3))) < 1e-3, "bias grad"
    fused = ops._tagged_gn_stats(y)
    assert fused is not None
    assert torch.allclose(fused, ops.gn_stats_raw(y), rtol=1e-5, atol=1e-3)
    # tap-major (arena) weight storage gives the same packs and receives the folded gradient in place
    flat = torch.zeros(wc.numel(), device=DEV)
    wt = flat.view(9, cout, cin).permute(1, 2, 0).unflatten(2, (3, 3))
    wt.copy_(w.to(DEV))
    pf, pd = ops.SubpixelPack().get(wt)
    rf, rd = ops.SubpixelPack().get(w.to(DEV))
    assert torch.equal(pf, rf) and torch.equal(pd, rd)


def test_vit_glue_scale_residual_bit_exact():
    """x += float(y) * gamma must equal the reference's two ATen passes (x + y * gamma with type promotion) bit for bit."""
    from dmvae_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(3)
    rows, D = 16 * 257, 1024
    x = torch.randn(rows, D, generator=g, device=DEV)
    y = torch.randn(rows, D, generator=g, device=DEV).bfloat16()
    gamma = torch.randn(D, generator=g, device=DEV) * 0.1
    ref = x + y * gamma
    out = x.clone()
    _lib.call("dmvae_scale_residual", _lib.ptr(out), _lib.ptr(y), _lib.ptr(gamma), rows, D)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("D", [384, 768, 1024])
def test_vit_glue_layernorm_bf16(D):
    """fp32 LayerNorm written as bf16 vs F.layer_norm(...).to(bf16): same fp32 value up to reassociation, so at most one bf16
    ulp apart and almost everywhere identical."""
    from dmvae_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(D)
    rows = 4112
    x = torch.randn(rows, D, generator=g, device=DEV) * 2 + 0.3
    w = torch.rand(D, generator=g, device=DEV) + 0.5
    b = torch.randn(D, generator=g, device=DEV) * 0.1
    ref = F.layer_norm(x, (D,), w, b, 1e-6)
    y = torch.empty(rows, D, dtype=torch.bfloat16, device=DEV)
    _lib.call("dmvae_layernorm_bf16", _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), rows, D, 1e-6)
    torch.cuda.synchronize()
    d = (y.float() - ref.to(torch.bfloat16).float()).abs()
    assert (d <= ref.abs() * 2 ** -7 + 1e-6).all(), d.max()
    assert (d > 0).float().mean() < 1e-3
    assert rel_err(y.float(), ref) < 3e-3            # bf16 rounding of the output itself


@pytest.mark.parametrize("D", [384, 768, 1024])
def test_vit_glue_scale_residual_layernorm_is_the_two_kernels_back_to_back(D):
    """dmvae_scale_residual_layernorm = dmvae_scale_residual followed by dmvae_layernorm_bf16, bit for bit (residual stream and the
    bf16 LayerNorm output), on an odd row count (16 x 257 tokens)."""
    from dmvae_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(D + 1)
    rows = 4112
    x0 = torch.randn(rows, D, generator=g, device=DEV) * 2 + 0.3
    y = torch.randn(rows, D, generator=g, device=DEV).bfloat16()
    gamma = torch.randn(D, generator=g, device=DEV) * 0.2
    w = torch.rand(D, generator=g, device=DEV) + 0.5
    b = torch.randn(D, generator=g, device=DEV) * 0.1
    xa, xb = x0.clone(), x0.clone()
    oa = torch.empty(rows, D, dtype=torch.bfloat16, device=DEV)
    ob = torch.empty_like(oa)
    _lib.call("dmvae_scale_residual", _lib.ptr(xa), _lib.ptr(y), _lib.ptr(gamma), rows, D)
    _lib.call("dmvae_layernorm_bf16", _lib.ptr(xa), _lib.ptr(w), _lib.ptr(b), _lib.ptr(oa), rows, D, 1e-6)
    _lib.call("dmvae_scale_residual_layernorm", _lib.ptr(xb), _lib.ptr(y), _lib.ptr(gamma), _lib.ptr(w), _lib.ptr(b), _lib.ptr(ob), rows, D, 1e-6)
    torch.cuda.synchronize()
    assert torch.equal(xa, xb) and torch.equal(oa, ob)
