/*
 * libdmvae_b200.so -- C ABI of the B200-native DMVAE training hot path.
 *
 * The reference (sen-ye/dmvae) has no FFI: its hot path is PyTorch nn.Module composition whose arithmetic runs in
 * cuDNN / ATen library kernels.  Each entry point below replaces one such library call site; the reference
 * file:line it stands in for is given per function.  Host code (dmvae_b200/*.py) binds these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch's allocator in practice); the library keeps
 *     no state except a TMA-descriptor cache keyed by (pointer, shape);
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises with the host inside a call;
 *   - activations are channels-last bf16:  x[b][h][w][c], c fastest ("NHWC");
 *   - packed weights are bf16 [tap][Cout][Cin] (see dmvae_pack_weights);
 *   - reductions are returned through caller-zeroed fp64 accumulators on the device (no host round trip);
 *   - return value: 0 on success, negative DMVAE_E* otherwise; dmvae_last_error() describes the failure.
 *   - dtype codes: DMVAE_F32 = 0, DMVAE_BF16 = 1.
 */
#ifndef DMVAE_B200_H
#define DMVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMVAE_OK 0
#define DMVAE_EINVAL (-1)
#define DMVAE_ECUDA (-2)
#define DMVAE_EUNSUPPORTED (-3)

#define DMVAE_F32 0
#define DMVAE_BF16 1

/* Bumped whenever a prototype below changes; the ctypes binding (dmvae_b200/_lib.py) refuses a library whose
 * dmvae_abi_version() differs from the table it was written against. */
#define DMVAE_ABI_VERSION 8

const char* dmvae_last_error(void);
int dmvae_abi_version(void);
/* 0 if the current CUDA device is sm_100 (B200); the host layer refuses to run otherwise. */
int dmvae_check_device(void);

/* ------------------------------------------------------------------ A3: DMD loss (train_dmd.py:204-230) ---- */

/* xt = t*z + (1-t)*x0 with the reference's per-op rounding.  Replaces ICPlan.plan / compute_mu_t
 * (diffusion/transport/path.py:114-136) as called from train_dmd.py:210.  z,x0,xt: [B][per_sample]; t: [B]. */
int dmvae_dmd_mix_xt(const void* z, const void* x0, const void* t, void* xt, int64_t B, int64_t per_sample,
                     int dtype, void* stream);

/* Fused CFG mix + pred + p_real/p_student + per-sample mean|p_real| normaliser + nan_to_num + 0.5*MSE surrogate
 * + latent gradient.  Replaces the ~25 ATen kernels of train_dmd.py:214-228 (normalize=1, cfg_scale>1) and the
 * toy variant toy_example_2d/dmd.py:349-360 (normalize=0, vT_u=vS_u=NULL).
 *   acc[0] += sum (z - target)^2   (loss = 0.5 * acc[0] / (B*per_sample))
 *   acc[1] += sum_b ||grad_b||_2   (dmd_gradient_norm = acc[1] / B)
 *   dz      = grad_scale * (z - target) / (B*per_sample)       stored as dz_dtype */
int dmvae_dmd_loss_fwd_bwd(const void* z, const void* xt, const void* t, const void* vT_c, const void* vT_u,
                           const void* vS_c, const void* vS_u, void* dz, double* acc, int64_t B,
                           int64_t per_sample, float cfg_scale, int normalize, float grad_scale, int dtype,
                           int dz_dtype, void* stream);

/* ------------------------------------------------------------------ A4: L1 + L2 (train_dmd.py:234-235) ----- */

/* acc[0] += sum|recon-image| ; acc[1] += sum (recon-image)^2.   Replaces F.l1_loss + F.mse_loss forward. */
int dmvae_l1l2_fwd(const float* recon, const float* image, double* acc, int64_t n, void* stream);
/* d_recon = (w_l1*g1*sign(d) + 2*w_l2*g2*d)/n ; g1,g2 device scalars (upstream grads) or NULL (=0 contribution). */
int dmvae_l1l2_bwd(const float* recon, const float* image, float* d_recon, const float* g1, const float* g2,
                   int64_t n, float w_l1, float w_l2, void* stream);
/* single pass: sums as dmvae_l1l2_fwd and d_recon = d(w_l1*L1 + w_l2*L2)/d recon. */
int dmvae_l1l2_fwd_bwd(const float* recon, const float* image, float* d_recon, double* acc, int64_t n,
                       float w_l1, float w_l2, void* stream);

/* ------------------------------------------------------------------ A5: LPIPS distance (utils/lpips.py:86-94) */

/* One VGG tap: acc[b] += sum_pixels sum_c w[c] * (f0/(|f0|+1e-10) - f1/(|f1|+1e-10))^2.   Replaces
 * normalize_tensor (:156), the squared difference (:89), NetLinLayer 1x1 conv (:91,107) and the spatial sum of
 * spatial_average (:161).  f0,f1: channels-last [B][HW][C]; lin_w: fp32 [C].  faithful=1 applies the bf16
 * roundings autocast inserts around the 1x1 conv. */
int dmvae_lpips_dist_fwd(const void* f0, const void* f1, const float* lin_w, double* acc, int64_t B, int64_t HW,
                         int C, int dtype, int faithful, void* stream);
/* df1 = scale * (*gout) * d(distance)/d f1   (the reference passes (images, recon): only f1 needs a gradient). */
int dmvae_lpips_dist_bwd(const void* f0, const void* f1, const float* lin_w, void* df1, const float* gout,
                         int64_t B, int64_t HW, int C, float scale, int dtype, void* stream);

/* ------------------------------------------------------------------ A8: reparameterize + KL (extension) ---- */

/* h rows = [mu | logvar], each `half` long; z = mu + exp(lv/2)*eps; acc[0] += 0.5*sum(mu^2 + e^lv - 1 - lv).
 * No reference implementation exists; oracle/dmvae_oracle.py:reparam_kl states the math. */
int dmvae_reparam_kl_fwd(const void* h, const void* eps, void* z, double* acc, int64_t rows, int64_t half,
                         int dtype, void* stream);
int dmvae_reparam_kl_bwd(const void* h, const void* eps, const void* dz, void* dh, const float* g_kl,
                         float kl_scale, int64_t rows, int64_t half, int dtype, void* stream);

/* ------------------------------------------------------------------ A1a: GroupNorm(32)+swish ---------------- */
/* models/flux_ae.py:21-22,30,62,64,157,236.  x,y,da,dx: channels-last bf16 [B][HW][C]; stats fp64 [B][32][2]. */

/* stats[b][g] += {sum x, sum x^2}  (caller zeroes). */
int dmvae_gn_stats(const void* x, double* stats, int64_t B, int64_t HW, int C, void* stream);
/* y = [swish]((x-mean)*rstd*gamma+beta) rounded to bf16. */
int dmvae_gn_apply(const void* x, const double* stats, const float* gamma, const float* beta, void* y, int64_t B,
                   int64_t HW, int C, float eps, int silu, void* stream);
/* backward of the above: dx (+= dres if given: the gradient arriving over the residual branch, :82), dgamma/dbeta
 * (fp32 [C], accumulated), gsum = fp64 scratch [B][32][2] (caller zeroes).  dx_colsum (optional, fp32 [C],
 * accumulated) receives sum_pixels dx: the bias gradient of the conv whose output x is. */
int dmvae_gn_bwd(const void* da, const void* x, const double* stats, const float* gamma, const float* beta,
                 double* gsum, float* dgamma, float* dbeta, const void* dres, void* dx, float* dx_colsum,
                 int64_t B, int64_t HW, int C, float eps, int silu, void* stream);

/* ------------------------------------------------------------------ A1/A2: convolutions --------------------- */

/* fp32 master weight [Cout][Cin][KH][KW] (the state_dict tensor) -> bf16 GEMM operands:
 *   w_fwd  [tap][Cout][Cin]            forward / wgrad layout
 *   w_dgrad[taps-1-tap][Cin][Cout]     data-gradient layout (either may be NULL). */
int dmvae_pack_weights(const float* w, void* w_fwd, void* w_dgrad, int Cout, int Cin, int KH, int KW,
                       void* stream);
/* w_dgrad from an existing bf16 w_fwd (the optimizer kernel maintains w_fwd for tap-major parameter arenas, see N2). */
int dmvae_pack_dgrad_bf16(const void* w_fwd, void* w_dgrad, int Cout, int Cin, int taps, void* stream);
/* The same for every conv weight of a flat bf16 parameter arena in ONE launch (issued by the optimizer after its update kernel).
 * desc: device int64 [n_desc][5] = {element offset of the parameter in both arenas, Cout, Cin, taps, index of its first 32x32
 * tile}, first-tile indices ascending from 0; total_tiles = sum over parameters of taps * ceil(Cout/32) * ceil(Cin/32). */
int dmvae_pack_dgrad_batched(const void* w_fwd_flat, void* w_dgrad_flat, const int64_t* desc, int n_desc, int64_t total_tiles,
                             void* stream);

/* 1 if the shape runs on the tcgen05 tile (stride 1, 3x3 pad 1 or 1x1, C%8==0, pixel tile divides H,W). */
int dmvae_conv_tc_supported(int B, int H, int W, int Cin, int Cout, int KH, int KW);

/* tcgen05 implicit GEMM.  Replaces cuDNN conv forward for nn.Conv2d at models/flux_ae.py:32-35,63,65,67,101,210,237
 * (and utils/lpips.py:116-153, the frozen VGG16) and -- fed dY and w_dgrad -- cuDNN's backward-data.
 * y = conv(x) + bias ; if residual: y = bf16(y) + residual.
 * gn_stats (optional, fp64 [B][32][2], caller-zeroed): the epilogue also accumulates {sum y, sum y^2} per (image,
 * GroupNorm group) of the stored bf16 values, i.e. the output of dmvae_gn_stats for the next GroupNorm(32).
 * flags: DMVAE_CONV_RELU  y = max(conv(x) + bias, 0) -- the nn.ReLU after every VGG16 conv (utils/lpips.py:116-153);
 *        DMVAE_CONV_MASK  `residual` is not added: it gates the output, y = residual > 0 ? y : 0 -- ATen's ReLU backward
 *                         (threshold_backward) for the ReLU whose output this data-gradient conv's input was. */
#define DMVAE_CONV_RELU 1
#define DMVAE_CONV_MASK 2
int dmvae_conv_tc_fwd(const void* x, const void* w_packed, const float* bias, const void* residual, void* y,
                      double* gn_stats, int B, int H, int W, int Cin, int Cout, int KH, int KW, int flags, void* stream);

/* Stride-2 3x3 convolution (flux_ae.Downsample, :85-95: pad (0,1,0,1) => pad_top = pad_left = 0) on the tcgen05 tile:
 * the A operand is sampled with TMA element strides, the right/bottom padding is the TMA out-of-bounds fill.
 * Forward, weight gradient (tap-major scratch as dmvae_conv_tc_wgrad), and -- via dmvae_zero_insert2x + a stride-1
 * dmvae_conv_tc_fwd with the dgrad weight pack -- the data gradient. */
int dmvae_conv_tc_strided_supported(int B, int IH, int IW, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride);
int dmvae_conv_tc_fwd_strided(const void* x, const void* w_packed, const float* bias, void* y, int B, int IH, int IW,
                              int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad_top, int pad_left,
                              void* stream);
int dmvae_conv_tc_wgrad_strided(const void* x, const void* dy, float* dw_tap_major, int B, int IH, int IW, int Cin,
                                int OH, int OW, int Cout, int KH, int KW, int stride, int pad_top, int pad_left,
                                void* stream);
/* dyz[b][2oh+1][2ow+1][c] = dy[b][oh][ow][c], zero elsewhere (bf16, C % 8 == 0). */
int dmvae_zero_insert2x(const void* dy, void* dyz, int64_t B, int OH, int OW, int C, void* stream);

/* Sub-pixel form of flux_ae.Upsample (models/flux_ae.py:98-107: F.interpolate(scale_factor=2, "nearest") + 3x3 conv): every
 * output phase (py, px) is a 2x2 conv on the LOW-RES input whose taps are sums of the 3x3 taps falling on the same input pixel --
 * 16 instead of 36 tap-GEMMs per low-res pixel, no 4x larger intermediate.  H, W below are the LOW-RES size; x [B][H][W][Cin],
 * y / dy [B][2H][2W][Cout]; Cin, Cout multiples of 256, W % 8 == 0, H % 16 == 0.
 *   dmvae_subpixel_pack        w3 (fp32, any element strides) -> wp_fwd[16][Cout][Cin], wp_dgrad[16][Cin][Cout] (bf16)
 *   dmvae_conv_up2x_fwd        y = conv(nearest2x(x)) + bias (+ GroupNorm statistics of y, as dmvae_conv_tc_fwd)
 *   dmvae_conv_up2x_dgrad      dx from dy (all four phases accumulated in one tile; dy sampled with TMA element stride 2)
 *   dmvae_conv_up2x_wgrad      dwp[16][Cout][Cin] (fp32) += per-phase-tap gradients
 *   dmvae_subpixel_fold_wgrad  dw3 (fp32, any element strides) += fold of dwp onto the nine 3x3 taps
 * Summing taps before the bf16 rounding changes the rounding points relative to the reference (one rounding of the tap sum
 * instead of one per tap): same bf16-level tolerance as every conv here, but not the same bits as the two-kernel form. */
int dmvae_conv_up2x_supported(int B, int H, int W, int Cin, int Cout);
int dmvae_subpixel_pack(const float* w3, int64_t stride_co, int64_t stride_ci, int64_t stride_tap, void* wp_fwd, void* wp_dgrad,
                        int Cout, int Cin, void* stream);
int dmvae_conv_up2x_fwd(const void* x, const void* wp_fwd, const float* bias, void* y, double* gn_stats, int B, int H, int W,
                        int Cin, int Cout, void* stream);
int dmvae_conv_up2x_dgrad(const void* dy, const void* wp_dgrad, void* dx, int B, int H, int W, int Cin, int Cout, void* stream);
int dmvae_conv_up2x_wgrad(const void* x, const void* dy, float* dwp, int B, int H, int W, int Cin, int Cout, void* stream);
int dmvae_subpixel_fold_wgrad(const float* dwp, float* dw3, int64_t stride_co, int64_t stride_ci, int64_t stride_tap, int Cout,
                              int Cin, void* stream);

/* Tuning / test hook (host only, process-wide).  The conv entry points pick a tile per shape:
 *   3x3, Cout % 256 == 0 ........ halo-resident CTA pair (one TMA halo serves all nine taps, cta_group::2, N = 256)
 *   3x3, Cout = 64 / 128 / <32 .. transposed tile (M = channels, N = 256 pixels out of one halo)
 *   1x1, stride 2, small images . per-tap operand fetch: CTA pair, or a single-CTA tile of 128 / 256 pixels
 * mode 0 = that heuristic; 1 / 2 = force single-CTA tiles of 128 / 256 pixels; 3 = force the per-tap CTA pair;
 * 4 / 5 = CTA pairs preferred / never; 6 / 7 = halo + transposed tiles off / on; 8 = halo + transposed tiles wherever the shape
 * allows, ignoring the occupancy thresholds (tests); 12 / 13 / 14 = taps per CTA (2 / 3 / 4) of the 128-channel weight gradient. */
int dmvae_conv_tc_set_tile_mode(int mode);
/* Measurement hook (host only): the tile kernel launched by the most recent tcgen05 conv entry point called on this thread:
 * 1 conv_tc_kernel, 2 conv_tc2_kernel (per-tap pair), 3 conv_tc2h_kernel (halo pair, incl. the sub-pixel modes),
 * 4 conv_tcT_kernel (transposed), 5 conv_tc_wgrad_kernel, 6 conv_tc_wgrad2_kernel (pair); 0 = none yet.  bench.py uses it to
 * attribute each timed launch to its kernel. */
int dmvae_conv_tc_last_kernel(void);

/* tcgen05 weight gradient (both operands MN-major): dw_tap_major[tap][Cout][Cin] (fp32, caller-zeroed or
 * accumulated) += sum_pixels dy[p][co] * x[p(+)tap][ci].  Replaces cuDNN backward-filter for the same call sites.
 * dmvae_wgrad_unpack moves the tap-major scratch into the state_dict layout dw[Cout][Cin][KH][KW]. */
int dmvae_conv_tc_wgrad_supported(int B, int H, int W, int Cin, int Cout, int KH, int KW);
int dmvae_conv_tc_wgrad(const void* x, const void* dy, float* dw_tap_major, int B, int H, int W, int Cin, int Cout,
                        int KH, int KW, void* stream);
int dmvae_wgrad_unpack(const float* dw_tap_major, float* dw, int Cout, int Cin, int taps, int accumulate,
                       void* stream);

/* Gradient patches for convs with a handful of output channels (the C->3 head, :237): P[pixel][tap*Cout+co] (bf16, 32
 * columns, zero padded) = dy[pixel - offset(tap)][co].  Both of that conv's gradients then run as 1x1 GEMMs on the tcgen05
 * tiles (dmvae_conv_tc_fwd with Cin=32 for dx, dmvae_conv_tc_wgrad with Cout=32 for dw). */
int dmvae_grad_patches(const void* dy, void* patches, int64_t B, int H, int W, int Cout, int KH, int KW, int pad_top,
                       int pad_left, void* stream);

/* CUDA-core path for the ragged layers (Cin=32 stem :272-275, Cout=3 head :237, Cin=3 encoder stem :133,
 * stride-2 Downsample :89-95) and the on-device cross-check of the tensor-core path. */
int dmvae_conv_direct_fwd(const void* x, const void* w_packed, const float* bias, const void* residual, void* y,
                          int B, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride,
                          int pad_top, int pad_left, int flags /* DMVAE_CONV_* */, void* stream);
int dmvae_conv_direct_dgrad_strided(const void* dy, const void* w_packed, void* dx, int B, int H, int W, int Cin,
                                    int OH, int OW, int Cout, int KH, int KW, int stride, int pad_top,
                                    int pad_left, void* stream);
int dmvae_conv_direct_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int OH, int OW,
                            int Cout, int KH, int KW, int stride, int pad_top, int pad_left, void* stream);
/* dbias[c] += sum_rows dy[row][c]. */
int dmvae_bias_grad(const void* dy, float* dbias, int64_t M, int C, void* stream);

/* ------------------------------------------------------------------ A1c / A6: data movement ---------------- */

/* F.interpolate(scale_factor=2, mode="nearest") (models/flux_ae.py:104) and its adjoint, channels-last bf16. */
int dmvae_upsample2x_fwd(const void* x, void* y, int64_t B, int H, int W, int C, void* stream);
int dmvae_upsample2x_bwd(const void* dy, void* dx, int64_t B, int H, int W, int C, void* stream);
/* NCHW (fp32|bf16) <-> channels-last bf16 at the module boundary (models/flux_ae.py:244-245, models/vae.py:97). */
int dmvae_nchw_to_nhwc(const void* src, void* dst, int64_t B, int C, int64_t HW, int src_dtype, void* stream);
int dmvae_nhwc_to_nchw(const void* src, void* dst, int64_t B, int C, int64_t HW, int dst_dtype, void* stream);
/* out = a + b (bf16, fp32 add, one rounding): gradient fan-in of the residual branches (:52, :82). */
int dmvae_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);

/* ------------------------------------------------------------------ N3: VGG16 glue of LPIPS ------------------ */
/* nn.MaxPool2d(2, 2) between the VGG16 slices (utils/lpips.py:127-153, torchvision vgg16.features[4,9,16,23]) on channels-last
 * bf16, C % 8 == 0: y [B][2*OH][2*OW][C] -> pooled [B][OH][OW][C]. */
int dmvae_maxpool2x2_fwd(const void* y, void* pooled, int64_t B, int OH, int OW, int C, void* stream);
/* Backward of a slice boundary  y -> (LPIPS tap = y, pooled = maxpool2x2(y)), y = relu(.):  one kernel instead of ATen's
 * max_pool2d backward + the autograd gradient add + ReLU's threshold_backward:
 *   dy = (route(d_pooled to the window's first maximum) + d_tap) * [y > 0]   (d_tap may be NULL; relu = 0 skips the gate). */
int dmvae_pool_tap_bwd(const void* y, const void* d_pooled, const void* d_tap, void* dy, int64_t B, int OH, int OW, int C,
                       int relu, void* stream);
/* out = y > 0 ? dy : 0 over n bf16 elements (n % 8 == 0): ReLU backward where no conv / pool epilogue can carry it (relu5_3). */
int dmvae_relu_mask(const void* y, const void* dy, void* out, int64_t n, void* stream);

/* ------------------------------------------------------------------ A6: frozen-encoder glue ------------------ */
/* The timm ViT of models/vae.py:34-53 under autocast + no_grad (train_tokenizer.py:295-297 freezes it): each block computes
 * x = x + ls(attn(norm1(x))); x = x + ls(mlp(norm2(x))) with an fp32 residual stream and bf16 Linear outputs.
 * dmvae_scale_residual: x[r][d] += float(y[r][d]) * gamma[d] -- LayerScale multiply + residual add in one pass, each rounded
 *   separately as the reference's two ATen kernels do (bit-identical).
 * dmvae_layernorm_bf16: LayerNorm(eps) over the last dim of fp32 x, written as bf16 (autocast runs layer_norm in fp32 and the
 *   next Linear casts its input to bf16). */
int dmvae_scale_residual(float* x, const void* y_bf16, const float* gamma, int64_t rows, int D, void* stream);
int dmvae_layernorm_bf16(const float* x, const float* weight, const float* bias, void* y_bf16, int64_t rows, int D, float eps, void* stream);
/* The two above back to back in one pass (the row stays in registers): x += float(y) * gamma, then out = bf16(LayerNorm(x)) with
 * the NEXT branch's LayerNorm parameters -- bit-identical to the two-call sequence, one launch and one read of x fewer. */
int dmvae_scale_residual_layernorm(float* x, const void* y_bf16, const float* gamma, const float* weight, const float* bias,
                                   void* out_bf16, int64_t rows, int D, float eps, void* stream);

/* ------------------------------------------------------------------ N1: LightningDiT glue (no-grad scoring passes) ---- */
/* The teacher / student velocity networks are evaluated without autograd four times per VAE turn (train_dmd.py:212-217); the
 * reference runs their elementwise chains through @torch.compile (diffusion/lightningdit/lightningdit.py:27,135,165,241,269).
 * dmvae_rmsnorm_modulate: y (bf16) = RMSNorm(x) * (1 + scale[b]) + shift[b] -- `modulate(norm(x), shift, scale)` of
 *   lightningdit.py:27-31,243-249 over rms_norm.py:34-77; shift / scale are bf16 rows of the adaLN output (row b at element
 *   offset b * mod_stride), either may be NULL; D in {256, 768, 1152}.
 * dmvae_qk_norm_rope: the qkv Linear's output [B][N][3][H][hd] -> q, k (RMSNorm over hd, then the 2-D rotary embedding of
 *   pos_embed.py:96-134) and v, each [B][H][N][hd] bf16, ready for SDPA (lightningdit.py:70-85). */
int dmvae_rmsnorm_modulate(const void* x, int x_dtype, const float* weight, const void* shift, const void* scale, int64_t mod_stride,
                           void* y, int64_t rows, int tokens, int D, float eps, void* stream);
int dmvae_qk_norm_rope(const void* qkv, const float* wq, const float* wk, const float* cos_tab, const float* sin_tab, void* q, void* k,
                       void* v, int64_t B, int N, int H, int hd, float eps, void* stream);

/* ------------------------------------------------------------------ N2: fused optimizer step ---------------- */
/* Replaces clip_grad_norm_ + AdamW.step + update_ema (train_tokenizer.py:140-150,415-417,437; train_dmd.py:540-544) on
 * flat fp32 arenas: sumsq[0] += sum g^2 ; then g *= min(1, max_norm/(||g||+1e-6)), AdamW (torch semantics), EMA.
 * w16 (optional): bf16 copy of the updated parameters in the same element order -- with 3x3 conv weights kept tap-major in the
 * arenas this is the conv tiles' packed forward operand, i.e. the weight re-pack after optimizer.step() rides in
 * this pass.  dmvae_cast_bf16 rebuilds it outside a step (construction, load_state_dict). */
int dmvae_grad_sumsq(const float* g, double* sumsq, int64_t n, void* stream);
int dmvae_adamw_ema_step(float* p, float* g, float* m, float* v, float* ema, void* w16, const double* sumsq, float* norm_out,
                         int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                         float max_norm, float ema_decay, void* stream);
int dmvae_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMVAE_B200_H */
