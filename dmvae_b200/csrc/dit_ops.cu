// LightningDiT glue kernels for the no-grad scoring passes of the DMD loss: train_dmd.py:212-217 runs the teacher
// and the student four times per VAE turn without autograd; in the reference those passes go through five @torch.compile sites
// (diffusion/lightningdit/lightningdit.py:27,135,165,241,269).  Two row kernels replace the elementwise chains around the GEMMs:
//
//   rmsnorm_modulate   y = RMSNorm(x) * (1 + scale) + shift  -> bf16      (lightningdit.py:27-31 `modulate` over rms_norm.py:34-77;
//                      the Linear that follows casts its input to bf16 under autocast, so bf16 is what leaves the kernel)
//   qk_norm_rope       qkv [B][N][3][H][hd] -> q, k: RMSNorm over hd, 2-D rotary embedding; v: copy; all as [B][H][N][hd] bf16
//                      (lightningdit.py:70-85: q_norm / k_norm, rope(q), rope(k), then SDPA, which casts q, k to bf16)
//
// Rounding points follow the reference under torch.autocast(bf16): the normalised value is cast back to the input's dtype before
// the fp32 gain (rms_norm.py:76 `.type_as(x)`), (1 + scale) is a bf16 tensor op, everything else is fp32 until the final cast.
#include "common.cuh"

template <typename TX, int VPL>      // VPL float4-sized vectors per lane: D = 128 * VPL
__global__ void __launch_bounds__(256) rmsnorm_modulate_kernel(const TX* __restrict__ x, const float* __restrict__ w,
                                                               const bf16* __restrict__ shift, const bf16* __restrict__ scale,
                                                               int64_t mod_stride, bf16* __restrict__ y, int64_t rows, int tokens, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    constexpr int D = 128 * VPL;
    const int64_t b = row / tokens;
    float v[VPL][4];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (lane + 32 * i) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[i][e] = ld_as_float(x, row * D + c + e); q = fmaf(v[i][e], v[i][e], q); }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q * (1.f / D) + eps);
    bf16* yr = y + row * D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (lane + 32 * i) * 4;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float t = rnd<TX>(v[i][e] * rstd) * __ldg(w + c + e);                          // .type_as(x) * weight
            if (scale) t *= bf16_round(1.f + __bfloat162float(scale[b * mod_stride + c + e]));  // (1 + scale): a bf16 tensor op
            if (shift) t += __bfloat162float(shift[b * mod_stride + c + e]);
            o[e] = t;
        }
        uint2 pk;
        pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]);
        *reinterpret_cast<uint2*>(yr + c) = pk;
    }
}

// y[rows][D] (bf16) = RMSNorm(x; weight, eps) * (1 + scale[b]) + shift[b],  b = row / tokens.  x: fp32 or bf16 [rows][D];
// shift / scale: bf16, row b at element offset b * mod_stride (views into the adaLN output), either may be NULL.  D in {256, 768, 1152}.
DMVAE_API int dmvae_rmsnorm_modulate(const void* x, int x_dtype, const float* weight, const void* shift, const void* scale,
                                     int64_t mod_stride, void* y, int64_t rows, int tokens, int D, float eps, void* stream) {
    DMVAE_CHECK_ARG(x && weight && y, "rmsnorm_modulate: null pointer");
    DMVAE_CHECK_ARG(rows >= 0 && tokens > 0, "rmsnorm_modulate: bad size");
    DMVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 7) == 0, "rmsnorm_modulate: buffers must be 8-byte aligned");
    if (D != 256 && D != 768 && D != 1152)
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "rmsnorm_modulate: D=%d not supported (256, 768, 1152)", D);
    if (x_dtype != DMVAE_F32 && x_dtype != DMVAE_BF16) return dmvae_set_error(DMVAE_EINVAL, "rmsnorm_modulate: bad dtype %d", x_dtype);
    if (rows == 0) return DMVAE_OK;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    const bf16 *sh = (const bf16*)shift, *sc = (const bf16*)scale;
#define RM(TX, V) rmsnorm_modulate_kernel<TX, V><<<grid, 256, 0, st>>>((const TX*)x, weight, sh, sc, mod_stride, (bf16*)y, rows, tokens, eps)
    if (x_dtype == DMVAE_F32) { if (D == 256) RM(float, 2); else if (D == 768) RM(float, 6); else RM(float, 9); }
    else { if (D == 256) RM(bf16, 2); else if (D == 768) RM(bf16, 6); else RM(bf16, 9); }
#undef RM
    DMVAE_CHECK_LAUNCH("rmsnorm_modulate_kernel");
    return DMVAE_OK;
}

// one warp per (b, n, head): q and k rows are normalised and rotated, the v row is copied
__global__ void __launch_bounds__(256) qk_norm_rope_kernel(const bf16* __restrict__ qkv, const float* __restrict__ wq, const float* __restrict__ wk,
                                                           const float* __restrict__ cosb, const float* __restrict__ sinb,
                                                           bf16* __restrict__ q, bf16* __restrict__ k, bf16* __restrict__ v,
                                                           int64_t B, int N, int H, int hd, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (b, n, h)
    if (r >= B * N * H) return;
    const int h = (int)(r % H);
    const int n = (int)((r / H) % N);
    const int64_t b = r / ((int64_t)H * N);
    const int64_t dst = ((b * H + h) * N + n) * hd;
    const int pairs = hd >> 1;
#pragma unroll
    for (int which = 0; which < 3; ++which) {
        const bf16* src = qkv + (((b * N + n) * 3 + which) * H + h) * (int64_t)hd;
        bf16* out = (which == 0 ? q : which == 1 ? k : v) + dst;
        if (which == 2) {
            for (int p = lane; p < pairs; p += 32) *reinterpret_cast<uint32_t*>(out + 2 * p) = *reinterpret_cast<const uint32_t*>(src + 2 * p);
            continue;
        }
        const float* w = which == 0 ? wq : wk;
        float a0[3], a1[3];                               // up to 3 pairs per lane: hd <= 192
        float ss = 0.f;
        int cnt = 0;
        for (int p = lane; p < pairs; p += 32, ++cnt) {
            const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(src + 2 * p);
            a0[cnt] = __low2float(t); a1[cnt] = __high2float(t);
            ss = fmaf(a0[cnt], a0[cnt], ss); ss = fmaf(a1[cnt], a1[cnt], ss);
        }
        ss = warp_sum(ss);
        const float rstd = w ? rsqrtf(ss / hd + eps) : 1.f;
        cnt = 0;
        for (int p = lane; p < pairs; p += 32, ++cnt) {
            float t0 = a0[cnt], t1 = a1[cnt];
            if (w) { t0 = bf16_round(t0 * rstd) * __ldg(w + 2 * p); t1 = bf16_round(t1 * rstd) * __ldg(w + 2 * p + 1); }
            if (cosb) {                                   // t * cos + rotate_half(t) * sin, rotate_half: (x0, x1) -> (-x1, x0)
                const float c0 = __ldg(cosb + (int64_t)n * hd + 2 * p), c1 = __ldg(cosb + (int64_t)n * hd + 2 * p + 1);
                const float s0 = __ldg(sinb + (int64_t)n * hd + 2 * p), s1 = __ldg(sinb + (int64_t)n * hd + 2 * p + 1);
                const float r0 = t0 * c0 - t1 * s0, r1 = t1 * c1 + t0 * s1;
                t0 = r0; t1 = r1;
            }
            *reinterpret_cast<uint32_t*>(out + 2 * p) = pack_bf16x2(t0, t1);
        }
    }
}

// qkv: bf16 [B][N][3][H][hd] (the qkv Linear's output) -> q, k, v: bf16 [B][H][N][hd].  wq / wk: RMSNorm gains [hd] (NULL: no
// qk-norm); cos / sin: fp32 [N][hd] rotary tables (NULL: no rope).  hd even, <= 192.
DMVAE_API int dmvae_qk_norm_rope(const void* qkv, const float* wq, const float* wk, const float* cos_tab, const float* sin_tab, void* q,
                                 void* k, void* v, int64_t B, int N, int H, int hd, float eps, void* stream) {
    DMVAE_CHECK_ARG(qkv && q && k && v, "qk_norm_rope: null pointer");
    DMVAE_CHECK_ARG((wq == nullptr) == (wk == nullptr) && (cos_tab == nullptr) == (sin_tab == nullptr), "qk_norm_rope: gains / tables come in pairs");
    DMVAE_CHECK_ARG(B >= 0 && N > 0 && H > 0 && hd > 0 && hd % 2 == 0 && hd <= 192, "qk_norm_rope: bad shape (hd even, <= 192)");
    DMVAE_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 3) == 0, "qk_norm_rope: buffers must be 4-byte aligned");
    const int64_t rows = B * N * H;
    if (rows == 0) return DMVAE_OK;
    qk_norm_rope_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)qkv, wq, wk, cos_tab, sin_tab, (bf16*)q,
                                                                                     (bf16*)k, (bf16*)v, B, N, H, hd, eps);
    DMVAE_CHECK_LAUNCH("qk_norm_rope_kernel");
    return DMVAE_OK;
}
