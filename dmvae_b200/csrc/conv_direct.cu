// CUDA-core convolution kernels on channels-last bf16 activations.
//
// These cover the layers whose GEMM dimensions are too small or too ragged for the tcgen05 tile in
// conv_tc.cu (Cin=32 stem, Cout=3 head, Cin=3 encoder stem, stride-2 Downsample), and they are the on-device
// cross-check for the tensor-core path in tests.  Same math as nn.Conv2d under autocast: bf16 operands,
// fp32 accumulate, fp32 bias, one rounding to bf16 at the store (models/flux_ae.py:63,65,67,89,101,210,237).
//
//   y[b][oh][ow][co] = bias[co] + sum_{kh,kw,ci} x[b][oh*s - pt + kh][ow*s - pl + kw][ci] * w[tap][co][ci]
//
// Out-of-range taps read zero, which gives both the symmetric pad=1 of the 3x3 convs and Downsample's
// one-sided pad (0,1,0,1) (:91-95) with pt = pl = 0.
#include "common.cuh"

struct ConvGeom {
    int B, H, W, Cin, OH, OW, Cout, KH, KW, stride, pt, pl;
    int flags;          // 1 = ReLU on the output, 2 = `res` gates the output (res > 0 ? y : 0) instead of being added (see conv_tc.cu)
};

// ---- generic 64x64x32 smem-tiled implicit GEMM ---------------------------------------------------------------
#define CD_BM 64
#define CD_BN 64
#define CD_BK 32

template <bool VEC>
__global__ void __launch_bounds__(256) conv_direct_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                              const float* __restrict__ bias, const bf16* __restrict__ res,
                                                              bf16* __restrict__ y, ConvGeom g) {
    __shared__ __align__(16) float sA[CD_BK][CD_BM + 4];
    __shared__ __align__(16) float sB[CD_BK][CD_BN + 4];
    const int t = threadIdx.x;
    const int64_t M = (int64_t)g.B * g.OH * g.OW;
    const int64_t m0 = (int64_t)blockIdx.x * CD_BM;
    const int n0 = blockIdx.y * CD_BN;
    // loader mapping: one 8-element k-vector per thread for A and for B
    const int lr = t >> 2, lk = (t & 3) * 8;
    const int64_t lm = m0 + lr;
    int lb = 0, loh = 0, low = 0;
    const bool lm_ok = lm < M;
    if (lm_ok) { low = (int)(lm % g.OW); loh = (int)((lm / g.OW) % g.OH); lb = (int)(lm / ((int64_t)g.OW * g.OH)); }
    const int ln = n0 + lr;
    // compute mapping: 4 pixels x 4 couts per thread
    const int tm = (t & 15) * 4, tn = (t >> 4) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int taps = g.KH * g.KW;
    for (int tap = 0; tap < taps; ++tap) {
        const int kh = tap / g.KW, kw = tap % g.KW;
        const int ih = loh * g.stride - g.pt + kh, iw = low * g.stride - g.pl + kw;
        const bool in_ok = lm_ok && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
        const bf16* xrow = x + (((int64_t)lb * g.H + (in_ok ? ih : 0)) * g.W + (in_ok ? iw : 0)) * g.Cin;
        const bf16* wrow = w + ((int64_t)tap * g.Cout + (ln < g.Cout ? ln : 0)) * g.Cin;
        for (int k0 = 0; k0 < g.Cin; k0 += CD_BK) {
            float fa[8], fb[8];
            const int k = k0 + lk;
            if (VEC) {
                if (in_ok && k < g.Cin) unpack_bf16x8(*reinterpret_cast<const uint4*>(xrow + k), fa);
                else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) fa[q] = 0.f;
                }
                if (ln < g.Cout && k < g.Cin) unpack_bf16x8(*reinterpret_cast<const uint4*>(wrow + k), fb);
                else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) fb[q] = 0.f;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    fa[q] = (in_ok && k + q < g.Cin) ? __bfloat162float(xrow[k + q]) : 0.f;
                    fb[q] = (ln < g.Cout && k + q < g.Cin) ? __bfloat162float(wrow[k + q]) : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) { sA[lk + q][lr] = fa[q]; sB[lk + q][lr] = fb[q]; }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < CD_BK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&sA[kk][tm]);
                const float4 b = *reinterpret_cast<const float4*>(&sB[kk][tn]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + tm + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn + j;
            if (n >= g.Cout) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (g.flags & 1) v = fmaxf(v, 0.f);
            if (res) {
                const float fr = __bfloat162float(res[m * g.Cout + n]);
                v = (g.flags & 2) ? (fr > 0.f ? v : 0.f) : bf16_round(v) + fr;    // bf16 conv output, then bf16 add (:82)
            }
            y[m * g.Cout + n] = __float2bfloat16_rn(v);
        }
    }
}

// ---- Cout <= 4 head (conv_out 128->3): one thread per output pixel -------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(128) conv_small_cout_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                  const float* __restrict__ bias, bf16* __restrict__ y,
                                                                  ConvGeom g) {
    extern __shared__ float sw[];   // [tap][CO][Cin]
    const int taps = g.KH * g.KW;
    for (int i = threadIdx.x; i < taps * CO * g.Cin; i += blockDim.x) sw[i] = __bfloat162float(w[i]);
    __syncthreads();
    const int64_t M = (int64_t)g.B * g.OH * g.OW;
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int ow = (int)(m % g.OW), oh = (int)((m / g.OW) % g.OH), b = (int)(m / ((int64_t)g.OW * g.OH));
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    for (int tap = 0; tap < taps; ++tap) {
        const int ih = oh * g.stride - g.pt + tap / g.KW, iw = ow * g.stride - g.pl + tap % g.KW;
        if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
        const bf16* xp = x + (((int64_t)b * g.H + ih) * g.W + iw) * g.Cin;
        const float* wp = sw + tap * CO * g.Cin;
        for (int k = 0; k < g.Cin; k += 8) {
            float f[8];
            unpack_bf16x8(*reinterpret_cast<const uint4*>(xp + k), f);
#pragma unroll
            for (int c = 0; c < CO; ++c)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[c] += f[q] * wp[c * g.Cin + k + q];
        }
    }
#pragma unroll
    for (int c = 0; c < CO; ++c) y[m * CO + c] = __float2bfloat16_rn(acc[c] + (bias ? bias[c] : 0.f));
}


// ---- thin-input conv: Cin <= 4 (encoder stem 3->128, and the data-gradient of the 128->3 head) ---------------------
// Each thread produces 8 output channels for two horizontally adjacent pixels; the 27 x 8 weights it needs are read
// from shared memory once per pixel pair.  3x3, stride 1, pad 1.
template <int CI>
__global__ void __launch_bounds__(256) conv_thin_in_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                               const float* __restrict__ bias, bf16* __restrict__ y, ConvGeom g) {
    extern __shared__ float sw[];                 // [tap*CI + ci][Cout]  (co fastest -> conflict-free LDS.128)
    for (int i = threadIdx.x; i < 9 * CI * g.Cout; i += blockDim.x) {
        const int co = i % g.Cout, tc = i / g.Cout;           // tc = tap*CI + ci
        sw[i] = __bfloat162float(w[((int64_t)(tc / CI) * g.Cout + co) * CI + (tc % CI)]);
    }
    __syncthreads();
    const int vc = g.Cout / 8;
    const int col = threadIdx.x % vc;
    const int64_t pairs = (int64_t)g.B * g.H * (g.W / 2);
    const int ppb = blockDim.x / vc;
    float bs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bs[k] = bias ? bias[col * 8 + k] : 0.f;
    for (int64_t pr = (int64_t)blockIdx.x * ppb + threadIdx.x / vc; pr < pairs; pr += (int64_t)gridDim.x * ppb) {
        const int w0 = (int)(pr % (g.W / 2)) * 2;
        const int h = (int)((pr / (g.W / 2)) % g.H);
        const int64_t b = pr / ((int64_t)(g.W / 2) * g.H);
        float a0[8], a1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { a0[k] = bs[k]; a1[k] = bs[k]; }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = h + kh - 1;
            if (ih < 0 || ih >= g.H) continue;
            const bf16* xrow = x + ((b * g.H + ih) * g.W) * CI;
            // the two pixels share columns w0-1 .. w0+2
            float xv[4][CI];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int iw = w0 - 1 + j;
                const bool ok = iw >= 0 && iw < g.W;
#pragma unroll
                for (int c = 0; c < CI; ++c) xv[j][c] = ok ? __bfloat162float(xrow[(int64_t)iw * CI + c]) : 0.f;
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int c = 0; c < CI; ++c) {
                    const float* wp = sw + ((kh * 3 + kw) * CI + c) * g.Cout + col * 8;
                    const float4 wa = *reinterpret_cast<const float4*>(wp), wb = *reinterpret_cast<const float4*>(wp + 4);
                    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                    const float x0 = xv[kw][c], x1 = xv[kw + 1][c];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { a0[k] += x0 * wv[k]; a1[k] += x1 * wv[k]; }
                }
        }
        if (g.flags & 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { a0[k] = fmaxf(a0[k], 0.f); a1[k] = fmaxf(a1[k], 0.f); }
        }
        bf16* yp = y + (((b * g.H + h) * g.W) + w0) * g.Cout + col * 8;
        st_stream16(yp, pack_bf16x8(a0));
        st_stream16(yp + g.Cout, pack_bf16x8(a1));
    }
}

// ---- thin weight gradient: one operand has <= 4 channels (the 128->3 head: dy thin; the 3->128 stem: x thin) ------
// Threads own one channel of the wide operand and keep 9 x T accumulators; the thin operand's rows (with halo) sit
// in shared memory and are read by warp-wide broadcast.  3x3, stride 1, pad 1.
//   THIN_IS_DY = true : dw[t][wide][tap] += dy[q][t] * x[p][wide],  q = p - (tap offset)   (loop over x pixels p)
//   THIN_IS_DY = false: dw[wide][t][tap] += x[p][t] * dy[q][wide],  p = q + (tap offset)   (loop over dy pixels q)
template <int T, bool THIN_IS_DY>
__global__ void __launch_bounds__(256) conv_thin_wgrad_kernel(const bf16* __restrict__ wide, const bf16* __restrict__ thin,
                                                              float* __restrict__ dw, int B, int H, int W, int CW, int rows_per_block) {
    extern __shared__ float sm[];                 // thin rows [rows_per_block + 2][W][T], then the fan-in buffer
    const int blocks_per_img = (H + rows_per_block - 1) / rows_per_block;
    const int b = blockIdx.x / blocks_per_img;
    const int r0 = (blockIdx.x % blocks_per_img) * rows_per_block;
    const int r1 = min(r0 + rows_per_block, H);
    const int nrows = r1 - r0 + 2;
    for (int i = threadIdx.x; i < nrows * W * T; i += blockDim.x) {
        const int rr = r0 - 1 + i / (W * T);
        sm[i] = (rr >= 0 && rr < H) ? __bfloat162float(thin[((int64_t)b * H + rr) * W * T + i % (W * T)]) : 0.f;
    }
    __syncthreads();
    const int halves = blockDim.x / CW > 0 ? blockDim.x / CW : 1;      // pixel-range split when CW < blockDim
    const int c = threadIdx.x % CW, part = threadIdx.x / CW;
    const int sgn = THIN_IS_DY ? -1 : 1;
    if (part < halves) {
        for (int cw = c; cw < CW; cw += blockDim.x) {
            float acc[9][T];
#pragma unroll
            for (int tp = 0; tp < 9; ++tp)
#pragma unroll
                for (int t = 0; t < T; ++t) acc[tp][t] = 0.f;
            for (int r = r0; r < r1; ++r) {
                const bf16* wrow = wide + (((int64_t)b * H + r) * W) * CW + cw;
                for (int wq = part; wq < W; wq += halves) {
                    const float v = __bfloat162float(wrow[(int64_t)wq * CW]);
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const int tr = r + sgn * (kh - 1);                 // thin row
                        if (tr < 0 || tr >= H) continue;
                        const float* trow = sm + (tr - r0 + 1) * W * T;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const int tw = wq + sgn * (kw - 1);
                            if (tw < 0 || tw >= W) continue;
#pragma unroll
                            for (int t = 0; t < T; ++t) acc[kh * 3 + kw][t] += trow[tw * T + t] * v;
                        }
                    }
                }
            }
#pragma unroll
            for (int tp = 0; tp < 9; ++tp)
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const int64_t idx = THIN_IS_DY ? (((int64_t)t * CW + cw) * 9 + tp) : (((int64_t)cw * T + t) * 9 + tp);
                    atomicAdd(&dw[idx], acc[tp][t]);
                }
        }
    }
}

// ---- weight gradient of the 3-channel head (conv_out, Cout = 3): sliding 3x3 window of x in registers ----------
// Each thread owns one input channel and walks along an image row keeping the 3x3 neighbourhood of x for that channel in
// registers (3 coalesced loads per step); dy for the row block sits in shared memory as float4 and is read by warp
// broadcast.  27 FMAs + 3 loads + 1 LDS.128 per (pixel, channel).   dw[co][ci][kh][kw] += dy[q][co] * x[q + (kh-1, kw-1)][ci]
__global__ void __launch_bounds__(256) conv_head_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                              float* __restrict__ dw, int B, int H, int W, int Cin, int rpb) {
    extern __shared__ float4 s_dy[];              // [rows][W] {d0, d1, d2, 0}
    const int blocks_per_img = (H + rpb - 1) / rpb;
    const int b = blockIdx.x / blocks_per_img;
    const int r0 = (blockIdx.x % blocks_per_img) * rpb;
    const int r1 = min(r0 + rpb, H);
    for (int i = threadIdx.x; i < (r1 - r0) * W; i += blockDim.x) {
        const bf16* p = dy + (((int64_t)b * H + r0) * W + i) * 3;
        s_dy[i] = make_float4(__bfloat162float(p[0]), __bfloat162float(p[1]), __bfloat162float(p[2]), 0.f);
    }
    __syncthreads();
    const int parts = blockDim.x / Cin;           // column ranges handled in parallel (Cin <= blockDim.x)
    const int ci = threadIdx.x % Cin, part = threadIdx.x / Cin;
    float acc[3][3][3];
#pragma unroll
    for (int i = 0; i < 27; ++i) (&acc[0][0][0])[i] = 0.f;
    if (part < parts) {
        const int ws = (int)((int64_t)W * part / parts), we = (int)((int64_t)W * (part + 1) / parts);
        for (int r = r0; r < r1; ++r) {
            const bf16* rowp[3];
            bool rok[3];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int rr = r + kh - 1;
                rok[kh] = rr >= 0 && rr < H;
                rowp[kh] = x + (((int64_t)b * H + (rok[kh] ? rr : 0)) * W) * Cin + ci;
            }
            float win[3][2];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                win[kh][0] = (rok[kh] && ws - 1 >= 0) ? __bfloat162float(rowp[kh][(int64_t)(ws - 1) * Cin]) : 0.f;
                win[kh][1] = rok[kh] ? __bfloat162float(rowp[kh][(int64_t)ws * Cin]) : 0.f;
            }
            const float4* drow = s_dy + (r - r0) * W;
            float nx[3];                                   // column w+1, loaded one iteration ahead of its use
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
                nx[kh] = (rok[kh] && ws + 1 < W) ? __bfloat162float(rowp[kh][(int64_t)(ws + 1) * Cin]) : 0.f;
            for (int w = ws; w < we; ++w) {
                float nn[3];                               // prefetch column w+2
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
                    nn[kh] = (rok[kh] && w + 2 < W) ? __bfloat162float(rowp[kh][(int64_t)(w + 2) * Cin]) : 0.f;
                const float4 d = drow[w];
                const float dv[3] = {d.x, d.y, d.z};
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const float xs[3] = {win[kh][0], win[kh][1], nx[kh]};
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                        for (int co = 0; co < 3; ++co) acc[kh][kw][co] = fmaf(dv[co], xs[kw], acc[kh][kw][co]);
                    win[kh][0] = win[kh][1];
                    win[kh][1] = nx[kh];
                    nx[kh] = nn[kh];
                }
            }
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int co = 0; co < 3; ++co) atomicAdd(&dw[((int64_t)co * Cin + ci) * 9 + kh * 3 + kw], acc[kh][kw][co]);
    }
}

static int check_geom(const ConvGeom& g, const char* who) {
    if (g.B < 0 || g.H <= 0 || g.W <= 0 || g.Cin <= 0 || g.Cout <= 0 || g.KH <= 0 || g.KW <= 0 || g.stride <= 0)
        return dmvae_set_error(DMVAE_EINVAL, "%s: bad geometry", who);
    if (g.OH <= 0 || g.OW <= 0) return dmvae_set_error(DMVAE_EINVAL, "%s: empty output", who);
    return DMVAE_OK;
}

DMVAE_API int dmvae_conv_direct_fwd(const void* x, const void* w_packed, const float* bias, const void* residual,
                                    void* y, int B, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW,
                                    int stride, int pad_top, int pad_left, int flags, void* stream) {
    DMVAE_CHECK_ARG(x && w_packed && y, "conv_direct_fwd: null pointer");
    DMVAE_CHECK_ARG((flags & ~3) == 0 && (!(flags & 2) || residual), "conv_direct_fwd: bad flags %d", flags);
    ConvGeom g = {B, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad_top, pad_left, flags};
    int rc = check_geom(g, "conv_direct_fwd");
    if (rc) return rc;
    if (B == 0) return DMVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = (int64_t)B * OH * OW;
    const bool vec = (Cin % 8 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)w_packed & 15) == 0);
    const size_t small_smem = (size_t)KH * KW * Cout * Cin * sizeof(float);
    if (Cout <= 4 && vec && !residual && !flags && small_smem <= 96 * 1024) {
        const unsigned grid = (unsigned)ceil_div64(M, 128);
#define SMALL(CO)                                                                                                    \
    {                                                                                                                \
        cudaFuncSetAttribute(conv_small_cout_fwd_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
        conv_small_cout_fwd_kernel<CO><<<grid, 128, small_smem, st>>>((const bf16*)x, (const bf16*)w_packed, bias, (bf16*)y, g); \
    }
        switch (Cout) { case 1: SMALL(1) break; case 2: SMALL(2) break; case 3: SMALL(3) break; default: SMALL(4) break; }
#undef SMALL
        DMVAE_CHECK_LAUNCH("conv_small_cout_fwd_kernel");
        return DMVAE_OK;
    }
    if (Cin == 3 && KH == 3 && KW == 3 && stride == 1 && pad_top == 1 && pad_left == 1 && OH == H && OW == W && !residual &&
        Cout % 8 == 0 && Cout <= 1024 && 256 % (Cout / 8) == 0 && W % 2 == 0 && (((uintptr_t)y & 15) == 0)) {
        const size_t smem = (size_t)27 * Cout * sizeof(float);
        cudaFuncSetAttribute(conv_thin_in_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        const int ppb = 256 / (Cout / 8);
        int64_t blocks = ceil_div64((int64_t)B * H * (W / 2), ppb);
        if (blocks > 148 * 16) blocks = 148 * 16;
        conv_thin_in_fwd_kernel<3><<<(unsigned)blocks, 256, smem, st>>>((const bf16*)x, (const bf16*)w_packed, bias, (bf16*)y, g);
        DMVAE_CHECK_LAUNCH("conv_thin_in_fwd_kernel");
        return DMVAE_OK;
    }
    dim3 grid((unsigned)ceil_div64(M, CD_BM), (unsigned)((Cout + CD_BN - 1) / CD_BN));
    if (vec) conv_direct_fwd_kernel<true><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)w_packed, bias, (const bf16*)residual, (bf16*)y, g);
    else     conv_direct_fwd_kernel<false><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)w_packed, bias, (const bf16*)residual, (bf16*)y, g);
    DMVAE_CHECK_LAUNCH("conv_direct_fwd_kernel");
    return DMVAE_OK;
}

// ---- data gradient for strided convs (gather form; stride-1 dgrad is a forward conv with flipped weights) ----
//   dx[b][ih][iw][ci] = sum_{kh,kw : (ih+pt-kh) % s == 0 ...} sum_co dy[b][oh][ow][co] * w[tap][co][ci]
__global__ void __launch_bounds__(256) conv_dgrad_strided_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ w,
                                                                 bf16* __restrict__ dx, ConvGeom g) {
    const int64_t n = (int64_t)g.B * g.H * g.W * g.Cin;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % g.Cin);
        int64_t p = i / g.Cin;
        const int iw = (int)(p % g.W); p /= g.W;
        const int ih = (int)(p % g.H);
        const int b = (int)(p / g.H);
        float acc = 0.f;
        for (int kh = 0; kh < g.KH; ++kh) {
            const int th = ih + g.pt - kh;
            if (th < 0 || th % g.stride) continue;
            const int oh = th / g.stride;
            if (oh >= g.OH) continue;
            for (int kw = 0; kw < g.KW; ++kw) {
                const int tw = iw + g.pl - kw;
                if (tw < 0 || tw % g.stride) continue;
                const int ow = tw / g.stride;
                if (ow >= g.OW) continue;
                const bf16* dyp = dy + (((int64_t)b * g.OH + oh) * g.OW + ow) * g.Cout;
                const bf16* wp = w + ((int64_t)(kh * g.KW + kw) * g.Cout) * g.Cin + ci;
                for (int co = 0; co < g.Cout; ++co) acc += __bfloat162float(dyp[co]) * __bfloat162float(wp[(int64_t)co * g.Cin]);
            }
        }
        dx[i] = __float2bfloat16_rn(acc);
    }
}

DMVAE_API int dmvae_conv_direct_dgrad_strided(const void* dy, const void* w_packed, void* dx, int B, int H, int W,
                                              int Cin, int OH, int OW, int Cout, int KH, int KW, int stride,
                                              int pad_top, int pad_left, void* stream) {
    DMVAE_CHECK_ARG(dy && w_packed && dx, "conv_direct_dgrad_strided: null pointer");
    ConvGeom g = {B, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad_top, pad_left};
    int rc = check_geom(g, "conv_direct_dgrad_strided");
    if (rc) return rc;
    if (B == 0) return DMVAE_OK;
    const int64_t n = (int64_t)B * H * W * Cin;
    int64_t grid = ceil_div64(n, 256);
    if (grid > 148 * 32) grid = 148 * 32;
    conv_dgrad_strided_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)w_packed, (bf16*)dx, g);
    DMVAE_CHECK_LAUNCH("conv_dgrad_strided_kernel");
    return DMVAE_OK;
}

// ---- weight gradient: dw[co][ci][tap] += sum_pixels dy[p][co] * x[p (+) tap][ci]   (fp32, atomically accumulated) ----
template <bool VEC>
__global__ void __launch_bounds__(256) conv_direct_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                                float* __restrict__ dw, ConvGeom g, int64_t m_per_split) {
    __shared__ __align__(16) float sD[CD_BK][CD_BM + 4];   // [pixel][co]
    __shared__ __align__(16) float sX[CD_BK][CD_BN + 4];   // [pixel][ci]
    const int t = threadIdx.x;
    const int taps = g.KH * g.KW;
    const int co0 = blockIdx.x * CD_BM;
    const int ci_tiles = (g.Cin + CD_BN - 1) / CD_BN;
    const int ci0 = (blockIdx.y % ci_tiles) * CD_BN;
    const int tap = blockIdx.y / ci_tiles;
    const int kh = tap / g.KW, kw = tap % g.KW;
    const int64_t M = (int64_t)g.B * g.OH * g.OW;
    const int64_t ms = (int64_t)blockIdx.z * m_per_split;
    const int64_t me = (ms + m_per_split < M) ? ms + m_per_split : M;
    const int lp = t >> 3, lc = (t & 7) * 8;          // loader: pixel row, channel vector
    const int tm = (t & 15) * 4, tn = (t >> 4) * 4;   // compute: 4 co x 4 ci
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t mb = ms; mb < me; mb += CD_BK) {
        const int64_t m = mb + lp;
        float fd[8], fx[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { fd[q] = 0.f; fx[q] = 0.f; }
        if (m < me) {
            const int ow = (int)(m % g.OW), oh = (int)((m / g.OW) % g.OH), b = (int)(m / ((int64_t)g.OW * g.OH));
            const bf16* dp = dy + m * g.Cout + co0 + lc;
            const int ih = oh * g.stride - g.pt + kh, iw = ow * g.stride - g.pl + kw;
            const bool in_ok = ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
            const bf16* xp = x + (((int64_t)b * g.H + (in_ok ? ih : 0)) * g.W + (in_ok ? iw : 0)) * g.Cin + ci0 + lc;
            if (VEC) {
                if (co0 + lc < g.Cout) unpack_bf16x8(*reinterpret_cast<const uint4*>(dp), fd);
                if (in_ok && ci0 + lc < g.Cin) unpack_bf16x8(*reinterpret_cast<const uint4*>(xp), fx);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (co0 + lc + q < g.Cout) fd[q] = __bfloat162float(dp[q]);
                    if (in_ok && ci0 + lc + q < g.Cin) fx[q] = __bfloat162float(xp[q]);
                }
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&sD[lp][lc]) = make_float4(fd[0], fd[1], fd[2], fd[3]);
        *reinterpret_cast<float4*>(&sD[lp][lc + 4]) = make_float4(fd[4], fd[5], fd[6], fd[7]);
        *reinterpret_cast<float4*>(&sX[lp][lc]) = make_float4(fx[0], fx[1], fx[2], fx[3]);
        *reinterpret_cast<float4*>(&sX[lp][lc + 4]) = make_float4(fx[4], fx[5], fx[6], fx[7]);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CD_BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&sD[kk][tm]);
            const float4 b = *reinterpret_cast<const float4*>(&sX[kk][tn]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + tm + i;
        if (co >= g.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci0 + tn + j;
            if (ci >= g.Cin) continue;
            atomicAdd(&dw[((int64_t)co * g.Cin + ci) * taps + tap], acc[i][j]);
        }
    }
}

DMVAE_API int dmvae_conv_direct_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int OH,
                                      int OW, int Cout, int KH, int KW, int stride, int pad_top, int pad_left,
                                      void* stream) {
    DMVAE_CHECK_ARG(x && dy && dw, "conv_direct_wgrad: null pointer");
    ConvGeom g = {B, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad_top, pad_left};
    int rc = check_geom(g, "conv_direct_wgrad");
    if (rc) return rc;
    if (B == 0) return DMVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = (int64_t)B * OH * OW;
    const int taps = KH * KW;
    const bool same3 = KH == 3 && KW == 3 && stride == 1 && pad_top == 1 && pad_left == 1 && OH == H && OW == W;
    if (same3 && (Cout == 3 || Cin == 3) && (Cout == 3 ? Cin : Cout) <= 256 && W <= 1024) {
        const int rpb = 4;
        const size_t smem = (size_t)(rpb + 2) * W * 3 * sizeof(float);
        const int blocks = B * ((H + rpb - 1) / rpb);
        if (Cout == 3 && 256 % Cin == 0 && (size_t)rpb * W * sizeof(float4) <= 96 * 1024) {
            const size_t smem4 = (size_t)rpb * W * sizeof(float4);
            cudaFuncSetAttribute(conv_head_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            conv_head_wgrad_kernel<<<blocks, 256, smem4, st>>>((const bf16*)x, (const bf16*)dy, dw, B, H, W, Cin, rpb);
            DMVAE_CHECK_LAUNCH("conv_head_wgrad_kernel");
            return DMVAE_OK;
        }
        if (Cout == 3) {
            cudaFuncSetAttribute(conv_thin_wgrad_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            conv_thin_wgrad_kernel<3, true><<<blocks, 256, smem, st>>>((const bf16*)x, (const bf16*)dy, dw, B, H, W, Cin, rpb);
        } else {
            cudaFuncSetAttribute(conv_thin_wgrad_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            conv_thin_wgrad_kernel<3, false><<<blocks, 256, smem, st>>>((const bf16*)dy, (const bf16*)x, dw, B, H, W, Cout, rpb);
        }
        DMVAE_CHECK_LAUNCH("conv_thin_wgrad_kernel");
        return DMVAE_OK;
    }
    const bool vec = (Cin % 8 == 0) && (Cout % 8 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)dy & 15) == 0);
    const int co_tiles = (Cout + CD_BM - 1) / CD_BM, ci_tiles = (Cin + CD_BN - 1) / CD_BN;
    const int64_t base_ctas = (int64_t)co_tiles * ci_tiles * taps;
    int64_t splits = (148 * 4 + base_ctas - 1) / base_ctas;
    const int64_t max_splits = ceil_div64(M, 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    int64_t mps = ceil_div64(ceil_div64(M, splits), CD_BK) * CD_BK;
    splits = ceil_div64(M, mps);
    dim3 grid((unsigned)co_tiles, (unsigned)(ci_tiles * taps), (unsigned)splits);
    if (vec) conv_direct_wgrad_kernel<true><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, g, mps);
    else     conv_direct_wgrad_kernel<false><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, g, mps);
    DMVAE_CHECK_LAUNCH("conv_direct_wgrad_kernel");
    return DMVAE_OK;
}

// ---- bias gradient: db[co] += sum_pixels dy[p][co] -----------------------------------------------------------------
__global__ void __launch_bounds__(256) bias_grad_kernel(const bf16* __restrict__ dy, float* __restrict__ db, int64_t M,
                                                        int C, int64_t m_per_block) {
    const int64_t ms = (int64_t)blockIdx.x * m_per_block;
    const int64_t me = (ms + m_per_block < M) ? ms + m_per_block : M;
    // threads tile (rows x channels): consecutive threads read consecutive channels
    const int cols = C < 256 ? C : 256;
    const int rows = 256 / cols;
    const int r = threadIdx.x / cols, c0 = threadIdx.x % cols;
    if (r >= rows) return;
    for (int c = c0; c < C; c += cols) {
        float a = 0.f;
        for (int64_t m = ms + r; m < me; m += rows) a += __bfloat162float(dy[m * C + c]);
        atomicAdd(&db[c], a);
    }
}

// vectorised variant: threads own a fixed 8-channel column, 4 x 16 B loads in flight, shared-memory fan-in
__global__ void __launch_bounds__(256) bias_grad_vec_kernel(const bf16* __restrict__ dy, float* __restrict__ db, int64_t M,
                                                            int C, int vc, int rows, int64_t m_per_block) {
    extern __shared__ float s_acc[];   // [C]
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_acc[c] = 0.f;
    __syncthreads();
    const int col = threadIdx.x % vc, row = threadIdx.x / vc;
    const int64_t ms = (int64_t)blockIdx.x * m_per_block;
    const int64_t me = (ms + m_per_block < M) ? ms + m_per_block : M;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bf16* base = dy + col * 8;
    constexpr int U = 4;
    for (int64_t m = ms + row; m < me; m += (int64_t)rows * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t mm = m + (int64_t)u * rows;
            v[u] = mm < me ? ld_stream16(base + mm * C) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float f[8];
            unpack_bf16x8(v[u], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += f[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[col * 8 + k], a[k]);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&db[c], s_acc[c]);
}

DMVAE_API int dmvae_bias_grad(const void* dy, float* dbias, int64_t M, int C, void* stream) {
    DMVAE_CHECK_ARG(dy && dbias, "bias_grad: null pointer");
    DMVAE_CHECK_ARG(M >= 0 && C > 0, "bias_grad: bad shape");
    if (M == 0) return DMVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0 && ((uintptr_t)dy & 15) == 0) {
        const int vc = C / 8, rows = 256 / vc;
        int64_t blocks = 148 * 8;
        int64_t mpb = ceil_div64(M, blocks);
        if (mpb < (int64_t)rows * 8) mpb = (int64_t)rows * 8;
        blocks = ceil_div64(M, mpb);
        bias_grad_vec_kernel<<<(unsigned)blocks, 256, C * sizeof(float), st>>>((const bf16*)dy, dbias, M, C, vc, rows, mpb);
        DMVAE_CHECK_LAUNCH("bias_grad_vec_kernel");
        return DMVAE_OK;
    }
    int64_t blocks = 148 * 4;
    int64_t mpb = ceil_div64(M, blocks);
    if (mpb < 64) mpb = 64;
    blocks = ceil_div64(M, mpb);
    bias_grad_kernel<<<(unsigned)blocks, 256, 0, st>>>((const bf16*)dy, dbias, M, C, mpb);
    DMVAE_CHECK_LAUNCH("bias_grad_kernel");
    return DMVAE_OK;
}
