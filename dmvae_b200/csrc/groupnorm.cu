// GroupNorm(32 groups, eps) + swish on channels-last bf16 activations, forward and backward.
// Reference: models/flux_ae.py:21-22 (swish), :30,62,64,157,236 (GroupNorm(32, C, eps=1e-6)),
// used as swish(norm(x)) in ResnetBlock.forward :69-77 and Decoder/Encoder tails :177-179,266-268;
// AttnBlock uses the norm without swish (:38).  Under autocast the reference runs GroupNorm and swish
// in fp32 and the following conv rounds its input to bf16: these kernels do the same (fp32 math,
// fp64 statistics, one bf16 rounding at the store).
//
// Layout: x[b][pixel][c], c fastest.  HBM-bound: stats = 2 B/elem read; apply = 2 B read + 2 B write.
#include "common.cuh"

#define GN_GROUPS 32
#define GN_THREADS 256
#define GN_UNROLL 4

struct GnGeom {
    int vc;        // 16-byte vectors per pixel  (C / 8)
    int rows;      // pixels handled per block pass (GN_THREADS / vc)
};

static int gn_geom(int C, GnGeom* g, const char* who) {
    if (C % GN_GROUPS != 0 || C % 8 != 0 || C > 2048 || (GN_THREADS % (C / 8)) != 0)
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "%s: unsupported channel count %d (need C%%32==0, C/8 | 256)", who, C);
    g->vc = C / 8;
    g->rows = GN_THREADS / g->vc;
    return DMVAE_OK;
}

#include <stdlib.h>
static int gn_ctas_per_sm() {          // tuning knob (DMVAE_GN_CTAS_PER_SM): grid = 148 x this many CTAs.  4 = two exactly full waves of the
    static int v = 0;                  // backward kernels (2 resident CTAs per SM); measured best of {3,4,6,8,16} for stats, apply and backward
    if (!v) { const char* e = getenv("DMVAE_GN_CTAS_PER_SM"); v = e ? atoi(e) : 4; if (v < 1) v = 4; }
    return v;
}
static int gn_min_passes() {           // minimum row-passes of work per CTA (amortises the per-CTA coefficient prologue)
    static int v = 0;
    if (!v) { const char* e = getenv("DMVAE_GN_MIN_PASSES"); v = e ? atoi(e) : 16; if (v < 1) v = 16; }
    return v;
}
static void gn_grid(int64_t B, int64_t HW, int rows, dim3* grid, int64_t* ppb) {
    // aim for several waves of CTAs, but at least 8 row-passes of work per CTA
    int64_t chunks = (148 * gn_ctas_per_sm() + B - 1) / B;
    int64_t max_chunks = ceil_div64(HW, (int64_t)rows * gn_min_passes());
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    *ppb = ceil_div64(HW, chunks);
    chunks = ceil_div64(HW, *ppb);
    *grid = dim3((unsigned)chunks, (unsigned)B);
}

// stats[b][g] = {sum, sumsq} (fp64, caller zero-initialised)
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const bf16* __restrict__ x, double* __restrict__ stats,
                                                              int64_t HW, int C, int vc, int rows, int64_t ppb) {
    __shared__ float s_part[2 * GN_THREADS * 8];      // per-thread {sum[8], sumsq[8]}
    const int b = blockIdx.y;
    const int col = threadIdx.x % vc, row = threadIdx.x / vc;
    const int cpg = C / GN_GROUPS;
    const int64_t p0 = (int64_t)blockIdx.x * ppb;
    const int64_t p1 = (p0 + ppb < HW) ? p0 + ppb : HW;
    float s[8], ss[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] = 0.f; ss[k] = 0.f; }
    const bf16* base = x + (int64_t)b * HW * C + col * 8;
    for (int64_t p = p0 + row; p < p1; p += (int64_t)rows * GN_UNROLL) {
        uint4 v[GN_UNROLL];
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {           // all loads first: GN_UNROLL x 16 B in flight per thread
            const int64_t pp = p + (int64_t)u * rows;
            v[u] = pp < p1 ? ld_stream16(base + pp * C) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
            float f[8];
            unpack_bf16x8(v[u], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { s[k] += f[k]; ss[k] += f[k] * f[k]; }
        }
    }
    // fixed-order fan-in (no floating-point atomics inside the CTA): the forward pass is run-to-run reproducible up to
    // the order of the per-CTA fp64 atomics below, which is far under fp32 resolution.
#pragma unroll
    for (int k = 0; k < 8; ++k) { s_part[threadIdx.x * 8 + k] = s[k]; s_part[(GN_THREADS + threadIdx.x) * 8 + k] = ss[k]; }
    __syncthreads();
    if (threadIdx.x < GN_GROUPS * 2) {
        const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
        const float* part = s_part + which * GN_THREADS * 8;
        float acc = 0.f;
        for (int r = 0; r < rows; ++r)
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) acc += part[(r * vc + (c >> 3)) * 8 + (c & 7)];
        atomicAdd(&stats[(int64_t)b * GN_GROUPS * 2 + threadIdx.x], (double)acc);
    }
}

__device__ __forceinline__ void gn_mean_rstd(const double* __restrict__ stats, int b, int g, double n, float eps,
                                             float& mean, float& rstd) {
    const double s = stats[((int64_t)b * GN_GROUPS + g) * 2], q = stats[((int64_t)b * GN_GROUPS + g) * 2 + 1];
    const double m = s / n;
    double var = q / n - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// y = [swish]( (x - mean) * rstd * gamma + beta )  -> bf16
template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const bf16* __restrict__ x, const double* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              bf16* __restrict__ y, int64_t HW, int C, int vc, int rows,
                                                              int64_t ppb, float eps) {
    extern __shared__ float s_ab[];   // scale[C], shift[C]
    const int b = blockIdx.y;
    const int cpg = C / GN_GROUPS;
    const double n = (double)HW * cpg;
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
        float mean, rstd;
        gn_mean_rstd(stats, b, c / cpg, n, eps, mean, rstd);
        const float sc = rstd * gamma[c];
        s_ab[c] = sc;
        s_ab[C + c] = beta[c] - mean * sc;
    }
    __syncthreads();
    const int col = threadIdx.x % vc, row = threadIdx.x / vc;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = s_ab[col * 8 + k]; sh[k] = s_ab[C + col * 8 + k]; }
    const int64_t p0 = (int64_t)blockIdx.x * ppb;
    const int64_t p1 = (p0 + ppb < HW) ? p0 + ppb : HW;
    const int64_t off = (int64_t)b * HW * C + col * 8;
    for (int64_t p = p0 + row; p < p1; p += (int64_t)rows * GN_UNROLL) {
        uint4 v[GN_UNROLL];
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
            const int64_t pp = p + (int64_t)u * rows;
            v[u] = pp < p1 ? ld_stream16(x + off + pp * C) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
            const int64_t pp = p + (int64_t)u * rows;
            if (pp >= p1) break;
            float f[8];
            unpack_bf16x8(v[u], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float t = f[k] * sc[k] + sh[k];
                f[k] = SILU ? t * sigmoidf_fast(t) : t;
            }
            st_stream16(y + off + pp * C, pack_bf16x8(f));
        }
    }
}

// Backward.  With y = a_c x + b_c (a = rstd*gamma, b = beta - mean*a) and xhat = r x + s (r = rstd, s = -mean*rstd):
//   dy = da * swish'(y),  swish'(y) = sg + y sg (1 - sg)
// pass 1 (reduce):  A_c = sum dy, B_c = sum dy*xhat ;  dbeta += A, dgamma += B, gsum[b][g] += {sum gamma A, sum gamma B}
// pass 2 (apply):   dx = rstd*(gamma*dy - (S1 + xhat*S2)/n) [+ dres] = a*dy + k0 + k1*x [+ dres]
//                   k1 = -rstd^2 * S2/n,  k0 = -rstd*(S1/n + s*S2/n)
// Per-channel coefficients are folded once per CTA so the per-element work is ~16 instructions and the register
// footprint allows 3 CTAs per SM.
template <bool SILU>
__device__ __forceinline__ float gn_dy(float da, float yv) {
    if (!SILU) return da;
    const float sg = sigmoidf_tanh(yv);
    return da * fmaf(yv * sg, 1.f - sg, sg);
}

template <bool SILU, int U>
__global__ void __launch_bounds__(GN_THREADS, U <= 2 ? 3 : 2) gn_bwd_reduce_kernel(
    const bf16* __restrict__ da, const bf16* __restrict__ x, const double* __restrict__ stats,
    const float* __restrict__ gamma, const float* __restrict__ beta, double* __restrict__ gsum,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t HW, int C, int vc, int rows, int64_t ppb, float eps) {
    extern __shared__ float s_mem[];   // a[C] b[C] r[C] s[C] accA[C] accB[C]
    float* s_a = s_mem; float* s_b = s_mem + C; float* s_r = s_mem + 2 * C; float* s_s = s_mem + 3 * C;
    float* s_A = s_mem + 4 * C; float* s_B = s_mem + 5 * C;
    const int b = blockIdx.y;
    const int cpg = C / GN_GROUPS;
    const double n = (double)HW * cpg;
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
        float mean, rstd;
        gn_mean_rstd(stats, b, c / cpg, n, eps, mean, rstd);
        const float a = rstd * gamma[c];
        s_a[c] = a; s_b[c] = beta[c] - mean * a; s_r[c] = rstd; s_s[c] = -mean * rstd;
        s_A[c] = 0.f; s_B[c] = 0.f;
    }
    __syncthreads();
    const int col = threadIdx.x % vc, row = threadIdx.x / vc;
    float ca[8], cb[8], cr[8], cs[8], A[8], Bv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = col * 8 + k;
        ca[k] = s_a[c]; cb[k] = s_b[c]; cr[k] = s_r[c]; cs[k] = s_s[c]; A[k] = 0.f; Bv[k] = 0.f;
    }
    const int64_t p0 = (int64_t)blockIdx.x * ppb;
    const int64_t p1 = (p0 + ppb < HW) ? p0 + ppb : HW;
    const int64_t off = (int64_t)b * HW * C + col * 8;
    for (int64_t p = p0 + row; p < p1; p += (int64_t)rows * U) {
        uint4 vx[U], vd[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t pp = p + (int64_t)u * rows;
            const bool ok = pp < p1;
            vx[u] = ok ? ld_stream16(x + off + pp * C) : make_uint4(0, 0, 0, 0);
            vd[u] = ok ? ld_stream16(da + off + pp * C) : make_uint4(0, 0, 0, 0);     // da = 0 contributes nothing
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float fx[8], fd[8];
            unpack_bf16x8(vx[u], fx);
            unpack_bf16x8(vd[u], fd);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float dy = gn_dy<SILU>(fd[k], fmaf(fx[k], ca[k], cb[k]));
                A[k] += dy;
                Bv[k] = fmaf(dy, fmaf(fx[k], cr[k], cs[k]), Bv[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s_A[col * 8 + k], A[k]); atomicAdd(&s_B[col * 8 + k], Bv[k]); }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
        atomicAdd(&dbeta[c], s_A[c]);
        atomicAdd(&dgamma[c], s_B[c]);
    }
    if (threadIdx.x < GN_GROUPS) {
        const int g = threadIdx.x;
        float a = 0.f, q = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += gamma[c] * s_A[c]; q += gamma[c] * s_B[c]; }
        atomicAdd(&gsum[((int64_t)b * GN_GROUPS + g) * 2], (double)a);
        atomicAdd(&gsum[((int64_t)b * GN_GROUPS + g) * 2 + 1], (double)q);
    }
}

template <bool SILU, int U>
__global__ void __launch_bounds__(GN_THREADS, U <= 2 ? 3 : 2) gn_bwd_apply_kernel(
    const bf16* __restrict__ da, const bf16* __restrict__ x, const double* __restrict__ stats,
    const float* __restrict__ gamma, const float* __restrict__ beta, const double* __restrict__ gsum,
    const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ colsum, int64_t HW, int C, int vc, int rows,
    int64_t ppb, float eps, int reverse) {
    extern __shared__ float s_mem[];   // a[C] b[C] k0[C] k1[C]
    float* s_a = s_mem; float* s_b = s_mem + C; float* s_k0 = s_mem + 2 * C; float* s_k1 = s_mem + 3 * C;
    const int b = blockIdx.y;
    const int cpg = C / GN_GROUPS;
    const double n = (double)HW * cpg;
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
        const int g = c / cpg;
        float mean, rstd;
        gn_mean_rstd(stats, b, g, n, eps, mean, rstd);
        const float m1 = (float)(gsum[((int64_t)b * GN_GROUPS + g) * 2] / n);
        const float m2 = (float)(gsum[((int64_t)b * GN_GROUPS + g) * 2 + 1] / n);
        const float a = rstd * gamma[c];
        s_a[c] = a; s_b[c] = beta[c] - mean * a;
        s_k1[c] = -rstd * rstd * m2;
        s_k0[c] = -rstd * (m1 - mean * rstd * m2);
    }
    __syncthreads();
    const int col = threadIdx.x % vc, row = threadIdx.x / vc;
    float ca[8], cb[8], k0[8], k1[8], cs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = col * 8 + k;
        ca[k] = s_a[c]; cb[k] = s_b[c]; k0[k] = s_k0[c]; k1[k] = s_k1[c]; cs[k] = 0.f;
    }
    const int64_t p0 = (int64_t)blockIdx.x * ppb;
    const int64_t p1 = (p0 + ppb < HW) ? p0 + ppb : HW;
    const int64_t off = (int64_t)b * HW * C + col * 8;
    // This pass walks its pixel range BACKWARDS: the reduce pass (same grid, same ranges) has just streamed da and x front to back,
    // so whatever the L2 still holds is the tail of every CTA's range -- read first here, before this pass's own traffic evicts it.
    // Measured on B200 (same-box A/B, DMVAE_GN_BWD_REVERSE=0/1): -0.05 ms per step only -- little of a 0.3-0.5 GB stream survives
    // in the 126 MB L2 next to the write-back traffic -- kept because it is free.
    const int64_t plast = p0 + p1 - 1;
    for (int64_t p = p0 + row; p < p1; p += (int64_t)rows * U) {
        uint4 vx[U], vd[U], vr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t pl = p + (int64_t)u * rows;
            const bool ok = pl < p1;
            const int64_t pp = reverse ? plast - pl : pl;
            vx[u] = ok ? ld_stream16(x + off + pp * C) : make_uint4(0, 0, 0, 0);
            vd[u] = ok ? ld_stream16(da + off + pp * C) : make_uint4(0, 0, 0, 0);
            vr[u] = (ok && dres) ? ld_stream16(dres + off + pp * C) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t pl = p + (int64_t)u * rows;
            if (pl >= p1) break;
            const int64_t pp = reverse ? plast - pl : pl;
            float fx[8], fd[8], fr[8];
            unpack_bf16x8(vx[u], fx);
            unpack_bf16x8(vd[u], fd);
            unpack_bf16x8(vr[u], fr);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float dy = gn_dy<SILU>(fd[k], fmaf(fx[k], ca[k], cb[k]));
                fx[k] = fmaf(ca[k], dy, fmaf(k1[k], fx[k], k0[k])) + fr[k];
            }
            const uint4 packed = pack_bf16x8(fx);
            st_stream16(dx + off + pp * C, packed);
            if (colsum) {                       // column sums of the bf16 values written: the bias gradient of the conv that made x
                float fo[8];
                unpack_bf16x8(packed, fo);
#pragma unroll
                for (int k = 0; k < 8; ++k) cs[k] += fo[k];
            }
        }
    }
    if (colsum) {
        __syncthreads();                        // coefficient arrays are dead: reuse s_a as the fan-in buffer
        for (int c = threadIdx.x; c < C; c += GN_THREADS) s_a[c] = 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s_a[col * 8 + k], cs[k]);
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += GN_THREADS) atomicAdd(&colsum[c], s_a[c]);
    }
}

DMVAE_API int dmvae_gn_stats(const void* x, double* stats, int64_t B, int64_t HW, int C, void* stream) {
    DMVAE_CHECK_ARG(x && stats, "gn_stats: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && HW >= 0, "gn_stats: negative size");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0, "gn_stats: x must be 16-byte aligned");
    GnGeom g;
    int rc = gn_geom(C, &g, "gn_stats");
    if (rc) return rc;
    if (B * HW == 0) return DMVAE_OK;
    dim3 grid; int64_t ppb;
    gn_grid(B, HW, g.rows, &grid, &ppb);
    gn_stats_kernel<<<grid, GN_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, stats, HW, C, g.vc, g.rows, ppb);
    DMVAE_CHECK_LAUNCH("gn_stats_kernel");
    return DMVAE_OK;
}

DMVAE_API int dmvae_gn_apply(const void* x, const double* stats, const float* gamma, const float* beta, void* y,
                             int64_t B, int64_t HW, int C, float eps, int silu, void* stream) {
    DMVAE_CHECK_ARG(x && stats && gamma && beta && y, "gn_apply: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && HW >= 0, "gn_apply: negative size");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "gn_apply: buffers must be 16-byte aligned");
    GnGeom g;
    int rc = gn_geom(C, &g, "gn_apply");
    if (rc) return rc;
    if (B * HW == 0) return DMVAE_OK;
    dim3 grid; int64_t ppb;
    gn_grid(B, HW, g.rows, &grid, &ppb);
    const size_t smem = (size_t)2 * C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (silu) gn_apply_kernel<true><<<grid, GN_THREADS, smem, st>>>((const bf16*)x, stats, gamma, beta, (bf16*)y, HW, C, g.vc, g.rows, ppb, eps);
    else      gn_apply_kernel<false><<<grid, GN_THREADS, smem, st>>>((const bf16*)x, stats, gamma, beta, (bf16*)y, HW, C, g.vc, g.rows, ppb, eps);
    DMVAE_CHECK_LAUNCH("gn_apply_kernel");
    return DMVAE_OK;
}

// gsum (fp64 [B][32][2]), dgamma, dbeta (fp32 [C]) are accumulated into: caller zero-initialises.
DMVAE_API int dmvae_gn_bwd(const void* da, const void* x, const double* stats, const float* gamma, const float* beta,
                           double* gsum, float* dgamma, float* dbeta, const void* dres, void* dx, float* dx_colsum,
                           int64_t B, int64_t HW, int C, float eps, int silu, void* stream) {
    DMVAE_CHECK_ARG(da && x && stats && gamma && beta && gsum && dgamma && dbeta && dx, "gn_bwd: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && HW >= 0, "gn_bwd: negative size");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)da & 15) == 0 && ((uintptr_t)dx & 15) == 0 &&
                    ((uintptr_t)dres & 15) == 0, "gn_bwd: buffers must be 16-byte aligned");
    GnGeom g;
    int rc = gn_geom(C, &g, "gn_bwd");
    if (rc) return rc;
    if (B * HW == 0) return DMVAE_OK;
    dim3 grid; int64_t ppb;
    gn_grid(B, HW, g.rows, &grid, &ppb);
    const size_t smem = (size_t)4 * C * sizeof(float);
    const size_t smem_r = (size_t)6 * C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    static int reverse = -1;         // DMVAE_GN_BWD_REVERSE=0 turns the backwards walk of the apply pass off (A/B measurements)
    if (reverse < 0) { const char* e = getenv("DMVAE_GN_BWD_REVERSE"); reverse = e ? (atoi(e) != 0) : 1; }
    static int U = 0;
    if (!U) { const char* e = getenv("DMVAE_GN_BWD_U"); U = e ? atoi(e) : 4; if (U != 2 && U != 3 && U != 4) U = 4; }     // 16-byte loads in flight per tensor per thread
#define GN_BWD_LAUNCH(S, UU)                                                                                                              \
    do {                                                                                                                                  \
        gn_bwd_reduce_kernel<S, UU><<<grid, GN_THREADS, smem_r, st>>>((const bf16*)da, (const bf16*)x, stats, gamma, beta, gsum, dgamma, \
                                                                      dbeta, HW, C, g.vc, g.rows, ppb, eps);                              \
        DMVAE_CHECK_LAUNCH("gn_bwd_reduce_kernel");                                                                                       \
        gn_bwd_apply_kernel<S, UU><<<grid, GN_THREADS, smem, st>>>((const bf16*)da, (const bf16*)x, stats, gamma, beta, gsum,            \
                                                                   (const bf16*)dres, (bf16*)dx, dx_colsum, HW, C, g.vc, g.rows, ppb, eps, \
                                                                   reverse);                                                              \
    } while (0)
    if (silu) {
        if (U == 2) GN_BWD_LAUNCH(true, 2); else if (U == 3) GN_BWD_LAUNCH(true, 3); else GN_BWD_LAUNCH(true, 4);
    } else {
        if (U == 2) GN_BWD_LAUNCH(false, 2); else if (U == 3) GN_BWD_LAUNCH(false, 3); else GN_BWD_LAUNCH(false, 4);
    }
#undef GN_BWD_LAUNCH
    DMVAE_CHECK_LAUNCH("gn_bwd_apply_kernel");
    return DMVAE_OK;
}
