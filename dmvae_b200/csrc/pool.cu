// VGG16 glue of LPIPS (utils/lpips.py:116-153: conv -> ReLU -> [tap] -> MaxPool2d(2, 2)) on channels-last bf16 activations.
// The ReLU forward lives in the conv epilogue (conv_tc.cu EPI_RELU) and its backward either in the consuming conv's
// data-gradient epilogue (EPI_MASK) or -- where a slice ends in a tap + pool -- in the pool backward below, which also adds the
// LPIPS tap gradient, so a VGG pass needs no elementwise ReLU / add / ATen pooling launches at all.
//   maxpool2x2_fwd      p[b][h][w][c]  = max over the 2x2 window of y                          8 B read + 2 B written per pooled... (10 B / 4 inputs)
//   pool_tap_bwd        dy[b][h][w][c] = ([y[h][w] is the window's first maximum] * dp[h/2][w/2] + dtap[h][w]) * [y[h][w] > 0]
//   relu_mask           dy = y > 0 ? dy : 0                                                    (the last tap, relu5_3: no pool after it)
// Ties go to the first maximum in (row, column) scan order, like ATen's max_pool2d (strict > while scanning).
#include "common.cuh"

static unsigned pool_grid(int64_t n) {
    int64_t g = ceil_div64(n, 256);
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (unsigned)g;
}

__global__ void __launch_bounds__(256) maxpool2x2_fwd_kernel(const uint4* __restrict__ y, uint4* __restrict__ p, int64_t B, int OH,
                                                             int OW, int vc) {
    const int64_t n = B * OH * OW * vc;
    const int W = 2 * OW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vc);
        int64_t q = i / vc;
        const int ow = (int)(q % OW); q /= OW;
        const int oh = (int)(q % OH);
        const int64_t b = q / OH;
        const int64_t base = ((b * 2 * OH + 2 * oh) * W + 2 * ow) * vc + v;
        float m[8], f[8];
        unpack_bf16x8(ld_stream16(y + base), m);
        unpack_bf16x8(ld_stream16(y + base + vc), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = f[k] > m[k] ? f[k] : m[k];
        unpack_bf16x8(ld_stream16(y + base + (int64_t)W * vc), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = f[k] > m[k] ? f[k] : m[k];
        unpack_bf16x8(ld_stream16(y + base + (int64_t)W * vc + vc), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = f[k] > m[k] ? f[k] : m[k];
        p[i] = pack_bf16x8(m);
    }
}

template <bool HAS_TAP, bool RELU>
__global__ void __launch_bounds__(256) pool_tap_bwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dp,
                                                           const uint4* __restrict__ dtap, uint4* __restrict__ dy, int64_t B, int OH,
                                                           int OW, int vc) {
    const int64_t n = B * OH * OW * vc;
    const int W = 2 * OW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vc);
        int64_t q = i / vc;
        const int ow = (int)(q % OW); q /= OW;
        const int oh = (int)(q % OH);
        const int64_t b = q / OH;
        const int64_t base = ((b * 2 * OH + 2 * oh) * W + 2 * ow) * vc + v;
        const int64_t off[4] = {base, base + vc, base + (int64_t)W * vc, base + (int64_t)W * vc + vc};
        float f[4][8], g[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack_bf16x8(ld_stream16(y + off[j]), f[j]);
        unpack_bf16x8(ld_stream16(dp + i), g);
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float m = f[0][k];
            int a = 0;
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (f[j][k] > m) { m = f[j][k]; a = j; }
            arg[k] = a;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float o[8];
            if (HAS_TAP) unpack_bf16x8(ld_stream16(dtap + off[j]), o);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float val = arg[k] == j ? g[k] : 0.f;
                if (HAS_TAP) val += o[k];                           // two bf16 gradients summed in fp32, one rounding (autograd adds in bf16 too)
                o[k] = (!RELU || f[j][k] > 0.f) ? val : 0.f;
            }
            st_stream16(dy + off[j], pack_bf16x8(o));
        }
    }
}

__global__ void __launch_bounds__(256) relu_mask_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dy, uint4* __restrict__ out,
                                                        int64_t nv) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        float f[8], g[8];
        unpack_bf16x8(ld_stream16(y + i), f);
        unpack_bf16x8(ld_stream16(dy + i), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = f[k] > 0.f ? g[k] : 0.f;
        st_stream16(out + i, pack_bf16x8(g));
    }
}

// nn.MaxPool2d(kernel_size=2, stride=2) on channels-last bf16: y [B][2*OH][2*OW][C] -> pooled [B][OH][OW][C]; C % 8 == 0
DMVAE_API int dmvae_maxpool2x2_fwd(const void* y, void* pooled, int64_t B, int OH, int OW, int C, void* stream) {
    DMVAE_CHECK_ARG(y && pooled, "maxpool2x2_fwd: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && OH > 0 && OW > 0 && C > 0 && C % 8 == 0, "maxpool2x2_fwd: bad shape (C must be a multiple of 8)");
    DMVAE_CHECK_ARG(((uintptr_t)y & 15) == 0 && ((uintptr_t)pooled & 15) == 0, "maxpool2x2_fwd: buffers must be 16-byte aligned");
    const int64_t n = B * OH * OW * (C / 8);
    if (n == 0) return DMVAE_OK;
    maxpool2x2_fwd_kernel<<<pool_grid(n), 256, 0, (cudaStream_t)stream>>>((const uint4*)y, (uint4*)pooled, B, OH, OW, C / 8);
    DMVAE_CHECK_LAUNCH("maxpool2x2_fwd_kernel");
    return DMVAE_OK;
}

// Backward of  y -> (tap = y, pooled = maxpool2x2(y))  with y = relu(.) when relu != 0:
//   dy = (route(d_pooled) + d_tap) * [y > 0]        d_tap may be null (no tap gradient)
DMVAE_API int dmvae_pool_tap_bwd(const void* y, const void* d_pooled, const void* d_tap, void* dy, int64_t B, int OH, int OW, int C,
                                 int relu, void* stream) {
    DMVAE_CHECK_ARG(y && d_pooled && dy, "pool_tap_bwd: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && OH > 0 && OW > 0 && C > 0 && C % 8 == 0, "pool_tap_bwd: bad shape (C must be a multiple of 8)");
    DMVAE_CHECK_ARG((((uintptr_t)y | (uintptr_t)d_pooled | (uintptr_t)d_tap | (uintptr_t)dy) & 15) == 0,
                    "pool_tap_bwd: buffers must be 16-byte aligned");
    const int64_t n = B * OH * OW * (C / 8);
    if (n == 0) return DMVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = pool_grid(n);
    const uint4 *py = (const uint4*)y, *pd = (const uint4*)d_pooled, *pt = (const uint4*)d_tap;
    uint4* po = (uint4*)dy;
    if (d_tap) {
        if (relu) pool_tap_bwd_kernel<true, true><<<grid, 256, 0, st>>>(py, pd, pt, po, B, OH, OW, C / 8);
        else pool_tap_bwd_kernel<true, false><<<grid, 256, 0, st>>>(py, pd, pt, po, B, OH, OW, C / 8);
    } else {
        if (relu) pool_tap_bwd_kernel<false, true><<<grid, 256, 0, st>>>(py, pd, pt, po, B, OH, OW, C / 8);
        else pool_tap_bwd_kernel<false, false><<<grid, 256, 0, st>>>(py, pd, pt, po, B, OH, OW, C / 8);
    }
    DMVAE_CHECK_LAUNCH("pool_tap_bwd_kernel");
    return DMVAE_OK;
}

// out = y > 0 ? dy : 0   (n bf16 elements, n % 8 == 0; out may alias dy)
DMVAE_API int dmvae_relu_mask(const void* y, const void* dy, void* out, int64_t n, void* stream) {
    DMVAE_CHECK_ARG(y && dy && out, "relu_mask: null pointer");
    DMVAE_CHECK_ARG(n >= 0 && n % 8 == 0, "relu_mask: element count must be a non-negative multiple of 8");
    DMVAE_CHECK_ARG((((uintptr_t)y | (uintptr_t)dy | (uintptr_t)out) & 15) == 0, "relu_mask: buffers must be 16-byte aligned");
    if (n == 0) return DMVAE_OK;
    relu_mask_kernel<<<pool_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)y, (const uint4*)dy, (uint4*)out, n / 8);
    DMVAE_CHECK_LAUNCH("relu_mask_kernel");
    return DMVAE_OK;
}
