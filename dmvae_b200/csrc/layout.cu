// Layout and small data-movement kernels around the conv tiles.
//   * NCHW <-> channels-last transposes with dtype conversion (module boundary: the reference's modules take and
//     return NCHW tensors, models/flux_ae.py:239-269; inside the library everything is [b][pixel][c] bf16)
//   * nearest-neighbour 2x upsample fwd/bwd               (models/flux_ae.py:103-104, F.interpolate(scale_factor=2, "nearest"))
//   * zero pad (0,1,0,1) helper is not needed: the stride-2 conv handles the bottom/right pad by bounds checks (:91-95)
//   * weight packing: fp32 [Cout][Cin][kh][kw] master weights -> bf16 tap-major GEMM operands
//   * bf16 elementwise add (gradient fan-in of residual branches, :82 / :52)
#include "common.cuh"

// src[b][R][C] (TI)  ->  dst[b][C][R] (TO), tiled through shared memory so both sides are coalesced
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) transpose_rc_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int64_t R, int64_t C) {
    __shared__ float tile[32][33];
    const int64_t b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const TI* s = src + b * R * C;
    TO* d = dst + b * R * C;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int64_t r = r0 + ty + i, c = c0 + tx;
        if (r < R && c < C) tile[ty + i][tx] = ld_as_float(s, r * C + c);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int64_t c = c0 + ty + i, r = r0 + tx;
        if (r < R && c < C) st_from_float(d, c * R + r, tile[tx][ty + i]);
    }
}

template <typename TI, typename TO>
static int launch_transpose(const void* src, void* dst, int64_t B, int64_t R, int64_t C, cudaStream_t st, const char* who) {
    if (B * R * C == 0) return DMVAE_OK;
    if (B > 65535 || ceil_div64(C, 32) > 65535) return dmvae_set_error(DMVAE_EUNSUPPORTED, "%s: grid too large", who);
    dim3 grid((unsigned)ceil_div64(R, 32), (unsigned)ceil_div64(C, 32), (unsigned)B);
    transpose_rc_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)src, (TO*)dst, R, C);
    DMVAE_CHECK_LAUNCH(who);
    return DMVAE_OK;
}

// NCHW (src_dtype) -> channels-last bf16.   [b][C][HW] -> [b][HW][C]
DMVAE_API int dmvae_nchw_to_nhwc(const void* src, void* dst, int64_t B, int C, int64_t HW, int src_dtype, void* stream) {
    DMVAE_CHECK_ARG(src && dst, "nchw_to_nhwc: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && C >= 0 && HW >= 0, "nchw_to_nhwc: negative size");
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == DMVAE_F32) return launch_transpose<float, bf16>(src, dst, B, C, HW, st, "nchw_to_nhwc");
    if (src_dtype == DMVAE_BF16) return launch_transpose<bf16, bf16>(src, dst, B, C, HW, st, "nchw_to_nhwc");
    return dmvae_set_error(DMVAE_EINVAL, "nchw_to_nhwc: bad dtype %d", src_dtype);
}

// channels-last bf16 -> NCHW (dst_dtype).   [b][HW][C] -> [b][C][HW]
DMVAE_API int dmvae_nhwc_to_nchw(const void* src, void* dst, int64_t B, int C, int64_t HW, int dst_dtype, void* stream) {
    DMVAE_CHECK_ARG(src && dst, "nhwc_to_nchw: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && C >= 0 && HW >= 0, "nhwc_to_nchw: negative size");
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == DMVAE_F32) return launch_transpose<bf16, float>(src, dst, B, HW, C, st, "nhwc_to_nchw");
    if (dst_dtype == DMVAE_BF16) return launch_transpose<bf16, bf16>(src, dst, B, HW, C, st, "nhwc_to_nchw");
    return dmvae_set_error(DMVAE_EINVAL, "nhwc_to_nchw: bad dtype %d", dst_dtype);
}

// ---- nearest 2x upsample, channels-last bf16, C % 8 == 0 -------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                             int64_t B, int H, int W, int vc) {
    // one thread per output 16-byte vector
    const int64_t n = B * (2 * H) * (2 * W) * vc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vc);
        int64_t p = i / vc;
        const int ow = (int)(p % (2 * W)); p /= (2 * W);
        const int oh = (int)(p % (2 * H));
        const int64_t b = p / (2 * H);
        y[i] = x[((b * H + (oh >> 1)) * W + (ow >> 1)) * vc + v];
    }
}
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                                             int64_t B, int H, int W, int vc) {
    const int64_t n = B * H * W * vc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vc);
        int64_t p = i / vc;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int64_t b = p / H;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int dh = 0; dh < 2; ++dh)
#pragma unroll
            for (int dw = 0; dw < 2; ++dw) {
                float f[8];
                unpack_bf16x8(ld_stream16(&dy[((b * 2 * H + 2 * h + dh) * 2 * W + 2 * w + dw) * vc + v]), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        dx[i] = pack_bf16x8(acc);
    }
}

// any channel count (tiny latent stems): one thread per output element
__global__ void __launch_bounds__(256) upsample2x_fwd_scalar_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                                    int64_t B, int H, int W, int C) {
    const int64_t n = B * (2 * H) * (2 * W) * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t p = i / C;
        const int ow = (int)(p % (2 * W)); p /= (2 * W);
        const int oh = (int)(p % (2 * H));
        const int64_t b = p / (2 * H);
        y[i] = x[((b * H + (oh >> 1)) * W + (ow >> 1)) * C + c];
    }
}
__global__ void __launch_bounds__(256) upsample2x_bwd_scalar_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                                                    int64_t B, int H, int W, int C) {
    const int64_t n = B * H * W * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t p = i / C;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int64_t b = p / H;
        float a = 0.f;
        for (int dh = 0; dh < 2; ++dh)
            for (int dw = 0; dw < 2; ++dw)
                a += __bfloat162float(dy[((b * 2 * H + 2 * h + dh) * 2 * W + 2 * w + dw) * C + c]);
        dx[i] = __float2bfloat16_rn(a);
    }
}

static unsigned ew_grid(int64_t n) {
    int64_t g = ceil_div64(n, 256);
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (unsigned)g;
}

DMVAE_API int dmvae_upsample2x_fwd(const void* x, void* y, int64_t B, int H, int W, int C, void* stream) {
    DMVAE_CHECK_ARG(x && y, "upsample2x_fwd: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0, "upsample2x_fwd: bad shape");
    if (C % 8 != 0 || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) {
        const int64_t ne = B * 4 * H * W * C;
        if (ne == 0) return DMVAE_OK;
        upsample2x_fwd_scalar_kernel<<<ew_grid(ne), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, H, W, C);
        DMVAE_CHECK_LAUNCH("upsample2x_fwd_scalar_kernel");
        return DMVAE_OK;
    }
    const int64_t n = B * 4 * H * W * (C / 8);
    if (n == 0) return DMVAE_OK;
    upsample2x_fwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, B, H, W, C / 8);
    DMVAE_CHECK_LAUNCH("upsample2x_fwd_kernel");
    return DMVAE_OK;
}

// dy: [B][2H][2W][C] -> dx: [B][H][W][C] (sum of the 4 replicas, fp32 accumulate)
DMVAE_API int dmvae_upsample2x_bwd(const void* dy, void* dx, int64_t B, int H, int W, int C, void* stream) {
    DMVAE_CHECK_ARG(dy && dx, "upsample2x_bwd: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0, "upsample2x_bwd: bad shape");
    if (C % 8 != 0 || ((uintptr_t)dy & 15) || ((uintptr_t)dx & 15)) {
        const int64_t ne = B * H * W * C;
        if (ne == 0) return DMVAE_OK;
        upsample2x_bwd_scalar_kernel<<<ew_grid(ne), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (bf16*)dx, B, H, W, C);
        DMVAE_CHECK_LAUNCH("upsample2x_bwd_scalar_kernel");
        return DMVAE_OK;
    }
    const int64_t n = B * H * W * (C / 8);
    if (n == 0) return DMVAE_OK;
    upsample2x_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (uint4*)dx, B, H, W, C / 8);
    DMVAE_CHECK_LAUNCH("upsample2x_bwd_kernel");
    return DMVAE_OK;
}

// ---- bf16 add ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                       bf16* __restrict__ o, int64_t n) {
    const int64_t nv = n >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float fa[8], fb[8];
        unpack_bf16x8(ld_stream16(a + 8 * i), fa);
        unpack_bf16x8(ld_stream16(b + 8 * i), fb);
#pragma unroll
        for (int k = 0; k < 8; ++k) fa[k] += fb[k];
        st_stream16(o + 8 * i, pack_bf16x8(fa));
    }
    for (int64_t i = (nv << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        o[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}

DMVAE_API int dmvae_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
    DMVAE_CHECK_ARG(a && b && out, "add_bf16: null pointer");
    DMVAE_CHECK_ARG(n >= 0, "add_bf16: negative size");
    DMVAE_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)out & 15) == 0, "add_bf16: buffers must be 16-byte aligned");
    if (n == 0) return DMVAE_OK;
    add_bf16_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n);
    DMVAE_CHECK_LAUNCH("add_bf16_kernel");
    return DMVAE_OK;
}

// ---- frozen-encoder glue (models/vae.py:34-53: the timm ViT under autocast, no_grad) ---------------------------------
// LayerScale + residual:  x[r][d] += float(y[r][d]) * gamma[d]   (x fp32 residual stream, y the bf16 Linear output).
// The reference does this as two ATen passes (a type-promoting multiply, then an add); product and sum are rounded
// separately here too (no FMA contraction), so the result is bit-identical.
__global__ void __launch_bounds__(256) scale_residual_kernel(float* __restrict__ x, const bf16* __restrict__ y,
                                                             const float* __restrict__ gamma, int64_t rows, int D) {
    const int vd = D >> 3;
    const int64_t nv = rows * vd;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const int c = (int)(i % vd) * 8;
        float fy[8];
        unpack_bf16x8(ld_stream16(y + 8 * i), fy);
        float4 a = *reinterpret_cast<const float4*>(x + 8 * i), b = *reinterpret_cast<const float4*>(x + 8 * i + 4);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        a.x = __fadd_rn(a.x, __fmul_rn(fy[0], g0.x)); a.y = __fadd_rn(a.y, __fmul_rn(fy[1], g0.y));
        a.z = __fadd_rn(a.z, __fmul_rn(fy[2], g0.z)); a.w = __fadd_rn(a.w, __fmul_rn(fy[3], g0.w));
        b.x = __fadd_rn(b.x, __fmul_rn(fy[4], g1.x)); b.y = __fadd_rn(b.y, __fmul_rn(fy[5], g1.y));
        b.z = __fadd_rn(b.z, __fmul_rn(fy[6], g1.z)); b.w = __fadd_rn(b.w, __fmul_rn(fy[7], g1.w));
        *reinterpret_cast<float4*>(x + 8 * i) = a;
        *reinterpret_cast<float4*>(x + 8 * i + 4) = b;
    }
}

DMVAE_API int dmvae_scale_residual(float* x, const void* y, const float* gamma, int64_t rows, int D, void* stream) {
    DMVAE_CHECK_ARG(x && y && gamma, "scale_residual: null pointer");
    DMVAE_CHECK_ARG(rows >= 0 && D > 0 && D % 8 == 0, "scale_residual: need rows >= 0 and D a positive multiple of 8");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)gamma & 15) == 0, "scale_residual: buffers must be 16-byte aligned");
    if (rows == 0) return DMVAE_OK;
    scale_residual_kernel<<<ew_grid(rows * (D / 8)), 256, 0, (cudaStream_t)stream>>>(x, (const bf16*)y, gamma, rows, D);
    DMVAE_CHECK_LAUNCH("scale_residual_kernel");
    return DMVAE_OK;
}

// LayerNorm over the last dimension of an fp32 [rows][D] tensor, output bf16 (what autocast hands the following Linear:
// layer_norm runs in fp32, the Linear casts its input to bf16).  One warp per row, two passes over registers
// (mean, then centred variance), D <= 2048.
template <int VPL>      // float4 vectors per lane: D = 128 * VPL
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, bf16* __restrict__ y, int64_t rows, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    constexpr int D = 128 * VPL;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { v[i] = xr[lane + 32 * i]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    s = warp_sum(s);
    const float mean = s * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + c * c) + (d * d + e * e);
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q * (1.f / D) + eps);
    bf16* yr = y + row * D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (lane + 32 * i) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(w + c)), be = __ldg(reinterpret_cast<const float4*>(b + c));
        const float o0 = (v[i].x - mean) * rstd * g.x + be.x, o1 = (v[i].y - mean) * rstd * g.y + be.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + be.z, o3 = (v[i].w - mean) * rstd * g.w + be.w;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(yr + c) = pk;
    }
}

DMVAE_API int dmvae_layernorm_bf16(const float* x, const float* weight, const float* bias, void* y, int64_t rows, int D, float eps, void* stream) {
    DMVAE_CHECK_ARG(x && weight && bias && y, "layernorm_bf16: null pointer");
    DMVAE_CHECK_ARG(rows >= 0, "layernorm_bf16: negative size");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 7) == 0 && ((uintptr_t)weight & 15) == 0 && ((uintptr_t)bias & 15) == 0,
                    "layernorm_bf16: buffers must be 16-byte aligned");
    if (D != 768 && D != 1024 && D != 384 && D != 512)
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "layernorm_bf16: D=%d not supported (384, 512, 768, 1024)", D);
    if (rows == 0) return DMVAE_OK;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    bf16* yp = (bf16*)y;
    switch (D) {
        case 384: layernorm_bf16_kernel<3><<<grid, 256, 0, st>>>(x, weight, bias, yp, rows, eps); break;
        case 512: layernorm_bf16_kernel<4><<<grid, 256, 0, st>>>(x, weight, bias, yp, rows, eps); break;
        case 768: layernorm_bf16_kernel<6><<<grid, 256, 0, st>>>(x, weight, bias, yp, rows, eps); break;
        default: layernorm_bf16_kernel<8><<<grid, 256, 0, st>>>(x, weight, bias, yp, rows, eps); break;
    }
    DMVAE_CHECK_LAUNCH("layernorm_bf16_kernel");
    return DMVAE_OK;
}

// LayerScale + residual of one branch fused with the LayerNorm that opens the next one:
//   x[r][d] += float(y[r][d]) * gamma[d]   (written back: the fp32 residual stream)   ;   out[r] = bf16(LayerNorm(x[r]; w, b, eps))
// Same arithmetic, in the same order and with the same roundings, as dmvae_scale_residual followed by dmvae_layernorm_bf16 (the row
// stays in registers between the two), i.e. bit-identical to them; one launch and one read of x fewer per branch (96 -> 49 glue
// launches in a ViT-L pass).
template <int VPL>
__global__ void __launch_bounds__(256) scale_residual_layernorm_kernel(float* __restrict__ x, const bf16* __restrict__ y,
                                                                       const float* __restrict__ gamma, const float* __restrict__ w,
                                                                       const float* __restrict__ b, bf16* __restrict__ out, int64_t rows,
                                                                       float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    constexpr int D = 128 * VPL;
    float4* xr = reinterpret_cast<float4*>(x + row * D);
    const bf16* yr = y + row * D;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (lane + 32 * i) * 4;
        v[i] = xr[lane + 32 * i];
        const uint2 yy = *reinterpret_cast<const uint2*>(yr + c);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        v[i].x = __fadd_rn(v[i].x, __fmul_rn(__uint_as_float(yy.x << 16), g.x));
        v[i].y = __fadd_rn(v[i].y, __fmul_rn(__uint_as_float(yy.x & 0xffff0000u), g.y));
        v[i].z = __fadd_rn(v[i].z, __fmul_rn(__uint_as_float(yy.y << 16), g.z));
        v[i].w = __fadd_rn(v[i].w, __fmul_rn(__uint_as_float(yy.y & 0xffff0000u), g.w));
        xr[lane + 32 * i] = v[i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    s = warp_sum(s);
    const float mean = s * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + c * c) + (d * d + e * e);
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q * (1.f / D) + eps);
    bf16* orow = out + row * D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = (lane + 32 * i) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(w + c)), be = __ldg(reinterpret_cast<const float4*>(b + c));
        const float o0 = (v[i].x - mean) * rstd * g.x + be.x, o1 = (v[i].y - mean) * rstd * g.y + be.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + be.z, o3 = (v[i].w - mean) * rstd * g.w + be.w;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(orow + c) = pk;
    }
}

DMVAE_API int dmvae_scale_residual_layernorm(float* x, const void* y, const float* gamma, const float* weight, const float* bias,
                                             void* out, int64_t rows, int D, float eps, void* stream) {
    DMVAE_CHECK_ARG(x && y && gamma && weight && bias && out, "scale_residual_layernorm: null pointer");
    DMVAE_CHECK_ARG(rows >= 0, "scale_residual_layernorm: negative size");
    DMVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)gamma | (uintptr_t)weight | (uintptr_t)bias) & 15) == 0 &&
                    (((uintptr_t)y | (uintptr_t)out) & 7) == 0, "scale_residual_layernorm: buffers must be 16-byte (bf16: 8-byte) aligned");
    if (D != 768 && D != 1024 && D != 384 && D != 512)
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "scale_residual_layernorm: D=%d not supported (384, 512, 768, 1024)", D);
    if (rows == 0) return DMVAE_OK;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    const bf16* yp = (const bf16*)y;
    bf16* op = (bf16*)out;
    switch (D) {
        case 384: scale_residual_layernorm_kernel<3><<<grid, 256, 0, st>>>(x, yp, gamma, weight, bias, op, rows, eps); break;
        case 512: scale_residual_layernorm_kernel<4><<<grid, 256, 0, st>>>(x, yp, gamma, weight, bias, op, rows, eps); break;
        case 768: scale_residual_layernorm_kernel<6><<<grid, 256, 0, st>>>(x, yp, gamma, weight, bias, op, rows, eps); break;
        default: scale_residual_layernorm_kernel<8><<<grid, 256, 0, st>>>(x, yp, gamma, weight, bias, op, rows, eps); break;
    }
    DMVAE_CHECK_LAUNCH("scale_residual_layernorm_kernel");
    return DMVAE_OK;
}

// ---- weight packing ------------------------------------------------------------------------------------------
// w[co][ci][kh][kw] fp32  ->  wf[tap][co][ci] bf16  (forward operand, K = ci contiguous)
//                         ->  wd[tap'][ci][co] bf16 (dgrad operand: tap' = flipped tap, K = co contiguous)
// One CTA owns a 32(co) x 32(ci) tile for all taps: coalesced fp32 reads, shared-memory transpose, 64-byte bf16 segments
// out in both layouts.  Runs once per optimizer step per conv (the packed operands are a cache of the fp32 parameter).
#define PK_T 32
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ wf,
                                                           bf16* __restrict__ wd, int Cout, int Cin, int taps) {
    extern __shared__ float tile[];                    // [PK_T co][PK_T ci * taps + 1]
    const int co0 = blockIdx.x * PK_T, ci0 = blockIdx.y * PK_T;
    const int row = PK_T * taps + 1;
    const int nci = min(PK_T, Cin - ci0), nco = min(PK_T, Cout - co0);
    // 32 x 8 thread layout, no integer division in the loops
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // load: for each co, the (ci0 .. ci0+nci) x taps block is contiguous in w
    for (int c = ty; c < nco; c += 8) {
        const float* src = w + ((int64_t)(co0 + c) * Cin + ci0) * taps;
        for (int r = tx; r < nci * taps; r += 32) tile[c * row + r] = src[r];
    }
    __syncthreads();
    // wf[tap][co][ci]: a warp writes 32 consecutive ci
    if (wf && tx < nci)
        for (int tap = 0; tap < taps; ++tap)
            for (int c = ty; c < nco; c += 8)
                wf[((int64_t)tap * Cout + co0 + c) * Cin + ci0 + tx] = __float2bfloat16_rn(tile[c * row + tx * taps + tap]);
    // wd[taps-1-tap][ci][co]: a warp writes 32 consecutive co
    if (wd && tx < nco)
        for (int tap = 0; tap < taps; ++tap)
            for (int ci = ty; ci < nci; ci += 8)
                wd[((int64_t)(taps - 1 - tap) * Cin + ci0 + ci) * Cout + co0 + tx] = __float2bfloat16_rn(tile[tx * row + ci * taps + tap]);
}

DMVAE_API int dmvae_pack_weights(const float* w, void* w_fwd, void* w_dgrad, int Cout, int Cin, int KH, int KW, void* stream) {
    DMVAE_CHECK_ARG(w && (w_fwd || w_dgrad), "pack_weights: null pointer");
    DMVAE_CHECK_ARG(Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "pack_weights: bad shape");
    const int taps = KH * KW;
    const size_t smem = (size_t)PK_T * (PK_T * taps + 1) * sizeof(float);
    DMVAE_CHECK_ARG(smem <= 96 * 1024, "pack_weights: filter too large (%d taps)", taps);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(pack_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
    dim3 grid((unsigned)((Cout + PK_T - 1) / PK_T), (unsigned)((Cin + PK_T - 1) / PK_T));
    pack_weights_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(w, (bf16*)w_fwd, (bf16*)w_dgrad, Cout, Cin, taps);
    DMVAE_CHECK_LAUNCH("pack_weights_kernel");
    return DMVAE_OK;
}

// bf16 forward operand wf[tap][co][ci]  ->  data-gradient operand wd[taps-1-tap][ci][co] (bf16 -> bf16, one 32x32 transpose per
// tile and tap).  Used when the forward operand is maintained by the optimizer kernel (tap-major parameter arenas, optim.py): the
// fp32 masters are then not read again for packing.
__global__ void __launch_bounds__(256) pack_dgrad_bf16_kernel(const bf16* __restrict__ wf, bf16* __restrict__ wd, int Cout, int Cin, int taps) {
    __shared__ bf16 tile[32][33];
    const int tap = blockIdx.z;
    const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const bf16* src = wf + (int64_t)tap * Cout * Cin;
    bf16* dst = wd + (int64_t)(taps - 1 - tap) * Cin * Cout;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int co = co0 + ty + i, ci = ci0 + tx;
        if (co < Cout && ci < Cin) tile[ty + i][tx] = src[(int64_t)co * Cin + ci];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int ci = ci0 + ty + i, co = co0 + tx;
        if (co < Cout && ci < Cin) dst[(int64_t)ci * Cout + co] = tile[tx][ty + i];
    }
}

DMVAE_API int dmvae_pack_dgrad_bf16(const void* w_fwd, void* w_dgrad, int Cout, int Cin, int taps, void* stream) {
    DMVAE_CHECK_ARG(w_fwd && w_dgrad, "pack_dgrad_bf16: null pointer");
    DMVAE_CHECK_ARG(Cout > 0 && Cin > 0 && taps > 0 && taps <= 64, "pack_dgrad_bf16: bad shape");
    dim3 grid((unsigned)((Cout + 31) / 32), (unsigned)((Cin + 31) / 32), (unsigned)taps);
    pack_dgrad_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)w_fwd, (bf16*)w_dgrad, Cout, Cin, taps);
    DMVAE_CHECK_LAUNCH("pack_dgrad_bf16_kernel");
    return DMVAE_OK;
}

// The same transpose for EVERY conv weight of a flat bf16 arena in one launch (the optimizer issues it right after its update
// kernel, optim.py): desc[i] = {element offset of the parameter in both arenas, Cout, Cin, taps, index of its first 32x32 tile};
// a CTA finds its descriptor by binary search over the tile starts.  Tiles of one parameter are ordered [tap][ci tile][co tile].
__global__ void __launch_bounds__(256) pack_dgrad_batched_kernel(const bf16* __restrict__ wf_flat, bf16* __restrict__ wd_flat,
                                                                 const int64_t* __restrict__ desc, int n_desc) {
    __shared__ bf16 tile[32][33];
    const int64_t t = blockIdx.x;
    int lo = 0, hi = n_desc - 1;
    while (lo < hi) {                                   // last descriptor whose first tile is <= t
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(desc + 5 * mid + 4) <= t) lo = mid; else hi = mid - 1;
    }
    const int64_t* d = desc + 5 * lo;
    const int64_t off = __ldg(d);
    const int Cout = (int)__ldg(d + 1), Cin = (int)__ldg(d + 2), taps = (int)__ldg(d + 3);
    const int tco = (Cout + 31) >> 5, tci = (Cin + 31) >> 5;
    int r = (int)(t - __ldg(d + 4));
    const int tap = r / (tco * tci);
    r -= tap * tco * tci;
    const int co0 = (r % tco) * 32, ci0 = (r / tco) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const bf16* src = wf_flat + off + (int64_t)tap * Cout * Cin;
    bf16* dst = wd_flat + off + (int64_t)(taps - 1 - tap) * Cin * Cout;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int co = co0 + ty + i, ci = ci0 + tx;
        if (co < Cout && ci < Cin) tile[ty + i][tx] = src[(int64_t)co * Cin + ci];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int ci = ci0 + ty + i, co = co0 + tx;
        if (co < Cout && ci < Cin) dst[(int64_t)ci * Cout + co] = tile[tx][ty + i];
    }
}

DMVAE_API int dmvae_pack_dgrad_batched(const void* w_fwd_flat, void* w_dgrad_flat, const int64_t* desc, int n_desc, int64_t total_tiles,
                                       void* stream) {
    DMVAE_CHECK_ARG(n_desc >= 0 && total_tiles >= 0 && total_tiles < (int64_t)1 << 31, "pack_dgrad_batched: bad size");
    if (n_desc == 0 || total_tiles == 0) return DMVAE_OK;
    DMVAE_CHECK_ARG(w_fwd_flat && w_dgrad_flat && desc, "pack_dgrad_batched: null pointer");
    pack_dgrad_batched_kernel<<<(unsigned)total_tiles, 256, 0, (cudaStream_t)stream>>>((const bf16*)w_fwd_flat, (bf16*)w_dgrad_flat, desc, n_desc);
    DMVAE_CHECK_LAUNCH("pack_dgrad_batched_kernel");
    return DMVAE_OK;
}

// tap-major wgrad scratch [tap][Cout][Cin] -> state_dict layout dw[co][ci][tap] (accumulate=1: +=).
// One CTA per (co, 256-wide ci chunk): taps coalesced row reads, shared-memory interleave, one contiguous write.
__global__ void __launch_bounds__(256) wgrad_unpack_kernel(const float* __restrict__ dwp, float* __restrict__ dw,
                                                           int Cout, int Cin, int taps, int accumulate) {
    extern __shared__ float s_t[];                     // [256 ci][taps] interleaved as the destination wants it
    const int co = blockIdx.x, ci0 = blockIdx.y * 256;
    const int nci = min(256, Cin - ci0);
    for (int tap = 0; tap < taps; ++tap)
        if (threadIdx.x < nci)
            s_t[threadIdx.x * taps + tap] = dwp[((int64_t)tap * Cout + co) * Cin + ci0 + threadIdx.x];
    __syncthreads();
    float* dst = dw + ((int64_t)co * Cin + ci0) * taps;
    for (int i = threadIdx.x; i < nci * taps; i += blockDim.x) dst[i] = accumulate ? dst[i] + s_t[i] : s_t[i];
}

DMVAE_API int dmvae_wgrad_unpack(const float* dw_tap_major, float* dw, int Cout, int Cin, int taps, int accumulate, void* stream) {
    DMVAE_CHECK_ARG(dw_tap_major && dw, "wgrad_unpack: null pointer");
    DMVAE_CHECK_ARG(Cout > 0 && Cin > 0 && taps > 0 && taps <= 32, "wgrad_unpack: bad shape");
    dim3 grid((unsigned)Cout, (unsigned)((Cin + 255) / 256));
    wgrad_unpack_kernel<<<grid, 256, (size_t)256 * taps * sizeof(float), (cudaStream_t)stream>>>(dw_tap_major, dw, Cout, Cin, taps, accumulate);
    DMVAE_CHECK_LAUNCH("wgrad_unpack_kernel");
    return DMVAE_OK;
}

// ---- sub-pixel form of nearest-2x + 3x3 (flux_ae.Upsample, :98-107; see dmvae_conv_up2x_fwd in conv_tc.cu) -------------------------
// Output phase p in {0,1} along one axis reads input offsets {-1, 0} (p = 0) or {0, +1} (p = 1); slot a in {0,1} of that 2-tap filter
// collects the 3x3 taps k in S(p, a):  S(0,0) = {0}, S(0,1) = {1,2}, S(1,0) = {0,1}, S(1,1) = {2}.
__host__ __device__ __forceinline__ bool subpixel_in_S(int p, int a, int k) {
    return p == 0 ? (a == 0 ? k == 0 : k >= 1) : (a == 0 ? k <= 1 : k == 2);
}

// w3 (fp32, element strides s_co / s_ci / s_tap: state_dict or tap-major storage) ->
//   wp_fwd  [4*(2py+px) + 2a + b][co][ci]  bf16 = sum of the 3x3 taps (kh in S(py,a), kw in S(px,b)), summed in fp32, rounded once
//   wp_dgrad[4*(2py+px) + (3 - (2a+b))][ci][co]  bf16: the same values transposed, the 2x2 filter flipped (data-gradient operand)
__global__ void __launch_bounds__(256) subpixel_pack_kernel(const float* __restrict__ w3, int64_t s_co, int64_t s_ci, int64_t s_tap,
                                                            bf16* __restrict__ wf, bf16* __restrict__ wd, int Cout, int Cin) {
    __shared__ float tile[32][33];
    const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 (ci) x 8 (co)
    float w[4][9];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty + 8 * i, ci = ci0 + tx;
#pragma unroll
        for (int k = 0; k < 9; ++k) w[i][k] = (co < Cout && ci < Cin) ? w3[co * s_co + ci * s_ci + k * s_tap] : 0.f;
    }
    for (int t = 0; t < 16; ++t) {
        const int py = t >> 3, px = (t >> 2) & 1, a = (t >> 1) & 1, b = t & 1;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = 0.f;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
                    if (subpixel_in_S(py, a, kh) && subpixel_in_S(px, b, kw)) v += w[i][kh * 3 + kw];
            const int co = co0 + ty + 8 * i, ci = ci0 + tx;
            if (co < Cout && ci < Cin) wf[((int64_t)t * Cout + co) * Cin + ci] = __float2bfloat16_rn(v);
            tile[ty + 8 * i][tx] = v;
        }
        __syncthreads();
        const int td = (t & ~3) | (3 - (t & 3));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ci = ci0 + ty + 8 * i, co = co0 + tx;
            if (co < Cout && ci < Cin) wd[((int64_t)td * Cin + ci) * Cout + co] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
        }
    }
}

DMVAE_API int dmvae_subpixel_pack(const float* w3, int64_t stride_co, int64_t stride_ci, int64_t stride_tap, void* wp_fwd, void* wp_dgrad,
                                  int Cout, int Cin, void* stream) {
    DMVAE_CHECK_ARG(w3 && wp_fwd && wp_dgrad, "subpixel_pack: null pointer");
    DMVAE_CHECK_ARG(Cout > 0 && Cin > 0, "subpixel_pack: bad shape");
    dim3 grid((unsigned)((Cout + 31) / 32), (unsigned)((Cin + 31) / 32));
    subpixel_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w3, stride_co, stride_ci, stride_tap, (bf16*)wp_fwd, (bf16*)wp_dgrad, Cout, Cin);
    DMVAE_CHECK_LAUNCH("subpixel_pack_kernel");
    return DMVAE_OK;
}

// dw3[co][ci][kh][kw] (element strides as above) += sum over the phase-taps that contain (kh, kw) of dwp[16][Cout][Cin]
__global__ void __launch_bounds__(256) subpixel_fold_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw3, int64_t s_co,
                                                                  int64_t s_ci, int64_t s_tap, int Cout, int Cin) {
    const int64_t n = (int64_t)Cout * Cin;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin), co = (int)(i / Cin);
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = dwp[(int64_t)t * n + i];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < 16; ++t)
                    if (subpixel_in_S(t >> 3, (t >> 1) & 1, kh) && subpixel_in_S((t >> 2) & 1, t & 1, kw)) acc += v[t];
                dw3[co * s_co + ci * s_ci + (kh * 3 + kw) * s_tap] += acc;
            }
    }
}

DMVAE_API int dmvae_subpixel_fold_wgrad(const float* dwp, float* dw3, int64_t stride_co, int64_t stride_ci, int64_t stride_tap, int Cout,
                                        int Cin, void* stream) {
    DMVAE_CHECK_ARG(dwp && dw3, "subpixel_fold_wgrad: null pointer");
    DMVAE_CHECK_ARG(Cout > 0 && Cin > 0, "subpixel_fold_wgrad: bad shape");
    subpixel_fold_wgrad_kernel<<<ew_grid((int64_t)Cout * Cin), 256, 0, (cudaStream_t)stream>>>(dwp, dw3, stride_co, stride_ci, stride_tap, Cout, Cin);
    DMVAE_CHECK_LAUNCH("subpixel_fold_wgrad_kernel");
    return DMVAE_OK;
}

// ---- gradient patches of a thin output (the C->3 head) ---------------------------------------------------------------
// P[pixel][j], j = tap*Cout + co  (padded with zeros to 32 columns):  P[p][j] = dy[p - offset(tap)][co], zero outside.
// With it both gradients of a "same" conv with a handful of output channels become plain 1x1 GEMMs on the tcgen05 tiles:
//   dx[p][ci]          = sum_j P[p][j] * Wm[j][ci]          (conv_tc_fwd, 1x1, Cin = 32)
//   dw[(tap,co)][ci]   = sum_p P[p][j] * x[p][ci]           (conv_tc_wgrad, 1x1, "Cout" = 32)
// instead of CUDA-core kernels that are bound by instruction issue (models/flux_ae.py:237 conv_out, 128 -> 3).
__global__ void __launch_bounds__(256) grad_patches_kernel(const bf16* __restrict__ dy, bf16* __restrict__ P, int64_t B, int H, int W,
                                                           int Cout, int KH, int KW, int pt, int pl) {
    const int64_t n = B * H * W * 4;                       // 4 x 16-byte vectors per 32-column row
    const int taps = KH * KW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i & 3);
        int64_t p = i >> 2;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int64_t b = p / H;
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int j = v * 8 + k;
            float val = 0.f;
            if (j < taps * Cout) {
                const int tap = j / Cout, co = j - tap * Cout;
                // y[q] uses x[q + (kh - pt, kw - pl)]  =>  x[p] meets dy[q] with q = p - (kh - pt, kw - pl)
                const int qh = h - (tap / KW - pt), qw = w - (tap % KW - pl);
                if (qh >= 0 && qh < H && qw >= 0 && qw < W) val = __bfloat162float(dy[((b * H + qh) * W + qw) * Cout + co]);
            }
            f[k] = val;
        }
        *reinterpret_cast<uint4*>(P + (i << 3)) = pack_bf16x8(f);
    }
}

DMVAE_API int dmvae_grad_patches(const void* dy, void* patches, int64_t B, int H, int W, int Cout, int KH, int KW, int pad_top,
                                 int pad_left, void* stream) {
    DMVAE_CHECK_ARG(dy && patches, "grad_patches: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0 && Cout > 0 && KH > 0 && KW > 0, "grad_patches: bad shape");
    DMVAE_CHECK_ARG(KH * KW * Cout <= 32, "grad_patches: taps*Cout = %d exceeds 32", KH * KW * Cout);
    DMVAE_CHECK_ARG(((uintptr_t)patches & 15) == 0, "grad_patches: output must be 16-byte aligned");
    const int64_t n = B * H * W * 4;
    if (n == 0) return DMVAE_OK;
    grad_patches_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (bf16*)patches, B, H, W, Cout, KH, KW, pad_top, pad_left);
    DMVAE_CHECK_LAUNCH("grad_patches_kernel");
    return DMVAE_OK;
}

// ---- zero insertion for the data gradient of a stride-2 conv (flux_ae.Downsample: pad (0,1,0,1), 3x3, stride 2) ----------
// dyz[b][2*oh+1][2*ow+1][c] = dy[b][oh][ow][c], zero elsewhere (H = 2*OH, W = 2*OW).  Then
//   dx = conv3x3_same(dyz, flipped/transposed weights)
// which runs on the tcgen05 tile (4x the minimal FLOPs of three small layers, instead of a CUDA-core gather).
__global__ void __launch_bounds__(256) zero_insert2x_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dyz, int64_t B,
                                                            int OH, int OW, int vc) {
    const int H = 2 * OH, W = 2 * OW;
    const int64_t n = B * H * W * vc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vc);
        int64_t p = i / vc;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int64_t b = p / H;
        uint4 val = make_uint4(0, 0, 0, 0);
        if ((h & 1) && (w & 1)) val = dy[((b * OH + (h >> 1)) * OW + (w >> 1)) * vc + v];
        dyz[i] = val;
    }
}

DMVAE_API int dmvae_zero_insert2x(const void* dy, void* dyz, int64_t B, int OH, int OW, int C, void* stream) {
    DMVAE_CHECK_ARG(dy && dyz, "zero_insert2x: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && OH > 0 && OW > 0 && C > 0 && C % 8 == 0, "zero_insert2x: bad shape (C must be a multiple of 8)");
    DMVAE_CHECK_ARG(((uintptr_t)dy & 15) == 0 && ((uintptr_t)dyz & 15) == 0, "zero_insert2x: buffers must be 16-byte aligned");
    const int64_t n = B * 4 * OH * OW * (C / 8);
    if (n == 0) return DMVAE_OK;
    zero_insert2x_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (uint4*)dyz, B, OH, OW, C / 8);
    DMVAE_CHECK_LAUNCH("zero_insert2x_kernel");
    return DMVAE_OK;
}
