// Fused optimizer step over flat fp32 arenas: global-norm clip + AdamW + EMA in two HBM passes
// instead of the reference's clip_grad_norm_ + AdamW.step + update_ema (train_tokenizer.py:415-417,437 / :140-150,
// train_dmd.py:540-544).  Arithmetic follows torch.optim.AdamW (decoupled weight decay, bias correction, eps added
// after the bias-corrected sqrt) and torch.nn.utils.clip_grad_norm_ (coef = min(1, max_norm / (norm + 1e-6))).
//   pass 1: acc[0] += sum g^2                                              4 B / parameter
//   pass 2: g *= coef ; m, v, p, ema updated                               20 B read + 16 B written / parameter
//           (+ 2 B written: w16 = bf16(p), the conv tiles' forward operand -- for arenas that keep 3x3 conv weights tap-major
//            ([tap][Cout][Cin], optim.py) this IS the packed bf16 weight, so no separate pack_weights pass reads the fp32 masters again)
#include "common.cuh"

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, double* __restrict__ acc, int64_t n) {
    __shared__ float red[32];
    float a[1] = {0.f};
    const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = *reinterpret_cast<const float4*>(g + 4 * i);
        a[0] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[0] += g[i] * g[i];
    block_sum<1>(a, red);
    if (threadIdx.x == 0) atomicAdd(acc, (double)a[0]);
}

struct AdamArgs {
    float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, max_norm, ema_decay;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* ema, const AdamArgs& a, float coef) {
    g *= coef;
    p *= 1.f - a.lr * a.wd;
    m = a.beta1 * m + (1.f - a.beta1) * g;
    v = a.beta2 * v + (1.f - a.beta2) * g * g;
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p -= (a.lr / a.bc1) * (m / denom);
    if (ema) *ema = a.ema_decay * (*ema) + (1.f - a.ema_decay) * p;
}

__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ ema, bf16* __restrict__ w16,
                                                        const double* __restrict__ sumsq, float* __restrict__ norm_out,
                                                        int64_t n, AdamArgs a) {
    float coef = 1.f;
    if (sumsq) {
        const float norm = (float)sqrt(*sumsq);
        if (a.max_norm > 0.f) coef = fminf(1.f, a.max_norm / (norm + 1e-6f));
        if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
    }
    const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = *reinterpret_cast<float4*>(p + 4 * i), G = *reinterpret_cast<const float4*>(g + 4 * i);
        float4 M = *reinterpret_cast<float4*>(m + 4 * i), V = *reinterpret_cast<float4*>(v + 4 * i);
        float4 E = ema ? *reinterpret_cast<float4*>(ema + 4 * i) : make_float4(0, 0, 0, 0);
        adam_one(P.x, G.x, M.x, V.x, ema ? &E.x : nullptr, a, coef);
        adam_one(P.y, G.y, M.y, V.y, ema ? &E.y : nullptr, a, coef);
        adam_one(P.z, G.z, M.z, V.z, ema ? &E.z : nullptr, a, coef);
        adam_one(P.w, G.w, M.w, V.w, ema ? &E.w : nullptr, a, coef);
        *reinterpret_cast<float4*>(p + 4 * i) = P;
        *reinterpret_cast<float4*>(m + 4 * i) = M;
        *reinterpret_cast<float4*>(v + 4 * i) = V;
        if (ema) *reinterpret_cast<float4*>(ema + 4 * i) = E;
        if (w16) {
            uint2 pk;
            pk.x = pack_bf16x2(P.x, P.y); pk.y = pack_bf16x2(P.z, P.w);
            *reinterpret_cast<uint2*>(w16 + 4 * i) = pk;
        }
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        adam_one(p[i], g[i], m[i], v[i], ema ? ema + i : nullptr, a, coef);
        if (w16) w16[i] = __float2bfloat16_rn(p[i]);
    }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
    const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 P = *reinterpret_cast<const float4*>(src + 4 * i);
        uint2 pk;
        pk.x = pack_bf16x2(P.x, P.y); pk.y = pack_bf16x2(P.z, P.w);
        *reinterpret_cast<uint2*>(dst + 4 * i) = pk;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}

static unsigned opt_grid(int64_t n) {
    int64_t b = ceil_div64(n / 4 + 1, 256);
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// sumsq[0] += sum_i g[i]^2   (fp64, caller-zeroed)
DMVAE_API int dmvae_grad_sumsq(const float* g, double* sumsq, int64_t n, void* stream) {
    DMVAE_CHECK_ARG(g && sumsq, "grad_sumsq: null pointer");
    DMVAE_CHECK_ARG(n >= 0 && ((uintptr_t)g & 15) == 0, "grad_sumsq: bad size or alignment");
    if (n == 0) return DMVAE_OK;
    sumsq_kernel<<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(g, sumsq, n);
    DMVAE_CHECK_LAUNCH("sumsq_kernel");
    return DMVAE_OK;
}

// One AdamW step on flat arenas.  sumsq (optional): device fp64 sum of squared gradients -> global-norm clip with
// max_norm (<= 0: no clip) and norm_out[0] = ||g||.  ema (optional): ema = decay*ema + (1-decay)*p_new.  step >= 1.
// w16 (optional): bf16 copy of the updated parameters, same element order.
DMVAE_API int dmvae_adamw_ema_step(float* p, float* g, float* m, float* v, float* ema, void* w16, const double* sumsq, float* norm_out,
                                   int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                   float max_norm, float ema_decay, void* stream) {
    DMVAE_CHECK_ARG(p && g && m && v, "adamw_ema_step: null pointer");
    DMVAE_CHECK_ARG(n >= 0 && step >= 1, "adamw_ema_step: bad size or step");
    DMVAE_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema | (uintptr_t)w16) & 15) == 0,
                    "adamw_ema_step: arenas must be 16-byte aligned");
    if (n == 0) return DMVAE_OK;
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.max_norm = max_norm; a.ema_decay = ema_decay;
    adamw_ema_kernel<<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, (bf16*)w16, sumsq, norm_out, n, a);
    DMVAE_CHECK_LAUNCH("adamw_ema_kernel");
    return DMVAE_OK;
}

// dst (bf16) = src (fp32), n elements: (re)build the bf16 copy of a parameter arena outside an optimizer step.
DMVAE_API int dmvae_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
    DMVAE_CHECK_ARG(src && dst, "cast_bf16: null pointer");
    DMVAE_CHECK_ARG(n >= 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "cast_bf16: bad size or alignment");
    if (n == 0) return DMVAE_OK;
    cast_bf16_kernel<<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
    DMVAE_CHECK_LAUNCH("cast_bf16_kernel");
    return DMVAE_OK;
}
