// HBM-bound fused loss kernels.
//
//   A3  DMD surrogate loss + latent gradient      reference: train_dmd.py:204-230,
//                                                 toy_example_2d/dmd.py:349-360,
//                                                 diffusion/transport/path.py:114-136
//   A4  L1 + L2 pixel loss fwd(+bwd)              reference: train_dmd.py:234-235
//   A5  LPIPS feature-distance reduction          reference: utils/lpips.py:86-94,156-161
//   A8  reparameterize + KL (extension, no reference implementation)
//
// All kernels: one pass over HBM with 16-byte loads, warp-shuffle reductions, per-block
// partials folded into fp64 accumulators with one atomic per block (order-insensitive at
// fp32 output precision).  Scalars never travel to the host inside a call.
#include "common.cuh"

// ------------------------------------------------------------------------------------
// A3: DMD
// ------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float type_max();
template <> __device__ __forceinline__ float type_max<float>() { return 3.402823466e+38f; }
template <> __device__ __forceinline__ float type_max<bf16>() { return 3.3895313892515355e+38f; }

template <typename T> __device__ __forceinline__ float nan_to_num(float g) {
    if (g != g) return 0.f;
    if (isinf(g)) return g > 0.f ? type_max<T>() : -type_max<T>();
    return g;
}

// 16-byte vector access for both storage types
template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
        const uint4 u = ld_stream16(p);
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ void store(float* p, const float (&f)[4]) {
        uint4 u = {__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])};
        st_stream16(p, u);
    }
};
template <> struct Vec<bf16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const bf16* p, float (&f)[8]) { unpack_bf16x8(ld_stream16(p), f); }
    static __device__ __forceinline__ void store(bf16* p, const float (&f)[8]) { st_stream16(p, pack_bf16x8(f)); }
};
// store VN floats as TD (TD may be wider than the input type)
template <typename TD, int VN> __device__ __forceinline__ void store_n(TD* p, const float (&f)[VN]);
template <> __device__ __forceinline__ void store_n<float, 4>(float* p, const float (&f)[4]) { Vec<float>::store(p, f); }
template <> __device__ __forceinline__ void store_n<bf16, 8>(bf16* p, const float (&f)[8]) { Vec<bf16>::store(p, f); }
template <> __device__ __forceinline__ void store_n<float, 8>(float* p, const float (&f)[8]) {
    const float a[4] = {f[0], f[1], f[2], f[3]}, b[4] = {f[4], f[5], f[6], f[7]};
    Vec<float>::store(p, a); Vec<float>::store(p + 4, b);
}

// xt = t*z + (1-t)*x0, rounded after every op like the reference's tensor-op chain
// (path.py:114-124 via compute_mu_t: alpha_t*x1 + sigma_t*x0, alpha=t, sigma=1-t).
// P % Vec<T>::N == 0 (host falls back to VN=1 semantics by passing P_vec) -- see launcher.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) dmd_mix_xt_kernel(const T* __restrict__ z, const T* __restrict__ x0,
                                                         const T* __restrict__ t, T* __restrict__ xt,
                                                         int64_t B, int64_t P) {
    constexpr int VN = VEC ? Vec<T>::N : 1;
    const int64_t n = B * P / VN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e0 = i * VN;
        const int64_t b = e0 / P;
        const float tb = ld_as_float(t, b);
        const float omt = rnd<T>(__fsub_rn(1.f, tb));
        float zf[VN], xf[VN], o[VN];
        if constexpr (VEC) { Vec<T>::load(z + e0, zf); Vec<T>::load(x0 + e0, xf); }
        else { zf[0] = ld_as_float(z, e0); xf[0] = ld_as_float(x0, e0); }
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            const float a = rnd<T>(__fmul_rn(tb, zf[k]));
            const float c = rnd<T>(__fmul_rn(omt, xf[k]));
            o[k] = rnd<T>(__fadd_rn(a, c));
        }
        if constexpr (VEC) Vec<T>::store(xt + e0, o);
        else st_from_float(xt, e0, o[0]);
    }
}

template <typename T>
__device__ __forceinline__ float cfg_mix(float vc, float vu, float cfg_m1) {
    // v + (s-1)*(v - v_uncond)    (train_dmd.py:216-217)
    const float d = rnd<T>(__fsub_rn(vc, vu));
    const float m = rnd<T>(__fmul_rn(cfg_m1, d));
    return rnd<T>(__fadd_rn(vc, m));
}

template <typename T>
__device__ __forceinline__ void dmd_point(float zf, float x, float vt, float vtu, float vs, float vsu, float omt,
                                          float cfg_m1, bool use_cfg, float& p_real, float& diff) {
    if (use_cfg) {
        vt = cfg_mix<T>(vt, vtu, cfg_m1);
        vs = cfg_mix<T>(vs, vsu, cfg_m1);
    }
    const float predT = rnd<T>(__fadd_rn(x, rnd<T>(__fmul_rn(vt, omt))));   // :218
    const float predS = rnd<T>(__fadd_rn(x, rnd<T>(__fmul_rn(vs, omt))));   // :219
    p_real = rnd<T>(__fsub_rn(zf, predT));                                  // :220
    const float p_student = rnd<T>(__fsub_rn(zf, predS));                   // :221
    diff = rnd<T>(__fsub_rn(p_real, p_student));                            // :223 numerator
}

template <typename T, bool VEC>
__device__ __forceinline__ void dmd_load(const T* __restrict__ z, const T* __restrict__ xt,
                                         const T* __restrict__ vTc, const T* __restrict__ vTu,
                                         const T* __restrict__ vSc, const T* __restrict__ vSu, int64_t e0,
                                         float omt, float cfg_m1, bool use_cfg,
                                         float (&zf)[VEC ? Vec<T>::N : 1], float (&pr)[VEC ? Vec<T>::N : 1],
                                         float (&df)[VEC ? Vec<T>::N : 1]) {
    constexpr int VN = VEC ? Vec<T>::N : 1;
    float x[VN], vt[VN], vs[VN], vtu[VN], vsu[VN];
    if constexpr (VEC) {
        Vec<T>::load(z + e0, zf); Vec<T>::load(xt + e0, x); Vec<T>::load(vTc + e0, vt); Vec<T>::load(vSc + e0, vs);
        if (use_cfg) { Vec<T>::load(vTu + e0, vtu); Vec<T>::load(vSu + e0, vsu); }
    } else {
        zf[0] = ld_as_float(z, e0); x[0] = ld_as_float(xt, e0); vt[0] = ld_as_float(vTc, e0); vs[0] = ld_as_float(vSc, e0);
        if (use_cfg) { vtu[0] = ld_as_float(vTu, e0); vsu[0] = ld_as_float(vSu, e0); }
    }
#pragma unroll
    for (int k = 0; k < VN; ++k)
        dmd_point<T>(zf[k], x[k], vt[k], use_cfg ? vtu[k] : 0.f, vs[k], use_cfg ? vsu[k] : 0.f, omt, cfg_m1, use_cfg, pr[k], df[k]);
}

// One CTA per sample: pass 1 builds p_real / (p_real - p_student) and the per-sample
// mean|p_real| normaliser (train_dmd.py:222); pass 2 forms grad, the MSE surrogate and dz.
// acc[0] += sum (z - target)^2 ; acc[1] += ||grad_b||_2
// CACHE keeps z and the numerator in shared memory so HBM is read exactly once.
template <typename T, typename TD, bool CACHE, bool VEC>
__global__ void __launch_bounds__(256) dmd_loss_kernel(
    const T* __restrict__ z, const T* __restrict__ xt, const T* __restrict__ t,
    const T* __restrict__ vTc, const T* __restrict__ vTu, const T* __restrict__ vSc, const T* __restrict__ vSu,
    TD* __restrict__ dz, double* __restrict__ acc, int64_t P, float cfg_m1, int use_cfg, int normalize,
    float dz_scale) {
    constexpr int VN = VEC ? Vec<T>::N : 1;
    // z and the numerator are values of type T (already rounded), so the cache holds them as T: 4 B/element for bf16
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    __shared__ float red[64];
    __shared__ float s_w;
    const int64_t b = blockIdx.x;
    const int64_t base = b * P;
    const float tb = ld_as_float(t, b);
    const float omt = rnd<T>(__fsub_rn(1.f, tb));
    T* s_z = reinterpret_cast<T*>(dyn_raw);
    T* s_d = s_z + (CACHE ? P : 0);
    const int64_t nv = P / VN;

    float v[1] = {0.f};
    for (int64_t j = threadIdx.x; j < nv; j += blockDim.x) {
        float zf[VN], pr[VN], df[VN];
        dmd_load<T, VEC>(z, xt, vTc, vTu, vSc, vSu, base + j * VN, omt, cfg_m1, use_cfg, zf, pr, df);
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            v[0] += fabsf(pr[k]);
            if (CACHE) { st_from_float(s_z, j * VN + k, zf[k]); st_from_float(s_d, j * VN + k, df[k]); }
        }
    }
    block_sum<1>(v, red);
    if (threadIdx.x == 0) s_w = rnd<T>(v[0] / (float)P);
    __syncthreads();
    const float w = s_w;

    float a[2] = {0.f, 0.f};
    for (int64_t j = threadIdx.x; j < nv; j += blockDim.x) {
        float zf[VN], df[VN], o[VN];
        if (CACHE) {
#pragma unroll
            for (int k = 0; k < VN; ++k) { zf[k] = ld_as_float(s_z, j * VN + k); df[k] = ld_as_float(s_d, j * VN + k); }
        } else {
            float pr[VN];
            dmd_load<T, VEC>(z, xt, vTc, vTu, vSc, vSu, base + j * VN, omt, cfg_m1, use_cfg, zf, pr, df);
        }
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            float g = normalize ? rnd<T>(__fdiv_rn(df[k], w)) : df[k];
            g = nan_to_num<T>(g);                                     // :224
            const float target = rnd<T>(__fsub_rn(zf[k], g));         // (latents - grad).detach()
            const float e = __fsub_rn(zf[k], target);                 // mse_loss runs in fp32
            a[0] += e * e;
            a[1] += g * g;
            o[k] = e * dz_scale;
        }
        if constexpr (VEC) store_n<TD, VN>(dz + base + j * VN, o);
        else st_from_float(dz, base + j, o[0]);
    }
    block_sum<2>(a, red);
    if (threadIdx.x == 0) {
        atomicAdd(&acc[0], (double)a[0]);
        atomicAdd(&acc[1], (double)sqrtf(a[1]));
    }
}

// ---- bf16 fast path: the same op-by-op bf16 arithmetic on packed pairs -------------------------------------------
// mul/add/sub.rn.bf16x2 round the exact result once.  That equals the reference's "compute in fp32, round to bf16"
// for these ops: a product of two bf16 values is exact in fp32, and a sum of two bf16 values is either exact in fp32
// or dominated by one operand beyond bf16 resolution.  ~3x fewer instructions than the scalar path, which is what
// the kernel was bound by (it issues ~90 instructions per element otherwise).
typedef __nv_bfloat162 bf2;
__device__ __forceinline__ bf2 u2b(uint32_t u) { return *reinterpret_cast<bf2*>(&u); }
__device__ __forceinline__ uint32_t b2u(bf2 v) { return *reinterpret_cast<uint32_t*>(&v); }

template <typename TD>
__global__ void __launch_bounds__(256) dmd_loss_bf16x2_kernel(
    const bf16* __restrict__ z, const bf16* __restrict__ xt, const bf16* __restrict__ t,
    const bf16* __restrict__ vTc, const bf16* __restrict__ vTu, const bf16* __restrict__ vSc, const bf16* __restrict__ vSu,
    TD* __restrict__ dz, double* __restrict__ acc, int64_t P, float cfg_m1, int use_cfg, int normalize, float dz_scale) {
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    __shared__ float red[64];
    __shared__ float s_w;
    uint4* s_z = reinterpret_cast<uint4*>(dyn_raw);
    uint4* s_d = s_z + P / 8;
    const int64_t b = blockIdx.x;
    const int64_t base = b * P;
    const float tb = __bfloat162float(t[b]);
    const bf16 omt1 = __float2bfloat16_rn(__fsub_rn(1.f, tb));
    const bf2 omt = __halves2bfloat162(omt1, omt1);
    const bf16 c1 = __float2bfloat16_rn(cfg_m1);
    const bf2 cm = __halves2bfloat162(c1, c1);
    const int nv = (int)(P / 8);

    float v[1] = {0.f};
    for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        const int64_t e0 = base + (int64_t)j * 8;
        const uint4 Z = ld_stream16(z + e0), X = ld_stream16(xt + e0), A = ld_stream16(vTc + e0), S = ld_stream16(vSc + e0);
        uint4 AU = make_uint4(0, 0, 0, 0), SU = AU;
        if (use_cfg) { AU = ld_stream16(vTu + e0); SU = ld_stream16(vSu + e0); }
        const uint32_t zq[4] = {Z.x, Z.y, Z.z, Z.w}, xq[4] = {X.x, X.y, X.z, X.w}, aq[4] = {A.x, A.y, A.z, A.w},
                       sq[4] = {S.x, S.y, S.z, S.w}, auq[4] = {AU.x, AU.y, AU.z, AU.w}, suq[4] = {SU.x, SU.y, SU.z, SU.w};
        uint32_t dq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bf2 vt = u2b(aq[q]), vs = u2b(sq[q]);
            if (use_cfg) {                                                   // v + (s-1)*(v - v_uncond)
                vt = __hadd2(vt, __hmul2(cm, __hsub2(vt, u2b(auq[q]))));
                vs = __hadd2(vs, __hmul2(cm, __hsub2(vs, u2b(suq[q]))));
            }
            const bf2 xx = u2b(xq[q]), zz = u2b(zq[q]);
            const bf2 pr = __hsub2(zz, __hadd2(xx, __hmul2(vt, omt)));       // z - (xt + v_T (1-t))
            const bf2 ps = __hsub2(zz, __hadd2(xx, __hmul2(vs, omt)));
            dq[q] = b2u(__hsub2(pr, ps));
            const bf2 ap = __habs2(pr);
            v[0] += __low2float(ap) + __high2float(ap);
        }
        s_z[j] = Z;
        s_d[j] = make_uint4(dq[0], dq[1], dq[2], dq[3]);
    }
    block_sum<1>(v, red);
    if (threadIdx.x == 0) s_w = bf16_round(v[0] / (float)P);
    __syncthreads();
    const float w = s_w;

    float a[2] = {0.f, 0.f};
    for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        const uint4 Z = s_z[j], D = s_d[j];
        const uint32_t zq[4] = {Z.x, Z.y, Z.z, Z.w}, dq[4] = {D.x, D.y, D.z, D.w};
        float o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float g0 = __low2float(u2b(dq[q])), g1 = __high2float(u2b(dq[q]));
            if (normalize) { g0 = bf16_round(__fdiv_rn(g0, w)); g1 = bf16_round(__fdiv_rn(g1, w)); }
            g0 = nan_to_num<bf16>(g0); g1 = nan_to_num<bf16>(g1);
            const bf2 zz = u2b(zq[q]);
            const bf2 tg = __hsub2(zz, __floats2bfloat162_rn(g0, g1));      // (latents - grad), bf16
            const float e0 = __fsub_rn(__low2float(zz), __low2float(tg)), e1 = __fsub_rn(__high2float(zz), __high2float(tg));
            a[0] += e0 * e0 + e1 * e1;
            a[1] += g0 * g0 + g1 * g1;
            o[2 * q] = e0 * dz_scale; o[2 * q + 1] = e1 * dz_scale;
        }
        store_n<TD, 8>(dz + base + (int64_t)j * 8, o);
    }
    block_sum<2>(a, red);
    if (threadIdx.x == 0) {
        atomicAdd(&acc[0], (double)a[0]);
        atomicAdd(&acc[1], (double)sqrtf(a[1]));
    }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <typename T, typename TD, bool VEC>
static int launch_dmd_v(const void* z, const void* xt, const void* t, const void* vTc, const void* vTu,
                        const void* vSc, const void* vSu, void* dz, double* acc, int64_t B, int64_t P,
                        float cfg_m1, int use_cfg, int normalize, float dz_scale, cudaStream_t st) {
    const size_t cache_bytes = (size_t)P * 2 * sizeof(T);
    if (cache_bytes <= 96 * 1024) {
        auto k = dmd_loss_kernel<T, TD, true, VEC>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        k<<<(unsigned)B, 256, cache_bytes, st>>>((const T*)z, (const T*)xt, (const T*)t, (const T*)vTc, (const T*)vTu,
                                                 (const T*)vSc, (const T*)vSu, (TD*)dz, acc, P, cfg_m1, use_cfg,
                                                 normalize, dz_scale);
    } else {
        dmd_loss_kernel<T, TD, false, VEC><<<(unsigned)B, 256, 0, st>>>(
            (const T*)z, (const T*)xt, (const T*)t, (const T*)vTc, (const T*)vTu, (const T*)vSc, (const T*)vSu,
            (TD*)dz, acc, P, cfg_m1, use_cfg, normalize, dz_scale);
    }
    DMVAE_CHECK_LAUNCH("dmd_loss_kernel");
    return DMVAE_OK;
}

template <typename T, typename TD>
static int launch_dmd(const void* z, const void* xt, const void* t, const void* vTc, const void* vTu,
                      const void* vSc, const void* vSu, void* dz, double* acc, int64_t B, int64_t P,
                      float cfg_scale, int normalize, float dz_scale, cudaStream_t st) {
    const int use_cfg = (cfg_scale > 1.f && vTu && vSu) ? 1 : 0;
    const float cfg_m1 = cfg_scale - 1.f;
    const bool vec = (P % Vec<T>::N == 0) && aligned16(z) && aligned16(xt) && aligned16(vTc) && aligned16(vSc) &&
                     aligned16(dz) && (!use_cfg || (aligned16(vTu) && aligned16(vSu)));
    if constexpr (sizeof(T) == 2) {
        // packed bf16x2 path: needs the CFG multiplier to be exactly representable in bf16 (4.0 for the default scale 5)
        const float c_r = __bfloat162float(__float2bfloat16_rn(cfg_m1));
        const size_t cache_bytes = (size_t)P * 4;
        if (vec && (!use_cfg || c_r == cfg_m1) && cache_bytes <= 96 * 1024) {
            auto k = dmd_loss_bf16x2_kernel<TD>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            k<<<(unsigned)B, 256, cache_bytes, st>>>((const bf16*)z, (const bf16*)xt, (const bf16*)t, (const bf16*)vTc, (const bf16*)vTu,
                                                     (const bf16*)vSc, (const bf16*)vSu, (TD*)dz, acc, P, cfg_m1, use_cfg, normalize, dz_scale);
            DMVAE_CHECK_LAUNCH("dmd_loss_bf16x2_kernel");
            return DMVAE_OK;
        }
    }
    return vec ? launch_dmd_v<T, TD, true>(z, xt, t, vTc, vTu, vSc, vSu, dz, acc, B, P, cfg_m1, use_cfg, normalize, dz_scale, st)
               : launch_dmd_v<T, TD, false>(z, xt, t, vTc, vTu, vSc, vSu, dz, acc, B, P, cfg_m1, use_cfg, normalize, dz_scale, st);
}

DMVAE_API int dmvae_dmd_mix_xt(const void* z, const void* x0, const void* t, void* xt, int64_t B,
                               int64_t per_sample, int dtype, void* stream) {
    DMVAE_CHECK_ARG(z && x0 && t && xt, "dmd_mix_xt: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && per_sample >= 0, "dmd_mix_xt: negative size");
    if (B * per_sample == 0) return DMVAE_OK;
    const int64_t n = B * per_sample;
    const unsigned grid = (unsigned)(ceil_div64(n, 256) < 148 * 16 ? ceil_div64(n, 256) : 148 * 16);
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(z) && aligned16(x0) && aligned16(xt);
    if (dtype == DMVAE_F32) {
        if (al && per_sample % 4 == 0)
            dmd_mix_xt_kernel<float, true><<<grid, 256, 0, st>>>((const float*)z, (const float*)x0, (const float*)t, (float*)xt, B, per_sample);
        else
            dmd_mix_xt_kernel<float, false><<<grid, 256, 0, st>>>((const float*)z, (const float*)x0, (const float*)t, (float*)xt, B, per_sample);
    } else if (dtype == DMVAE_BF16) {
        if (al && per_sample % 8 == 0)
            dmd_mix_xt_kernel<bf16, true><<<grid, 256, 0, st>>>((const bf16*)z, (const bf16*)x0, (const bf16*)t, (bf16*)xt, B, per_sample);
        else
            dmd_mix_xt_kernel<bf16, false><<<grid, 256, 0, st>>>((const bf16*)z, (const bf16*)x0, (const bf16*)t, (bf16*)xt, B, per_sample);
    } else
        return dmvae_set_error(DMVAE_EINVAL, "dmd_mix_xt: bad dtype %d", dtype);
    DMVAE_CHECK_LAUNCH("dmd_mix_xt_kernel");
    return DMVAE_OK;
}

DMVAE_API int dmvae_dmd_loss_fwd_bwd(const void* z, const void* xt, const void* t, const void* vT_c,
                                     const void* vT_u, const void* vS_c, const void* vS_u, void* dz,
                                     double* acc, int64_t B, int64_t per_sample, float cfg_scale,
                                     int normalize, float grad_scale, int dtype, int dz_dtype, void* stream) {
    DMVAE_CHECK_ARG(z && xt && t && vT_c && vS_c && dz && acc, "dmd_loss_fwd_bwd: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && per_sample >= 0, "dmd_loss_fwd_bwd: negative size");
    if (B * per_sample == 0) return DMVAE_OK;
    // loss = 0.5 * mean(e^2)  =>  dL/dz = e / N
    const float dz_scale = grad_scale / (float)((double)B * (double)per_sample);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DMVAE_F32 && dz_dtype == DMVAE_F32)
        return launch_dmd<float, float>(z, xt, t, vT_c, vT_u, vS_c, vS_u, dz, acc, B, per_sample, cfg_scale, normalize, dz_scale, st);
    if (dtype == DMVAE_BF16 && dz_dtype == DMVAE_F32)
        return launch_dmd<bf16, float>(z, xt, t, vT_c, vT_u, vS_c, vS_u, dz, acc, B, per_sample, cfg_scale, normalize, dz_scale, st);
    if (dtype == DMVAE_BF16 && dz_dtype == DMVAE_BF16)
        return launch_dmd<bf16, bf16>(z, xt, t, vT_c, vT_u, vS_c, vS_u, dz, acc, B, per_sample, cfg_scale, normalize, dz_scale, st);
    return dmvae_set_error(DMVAE_EINVAL, "dmd_loss_fwd_bwd: unsupported dtype pair (%d,%d)", dtype, dz_dtype);
}

// ------------------------------------------------------------------------------------
// A4: L1 + L2
// ------------------------------------------------------------------------------------
// acc[0] += sum|r-x| ; acc[1] += sum (r-x)^2 ; optional d_recon = (c1*sign(d) + c2*d)
template <bool WRITE_GRAD, bool DEV_COEF>
__global__ void __launch_bounds__(256) l1l2_kernel(const float* __restrict__ recon, const float* __restrict__ image,
                                                   float* __restrict__ d_recon, double* __restrict__ acc,
                                                   int64_t n, float c1, float c2, const float* __restrict__ g1,
                                                   const float* __restrict__ g2) {
    __shared__ float red[64];
    if (DEV_COEF) { c1 *= g1 ? *g1 : 0.f; c2 *= g2 ? *g2 : 0.f; }
    float a[2] = {0.f, 0.f};
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const uint4 ur = ld_stream16(recon + 4 * i), ui = ld_stream16(image + 4 * i);
        const float r[4] = {__uint_as_float(ur.x), __uint_as_float(ur.y), __uint_as_float(ur.z), __uint_as_float(ur.w)};
        const float x[4] = {__uint_as_float(ui.x), __uint_as_float(ui.y), __uint_as_float(ui.z), __uint_as_float(ui.w)};
        float g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d = __fsub_rn(r[k], x[k]);
            a[0] += fabsf(d);
            a[1] += d * d;
            const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
            g[k] = c1 * sgn + c2 * d;
        }
        if (WRITE_GRAD) {
            uint4 o = {__float_as_uint(g[0]), __float_as_uint(g[1]), __float_as_uint(g[2]), __float_as_uint(g[3])};
            st_stream16(d_recon + 4 * i, o);
        }
    }
    // ragged tail (n % 4)
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = __fsub_rn(recon[i], image[i]);
        a[0] += fabsf(d);
        a[1] += d * d;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        if (WRITE_GRAD) d_recon[i] = c1 * sgn + c2 * d;
    }
    if (acc) {
        block_sum<2>(a, red);
        if (threadIdx.x == 0) { atomicAdd(&acc[0], (double)a[0]); atomicAdd(&acc[1], (double)a[1]); }
    }
}

static unsigned stream_grid(int64_t n_vec) {
    int64_t g = ceil_div64(n_vec, 256 * 4);
    if (g < 1) g = 1;
    if (g > 148 * 8) g = 148 * 8;
    return (unsigned)g;
}

DMVAE_API int dmvae_l1l2_fwd(const float* recon, const float* image, double* acc, int64_t n, void* stream) {
    DMVAE_CHECK_ARG(recon && image && acc, "l1l2_fwd: null pointer");
    DMVAE_CHECK_ARG(n >= 0, "l1l2_fwd: negative size");
    DMVAE_CHECK_ARG(((uintptr_t)recon & 15) == 0 && ((uintptr_t)image & 15) == 0, "l1l2_fwd: inputs must be 16-byte aligned");
    if (n == 0) return DMVAE_OK;
    l1l2_kernel<false, false><<<stream_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(recon, image, nullptr, acc, n, 0.f, 0.f, nullptr, nullptr);
    DMVAE_CHECK_LAUNCH("l1l2_kernel<fwd>");
    return DMVAE_OK;
}

// d_recon = (w_l1 * g1 * sign(d) + 2 * w_l2 * g2 * d) / n ; g1/g2 are device scalars (upstream grads) or NULL (=1)
DMVAE_API int dmvae_l1l2_bwd(const float* recon, const float* image, float* d_recon, const float* g1,
                             const float* g2, int64_t n, float w_l1, float w_l2, void* stream) {
    DMVAE_CHECK_ARG(recon && image && d_recon, "l1l2_bwd: null pointer");
    DMVAE_CHECK_ARG(n >= 0, "l1l2_bwd: negative size");
    DMVAE_CHECK_ARG(((uintptr_t)recon & 15) == 0 && ((uintptr_t)image & 15) == 0 && ((uintptr_t)d_recon & 15) == 0,
                    "l1l2_bwd: buffers must be 16-byte aligned");
    if (n == 0) return DMVAE_OK;
    const float c1 = w_l1 / (float)n, c2 = 2.f * w_l2 / (float)n;
    l1l2_kernel<true, true><<<stream_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(recon, image, d_recon, nullptr, n, c1, c2, g1, g2);
    DMVAE_CHECK_LAUNCH("l1l2_kernel<bwd>");
    return DMVAE_OK;
}

// fused single pass: both means and d(w_l1*L1 + w_l2*L2)/d_recon (upstream grad folded into the weights)
DMVAE_API int dmvae_l1l2_fwd_bwd(const float* recon, const float* image, float* d_recon, double* acc,
                                 int64_t n, float w_l1, float w_l2, void* stream) {
    DMVAE_CHECK_ARG(recon && image && d_recon && acc, "l1l2_fwd_bwd: null pointer");
    DMVAE_CHECK_ARG(n >= 0, "l1l2_fwd_bwd: negative size");
    DMVAE_CHECK_ARG(((uintptr_t)recon & 15) == 0 && ((uintptr_t)image & 15) == 0 && ((uintptr_t)d_recon & 15) == 0,
                    "l1l2_fwd_bwd: buffers must be 16-byte aligned");
    if (n == 0) return DMVAE_OK;
    const float c1 = w_l1 / (float)n, c2 = 2.f * w_l2 / (float)n;
    l1l2_kernel<true, false><<<stream_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(recon, image, d_recon, acc, n, c1, c2, nullptr, nullptr);
    DMVAE_CHECK_LAUNCH("l1l2_kernel<fwd_bwd>");
    return DMVAE_OK;
}

// ------------------------------------------------------------------------------------
// A5: LPIPS feature distance on channels-last features  f[b][pixel][c]
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float seg_sum(float v, int lanes) {
    for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LP lanes cooperate on one pixel, NV 16-byte vectors per lane:  C = LP * NV * Vec<T>::N.
// FAITHFUL reproduces the bf16 rounding autocast applies to the 1x1 "lin" conv
// (utils/lpips.py:91: diffs and weights cast to bf16, per-pixel output rounded to bf16).
template <typename T, int NV, bool BWD, bool FAITHFUL>
__global__ void __launch_bounds__(256) lpips_dist_kernel(const T* __restrict__ f0, const T* __restrict__ f1,
                                                         const float* __restrict__ lin_w, T* __restrict__ df1,
                                                         double* __restrict__ acc, const float* __restrict__ gout,
                                                         int64_t HW, int C, int LP, float scale) {
    constexpr int VN = Vec<T>::N;
    __shared__ float red[32];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LP;                 // lane within the pixel group
    const int ppw = 32 / LP;                   // pixels per warp per iteration
    const int warps_per_block = blockDim.x >> 5;
    const int64_t pix0 = ((int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * ppw + lane / LP;
    const int64_t pstride = (int64_t)gridDim.x * warps_per_block * ppw;
    const float eps = 1e-10f;

    float w[NV][VN];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            const float wv = lin_w[(v * LP + sub) * VN + k];
            w[v][k] = FAITHFUL ? bf16_round(wv) : wv;
        }
    float g_up = scale;
    if (BWD && gout) g_up *= *gout;

    float total = 0.f;
    const int64_t iters = (HW + pstride - 1) / pstride;   // uniform trip count keeps shuffles converged
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t p = pix0 + it * pstride;
        const bool live = p < HW;
        const int64_t off = ((int64_t)b * HW + (live ? p : 0)) * C;
        float a0[NV][VN], a1[NV][VN];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            Vec<T>::load(f0 + off + (v * LP + sub) * VN, a0[v]);
            Vec<T>::load(f1 + off + (v * LP + sub) * VN, a1[v]);
#pragma unroll
            for (int k = 0; k < VN; ++k) { s0 += a0[v][k] * a0[v][k]; s1 += a1[v][k] * a1[v][k]; }
        }
        s0 = seg_sum(s0, LP); s1 = seg_sum(s1, LP);
        const float n0 = sqrtf(s0), n1 = sqrtf(s1);
        const float r0 = 1.f / (n0 + eps), r1 = 1.f / (n1 + eps);
        if (!BWD) {
            float d = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int k = 0; k < VN; ++k) {
                    const float u = a0[v][k] * r0 - a1[v][k] * r1;
                    float sq = u * u;
                    if (FAITHFUL) sq = bf16_round(sq);
                    d += w[v][k] * sq;
                }
            d = seg_sum(d, LP);
            if (FAITHFUL) d = bf16_round(d);
            if (live && sub == 0) total += d;
        } else {
            float g[NV][VN];
            float dot = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int k = 0; k < VN; ++k) {
                    const float u = a0[v][k] * r0 - a1[v][k] * r1;
                    g[v][k] = -2.f * w[v][k] * u * g_up;
                    dot += g[v][k] * a1[v][k];
                }
            dot = seg_sum(dot, LP);
            // d u1_k / d f1_c = delta/(n1+eps) - f1_k f1_c / (n1 (n1+eps)^2)
            const float c2 = n1 > 0.f ? dot * r1 * r1 / n1 : 0.f;
            if (live) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float o[VN];
#pragma unroll
                    for (int k = 0; k < VN; ++k) o[k] = g[v][k] * r1 - c2 * a1[v][k];
                    Vec<T>::store(df1 + off + (v * LP + sub) * VN, o);
                }
            }
        }
    }
    if (!BWD) {
        float v[1] = {total};
        block_sum<1>(v, red);
        if (threadIdx.x == 0) atomicAdd(&acc[b], (double)v[0]);
    }
}

template <typename T, bool BWD, bool FAITHFUL>
static int launch_lpips(const void* f0, const void* f1, const float* lin_w, void* df1, double* acc,
                        const float* gout, int64_t B, int64_t HW, int C, float scale, cudaStream_t st) {
    constexpr int VN = Vec<T>::N;
    const int nvec = C / VN;
    int LP = nvec < 32 ? nvec : 32;
    const int NV = nvec / LP;
    if (C % VN != 0 || (LP & (LP - 1)) != 0 || NV * LP != nvec || NV > 4 || NV == 3)
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "lpips_dist: unsupported channel count %d", C);
    const int ppw = 32 / LP;
    int64_t gx = ceil_div64(HW, (int64_t)8 * ppw * 4);
    int64_t cap = (148 * 8 + B - 1) / B;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)B);
#define LPIPS_LAUNCH(NVV)                                                                                          \
    lpips_dist_kernel<T, NVV, BWD, FAITHFUL><<<grid, 256, 0, st>>>((const T*)f0, (const T*)f1, lin_w, (T*)df1, acc, \
                                                                   gout, HW, C, LP, scale)
    switch (NV) {
        case 1: LPIPS_LAUNCH(1); break;
        case 2: LPIPS_LAUNCH(2); break;
        case 4: LPIPS_LAUNCH(4); break;
    }
#undef LPIPS_LAUNCH
    DMVAE_CHECK_LAUNCH("lpips_dist_kernel");
    return DMVAE_OK;
}

// acc[b] += sum_pixels sum_c w_c (f0^ - f1^)^2      (caller divides by HW and averages over b)
DMVAE_API int dmvae_lpips_dist_fwd(const void* f0, const void* f1, const float* lin_w, double* acc, int64_t B,
                                   int64_t HW, int C, int dtype, int faithful, void* stream) {
    DMVAE_CHECK_ARG(f0 && f1 && lin_w && acc, "lpips_dist_fwd: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && HW >= 0 && C > 0, "lpips_dist_fwd: bad shape");
    DMVAE_CHECK_ARG(((uintptr_t)f0 & 15) == 0 && ((uintptr_t)f1 & 15) == 0, "lpips_dist_fwd: features must be 16-byte aligned");
    if (B * HW == 0) return DMVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DMVAE_F32) return launch_lpips<float, false, false>(f0, f1, lin_w, nullptr, acc, nullptr, B, HW, C, 1.f, st);
    if (dtype == DMVAE_BF16)
        return faithful ? launch_lpips<bf16, false, true>(f0, f1, lin_w, nullptr, acc, nullptr, B, HW, C, 1.f, st)
                        : launch_lpips<bf16, false, false>(f0, f1, lin_w, nullptr, acc, nullptr, B, HW, C, 1.f, st);
    return dmvae_set_error(DMVAE_EINVAL, "lpips_dist_fwd: bad dtype %d", dtype);
}

// df1 = scale * (*gout) * d/df1 [ sum_c w_c (f0^ - f1^)^2 ]   (gradient flows to the second argument only)
DMVAE_API int dmvae_lpips_dist_bwd(const void* f0, const void* f1, const float* lin_w, void* df1,
                                   const float* gout, int64_t B, int64_t HW, int C, float scale, int dtype,
                                   void* stream) {
    DMVAE_CHECK_ARG(f0 && f1 && lin_w && df1, "lpips_dist_bwd: null pointer");
    DMVAE_CHECK_ARG(B >= 0 && HW >= 0 && C > 0, "lpips_dist_bwd: bad shape");
    DMVAE_CHECK_ARG(((uintptr_t)f0 & 15) == 0 && ((uintptr_t)f1 & 15) == 0 && ((uintptr_t)df1 & 15) == 0,
                    "lpips_dist_bwd: buffers must be 16-byte aligned");
    if (B * HW == 0) return DMVAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DMVAE_F32) return launch_lpips<float, true, false>(f0, f1, lin_w, df1, nullptr, gout, B, HW, C, scale, st);
    if (dtype == DMVAE_BF16) return launch_lpips<bf16, true, false>(f0, f1, lin_w, df1, nullptr, gout, B, HW, C, scale, st);
    return dmvae_set_error(DMVAE_EINVAL, "lpips_dist_bwd: bad dtype %d", dtype);
}

// ------------------------------------------------------------------------------------
// A8 (extension): reparameterize + KL(N(mu, e^lv) || N(0, I))
//   h rows hold [mu (half) | logvar (half)];  z = mu + exp(0.5 lv) * eps ;  acc[0] += 0.5*sum(mu^2 + e^lv - 1 - lv)
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) reparam_kl_fwd_kernel(const T* __restrict__ h, const T* __restrict__ eps,
                                                             T* __restrict__ z, double* __restrict__ acc,
                                                             int64_t rows, int64_t half) {
    __shared__ float red[32];
    const int64_t n = rows * half;
    float a[1] = {0.f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / half, j = i - r * half;
        const float mu = ld_as_float(h, r * 2 * half + j);
        const float lv = ld_as_float(h, r * 2 * half + half + j);
        const float e = ld_as_float(eps, i);
        const float sd = expf(0.5f * lv);
        st_from_float(z, i, mu + sd * e);
        a[0] += 0.5f * (mu * mu + sd * sd - 1.f - lv);
    }
    block_sum<1>(a, red);
    if (threadIdx.x == 0) atomicAdd(&acc[0], (double)a[0]);
}

// dh = [ dz + g*mu | dz*eps*0.5*sd + g*0.5*(sd^2 - 1) ],  g = kl_scale * (*g_kl)
template <typename T>
__global__ void __launch_bounds__(256) reparam_kl_bwd_kernel(const T* __restrict__ h, const T* __restrict__ eps,
                                                             const T* __restrict__ dz, T* __restrict__ dh,
                                                             const float* __restrict__ g_kl, float kl_scale,
                                                             int64_t rows, int64_t half) {
    const int64_t n = rows * half;
    const float g = kl_scale * (g_kl ? *g_kl : 1.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / half, j = i - r * half;
        const float mu = ld_as_float(h, r * 2 * half + j);
        const float lv = ld_as_float(h, r * 2 * half + half + j);
        const float e = ld_as_float(eps, i);
        const float d = dz ? ld_as_float(dz, i) : 0.f;
        const float sd = expf(0.5f * lv);
        st_from_float(dh, r * 2 * half + j, d + g * mu);
        st_from_float(dh, r * 2 * half + half + j, d * e * 0.5f * sd + g * 0.5f * (sd * sd - 1.f));
    }
}

// 16-byte vector variants (half % Vec<T>::N == 0, aligned): one thread handles Vec<T>::N consecutive latents of a row
template <typename T>
__global__ void __launch_bounds__(256) reparam_kl_fwd_vec_kernel(const T* __restrict__ h, const T* __restrict__ eps,
                                                                 T* __restrict__ z, double* __restrict__ acc,
                                                                 int64_t rows, int64_t half) {
    constexpr int VN = Vec<T>::N;
    __shared__ float red[32];
    const int64_t hv = half / VN, n = rows * hv;
    float a[1] = {0.f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / hv, j = (i - r * hv) * VN;
        float mu[VN], lv[VN], e[VN], o[VN];
        Vec<T>::load(h + r * 2 * half + j, mu);
        Vec<T>::load(h + r * 2 * half + half + j, lv);
        Vec<T>::load(eps + r * half + j, e);
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            const float sd = __expf(0.5f * lv[k]);
            o[k] = fmaf(sd, e[k], mu[k]);
            a[0] += 0.5f * (fmaf(mu[k], mu[k], sd * sd) - 1.f - lv[k]);
        }
        Vec<T>::store(z + r * half + j, o);
    }
    block_sum<1>(a, red);
    if (threadIdx.x == 0) atomicAdd(&acc[0], (double)a[0]);
}

template <typename T>
__global__ void __launch_bounds__(256) reparam_kl_bwd_vec_kernel(const T* __restrict__ h, const T* __restrict__ eps,
                                                                 const T* __restrict__ dz, T* __restrict__ dh,
                                                                 const float* __restrict__ g_kl, float kl_scale,
                                                                 int64_t rows, int64_t half) {
    constexpr int VN = Vec<T>::N;
    const int64_t hv = half / VN, n = rows * hv;
    const float g = kl_scale * (g_kl ? *g_kl : 1.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / hv, j = (i - r * hv) * VN;
        float mu[VN], lv[VN], e[VN], d[VN], o1[VN], o2[VN];
        Vec<T>::load(h + r * 2 * half + j, mu);
        Vec<T>::load(h + r * 2 * half + half + j, lv);
        Vec<T>::load(eps + r * half + j, e);
        if (dz) Vec<T>::load(dz + r * half + j, d);
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            const float dk = dz ? d[k] : 0.f;
            const float sd = __expf(0.5f * lv[k]);
            o1[k] = fmaf(g, mu[k], dk);
            o2[k] = dk * e[k] * 0.5f * sd + g * 0.5f * (sd * sd - 1.f);
        }
        Vec<T>::store(dh + r * 2 * half + j, o1);
        Vec<T>::store(dh + r * 2 * half + half + j, o2);
    }
}

DMVAE_API int dmvae_reparam_kl_fwd(const void* h, const void* eps, void* z, double* acc, int64_t rows,
                                   int64_t half, int dtype, void* stream) {
    DMVAE_CHECK_ARG(h && eps && z && acc, "reparam_kl_fwd: null pointer");
    DMVAE_CHECK_ARG(rows >= 0 && half >= 0, "reparam_kl_fwd: negative size");
    const int64_t n = rows * half;
    if (n == 0) return DMVAE_OK;
    const unsigned grid = stream_grid(n / 2);
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(h) && aligned16(eps) && aligned16(z);
    if (dtype == DMVAE_F32) {
        if (al && half % 4 == 0) reparam_kl_fwd_vec_kernel<float><<<stream_grid(n / 4), 256, 0, st>>>((const float*)h, (const float*)eps, (float*)z, acc, rows, half);
        else reparam_kl_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)h, (const float*)eps, (float*)z, acc, rows, half);
    } else if (dtype == DMVAE_BF16) {
        if (al && half % 8 == 0) reparam_kl_fwd_vec_kernel<bf16><<<stream_grid(n / 8), 256, 0, st>>>((const bf16*)h, (const bf16*)eps, (bf16*)z, acc, rows, half);
        else reparam_kl_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)h, (const bf16*)eps, (bf16*)z, acc, rows, half);
    } else
        return dmvae_set_error(DMVAE_EINVAL, "reparam_kl_fwd: bad dtype %d", dtype);
    DMVAE_CHECK_LAUNCH("reparam_kl_fwd_kernel");
    return DMVAE_OK;
}

DMVAE_API int dmvae_reparam_kl_bwd(const void* h, const void* eps, const void* dz, void* dh, const float* g_kl,
                                   float kl_scale, int64_t rows, int64_t half, int dtype, void* stream) {
    DMVAE_CHECK_ARG(h && eps && dh, "reparam_kl_bwd: null pointer");
    DMVAE_CHECK_ARG(rows >= 0 && half >= 0, "reparam_kl_bwd: negative size");
    const int64_t n = rows * half;
    if (n == 0) return DMVAE_OK;
    const unsigned grid = stream_grid(n / 2);
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(h) && aligned16(eps) && aligned16(dh) && aligned16(dz);
    if (dtype == DMVAE_F32) {
        if (al && half % 4 == 0) reparam_kl_bwd_vec_kernel<float><<<stream_grid(n / 4), 256, 0, st>>>((const float*)h, (const float*)eps, (const float*)dz, (float*)dh, g_kl, kl_scale, rows, half);
        else reparam_kl_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)h, (const float*)eps, (const float*)dz, (float*)dh, g_kl, kl_scale, rows, half);
    } else if (dtype == DMVAE_BF16) {
        if (al && half % 8 == 0) reparam_kl_bwd_vec_kernel<bf16><<<stream_grid(n / 8), 256, 0, st>>>((const bf16*)h, (const bf16*)eps, (const bf16*)dz, (bf16*)dh, g_kl, kl_scale, rows, half);
        else reparam_kl_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)h, (const bf16*)eps, (const bf16*)dz, (bf16*)dh, g_kl, kl_scale, rows, half);
    } else
        return dmvae_set_error(DMVAE_EINVAL, "reparam_kl_bwd: bad dtype %d", dtype);
    DMVAE_CHECK_LAUNCH("reparam_kl_bwd_kernel");
    return DMVAE_OK;
}
