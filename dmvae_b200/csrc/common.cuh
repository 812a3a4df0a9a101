// Shared device/host helpers for libdmvae_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define DMVAE_OK 0
#define DMVAE_EINVAL (-1)
#define DMVAE_ECUDA (-2)
#define DMVAE_EUNSUPPORTED (-3)

#define DMVAE_F32 0
#define DMVAE_BF16 1

#define DMVAE_API extern "C" __attribute__((visibility("default")))

// error string shared by all translation units (defined in api.cu)
int dmvae_set_error(int code, const char* fmt, ...);

#define DMVAE_CHECK_ARG(cond, ...)                                   \
    do {                                                             \
        if (!(cond)) return dmvae_set_error(DMVAE_EINVAL, __VA_ARGS__); \
    } while (0)

#define DMVAE_CHECK_LAUNCH(name)                                                      \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess)                                                       \
            return dmvae_set_error(DMVAE_ECUDA, "%s: %s", name, cudaGetErrorString(e__)); \
    } while (0)

typedef __nv_bfloat16 bf16;

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p, int64_t i);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ld_as_float<bf16>(const bf16* p, int64_t i) { return __bfloat162float(p[i]); }

template <typename T> __device__ __forceinline__ void st_from_float(T* p, int64_t i, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, int64_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void st_from_float<bf16>(bf16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// round to the storage type's precision (identity for fp32): reproduces PyTorch's
// per-op rounding of bf16 elementwise chains.
template <typename T> __device__ __forceinline__ float rnd(float x);
template <> __device__ __forceinline__ float rnd<float>(float x) { return x; }
template <> __device__ __forceinline__ float rnd<bf16>(float x) { return bf16_round(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NV values per thread; result valid in thread 0. smem: NV*32 floats.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float x = lane < nw ? smem[i * 32 + lane] : 0.f;
            v[i] = warp_sum(x);
        }
    }
}

// 16-byte streaming loads / stores (read-once data: do not pollute L1)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 256-bit global accesses (sm_100): one full 32-byte sector per thread instead of two half-sector requests
__device__ __forceinline__ void ld_stream32(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void st_global32(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i]     = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}

// sigmoid via one MUFU.TANH: 0.5*tanh(0.5x)+0.5, abs. error ~1.2e-4 -- used only where the result is a gradient factor
// that is rounded to bf16 afterwards (GroupNorm/swish backward); the forward pass keeps the accurate form below.
__device__ __forceinline__ float sigmoidf_tanh(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}
// sigmoid via MUFU.EX2 + MUFU.RCP (rel. error ~1e-6; no IEEE division sequence)
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
