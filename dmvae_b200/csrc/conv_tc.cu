// tcgen05 implicit-GEMM convolution for sm_100a: the conv tiles of flux_ae's ResnetBlock / Upsample / AttnBlock
// (models/flux_ae.py:32-35,63-67,101,210,237) and of LPIPS' frozen VGG16 (utils/lpips.py:116-153) in forward,
// data-gradient and weight-gradient form.
//
//   D[pixel, co] = sum_{tap, ci} X[pixel (+) tap, ci] * Wp[tap][co][ci]          M = B*H*W, N = Cout, K = taps*Cin
//
// Kernels in this file (dispatch: dmvae_conv_tc_fwd / dmvae_conv_tc_wgrad at the bottom):
//   conv_tc2h_kernel<BN>      3x3, Cout % 256 == 0: halo-resident CTA pair -- ONE (16+2)x(8+2)-pixel TMA halo per 64-channel
//                             chunk serves all nine taps (descriptor row offsets), cta_group::2, UMMA 256 x 256 x 16
//   conv_tcT_kernel           3x3, Cout = 64 / 128 / < 32: transposed tile, M = 128 channels, N = 256 pixels out of one halo
//   conv_tc2_kernel<BN>       per-tap operand fetch, CTA pair (1x1, stride 2, small images)
//   conv_tc_kernel<BN,MT>     per-tap operand fetch, single CTA, MT x 128 pixels x BN in {32,128,256}
//   conv_tc_wgrad2_kernel<MT> weight gradient, CTA pair sharing the x tile (Cin, Cout multiples of 256)
//   conv_tc_wgrad_kernel<..>  weight gradient, single CTA (remaining shapes)
// Common structure:
// * activations (channels-last bf16 [B][H][W][C]) are fetched by 4-D tiled TMA boxes; out-of-image coordinates are
//   zero-filled by the TMA unit, which is exactly the conv's zero padding.  No im2col buffer exists anywhere.
// * weights are tap-major packed [tap][Cout][Cin], fetched by 3-D TMA boxes {64, rows, 1}.
// * both land in shared memory in the 128-byte-swizzled layout tcgen05.mma consumes directly (K-major for forward /
//   dgrad, MN-major for wgrad).
// * one elected thread issues tcgen05.mma (bf16 x bf16 -> fp32) into TMEM accumulators; two accumulator buffers let the
//   epilogue of tile i overlap the main loop of tile i+1.
// * persistent CTAs (one per SM, or one pair per TPC) walk tiles round-robin.  Warp 0 = TMA producer, warp 1 = MMA issuer
//   (+TMEM alloc), the remaining warps = epilogue (tcgen05.ld -> +bias -> bf16 [-> +residual] -> global stores).
//
// dgrad of a stride-1 "same" conv is the same kernel fed dY and the flipped/transposed weight pack.
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>

namespace {

// kernel ids reported by dmvae_conv_tc_last_kernel (include/dmvae_b200.h)
constexpr int DMVAE_KERNEL_CONV_TC = 1, DMVAE_KERNEL_CONV_TC2 = 2, DMVAE_KERNEL_CONV_TC2H = 3, DMVAE_KERNEL_CONV_TCT = 4,
              DMVAE_KERNEL_WGRAD = 5, DMVAE_KERNEL_WGRAD2 = 6;

constexpr int BM = 128;        // pixels per tile (UMMA M)
constexpr int BK = 64;         // channels per pipeline stage (one 128-byte swizzle row)
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---------------------------------------------------------------- epilogue helpers
// GroupNorm statistics of the tile being stored (sum, sum of squares per (image, group)) are reduced here -- per thread
// over its 32 channels, across the warp's 32 pixels by shuffles, then one fp64 atomic per group -- so the following
// GroupNorm needs no separate statistics pass over the activation (2 B/element of HBM traffic saved per GN).
template <int CPG>   // channels per group inside a 32-channel chunk (CPG = 32 means the chunk lies in one group)
__device__ __forceinline__ void chunk_group_stats(const float (&r)[32], double* __restrict__ st, int g0, int lane) {
    constexpr int NG = 32 / CPG;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
        float sm = 0.f, sq = 0.f;
#pragma unroll
        for (int e = 0; e < CPG; ++e) { const float x = r[gi * CPG + e]; sm += x; sq = fmaf(x, x, sq); }
        sm = warp_sum(sm);
        sq = warp_sum(sq);
        if (lane == 0) { atomicAdd(&st[(g0 + gi) * 2], (double)sm); atomicAdd(&st[(g0 + gi) * 2 + 1], (double)sq); }
    }
}

// Epilogue flags (dmvae_conv_tc_fwd `flags`): EPI_RELU = y = max(conv + bias, 0) (VGG16's conv -> ReLU, utils/lpips.py:116-153);
// EPI_MASK = the `residual` tensor is not added but gates the output, y = residual > 0 ? y : 0 -- the backward of the ReLU that
// produced this (data-gradient) conv's input, applied where the gradient is written instead of in a separate pass.
constexpr int EPI_RELU = 1, EPI_MASK = 2;

// One 32-column chunk of one accumulator row: + bias, bf16 rounding, optional residual add, 64-byte store, optional stats.
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], int n, int64_t pix, const float* __restrict__ bias,
                                               const bf16* __restrict__ res, bf16* __restrict__ out, int Cout,
                                               double* __restrict__ st /* stats of this image or null */, int cpg, int lane,
                                               int flags = 0) {
    bf16* op = out + pix * Cout + n;
    const bf16* rp = res ? res + pix * Cout + n : nullptr;
    if (n + 32 <= Cout && (Cout & 7) == 0) {
        float r[32];
        // rows 32-byte aligned: 256-bit accesses (one full sector per thread instead of two half-sector requests)
        const bool wide = (Cout & 15) == 0 && (((uintptr_t)out | (uintptr_t)res) & 31) == 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint4 pk[2], rr[2];
            if (rp) {
                if (wide) ld_stream32(rp + h * 16, rr[0], rr[1]);
                else { rr[0] = *reinterpret_cast<const uint4*>(rp + h * 16); rr[1] = *reinterpret_cast<const uint4*>(rp + h * 16 + 8); }
            }
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
                const int q = h * 2 + qq;
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]) + (bias ? __ldg(bias + n + q * 8 + e) : 0.f);
                if (flags & EPI_RELU) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                }
                if (rp) {
                    float fr[8];
                    unpack_bf16x8(rr[qq], fr);
                    if (flags & EPI_MASK) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = fr[e] > 0.f ? f[e] : 0.f;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = bf16_round(f[e]) + fr[e];   // bf16 conv output, then bf16 add
                    }
                }
                pk[qq] = pack_bf16x8(f);
                if (st) {
                    float fo[8];
                    unpack_bf16x8(pk[qq], fo);
#pragma unroll
                    for (int e = 0; e < 8; ++e) r[q * 8 + e] = fo[e];
                }
            }
            if (wide) st_global32(op + h * 16, pk[0], pk[1]);
            else { *reinterpret_cast<uint4*>(op + h * 16) = pk[0]; *reinterpret_cast<uint4*>(op + h * 16 + 8) = pk[1]; }
        }
        if (st) {
            switch (cpg) {
                case 1: chunk_group_stats<1>(r, st, n, lane); break;
                case 2: chunk_group_stats<2>(r, st, n / 2, lane); break;
                case 4: chunk_group_stats<4>(r, st, n / 4, lane); break;
                case 8: chunk_group_stats<8>(r, st, n / 8, lane); break;
                case 16: chunk_group_stats<16>(r, st, n / 16, lane); break;
                default: chunk_group_stats<32>(r, st, n / cpg, lane); break;
            }
        }
    } else {
        for (int e = 0; e < 32 && n + e < Cout; ++e) {
            float f = __uint_as_float(v[e]) + (bias ? __ldg(bias + n + e) : 0.f);
            if (flags & EPI_RELU) f = fmaxf(f, 0.f);
            if (rp) {
                const float fr = __bfloat162float(rp[e]);
                f = (flags & EPI_MASK) ? (fr > 0.f ? f : 0.f) : bf16_round(f) + fr;
            }
            op[e] = __float2bfloat16_rn(f);
        }
    }
}

// Same as epilogue_chunk for full 32-channel chunks, with the residual values already in registers (prefetched by the caller
// while the tile's main loop was still running / while the previous chunk was being stored).
__device__ __forceinline__ void epilogue_chunk_pre(const uint32_t (&v)[32], int n, int64_t pix, const float* __restrict__ bias,
                                                   const uint4 (&rv)[4], bool has_res, bf16* __restrict__ out, int Cout,
                                                   double* __restrict__ st, int cpg, int lane, int flags = 0) {
    bf16* op = out + pix * Cout + n;
    float r[32];
    const bool wide = (Cout & 15) == 0 && ((uintptr_t)out & 31) == 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint4 pk[2];
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int q = h * 2 + qq;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]) + (bias ? __ldg(bias + n + q * 8 + e) : 0.f);
            if (flags & EPI_RELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            if (has_res) {
                float fr[8];
                unpack_bf16x8(rv[q], fr);
                if (flags & EPI_MASK) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = fr[e] > 0.f ? f[e] : 0.f;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = bf16_round(f[e]) + fr[e];
                }
            }
            pk[qq] = pack_bf16x8(f);
            if (st) {
                float fo[8];
                unpack_bf16x8(pk[qq], fo);
#pragma unroll
                for (int e = 0; e < 8; ++e) r[q * 8 + e] = fo[e];
            }
        }
        if (wide) st_global32(op + h * 16, pk[0], pk[1]);
        else { *reinterpret_cast<uint4*>(op + h * 16) = pk[0]; *reinterpret_cast<uint4*>(op + h * 16 + 8) = pk[1]; }
    }
    if (st) {
        switch (cpg) {
            case 1: chunk_group_stats<1>(r, st, n, lane); break;
            case 2: chunk_group_stats<2>(r, st, n / 2, lane); break;
            case 4: chunk_group_stats<4>(r, st, n / 4, lane); break;
            case 8: chunk_group_stats<8>(r, st, n / 8, lane); break;
            case 16: chunk_group_stats<16>(r, st, n / 16, lane); break;
            default: chunk_group_stats<32>(r, st, n / cpg, lane); break;
        }
    }
}

// Line-coalesced epilogue for one warp: 32 accumulator rows (pixels) x 64 columns (channels).
// tcgen05.ld hands every thread ITS pixel's 32 consecutive channels, so direct stores touch 32 different 128-byte lines per
// instruction (ncu on the 1x1 layers: l1tex 66 % busy, tensor pipe 8 %, 12 k cycles per tile in the epilogue).  Here the block
// is first written to a 4 KB shared-memory patch (row = one pixel's 128 bytes, 16-byte chunks XOR-swizzled by the row so both
// directions are conflict-free) and read back so that 8 consecutive lanes cover one pixel's full line: 4 lines per instruction,
// for the residual loads as well.  The residual block is fetched before the accumulator is read, so its latency is hidden.
struct EpiGeom {
    int64_t tile_pix0;     // pixel index of tile row 0
    int W, bw_shift, bw_mask, Cout;
    int pix_step;          // output pixels between horizontally / vertically adjacent tile pixels (1; 2 for the sub-pixel upsampling conv,
};                         // whose tile is one phase of the 2x finer output grid -- W is then the OUTPUT image width)
// row offsets (elements, relative to tile_pix0 * Cout) of the 8 pixel rows this lane touches in the read-back phase
__device__ __forceinline__ void epilogue_row_offsets(int row0, const EpiGeom& eg, int lane, int (&off)[8]) {
    const int rsub = lane >> 3;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int rr = row0 + it * 4 + rsub;
        off[it] = ((rr >> eg.bw_shift) * eg.W + (rr & eg.bw_mask)) * eg.pix_step * eg.Cout + (lane & 7) * 8;
    }
}
// the residual values of one 32-row x 64-channel block, line-coalesced (8 lanes = one pixel's 128 bytes)
__device__ __forceinline__ void epilogue_load_residual(const bf16* __restrict__ res_tile, int n, const int (&off)[8], uint4 (&rv)[8]) {
#pragma unroll
    for (int it = 0; it < 8; ++it) rv[it] = ld_stream16(res_tile + off[it] + n);
}
// GroupNorm statistics in the line-coalesced epilogue.  In the read-back phase a lane owns 8 consecutive channels of 8 pixel rows,
// and for the decoder's widths those 8 channels are whole groups (C = 128: two groups of 4, C = 256: one group of 8) or half of one
// (C = 512: 16 channels = two adjacent lanes), so the per-group {sum, sum of squares} cost 8 FADD/FFMA per 16-byte vector, two or
// three shuffles per warp and block, and one fp64 atomic pair from 4..16 lanes -- instead of the 10 shuffles per group and
// 32-channel chunk of the per-thread-row form (chunk_group_stats), which made fused statistics a loss for narrow groups.
__device__ __forceinline__ bool line_stats_ok(int cpg) { return cpg == 4 || cpg == 8 || cpg == 16; }
__device__ __forceinline__ void block64_group_stats(const float (&gs)[8], const float (&gq)[8], double* __restrict__ st, int n,
                                                    int cpg, int lane) {
    const int c = lane & 7;                              // this lane's channels: n + 8c .. n + 8c + 7
    float s0 = (gs[0] + gs[1]) + (gs[2] + gs[3]), s1 = (gs[4] + gs[5]) + (gs[6] + gs[7]);
    float q0 = (gq[0] + gq[1]) + (gq[2] + gq[3]), q1 = (gq[4] + gq[5]) + (gq[6] + gq[7]);
    if (cpg != 4) { s0 += s1; q0 += q1; }
    if (cpg == 16) { s0 += __shfl_xor_sync(0xffffffffu, s0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 1); }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {                  // the four lanes that hold the same channels of different rows
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        if (cpg == 4) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o); }
    }
    if (lane < 8) {
        const int ch = n + 8 * c;
        if (cpg == 4) {
            atomicAdd(&st[(ch >> 2) * 2], (double)s0); atomicAdd(&st[(ch >> 2) * 2 + 1], (double)q0);
            atomicAdd(&st[((ch >> 2) + 1) * 2], (double)s1); atomicAdd(&st[((ch >> 2) + 1) * 2 + 1], (double)q1);
        } else if (cpg == 8) {
            atomicAdd(&st[(ch >> 3) * 2], (double)s0); atomicAdd(&st[(ch >> 3) * 2 + 1], (double)q0);
        } else if ((c & 1) == 0) {
            atomicAdd(&st[(ch >> 4) * 2], (double)s0); atomicAdd(&st[(ch >> 4) * 2 + 1], (double)q0);
        }
    }
}

__device__ __forceinline__ void epilogue_block64(uint32_t taddr, int n, const int (&off)[8], const float* __restrict__ bias,
                                                 const uint4 (&rv)[8], bool has_res, bf16* __restrict__ out_tile,
                                                 uint4* __restrict__ patch, int lane, int flags = 0,
                                                 double* __restrict__ st = nullptr, int cpg = 0) {
    const int c = lane & 7, rsub = lane >> 3;
    float gs[8] = {0, 0, 0, 0, 0, 0, 0, 0}, gq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(taddr + half * 32, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float f[8];
            if (bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n + half * 32 + q * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n + half * 32 + q * 8 + 4));
                f[0] = __uint_as_float(v[q * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[q * 8 + 1]) + b0.y;
                f[2] = __uint_as_float(v[q * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[q * 8 + 3]) + b0.w;
                f[4] = __uint_as_float(v[q * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[q * 8 + 5]) + b1.y;
                f[6] = __uint_as_float(v[q * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[q * 8 + 7]) + b1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
            }
            if (flags & EPI_RELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            patch[lane * 8 + ((half * 4 + q) ^ (lane & 7))] = pack_bf16x8(f);
        }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + rsub;
        uint4 pk = patch[r * 8 + (c ^ (r & 7))];
        if (has_res) {
            float f[8], fr[8];
            unpack_bf16x8(pk, f);
            unpack_bf16x8(rv[it], fr);
            if (flags & EPI_MASK) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fr[e] > 0.f ? f[e] : 0.f;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += fr[e];           // bf16 conv output + bf16 residual, one rounding
            }
            pk = pack_bf16x8(f);
        }
        *reinterpret_cast<uint4*>(out_tile + off[it] + n) = pk;
        if (st) {                                                     // statistics of the values as stored (bf16)
            float fo[8];
            unpack_bf16x8(pk, fo);
#pragma unroll
            for (int e = 0; e < 8; ++e) { gs[e] += fo[e]; gq[e] = fmaf(fo[e], fo[e], gq[e]); }
        }
    }
    if (st) block64_group_stats(gs, gq, st, n, cpg, lane);
    __syncwarp();
}
// all 64-channel blocks of one warp's 32 rows: the residual of block i+1 is in flight while block i is processed
template <int BN>
__device__ __forceinline__ void epilogue_rows(uint32_t taddr, int n0, int n_end, const int (&off)[8], const float* __restrict__ bias,
                                              const bf16* __restrict__ res_tile, uint4 (&rv)[8], bf16* __restrict__ out_tile,
                                              uint4* __restrict__ patch, int lane, int flags = 0,
                                              double* __restrict__ st = nullptr, int cpg = 0) {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 64) {
        if (n0 + c0 >= n_end) break;
        uint4 rn[8];
        const bool more = res_tile && c0 + 64 < BN && n0 + c0 + 64 < n_end;
        if (more) epilogue_load_residual(res_tile, n0 + c0 + 64, off, rn);
        epilogue_block64(taddr + c0, n0 + c0, off, bias, rv, res_tile != nullptr, out_tile, patch, lane, flags, st, cpg);
        if (more) {
#pragma unroll
            for (int it = 0; it < 8; ++it) rv[it] = rn[it];
        }
    }
}

struct TcGeom {
    int B, H, W, Cin, Cout, KH, KW, pt, pl;      // H, W: OUTPUT image size (tiles live on the output grid)
    int stride, IH, IW;                          // conv stride and INPUT image size (IH = H, IW = W when stride == 1)
    int cpg;               // Cout / 32 (GroupNorm group width) when output statistics are requested
    int flags;             // EPI_RELU | EPI_MASK
    int up;                // halo pair kernel only.  1 = sub-pixel form of nearest-2x + 3x3 (flux_ae.Upsample), forward: KH = KW = 2, the
                           // tile list runs over the 4 output phases (py, px), phase p reads taps 4p..4p+3 of a 16-tap pack and its
                           // halo starts at (-1 + py, -1 + px); outputs land on every other pixel of the 2H x 2W image.
                           // 2 = its data gradient: all 4 phases are accumulated into one tile of the H x W input gradient, the A
                           // operand being the phase's pixels of the 2H x 2W output gradient (TMA element stride 2, start (-py, -px)).
    int BW, BH;            // pixel tile = BH rows x BW cols of one image (BH*BW = 128)
    int tiles_w, tiles_h;  // W/BW, H/BH
    int m_tiles, n_tiles, k_chunks;
};

// MT = 128-pixel sub-tiles per CTA.  MT = 2 halves the L2->SM operand traffic per FLOP (one weight tile feeds two
// pixel tiles: 64 B/cycle/SM instead of 94-128), which is what bounds the 128-pixel tile (ncu: tensor pipe 45 % active).
template <int BN, int MT> struct Cfg {
    static constexpr int A_BYTES = MT * BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int ACC_COLS = MT * BN;                             // one accumulator set
    static constexpr int NBUF = (2 * ACC_COLS <= 512) ? 2 : 1;           // double-buffer when TMEM allows
    static constexpr uint32_t TMEM_COLS = (NBUF * ACC_COLS < 32) ? 32 : NBUF * ACC_COLS;
    static constexpr int EPI_WARPS = 4 * MT;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

template <int BN, int MT>
__global__ void __launch_bounds__(Cfg<BN, MT>::THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const float* __restrict__ bias, const bf16* __restrict__ res, bf16* __restrict__ out,
               double* __restrict__ stats, TcGeom g) {
    using C = Cfg<BN, MT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + C::STAGES;          // [STAGES]
    uint64_t* tfull = bars + 2 * C::STAGES;      // [2]
    uint64_t* tempty = bars + 2 * C::STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = g.m_tiles * g.n_tiles;
    const int k_iters = g.KH * g.KW * g.k_chunks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], C::EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile / g.n_tiles, nt = tile % g.n_tiles;
                const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, b = mt / (g.tiles_w * g.tiles_h);
                const int w0 = tw * g.BW * g.stride - g.pl, h0 = th * g.BH * g.stride - g.pt, n0 = nt * BN;
                for (int tap = 0; tap < g.KH * g.KW; ++tap) {
                    const int kh = tap / g.KW, kw = tap % g.KW;
                    for (int kc = 0; kc < g.k_chunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * C::STAGE_BYTES;
                        uint8_t* sb = sa + C::A_BYTES;
                        mbar_expect_tx(&full[stage], C::STAGE_BYTES);
                        tma_load_4d(sa, &map_a, &full[stage], kc * BK, w0 + kw, h0 + kh, b);   // MT*128 pixels x 64 ch
                        tma_load_3d(sb, &map_b, &full[stage], kc * BK, n0, tap);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int buf = (C::NBUF == 2) ? (it & 1) : 0;
                const uint32_t par = (C::NBUF == 2) ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(&tempty[buf], par ^ 1);                   // epilogue has drained this accumulator set
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * C::ACC_COLS;
                for (int k = 0; k < k_iters; ++k) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t bdesc = make_kmajor_sw128_desc(sa + C::A_BYTES);
#pragma unroll
                    for (int sub = 0; sub < MT; ++sub) {
                        const uint64_t adesc = make_kmajor_sw128_desc(sa + sub * (BM * BK * 2));
#pragma unroll
                        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                            // advance 16 elements (32 B) along K inside the swizzled row: +2 in 16-byte units
                            umma_bf16(d_tmem + sub * BN, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0);
                        }
                    }
                    umma_commit(&empty[stage]);                     // frees the smem slot when these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[buf]);                           // accumulators complete -> epilogue
            }
        }
    } else {
        // ============================== epilogue (warps 2 .. 2+4*MT) ==============================
        const int lg = warp & 3;                 // TMEM lane group this warp may access: lanes [32*lg, 32*lg+32)
        const int sub = (warp - 2) >> 2;         // which 128-pixel sub-tile
        const int r = sub * BM + lg * 32 + lane; // row of the tile = pixel index inside the tile
        const int dh = r / g.BW, dw = r % g.BW;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int buf = (C::NBUF == 2) ? (it & 1) : 0;
            const uint32_t par = (C::NBUF == 2) ? ((it >> 1) & 1) : (it & 1);
            const int mt = tile / g.n_tiles, nt = tile % g.n_tiles;
            const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, b = mt / (g.tiles_w * g.tiles_h);
            const int64_t pix = ((int64_t)b * g.H + th * g.BH + dh) * g.W + tw * g.BW + dw;
            const int n0 = nt * BN;
            mbar_wait(&tfull[buf], par);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * C::ACC_COLS + sub * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                const int n = n0 + c0;
                if (n >= g.Cout) continue;                            // warp-uniform
                epilogue_chunk(v, n, pix, bias, res, out, g.Cout, stats ? stats + (int64_t)b * 64 : nullptr, g.cpg, lane, g.flags);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------- CTA-pair variant (cta_group::2)
// Two CTAs of a cluster (one TPC) cooperate on a 256-pixel x BN tile: each loads its own 128-pixel A tile and HALF of
// the weight tile; the leader issues tcgen05.mma.cta_group::2 (UMMA M = 256), which reads A from each CTA's own shared
// memory and the two B halves from both.  Per SM this halves the weight bytes pulled from L2 and read from shared
// memory (64 B/cycle of tensor-core reads instead of 96), which is what bounds the single-CTA tile.
//   * TMA loads in both CTAs complete on the LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2).
//   * tcgen05.commit ... multicast::cluster releases the smem slot / publishes the accumulator in BOTH CTAs.
//   * each CTA's epilogue drains its own 128 TMEM lanes; the peer arrives remotely on the leader's tmem-empty barrier.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m256(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN> struct Cfg2 {
    static constexpr int A_BYTES = BM * BK * 2;                 // this CTA's 128 pixels
    static constexpr int B_BYTES = (BN / 2) * BK * 2;           // this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 4 * 4096;     // + one 4 KB store patch per epilogue warp
    static constexpr uint32_t TMEM_COLS = 2 * BN;               // two accumulator buffers of BN columns
    static constexpr int THREADS = 192;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const float* __restrict__ bias, const bf16* __restrict__ res, bf16* __restrict__ out,
                double* __restrict__ stats, TcGeom g) {
    using C = Cfg2<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;                       // [STAGES]  (the leader's copy is the live one)
    uint64_t* empty = bars + C::STAGES;          // [STAGES]
    uint64_t* tfull = bars + 2 * C::STAGES;      // [2]
    uint64_t* tempty = bars + 2 * C::STAGES + 2; // [2]       (leader's copy is the live one)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int pair_tiles = (g.m_tiles >> 1) * g.n_tiles;
    const int k_iters = g.KH * g.KW * g.k_chunks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();             // CTA-level ordering of the barrier inits / TMEM slot (the cluster barrier below subsumes it;
    cluster_sync_all();          // kept explicit so compute-sanitizer racecheck sees the dependency)
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer (both CTAs) ==============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters) {
                const int mt = 2 * (pt / g.n_tiles) + (int)rank, nt = pt % g.n_tiles;
                const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, b = mt / (g.tiles_w * g.tiles_h);
                const int w0 = tw * g.BW * g.stride - g.pl, h0 = th * g.BH * g.stride - g.pt, n0 = nt * BN + (int)rank * (BN / 2);
                for (int tap = 0; tap < g.KH * g.KW; ++tap) {
                    const int kh = tap / g.KW, kw = tap % g.KW;
                    for (int kc = 0; kc < g.k_chunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * C::STAGE_BYTES;
                        uint8_t* sb = sa + C::A_BYTES;
                        if (leader) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);      // bytes of both CTAs
                        const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
                        tma_load_4d_2sm(sa, &map_a, lbar, kc * BK, w0 + kw, h0 + kh, b);
                        tma_load_3d_2sm(sb, &map_b, lbar, kc * BK, n0, tap);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (leader only) ==============================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_m256(BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);      // both CTAs' epilogues drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int k = 0; k < k_iters; ++k) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t adesc = make_kmajor_sw128_desc(sa);
                    const uint64_t bdesc = make_kmajor_sw128_desc(sa + C::A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk)
                        umma_bf16_2sm(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0);
                    umma_commit_2sm(&empty[stage]);                 // frees the slot in both CTAs
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tfull[buf]);                       // accumulator ready in both CTAs
            }
        }
    } else {
        // ============================== epilogue (warps 2..5, both CTAs) ==============================
        const int lg = warp & 3;
        const int r = lg * 32 + lane;
        const int dh = r / g.BW, dw = r % g.BW;
        int it = 0;
        for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++it) {
            const int buf = it & 1;
            const int mt = 2 * (pt / g.n_tiles) + (int)rank, nt = pt % g.n_tiles;
            const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, b = mt / (g.tiles_w * g.tiles_h);
            const int64_t pix = ((int64_t)b * g.H + th * g.BH + dh) * g.W + tw * g.BW + dw;
            const int n0 = nt * BN;
            mbar_wait(&tfull[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * BN;
            if ((!stats || line_stats_ok(g.cpg)) && (g.Cout & 63) == 0) {
                EpiGeom eg;
                eg.tile_pix0 = ((int64_t)b * g.H + th * g.BH) * g.W + tw * g.BW;
                eg.W = g.W; eg.bw_shift = 31 - __clz(g.BW); eg.bw_mask = g.BW - 1; eg.Cout = g.Cout; eg.pix_step = 1;
                uint4* patch = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(bars) + 256) + (warp - 2) * 256;
                int off[8];
                epilogue_row_offsets(lg * 32, eg, lane, off);
                const bf16* res_tile = res ? res + eg.tile_pix0 * g.Cout : nullptr;
                uint4 rv[8];
                if (res_tile) epilogue_load_residual(res_tile, n0, off, rv);
                epilogue_rows<BN>(taddr, n0, g.Cout, off, bias, res_tile, rv, out + eg.tile_pix0 * g.Cout, patch, lane, g.flags,
                                  stats ? stats + (int64_t)b * 64 : nullptr, g.cpg);
            } else {
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    const int n = n0 + c0;
                    if (n >= g.Cout) continue;
                    epilogue_chunk(v, n, pix, bias, res, out, g.Cout, stats ? stats + (int64_t)b * 64 : nullptr, g.cpg, lane, g.flags);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------- halo-resident CTA-pair variant (3x3 and larger filters)
// tcgen05 applies the 128-byte swizzle to ABSOLUTE shared-memory address bits (probed on B200: a K-major SW128 operand may
// start at any 128-byte row of a TMA-written tile and use a stride-byte-offset that is not a multiple of 1024 B --
// scripts/experiments/umma_rowshift_probe.cu).  So ONE (BH+KH-1) x (BW+KW-1) pixel halo of 64 channels, fetched once, serves
// all KH*KW filter taps: tap (kh, kw) is the same tile read from row (kh*HP + kw) on, with the 8-pixel groups HP*128 B apart
// (pixel tile = BH rows of exactly 8 pixels, HP = 8 + KW - 1).  Per 64-channel chunk of a 3x3 conv each CTA writes
// 23 KB (halo) + 9 weight half-tiles into shared memory instead of 9 x (16 KB + weight half-tile): for N = 256 the operand
// bytes pulled from L2 and written to shared memory drop by 42 %, for N = 128 by 57 %, which is what bounds these tiles.
// Pipeline: a ring of halo slots (one per tile x channel chunk) and a ring of weight slots (one per tap), each with its
// own full/empty barriers; loop order is chunk-outer / tap-inner.  Otherwise identical to conv_tc2_kernel.
constexpr int HALO_W = 8;       // pixels per tile row = rows per swizzle group
constexpr int HALO_H = 16;

template <int BN> struct CfgH {
    static constexpr int A_SLOT3 = 23 * 1024;                   // 3x3: 18 x 10 px x 128 B = 23040 B, rounded to 1024
    static constexpr int B_BYTES = (BN / 2) * BK * 2;           // this CTA's half of one tap's weight tile
    static constexpr int SA = BN == 256 ? 3 : 4;
    static constexpr int SB = (200 * 1024 - SA * A_SLOT3) / B_BYTES > 12 ? 12 : (200 * 1024 - SA * A_SLOT3) / B_BYTES;
    static constexpr int SMEM_BYTES = SA * A_SLOT3 + SB * B_BYTES + 1024 + 512 + 4 * 4096;   // + one 4 KB store patch per epilogue warp
    static constexpr uint32_t TMEM_COLS = 2 * BN;
    static constexpr int THREADS = 192;
};

__device__ __forceinline__ uint64_t make_kmajor_sw128_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
conv_tc2h_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const float* __restrict__ bias, const bf16* __restrict__ res, bf16* __restrict__ out,
                 double* __restrict__ stats, TcGeom g) {
    using C = CfgH<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem + C::SA * C::A_SLOT3;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + C::SB * C::B_BYTES);
    uint64_t* afull = bars;                            // [SA]   (leader's copies are the live ones for full / tempty)
    uint64_t* aempty = afull + C::SA;                  // [SA]
    uint64_t* bfull = aempty + C::SA;                  // [SB]
    uint64_t* bempty = bfull + C::SB;                  // [SB]
    uint64_t* tfull = bempty + C::SB;                  // [2]
    uint64_t* tempty = tfull + 2;                      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int pair_tiles = (g.m_tiles >> 1) * g.n_tiles * (g.up == 1 ? 4 : 1);
    const int phases_in = g.up == 2 ? 4 : 1;                       // output phases accumulated inside one tile (sub-pixel data gradient)
    const int taps = g.KH * g.KW;
    const int HP = HALO_W + g.KW - 1;                              // halo pitch in pixels
    const uint32_t halo_bytes = (uint32_t)((HALO_H + g.KH - 1) * HP * 128);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int s = 0; s < C::SA; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < C::SB; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer (both CTAs) ==============================
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters) {
                const int q = g.up == 1 ? (pt >> 2) : pt;
                const int mt = 2 * (q / g.n_tiles) + (int)rank, nt = q % g.n_tiles;
                const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, b = mt / (g.tiles_w * g.tiles_h);
                const int n0 = nt * BN + (int)rank * (BN / 2);
                for (int pi = 0; pi < phases_in; ++pi) {
                    const int ph = g.up == 1 ? (pt & 3) : pi, py = ph >> 1, px = ph & 1;
                    int w0 = tw * HALO_W - g.pl, h0 = th * HALO_H - g.pt, tap0 = 0;
                    if (g.up == 1) { w0 = tw * HALO_W - 1 + px; h0 = th * HALO_H - 1 + py; tap0 = ph * 4; }
                    if (g.up == 2) { w0 = 2 * tw * HALO_W - px; h0 = 2 * th * HALO_H - py; tap0 = ph * 4; }    // element-stride-2 map
                    for (int kc = 0; kc < g.k_chunks; ++kc) {
                        mbar_wait(&aempty[sa], pa ^ 1);
                        if (leader) mbar_expect_tx(&afull[sa], 2 * halo_bytes);
                        tma_load_4d_2sm(smem + sa * C::A_SLOT3, &map_a, mapa_u32(smem_u32(&afull[sa]), 0), kc * BK, w0, h0, b);
                        if (++sa == C::SA) { sa = 0; pa ^= 1; }
                        for (int tap = 0; tap < taps; ++tap) {
                            mbar_wait(&bempty[sb], pb ^ 1);
                            if (leader) mbar_expect_tx(&bfull[sb], 2 * C::B_BYTES);
                            tma_load_3d_2sm(smem_b + sb * C::B_BYTES, &map_b, mapa_u32(smem_u32(&bfull[sb]), 0), kc * BK, n0, tap0 + tap);
                            if (++sb == C::SB) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (leader only) ==============================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_m256(BN);
            const uint32_t sbo = (uint32_t)HP * 128;
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int pk = 0; pk < phases_in * g.k_chunks; ++pk) {          // (phase, channel chunk): one halo each
                    mbar_wait(&afull[sa], pa);
                    const uint32_t a_base = smem_u32(smem + sa * C::A_SLOT3);
                    int kh = 0, kw = 0;
                    for (int tap = 0; tap < taps; ++tap) {
                        mbar_wait(&bfull[sb], pb);
                        tc_fence_after();
                        const uint64_t adesc = make_kmajor_sw128_desc_sbo(a_base + (uint32_t)(kh * HP + kw) * 128, sbo);
                        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem_b + sb * C::B_BYTES));
#pragma unroll
                        for (int kk = 0; kk < BK / UMMA_K; ++kk)
                            umma_bf16_2sm(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (pk | tap | kk) != 0);
                        umma_commit_2sm(&bempty[sb]);
                        if (++sb == C::SB) { sb = 0; pb ^= 1; }
                        if (++kw == g.KW) { kw = 0; ++kh; }
                    }
                    umma_commit_2sm(&aempty[sa]);                   // halo slot free once the last tap's MMAs retire
                    if (++sa == C::SA) { sa = 0; pa ^= 1; }
                }
                umma_commit_2sm(&tfull[buf]);
            }
        }
    } else {
        // ============================== epilogue (warps 2..5, both CTAs) ==============================
        const int lg = warp & 3;
        const int r = lg * 32 + lane;
        const int dh = r / HALO_W, dw = r % HALO_W;
        int it = 0;
        for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++it) {
            const int buf = it & 1;
            const int q = g.up == 1 ? (pt >> 2) : pt;
            const int mt = 2 * (q / g.n_tiles) + (int)rank, nt = q % g.n_tiles;
            const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, b = mt / (g.tiles_w * g.tiles_h);
            const int64_t pix = ((int64_t)b * g.H + th * HALO_H + dh) * g.W + tw * HALO_W + dw;
            const int n0 = nt * BN;
            if (!stats || line_stats_ok(g.cpg) || g.up == 1) {
                // residual rows of the first block are fetched (line-coalesced) while the main loop of this tile still runs
                EpiGeom eg;
                eg.tile_pix0 = ((int64_t)b * g.H + th * HALO_H) * g.W + tw * HALO_W;
                eg.W = g.W; eg.bw_shift = 3; eg.bw_mask = HALO_W - 1; eg.Cout = g.Cout; eg.pix_step = 1;
                if (g.up == 1) {                 // this tile is phase (py, px) of the 2H x 2W output image
                    const int py = (pt & 3) >> 1, px = pt & 1;
                    eg.W = 2 * g.W; eg.pix_step = 2;
                    eg.tile_pix0 = ((int64_t)b * 2 * g.H + 2 * th * HALO_H + py) * (2 * g.W) + 2 * tw * HALO_W + px;
                }
                uint4* patch = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(bars) + 512) + (warp - 2) * 256;
                int off[8];
                epilogue_row_offsets(lg * 32, eg, lane, off);
                const bf16* res_tile = res ? res + eg.tile_pix0 * g.Cout : nullptr;
                uint4 rv[8];
                if (res_tile) epilogue_load_residual(res_tile, n0, off, rv);
                mbar_wait(&tfull[buf], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * BN;
                epilogue_rows<BN>(taddr, n0, g.Cout, off, bias, res_tile, rv, out + eg.tile_pix0 * g.Cout, patch, lane, g.flags,
                                  stats ? stats + (int64_t)b * 64 : nullptr, g.cpg);
            } else {
                // fused GroupNorm statistics: per-thread rows (the statistics are reduced per pixel row), residual one chunk ahead
                const bf16* rp = res ? res + pix * g.Cout + n0 : nullptr;
                uint4 rv[4], rn[4];
                if (rp) {
#pragma unroll
                    for (int q = 0; q < 4; q += 2) ld_stream32(rp + q * 8, rv[q], rv[q + 1]);     // Cout % 128 == 0 here: 32-byte aligned
                }
                mbar_wait(&tfull[buf], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * BN;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    if (rp && c0 + 32 < BN) {
#pragma unroll
                        for (int q = 0; q < 4; q += 2) ld_stream32(rp + c0 + 32 + q * 8, rn[q], rn[q + 1]);
                    }
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    epilogue_chunk_pre(v, n0 + c0, pix, bias, rv, rp != nullptr, out, g.Cout, stats + (int64_t)b * 64, g.cpg, lane, g.flags);
#pragma unroll
                    for (int q = 0; q < 4; ++q) rv[q] = rn[q];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------- transposed halo tile for 128-channel outputs
// Measured on B200 (scripts/experiments/umma_pace_probe.cu): a shared-memory-operand tcgen05.mma with M = 128 rows per CTA
// retires one instruction per ~140 cycles whatever N is (the 128 x 32 B A-operand read paces it), so an N = 128 instruction
// runs the tensor pipe at half rate and only N = 256 reaches the peak.  For Cout = 128 layers the GEMM is therefore issued
// transposed:   D^T[co, pixel] = sum_{tap, ci} Wp[tap][co][ci] * X[pixel (+) tap, ci]        M = 128 (Cout), N = 256 (pixels)
// A = the weight tile of one tap (the same TMA box as before, now the A operand), B = 256 pixels = 32 image rows x 8 pixels
// read straight out of one (32+2) x (8+2) pixel halo (see conv_tc2h_kernel) at the tap's row offset.  The accumulator holds
// channels in TMEM lanes and pixels in columns; the epilogue writes it back to channels-last memory as 64-byte segments
// (32 consecutive channels of one pixel per warp store).
constexpr int THALO_H = 32;
struct CfgT {
    static constexpr int A_SLOT = 43 * 1024;                    // (32+2) x (8+2) px x 128 B = 43520 B
    static constexpr int W_BYTES = BM * BK * 2;                 // one tap's weight tile: 128 co x 64 ci
    static constexpr int SA = 2, SW = 7;
    static constexpr int SMEM_BYTES = SA * A_SLOT + SW * W_BYTES + 1024 + 512 + 8 * 2048;   // + one 2 KB transpose patch per epilogue warp
    static constexpr uint32_t TMEM_COLS = 512;                  // two accumulator buffers of 256 pixel columns
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

__global__ void __launch_bounds__(CfgT::THREADS, 1)
conv_tcT_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                const float* __restrict__ bias, const bf16* __restrict__ res, bf16* __restrict__ out,
                double* __restrict__ stats /* GroupNorm(32) statistics of the output, group width 4 (Cout = 128), or null */, TcGeom g) {
    using C = CfgT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_w = smem + C::SA * C::A_SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + C::SW * C::W_BYTES);
    uint64_t* afull = bars;
    uint64_t* aempty = afull + C::SA;
    uint64_t* wfull = aempty + C::SA;
    uint64_t* wempty = wfull + C::SW;
    uint64_t* tfull = wempty + C::SW;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = g.m_tiles * g.n_tiles;              // m_tiles: 256-pixel tiles, n_tiles: 128-channel tiles
    const int taps = g.KH * g.KW;
    const int HP = HALO_W + g.KW - 1;
    const uint32_t halo_bytes = (uint32_t)((THALO_H + g.KH - 1) * HP * 128);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_w);
        for (int s = 0; s < C::SA; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < C::SW; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], C::EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int sa = 0, sw = 0; uint32_t pa = 0, pw = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mt = tile / g.n_tiles, nt = tile % g.n_tiles;
                const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, b = mt / (g.tiles_w * g.tiles_h);
                const int w0 = tw * HALO_W - g.pl, h0 = th * THALO_H - g.pt, co0 = nt * BM;
                for (int kc = 0; kc < g.k_chunks; ++kc) {
                    mbar_wait(&aempty[sa], pa ^ 1);
                    mbar_expect_tx(&afull[sa], halo_bytes);
                    tma_load_4d(smem + sa * C::A_SLOT, &map_x, &afull[sa], kc * BK, w0, h0, b);
                    if (++sa == C::SA) { sa = 0; pa ^= 1; }
                    for (int tap = 0; tap < taps; ++tap) {
                        mbar_wait(&wempty[sw], pw ^ 1);
                        mbar_expect_tx(&wfull[sw], C::W_BYTES);
                        tma_load_3d(smem_w + sw * C::W_BYTES, &map_w, &wfull[sw], kc * BK, co0, tap);
                        if (++sw == C::SW) { sw = 0; pw ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(256);           // M = 128 channels, N = 256 pixels
            const uint32_t sbo = (uint32_t)HP * 128;
            int sa = 0, sw = 0; uint32_t pa = 0, pw = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int kc = 0; kc < g.k_chunks; ++kc) {
                    mbar_wait(&afull[sa], pa);
                    const uint32_t x_base = smem_u32(smem + sa * C::A_SLOT);
                    int kh = 0, kw = 0;
                    for (int tap = 0; tap < taps; ++tap) {
                        mbar_wait(&wfull[sw], pw);
                        tc_fence_after();
                        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem_w + sw * C::W_BYTES));
                        const uint64_t bdesc = make_kmajor_sw128_desc_sbo(x_base + (uint32_t)(kh * HP + kw) * 128, sbo);
#pragma unroll
                        for (int kk = 0; kk < BK / UMMA_K; ++kk)
                            umma_bf16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kc | tap | kk) != 0);
                        umma_commit(&wempty[sw]);
                        if (++sw == C::SW) { sw = 0; pw ^= 1; }
                        if (++kw == g.KW) { kw = 0; ++kh; }
                    }
                    umma_commit(&aempty[sa]);
                    if (++sa == C::SA) { sa = 0; pa ^= 1; }
                }
                umma_commit(&tfull[buf]);
            }
        }
    } else {
        // ============================== epilogue: lane = channel, register = pixel ==============================
        // Each warp turns its 32 channels x 32 pixels block around through a private 2 KB shared-memory patch so that global
        // memory sees 16-byte accesses: 4 lanes cover the 64 contiguous bytes (32 channels) of one pixel.
        const int lg = warp & 3;                       // TMEM lanes [32*lg, 32*lg+32) = channels co0 + 32*lg + lane
        const int half = (warp - 2) >> 2;              // pixel columns [128*half, 128*half+128)
        bf16* patch = reinterpret_cast<bf16*>(smem_w + C::SW * C::W_BYTES + 512) + (warp - 2) * 1024;
        const int px_l = lane >> 2, grp = lane & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int mt = tile / g.n_tiles, nt = tile % g.n_tiles;
            const int tw = mt % g.tiles_w, th = (mt / g.tiles_w) % g.tiles_h, b = mt / (g.tiles_w * g.tiles_h);
            const int cw = nt * BM + lg * 32;          // first channel of this warp
            const bool live = cw < g.Cout;             // Cout = 64: the upper half of the 128-row tile is TMA zero fill, nothing to store
            const float bv = (bias && live && cw + lane < g.Cout) ? __ldg(bias + cw + lane) : 0.f;
            const int64_t pix0 = ((int64_t)b * g.H + th * THALO_H) * g.W + tw * HALO_W;
            // the residual tile is fetched while the main loop of this tile is still running (the warp would only wait)
            uint4 rv[16];
            if (res && live && (g.Cout & 31) == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int n = half * 128 + i * 8 + px_l;
                    rv[i] = ld_stream16(res + (pix0 + (int64_t)(n >> 3) * g.W + (n & 7)) * g.Cout + cw + grp * 8);
                }
            }
            mbar_wait(&tfull[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * 256 + half * 128;
            if (live && (g.Cout & 31) != 0) {
                // thin output (the 128 -> 3 head, the 64 -> 3 stem gradient): a few live channel lanes, scalar bf16 stores
                const bool lane_live = cw + lane < g.Cout;
                const float bs = (bias && lane_live) ? __ldg(bias + cw + lane) : 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    if (lane_live) {
                        const int n0 = half * 128 + c0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int64_t idx = (pix0 + (int64_t)((n0 + j) >> 3) * g.W + ((n0 + j) & 7)) * g.Cout + cw + lane;
                            float f = __uint_as_float(v[j]) + bs;
                            if (g.flags & EPI_RELU) f = fmaxf(f, 0.f);
                            if (res) {
                                const float fr = __bfloat162float(res[idx]);
                                f = (g.flags & EPI_MASK) ? (fr > 0.f ? f : 0.f) : bf16_round(f) + fr;
                            }
                            out[idx] = __float2bfloat16_rn(f);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);
                continue;
            }
            float ts0 = 0.f, tq0 = 0.f, ts1 = 0.f, tq1 = 0.f;      // this lane's two groups of 4 channels (fused GroupNorm statistics)
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                if (!live) break;                      // warp-uniform
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float f = __uint_as_float(v[j]) + bv;
                    patch[j * 32 + lane] = __float2bfloat16_rn((g.flags & EPI_RELU) ? fmaxf(f, 0.f) : f);
                }
                __syncwarp();
                const int n0 = half * 128 + c0;        // 32 pixel columns = 4 image rows of 8 pixels
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int n = n0 + q * 8 + px_l;
                    const int64_t idx = (pix0 + (int64_t)(n >> 3) * g.W + (n & 7)) * g.Cout + cw + grp * 8;
                    uint4 pk = *reinterpret_cast<const uint4*>(patch + (q * 8 + px_l) * 32 + grp * 8);
                    if (res) {
                        float f[8], fr[8];
                        unpack_bf16x8(pk, f);
                        unpack_bf16x8(rv[(c0 >> 3) + q], fr);
                        if (g.flags & EPI_MASK) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = fr[e] > 0.f ? f[e] : 0.f;
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] += fr[e];           // bf16 conv output + bf16 residual, one rounding
                        }
                        pk = pack_bf16x8(f);
                    }
                    *reinterpret_cast<uint4*>(out + idx) = pk;
                    if (stats) {
                        float fo[8];
                        unpack_bf16x8(pk, fo);
                        ts0 += (fo[0] + fo[1]) + (fo[2] + fo[3]);
                        tq0 += (fo[0] * fo[0] + fo[1] * fo[1]) + (fo[2] * fo[2] + fo[3] * fo[3]);
                        ts1 += (fo[4] + fo[5]) + (fo[6] + fo[7]);
                        tq1 += (fo[4] * fo[4] + fo[5] * fo[5]) + (fo[6] * fo[6] + fo[7] * fo[7]);
                    }
                }
                __syncwarp();
            }
            if (stats && live) {                       // lanes with the same `grp` hold the same channels of different pixels
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    ts0 += __shfl_xor_sync(0xffffffffu, ts0, o); tq0 += __shfl_xor_sync(0xffffffffu, tq0, o);
                    ts1 += __shfl_xor_sync(0xffffffffu, ts1, o); tq1 += __shfl_xor_sync(0xffffffffu, tq1, o);
                }
                if (px_l == 0) {
                    double* st = stats + (int64_t)b * 64 + ((cw + grp * 8) >> 2) * 2;
                    atomicAdd(st, (double)ts0); atomicAdd(st + 1, (double)tq0);
                    atomicAdd(st + 2, (double)ts1); atomicAdd(st + 3, (double)tq1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------- weight gradient
// dWp[tap][co][ci] += sum_pixels dY[p][co] * X[p (+) tap][ci]          M = Cout, N = Cin, K = pixels (split)
//
// Both operands are "MN-major" for this GEMM (the contraction index -- pixels -- is the slow index of a
// channels-last tensor), which tcgen05 consumes natively: the same {64 ch, BW, BH, 1} TMA boxes as the forward
// pass land as [pixel][64 ch] 128-byte-swizzled rows, described to the MMA as MN-major atoms (8 x 128 B) with
// LBO = distance between 64-channel boxes and SBO = 1024 B between 8-pixel groups.
constexpr int WG_PIX = 64;                 // pixels (K) per pipeline stage
constexpr int WG_BOX_BYTES = WG_PIX * 128; // one {64 ch x 64 px} box

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(WG_BOX_BYTES >> 4) << 16;       // LBO: next 64-channel atom along M/N
    d |= (uint64_t)(1024 >> 4) << 32;               // SBO: next 8-pixel group along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) {
    return make_idesc(n) | (1u << 15) | (1u << 16);   // A and B both MN-major
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WgGeom {
    int B, H, W, Cin, Cout, KH, KW, pt, pl;      // H, W: dy (output) image size
    int stride, IH, IW;                          // conv stride and x (input) image size
    int BW, BH, tiles_w, tiles_h;
    int pix_tiles;          // B * tiles_h * tiles_w
    int co_tiles, ci_tiles, tap_groups, splits, tiles_per_split;
    int up;                 // CTA-pair kernel only: 1 = sub-pixel Upsample (see dmvae_conv_up2x_fwd): 16 "taps" t = 4*phase + 2a + b; H, W are
                            // the LOW-RES size; dy is the phase's pixels of the 2H x 2W gradient (element-stride-2 map), x the low-res input
};

// One CTA owns an (MT*128 co) x (TPC taps x BN ci) block of the tap-major gradient and walks a contiguous range of
// 64-pixel tiles (split-K).  MT = 2 and/or TPC > 1 raise the MACs per operand byte fetched from L2
// (43.7 -> 65 MAC/B for 512x512, 32 -> 48 for the 128-channel layers), which is what bounds this kernel.
template <int BN, int MT, int TPC> struct WgCfg {
    static constexpr int A_BYTES = MT * 2 * WG_BOX_BYTES;
    static constexpr int B_BYTES_TAP = (BN / 64) * WG_BOX_BYTES;
    static constexpr int B_BYTES = TPC * B_BYTES_TAP;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr int ACC_COLS = MT * TPC * BN;
    static constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
    static constexpr int EPI_WARPS = 4 * MT;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
};

template <int BN, int MT, int TPC>
__global__ void __launch_bounds__(WgCfg<BN, MT, TPC>::THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                     float* __restrict__ dwp, WgGeom g) {
    using C = WgCfg<BN, MT, TPC>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tfull = bars + 2 * C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x -> (co tile, ci tile, tap group) ; blockIdx.y -> pixel split
    int id = blockIdx.x;
    const int co_t = id % g.co_tiles; id /= g.co_tiles;
    const int ci_t = id % g.ci_tiles; id /= g.ci_tiles;
    const int tap0 = id * TPC;
    const int taps = g.KH * g.KW;
    const int ntap = (taps - tap0 < TPC) ? taps - tap0 : TPC;     // last group may be partial
    const int co0 = co_t * (BM * MT), ci0 = ci_t * BN;
    const int t_begin = blockIdx.y * g.tiles_per_split;
    const int t_end = (t_begin + g.tiles_per_split < g.pix_tiles) ? t_begin + g.tiles_per_split : g.pix_tiles;
    const int k_iters = t_end - t_begin;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_dy);
        prefetch_tmap(&map_x);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx = C::A_BYTES + ntap * C::B_BYTES_TAP;
            for (int t = t_begin; t < t_end; ++t) {
                const int tw = t % g.tiles_w, th = (t / g.tiles_w) % g.tiles_h, b = t / (g.tiles_w * g.tiles_h);
                const int w0 = tw * g.BW, h0 = th * g.BH;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * C::STAGE_BYTES;
                uint8_t* sb = sa + C::A_BYTES;
                mbar_expect_tx(&full[stage], tx);
#pragma unroll
                for (int j = 0; j < 2 * MT; ++j) tma_load_4d(sa + j * WG_BOX_BYTES, &map_dy, &full[stage], co0 + j * 64, w0, h0, b);
                for (int tp = 0; tp < ntap; ++tp) {
                    const int kh = (tap0 + tp) / g.KW, kw = (tap0 + tp) % g.KW;
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_4d(sb + tp * C::B_BYTES_TAP + j * WG_BOX_BYTES, &map_x, &full[stage], ci0 + j * 64,
                                    w0 * g.stride + kw - g.pl, h0 * g.stride + kh - g.pt, b);
                }
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_mn(BN);
            int stage = 0; uint32_t phase = 0;
            for (int k = 0; k < k_iters; ++k) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                for (int tp = 0; tp < ntap;) {
                    // An M = 128 instruction takes ~140 cycles whatever N is (umma_pace_probe.cu), so 128-channel taps are issued
                    // two at a time: their boxes are consecutive 64-channel atoms LBO apart, i.e. one N = 256 operand, and
                    // their accumulators are adjacent TMEM column ranges.
                    const bool two = (BN == 128) && (tp + 1 < ntap);
                    const uint32_t id = two ? make_idesc_mn(256) : idesc;
                    const uint64_t bdesc = make_mnmajor_sw128_desc(sa + C::A_BYTES + tp * C::B_BYTES_TAP);
#pragma unroll
                    for (int sub = 0; sub < MT; ++sub) {
                        const uint64_t adesc = make_mnmajor_sw128_desc(sa + sub * 2 * WG_BOX_BYTES);
                        const uint32_t d_tmem = tmem_base + (sub * TPC + tp) * BN;
#pragma unroll
                        for (int kk = 0; kk < WG_PIX / UMMA_K; ++kk) {
                            // 16 pixels = 16 rows of 128 B = 2048 B along K: +128 in 16-byte units
                            umma_bf16(d_tmem, adesc + 128 * kk, bdesc + 128 * kk, id, (k | kk) != 0);
                        }
                    }
                    tp += two ? 2 : 1;
                }
                umma_commit(&empty[stage]);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(tfull);
        }
    } else {
        const int lg = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int co = co0 + sub * BM + lg * 32 + lane;
        if (k_iters > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
            for (int tp = 0; tp < ntap; ++tp) {
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (sub * TPC + tp) * BN;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    const int ci = ci0 + c0;
                    if (co < g.Cout && ci < g.Cin) {
                        float* dst = dwp + ((int64_t)(tap0 + tp) * g.Cout + co) * g.Cin + ci;
                        if (ci + 32 <= g.Cin) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                red_add_v4(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                           __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                        } else {
                            for (int e = 0; e < 32 && ci + e < g.Cin; ++e) atomicAdd(dst + e, __uint_as_float(v[e]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------- weight gradient, CTA-pair variant (cta_group::2)
// ncu on the single-CTA kernel: tensor pipe 76 % active with sm__memory_throughput at the same 76 % -- the shared-memory
// pipe (TMA writes + operand reads, ~146 B/cycle demanded of 128) is what holds it.  A CTA pair shares the x tile: each CTA
// loads half of its 256 input channels and MT x 128 rows of dy; tcgen05.mma.cta_group::2 (M = 256) reads the two x halves
// from both CTAs, so per SM the x bytes written to and read from shared memory halve (~101 B/cycle for MT = 2).
template <int MT> struct Wg2Cfg {
    static constexpr int A_BYTES = MT * 2 * WG_BOX_BYTES;       // dy: MT x 128 output channels of this CTA
    static constexpr int B_BYTES = 2 * WG_BOX_BYTES;            // x: this CTA's 128 of the 256 input channels
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr uint32_t TMEM_COLS = MT * 256;
    static constexpr int EPI_WARPS = 4 * MT;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

template <int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Wg2Cfg<MT>::THREADS, 1)
conv_tc_wgrad2_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                      float* __restrict__ dwp, WgGeom g) {
    using C = Wg2Cfg<MT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;                        // leader's copies are live
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tfull = bars + 2 * C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    int id = blockIdx.x >> 1;
    const int co_t = id % g.co_tiles; id /= g.co_tiles;       // co tile of the PAIR: 2 * MT * 128 output channels
    const int ci_t = id % g.ci_tiles; id /= g.ci_tiles;
    const int tap = id;
    const int kh = tap / g.KW, kw = tap % g.KW;
    // offsets of the dy box and the x box relative to the pixel tile's origin (w0, h0); the dy origin is scaled by dys
    int dys = 1, dyw = 0, dyh = 0, xw = kw - g.pl, xh = kh - g.pt;
    if (g.up) {
        const int py = tap >> 3, px = (tap >> 2) & 1, a = (tap >> 1) & 1, bb = tap & 1;
        dys = 2; dyw = px; dyh = py; xw = bb - 1 + px; xh = a - 1 + py;
    }
    const int co0 = co_t * (2 * MT * BM) + (int)rank * (MT * BM);
    const int ci0 = ci_t * 256;
    const int t_begin = blockIdx.y * g.tiles_per_split;
    const int t_end = (t_begin + g.tiles_per_split < g.pix_tiles) ? t_begin + g.tiles_per_split : g.pix_tiles;
    const int k_iters = t_end - t_begin;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_dy);
        prefetch_tmap(&map_x);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int tw = t % g.tiles_w, th = (t / g.tiles_w) % g.tiles_h, b = t / (g.tiles_w * g.tiles_h);
                const int w0 = tw * g.BW, h0 = th * g.BH;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * C::STAGE_BYTES;
                uint8_t* sb = sa + C::A_BYTES;
                if (leader) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
                const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
#pragma unroll
                for (int j = 0; j < 2 * MT; ++j)
                    tma_load_4d_2sm(sa + j * WG_BOX_BYTES, &map_dy, lbar, co0 + j * 64, w0 * dys + dyw, h0 * dys + dyh, b);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    tma_load_4d_2sm(sb + j * WG_BOX_BYTES, &map_x, lbar, ci0 + (int)rank * 128 + j * 64,
                                    w0 * g.stride + xw, h0 * g.stride + xh, b);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_m256(256) | (1u << 15) | (1u << 16);      // M = 256 (pair), N = 256, both MN-major
            int stage = 0; uint32_t phase = 0;
            for (int k = 0; k < k_iters; ++k) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                const uint64_t bdesc = make_mnmajor_sw128_desc(sa + C::A_BYTES);
#pragma unroll
                for (int sub = 0; sub < MT; ++sub) {
                    const uint64_t adesc = make_mnmajor_sw128_desc(sa + sub * 2 * WG_BOX_BYTES);
#pragma unroll
                    for (int kk = 0; kk < WG_PIX / UMMA_K; ++kk)
                        umma_bf16_2sm(tmem_base + sub * 256, adesc + 128 * kk, bdesc + 128 * kk, idesc, (k | kk) != 0);
                }
                umma_commit_2sm(&empty[stage]);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit_2sm(tfull);
        }
    } else {
        const int lg = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int co = co0 + sub * BM + lg * 32 + lane;
        if (k_iters > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + sub * 256;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                const int ci = ci0 + c0;
                if (co < g.Cout && ci + 32 <= g.Cin) {
                    float* dst = dwp + ((int64_t)tap * g.Cout + co) * g.Cin + ci;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        red_add_v4(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                   __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------- host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

struct MapKey {
    uint64_t v[8];
    bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 8; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

// bf16 tensor, rank 3 or 4, innermost dim contiguous, 128B swizzle, zero OOB fill
// `box` counts elements landing in shared memory per dimension; `estr` (optional) is the traversal stride per dimension
// (a stride-2 conv samples every other pixel: the TMA unit does the subsampling).
int get_tensor_map(const void* ptr, int rank, const uint64_t* dims, const uint32_t* box, CUtensorMap* out, const uint32_t* estr = nullptr) {
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.v[0] = (uint64_t)(uintptr_t)ptr;
    key.v[1] = (uint64_t)rank;
    for (int i = 0; i < rank; ++i) key.v[2 + i] = dims[i] | ((uint64_t)box[i] << 40) | ((uint64_t)(estr ? estr[i] : 1) << 56);
    {
        std::lock_guard<std::mutex> lk(g_map_mu);
        auto itc = g_map_cache.find(key);
        if (itc != g_map_cache.end()) { *out = itc->second; return DMVAE_OK; }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return dmvae_set_error(DMVAE_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    // cuTensorMapEncodeTiled is a driver call and needs a current context on THIS thread; an autograd worker thread that has not
    // issued a runtime call through this library yet may have none (CUDA_ERROR_INVALID_CONTEXT).  Bind the primary context of the
    // device that owns the tensor (not "device 0": one process per GPU, but the rank's GPU need not be ordinal 0).
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) cudaSetDevice(at.device);
        else cudaGetLastError();
    }
    cuuint64_t gdim[4]; cuuint64_t gstride[3]; cuuint32_t bdim[4]; cuuint32_t estride[4];
    uint64_t stride = 2;
    for (int i = 0; i < rank; ++i) {
        estride[i] = estr ? estr[i] : 1;
        gdim[i] = dims[i]; bdim[i] = box[i] * estride[i];
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride, bdim, estride,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return dmvae_set_error(DMVAE_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_map_cache.size() > 4096) g_map_cache.clear();
    g_map_cache.emplace(key, *out);
    return DMVAE_OK;
}

int pick_pixel_tile_n(int H, int W, int pixels, int* BW, int* BH) {
    int bw = 1;
    while (bw * 2 <= pixels && bw * 2 <= 256 && W % (bw * 2) == 0) bw *= 2;
    const int bh = pixels / bw;
    if (bw < 8 || H % bh != 0 || bh > 256) return 0;
    *BW = bw; *BH = bh;
    return 1;
}
int pick_pixel_tile(int H, int W, int* BW, int* BH) { return pick_pixel_tile_n(H, W, BM, BW, BH); }

int pick_pixel_tile64(int H, int W, int* BW, int* BH);
// which kernel the last conv entry point called on this thread launched (dmvae_conv_tc_last_kernel: measurement hook)
thread_local int g_last_kernel = 0;
int g_num_sms = 0;
int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int BN, int MT>
int launch_conv_tc(const void* x, const void* w, const float* bias, const void* res, void* y, double* stats, TcGeom g, cudaStream_t st) {
    using C = Cfg<BN, MT>;
    CUtensorMap ma, mb;
    const uint64_t adims[4] = {(uint64_t)g.Cin, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.B};
    const uint32_t abox[4] = {BK, (uint32_t)g.BW, (uint32_t)g.BH, 1};
    const uint32_t astr[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
    int rc = get_tensor_map(x, 4, adims, abox, &ma, astr);
    if (rc) return rc;
    const uint64_t bdims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, (uint64_t)(g.KH * g.KW)};
    const uint32_t bbox[3] = {BK, BN, 1};
    rc = get_tensor_map(w, 3, bdims, bbox, &mb);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return dmvae_set_error(DMVAE_ECUDA, "conv_tc: smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const int tiles = g.m_tiles * g.n_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    conv_tc_kernel<BN, MT><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(ma, mb, bias, (const bf16*)res, (bf16*)y, stats, g);
    g_last_kernel = DMVAE_KERNEL_CONV_TC;
    DMVAE_CHECK_LAUNCH("conv_tc_kernel");
    return DMVAE_OK;
}

template <int BN>
int launch_conv_tc2(const void* x, const void* w, const float* bias, const void* res, void* y, double* stats, TcGeom g, cudaStream_t st) {
    using C = Cfg2<BN>;
    CUtensorMap ma, mb;
    const uint64_t adims[4] = {(uint64_t)g.Cin, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.B};
    const uint32_t abox[4] = {BK, (uint32_t)g.BW, (uint32_t)g.BH, 1};
    const uint32_t astr[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
    int rc = get_tensor_map(x, 4, adims, abox, &ma, astr);
    if (rc) return rc;
    const uint64_t bdims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, (uint64_t)(g.KH * g.KW)};
    const uint32_t bbox[3] = {BK, BN / 2, 1};
    rc = get_tensor_map(w, 3, bdims, bbox, &mb);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return dmvae_set_error(DMVAE_ECUDA, "conv_tc2: smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const int pair_tiles = (g.m_tiles / 2) * g.n_tiles;
    const int clusters = pair_tiles < num_sms() / 2 ? pair_tiles : num_sms() / 2;
    conv_tc2_kernel<BN><<<2 * clusters, C::THREADS, C::SMEM_BYTES, st>>>(ma, mb, bias, (const bf16*)res, (bf16*)y, stats, g);
    g_last_kernel = DMVAE_KERNEL_CONV_TC2;
    DMVAE_CHECK_LAUNCH("conv_tc2_kernel");
    return DMVAE_OK;
}

template <int BN>
int launch_conv_tc2h(const void* x, const void* w, const float* bias, const void* res, void* y, double* stats, TcGeom g, cudaStream_t st) {
    using C = CfgH<BN>;
    CUtensorMap ma, mb;
    const uint64_t adims[4] = {(uint64_t)g.Cin, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.B};
    const uint32_t abox[4] = {BK, (uint32_t)(HALO_W + g.KW - 1), (uint32_t)(HALO_H + g.KH - 1), 1};
    int rc = get_tensor_map(x, 4, adims, abox, &ma);
    if (rc) return rc;
    const uint64_t bdims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, (uint64_t)(g.KH * g.KW)};
    const uint32_t bbox[3] = {BK, BN / 2, 1};
    rc = get_tensor_map(w, 3, bdims, bbox, &mb);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2h_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return dmvae_set_error(DMVAE_ECUDA, "conv_tc2h: smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const int pair_tiles = (g.m_tiles / 2) * g.n_tiles;
    const int clusters = pair_tiles < num_sms() / 2 ? pair_tiles : num_sms() / 2;
    conv_tc2h_kernel<BN><<<2 * clusters, C::THREADS, C::SMEM_BYTES, st>>>(ma, mb, bias, (const bf16*)res, (bf16*)y, stats, g);
    g_last_kernel = DMVAE_KERNEL_CONV_TC2H;
    DMVAE_CHECK_LAUNCH("conv_tc2h_kernel");
    return DMVAE_OK;
}

// Sub-pixel modes of the halo pair kernel (TcGeom::up).  `a`: the low-res input x (up = 1) or the 2H x 2W output gradient
// (up = 2); `w16`: the 16-tap pack [phase*4 + tap][N][K]; g.H, g.W: the LOW-RES grid the tiles live on; g.Cin = K, g.Cout = N.
template <int BN>
int launch_conv_tc2h_up(const void* a, const void* w16, const float* bias, void* y, double* stats, TcGeom g, cudaStream_t st) {
    using C = CfgH<BN>;
    CUtensorMap ma, mb;
    const int s = g.up == 2 ? 2 : 1;
    const uint64_t adims[4] = {(uint64_t)g.Cin, (uint64_t)(s * g.W), (uint64_t)(s * g.H), (uint64_t)g.B};
    const uint32_t abox[4] = {BK, (uint32_t)(HALO_W + 1), (uint32_t)(HALO_H + 1), 1};
    const uint32_t astr[4] = {1, (uint32_t)s, (uint32_t)s, 1};
    int rc = get_tensor_map(a, 4, adims, abox, &ma, astr);
    if (rc) return rc;
    const uint64_t bdims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, 16};
    const uint32_t bbox[3] = {BK, BN / 2, 1};
    rc = get_tensor_map(w16, 3, bdims, bbox, &mb);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2h_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return dmvae_set_error(DMVAE_ECUDA, "conv_tc2h: smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const int pair_tiles = (g.m_tiles / 2) * g.n_tiles * (g.up == 1 ? 4 : 1);
    const int clusters = pair_tiles < num_sms() / 2 ? pair_tiles : num_sms() / 2;
    conv_tc2h_kernel<BN><<<2 * clusters, C::THREADS, C::SMEM_BYTES, st>>>(ma, mb, bias, nullptr, (bf16*)y, stats, g);
    g_last_kernel = DMVAE_KERNEL_CONV_TC2H;
    DMVAE_CHECK_LAUNCH("conv_tc2h_kernel (sub-pixel)");
    return DMVAE_OK;
}

int launch_conv_tcT(const void* x, const void* w, const float* bias, const void* res, void* y, double* stats, TcGeom g, cudaStream_t st) {
    using C = CfgT;
    CUtensorMap mx, mw;
    const uint64_t xdims[4] = {(uint64_t)g.Cin, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.B};
    const uint32_t xbox[4] = {BK, (uint32_t)(HALO_W + g.KW - 1), (uint32_t)(THALO_H + g.KH - 1), 1};
    int rc = get_tensor_map(x, 4, xdims, xbox, &mx);
    if (rc) return rc;
    const uint64_t wdims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, (uint64_t)(g.KH * g.KW)};
    const uint32_t wbox[3] = {BK, BM, 1};
    rc = get_tensor_map(w, 3, wdims, wbox, &mw);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tcT_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return dmvae_set_error(DMVAE_ECUDA, "conv_tcT: smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const int tiles = g.m_tiles * g.n_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    conv_tcT_kernel<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(mx, mw, bias, (const bf16*)res, (bf16*)y, stats, g);
    g_last_kernel = DMVAE_KERNEL_CONV_TCT;
    DMVAE_CHECK_LAUNCH("conv_tcT_kernel");
    return DMVAE_OK;
}

// 0 = let the heuristic decide, 1 / 2 = force the number of 128-pixel sub-tiles per CTA, 3 = force CTA pairs
int g_force_mt = 0;
int g_pair_default = 1;      // CTA pairs for the N = 256 tiles (measured +10..30 % over the single-CTA tiles); mode 5 turns it off
int g_wgrad_tpc = 3;         // taps per CTA of the 128-channel weight-gradient tile (modes 12 / 13 / 14 select 2 / 3 / 4)
int g_halo = 1;              // halo-resident pair tiles for 3x3 filters (modes 6 / 7 turn it off / on)

}  // namespace

// Measurement hook (host only): which tile kernel the most recent tcgen05 conv entry point called on this thread launched --
// 1 conv_tc_kernel, 2 conv_tc2_kernel, 3 conv_tc2h_kernel, 4 conv_tcT_kernel, 5 conv_tc_wgrad_kernel, 6 conv_tc_wgrad2_kernel, 0 none yet.
DMVAE_API int dmvae_conv_tc_last_kernel(void) { return g_last_kernel; }

// 1 if (shape) can run on the tcgen05 tile, else 0 (caller uses conv_direct)
DMVAE_API int dmvae_conv_tc_supported(int B, int H, int W, int Cin, int Cout, int KH, int KW) {
    int bw, bh;
    // Cout may be tiny (the 128->3 head runs as an N=32 tile with TMA zero-filled weight rows and masked stores)
    if (B <= 0 || Cin % 8 != 0 || Cin < 32 || Cout < 1 || (Cout > 32 && Cout % 8 != 0)) return 0;
    if (!((KH == 3 && KW == 3) || (KH == 1 && KW == 1))) return 0;
    return pick_pixel_tile(H, W, &bw, &bh);
}

// stride-1 "same" convolution (3x3 pad 1 or 1x1 pad 0) on the tensor cores.
//   y = conv(x, w_packed) + bias [; y = bf16(y) + residual]
DMVAE_API int dmvae_conv_tc_fwd(const void* x, const void* w_packed, const float* bias, const void* residual, void* y,
                                double* gn_stats, int B, int H, int W, int Cin, int Cout, int KH, int KW, int flags, void* stream) {
    DMVAE_CHECK_ARG((flags & ~(EPI_RELU | EPI_MASK)) == 0, "conv_tc_fwd: unknown flags %d", flags);
    DMVAE_CHECK_ARG(!(flags & EPI_MASK) || residual, "conv_tc_fwd: the mask flag needs the mask tensor in `residual`");
    DMVAE_CHECK_ARG(x && w_packed && y, "conv_tc_fwd: null pointer");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                    ((uintptr_t)residual & 15) == 0, "conv_tc_fwd: buffers must be 16-byte aligned");
    if (!dmvae_conv_tc_supported(B, H, W, Cin, Cout, KH, KW))
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_tc_fwd: shape B=%d H=%d W=%d Cin=%d Cout=%d k=%dx%d not supported", B, H, W, Cin, Cout, KH, KW);
    TcGeom g;
    g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.KH = KH; g.KW = KW;
    g.pt = (KH - 1) / 2; g.pl = (KW - 1) / 2;
    g.stride = 1; g.IH = H; g.IW = W;
    g.flags = flags; g.up = 0;
    double* stats = nullptr;
    g.cpg = 0;
    if (gn_stats) {
        // fused GroupNorm(32) statistics need whole 32-channel chunks that do not straddle groups
        const int cpg = Cout / 32;
        if (Cout % 32 != 0 || !(cpg == 1 || cpg == 2 || cpg == 4 || cpg == 8 || cpg == 16 || cpg % 32 == 0))
            return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_tc_fwd: fused GroupNorm statistics unsupported for Cout=%d", Cout);
        stats = gn_stats;
        g.cpg = cpg;
    }
    g.k_chunks = (Cin + BK - 1) / BK;
    cudaStream_t st = (cudaStream_t)stream;
    const int bn = Cout <= 32 ? 32 : ((Cout % 256 == 0 || Cout > 256) ? 256 : 128);
    g.n_tiles = (Cout + bn - 1) / bn;
    // 256-pixel tiles when the image allows them and there is at least ~one wave of them
    int mt = 1, bw2, bh2;
    if (bn >= 128 && pick_pixel_tile_n(H, W, 2 * BM, &bw2, &bh2)) {
        const int tiles2 = B * (W / bw2) * (H / bh2) * g.n_tiles;
        if (tiles2 >= (num_sms() * 3) / 4) mt = 2;
    }
    if (g_force_mt == 1) mt = 1;
    if (g_force_mt == 2 && bn >= 128 && pick_pixel_tile_n(H, W, 2 * BM, &bw2, &bh2)) mt = 2;
    // transposed halo tiles for 128-channel outputs (an N = 128 instruction only half-fills the tensor pipe)
    if (g_halo && (g_force_mt == 0 || g_halo == 2) && KH == 3 && KW == 3 && ((bn == 128 && Cout % 64 == 0) || (bn == 32 && Cout < 32)) &&
        Cin % BK == 0 && (!stats || (Cout == 128 && g.cpg == 4)) && W % HALO_W == 0 && H % THALO_H == 0 &&
        (((uintptr_t)y | (uintptr_t)residual) & 15) == 0) {
        TcGeom gt = g;
        gt.BW = HALO_W; gt.BH = THALO_H;
        gt.tiles_w = W / HALO_W; gt.tiles_h = H / THALO_H;
        gt.m_tiles = B * gt.tiles_w * gt.tiles_h;
        gt.n_tiles = (Cout + BM - 1) / BM;
        if (g_halo == 2 || gt.m_tiles * gt.n_tiles >= (num_sms() * 3) / 4)
            return launch_conv_tcT(x, w_packed, bias, residual, y, stats, gt, st);
    }
    // halo-resident CTA pairs: 3x3 filters, 16 x 8 pixel tiles, full N tiles, at least ~3/8 of a wave of pair tiles
    if (g_halo && (g_force_mt == 0 || g_halo == 2) && KH == 3 && KW == 3 && bn >= 128 && Cout % bn == 0 && Cin % BK == 0 && W % HALO_W == 0 && H % HALO_H == 0 &&
        (((uintptr_t)y | (uintptr_t)residual) & 31) == 0) {
        TcGeom gh = g;
        gh.BW = HALO_W; gh.BH = HALO_H;
        gh.tiles_w = W / HALO_W; gh.tiles_h = H / HALO_H;
        gh.m_tiles = B * gh.tiles_w * gh.tiles_h;
        if (gh.m_tiles % 2 == 0 && (g_halo == 2 || (gh.m_tiles / 2) * gh.n_tiles >= (num_sms() * 3) / 8))
            return bn == 256 ? launch_conv_tc2h<256>(x, w_packed, bias, residual, y, stats, gh, st)
                             : launch_conv_tc2h<128>(x, w_packed, bias, residual, y, stats, gh, st);
    }
    // CTA pairs: 128-pixel tiles per CTA, an even number of them, full N tiles
    if (bn >= 128 && Cout % bn == 0 && (g_force_mt == 3 || (g_force_mt == 0 && g_pair_default && bn == 256))) {
        TcGeom gp = g;
        if (pick_pixel_tile_n(H, W, BM, &gp.BW, &gp.BH)) {
            gp.tiles_w = W / gp.BW; gp.tiles_h = H / gp.BH;
            gp.m_tiles = B * gp.tiles_w * gp.tiles_h;
            if (gp.m_tiles % 2 == 0 && (g_force_mt == 3 || (gp.m_tiles / 2) * gp.n_tiles >= (num_sms() * 3) / 8)) {
                return bn == 256 ? launch_conv_tc2<256>(x, w_packed, bias, residual, y, stats, gp, st)
                                 : launch_conv_tc2<128>(x, w_packed, bias, residual, y, stats, gp, st);
            }
        }
    }
    pick_pixel_tile_n(H, W, mt * BM, &g.BW, &g.BH);
    g.tiles_w = W / g.BW; g.tiles_h = H / g.BH;
    g.m_tiles = B * g.tiles_w * g.tiles_h;
    if (bn == 32) return launch_conv_tc<32, 1>(x, w_packed, bias, residual, y, stats, g, st);
    if (bn == 256) return mt == 2 ? launch_conv_tc<256, 2>(x, w_packed, bias, residual, y, stats, g, st)
                                  : launch_conv_tc<256, 1>(x, w_packed, bias, residual, y, stats, g, st);
    return mt == 2 ? launch_conv_tc<128, 2>(x, w_packed, bias, residual, y, stats, g, st)
                   : launch_conv_tc<128, 1>(x, w_packed, bias, residual, y, stats, g, st);
}

namespace {
int pick_pixel_tile64(int H, int W, int* BW, int* BH) {
    int bw = 1;
    while (bw * 2 <= WG_PIX && W % (bw * 2) == 0) bw *= 2;
    const int bh = WG_PIX / bw;
    if (bw < 8 || H % bh != 0) return 0;
    if (BW) *BW = bw;
    if (BH) *BH = bh;
    return 1;
}

template <int BN, int MT, int TPC>
int launch_wgrad_tc(const void* x, const void* dy, float* dwp, WgGeom g, cudaStream_t st) {
    using C = WgCfg<BN, MT, TPC>;
    CUtensorMap mdy, mx;
    const uint32_t box[4] = {64, (uint32_t)g.BW, (uint32_t)g.BH, 1};
    const uint64_t ddims[4] = {(uint64_t)g.Cout, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    int rc = get_tensor_map(dy, 4, ddims, box, &mdy);
    if (rc) return rc;
    const uint64_t xdims[4] = {(uint64_t)g.Cin, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.B};
    const uint32_t xstr[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
    rc = get_tensor_map(x, 4, xdims, box, &mx, xstr);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<BN, MT, TPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return dmvae_set_error(DMVAE_ECUDA, "conv_tc_wgrad: smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const int taps = g.KH * g.KW;
    g.co_tiles = (g.Cout + BM * MT - 1) / (BM * MT);
    g.ci_tiles = (g.Cin + BN - 1) / BN;
    g.tap_groups = (taps + TPC - 1) / TPC;
    const int base = g.co_tiles * g.ci_tiles * g.tap_groups;
    int splits = num_sms() / base;                 // fill one wave without spilling into a second
    if (splits > g.pix_tiles) splits = g.pix_tiles;
    if (splits < 1) splits = 1;
    g.tiles_per_split = (g.pix_tiles + splits - 1) / splits;
    g.splits = (g.pix_tiles + g.tiles_per_split - 1) / g.tiles_per_split;
    dim3 grid((unsigned)base, (unsigned)g.splits);
    conv_tc_wgrad_kernel<BN, MT, TPC><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(mdy, mx, dwp, g);
    g_last_kernel = DMVAE_KERNEL_WGRAD;
    DMVAE_CHECK_LAUNCH("conv_tc_wgrad_kernel");
    return DMVAE_OK;
}
template <int MT>
int launch_wgrad_tc2(const void* x, const void* dy, float* dwp, WgGeom g, cudaStream_t st) {
    using C = Wg2Cfg<MT>;
    CUtensorMap mdy, mx;
    const uint32_t box[4] = {64, (uint32_t)g.BW, (uint32_t)g.BH, 1};
    const int dys = g.up ? 2 : 1;
    const uint64_t ddims[4] = {(uint64_t)g.Cout, (uint64_t)(dys * g.W), (uint64_t)(dys * g.H), (uint64_t)g.B};
    const uint32_t dstr[4] = {1, (uint32_t)dys, (uint32_t)dys, 1};
    int rc = get_tensor_map(dy, 4, ddims, box, &mdy, dstr);
    if (rc) return rc;
    const uint64_t xdims[4] = {(uint64_t)g.Cin, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.B};
    const uint32_t xstr[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
    rc = get_tensor_map(x, 4, xdims, box, &mx, xstr);
    if (rc) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad2_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return dmvae_set_error(DMVAE_ECUDA, "conv_tc_wgrad2: smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    g.co_tiles = g.Cout / (2 * MT * BM);
    g.ci_tiles = g.Cin / 256;
    g.tap_groups = g.up ? 16 : g.KH * g.KW;
    const int base = g.co_tiles * g.ci_tiles * g.tap_groups;           // clusters before the pixel split
    int splits = (num_sms() / 2) / base;
    if (splits > g.pix_tiles) splits = g.pix_tiles;
    if (splits < 1) splits = 1;
    g.tiles_per_split = (g.pix_tiles + splits - 1) / splits;
    g.splits = (g.pix_tiles + g.tiles_per_split - 1) / g.tiles_per_split;
    dim3 grid((unsigned)(2 * base), (unsigned)g.splits);
    conv_tc_wgrad2_kernel<MT><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(mdy, mx, dwp, g);
    g_last_kernel = DMVAE_KERNEL_WGRAD2;
    DMVAE_CHECK_LAUNCH("conv_tc_wgrad2_kernel");
    return DMVAE_OK;
}
// CTA pairs for the weight gradient: 256-channel input tiles, whole 256- or 512-row output-channel blocks
int wgrad_pair_mt(int Cin, int Cout) {
    if (!g_pair_default || g_force_mt == 1 || g_force_mt == 2 || Cin % 256 != 0) return 0;
    if (Cout % 512 == 0) return 2;
    if (Cout % 256 == 0) return 1;
    return 0;
}
}  // namespace

DMVAE_API int dmvae_conv_tc_wgrad_supported(int B, int H, int W, int Cin, int Cout, int KH, int KW) {
    int bw, bh;
    if (B <= 0 || Cin % 8 != 0 || Cout % 8 != 0 || Cin < 32 || Cout < 32) return 0;
    if (!((KH == 3 && KW == 3) || (KH == 1 && KW == 1))) return 0;
    return pick_pixel_tile64(H, W, &bw, &bh);
}

// dw_tap_major[tap][Cout][Cin] (fp32) += sum_pixels dy[p][co] * x[p (+) tap][ci]
DMVAE_API int dmvae_conv_tc_wgrad(const void* x, const void* dy, float* dw_tap_major, int B, int H, int W, int Cin,
                                  int Cout, int KH, int KW, void* stream) {
    DMVAE_CHECK_ARG(x && dy && dw_tap_major, "conv_tc_wgrad: null pointer");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dw_tap_major & 15) == 0,
                    "conv_tc_wgrad: buffers must be 16-byte aligned");
    if (!dmvae_conv_tc_wgrad_supported(B, H, W, Cin, Cout, KH, KW) || Cin % 4 != 0)
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_tc_wgrad: shape B=%d H=%d W=%d Cin=%d Cout=%d k=%dx%d not supported", B, H, W, Cin, Cout, KH, KW);
    WgGeom g;
    g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.KH = KH; g.KW = KW;
    g.pt = (KH - 1) / 2; g.pl = (KW - 1) / 2;
    g.stride = 1; g.IH = H; g.IW = W; g.up = 0;
    pick_pixel_tile64(H, W, &g.BW, &g.BH);
    g.tiles_w = W / g.BW; g.tiles_h = H / g.BH;
    g.pix_tiles = B * g.tiles_w * g.tiles_h;
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide_m = Cout >= 256 && g_force_mt != 1;          // two 128-row co tiles per CTA
    if (const int pmt = wgrad_pair_mt(Cin, Cout))
        return pmt == 2 ? launch_wgrad_tc2<2>(x, dy, dw_tap_major, g, st) : launch_wgrad_tc2<1>(x, dy, dw_tap_major, g, st);
    if (Cin >= 256) {
        return wide_m ? launch_wgrad_tc<256, 2, 1>(x, dy, dw_tap_major, g, st)
                      : launch_wgrad_tc<256, 1, 1>(x, dy, dw_tap_major, g, st);
    }
    if (wide_m) return launch_wgrad_tc<128, 2, 2>(x, dy, dw_tap_major, g, st);
    if (g_force_mt == 1) return launch_wgrad_tc<128, 1, 1>(x, dy, dw_tap_major, g, st);
    if (g_wgrad_tpc == 2) return launch_wgrad_tc<128, 1, 2>(x, dy, dw_tap_major, g, st);
    if (g_wgrad_tpc == 4) return launch_wgrad_tc<128, 1, 4>(x, dy, dw_tap_major, g, st);
    return launch_wgrad_tc<128, 1, 3>(x, dy, dw_tap_major, g, st);
}

// test / tuning hook: 0 = heuristic, 1 = 128-pixel tiles, 2 = 256-pixel tiles where the shape allows
DMVAE_API int dmvae_conv_tc_set_tile_mode(int mode) {
    if (mode == 4) { g_pair_default = 1; g_force_mt = 0; return DMVAE_OK; }      // heuristic, CTA pairs preferred
    if (mode == 5) { g_pair_default = 0; g_force_mt = 0; return DMVAE_OK; }      // heuristic, single-CTA tiles only
    if (mode == 6) { g_halo = 0; return DMVAE_OK; }                              // per-tap operand fetch (no halo reuse)
    if (mode == 7) { g_halo = 1; return DMVAE_OK; }
    if (mode == 8) { g_halo = 2; return DMVAE_OK; }
    if (mode >= 12 && mode <= 14) { g_wgrad_tpc = mode - 10; return DMVAE_OK; }                              // halo tiles wherever the shape allows (tests)
    g_force_mt = (mode >= 1 && mode <= 3) ? mode : 0;
    return DMVAE_OK;
}

// ---------------------------------------------------------------- sub-pixel Upsample (flux_ae.Upsample, :98-107)
// F.interpolate(x, 2, "nearest") followed by a 3x3 / pad 1 conv equals, for each of the four output phases (py, px),
// a 2x2 conv on the LOW-RES input with the 3x3 taps that fall on the same input pixel summed:
//     y[2i+py, 2j+px] = sum_{a,b in {0,1}} Wp[py][px][a][b] . x[i + a - 1 + py, j + b - 1 + px]
//     Wp[py][.][a][.] = sum_{kh in S(py,a)} W[kh][.],   S(0,0) = {0}, S(0,1) = {1,2}, S(1,0) = {0,1}, S(1,1) = {2}   (same for kw)
// i.e. 16 instead of 36 tap-GEMMs per low-res pixel (2.25x fewer FLOPs), no 4x larger intermediate, no upsample kernels.
// All three passes run on the halo-resident CTA-pair tile (forward, data gradient) / the CTA-pair weight-gradient tile.
namespace {
int up2x_geom(TcGeom* g, int B, int H, int W, int K, int N, int up) {
    g->B = B; g->H = H; g->W = W; g->Cin = K; g->Cout = N; g->KH = 2; g->KW = 2; g->pt = 0; g->pl = 0;
    g->stride = 1; g->IH = H; g->IW = W; g->cpg = 0; g->flags = 0; g->up = up;
    g->BW = HALO_W; g->BH = HALO_H; g->tiles_w = W / HALO_W; g->tiles_h = H / HALO_H;
    g->m_tiles = B * g->tiles_w * g->tiles_h;
    g->k_chunks = K / BK;
    g->n_tiles = N / 256;
    return DMVAE_OK;
}
}  // namespace

// 1 if nearest-2x + 3x3 with these sizes runs in sub-pixel form on the tensor cores (H, W: the LOW-RES input size)
DMVAE_API int dmvae_conv_up2x_supported(int B, int H, int W, int Cin, int Cout) {
    if (B <= 0 || Cin % 256 != 0 || Cout % 256 != 0 || W % HALO_W != 0 || H % HALO_H != 0) return 0;
    return ((B * (W / HALO_W) * (H / HALO_H)) % 2) == 0 && pick_pixel_tile64(H, W, nullptr, nullptr);
}

// y[B][2H][2W][Cout] = conv3x3(nearest2x(x[B][H][W][Cin])) + bias with the 16-tap sub-pixel pack wp_fwd[16][Cout][Cin]
// (dmvae_subpixel_pack).  gn_stats as in dmvae_conv_tc_fwd (group widths 4 / 8 / 16).
DMVAE_API int dmvae_conv_up2x_fwd(const void* x, const void* wp_fwd, const float* bias, void* y, double* gn_stats, int B, int H,
                                  int W, int Cin, int Cout, void* stream) {
    DMVAE_CHECK_ARG(x && wp_fwd && y, "conv_up2x_fwd: null pointer");
    DMVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)wp_fwd | (uintptr_t)y) & 31) == 0, "conv_up2x_fwd: buffers must be 32-byte aligned");
    if (!dmvae_conv_up2x_supported(B, H, W, Cin, Cout))
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_up2x_fwd: shape B=%d H=%d W=%d Cin=%d Cout=%d not supported", B, H, W, Cin, Cout);
    TcGeom g;
    up2x_geom(&g, B, H, W, Cin, Cout, 1);
    if (gn_stats) {
        g.cpg = Cout / 32;
        if (!(g.cpg == 4 || g.cpg == 8 || g.cpg == 16))
            return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_up2x_fwd: fused GroupNorm statistics unsupported for Cout=%d", Cout);
    }
    return launch_conv_tc2h_up<256>(x, wp_fwd, bias, y, gn_stats, g, (cudaStream_t)stream);
}

// dx[B][H][W][Cin] = data gradient of the above, from dy[B][2H][2W][Cout] and wp_dgrad[16][Cin][Cout] (dmvae_subpixel_pack).
DMVAE_API int dmvae_conv_up2x_dgrad(const void* dy, const void* wp_dgrad, void* dx, int B, int H, int W, int Cin, int Cout,
                                    void* stream) {
    DMVAE_CHECK_ARG(dy && wp_dgrad && dx, "conv_up2x_dgrad: null pointer");
    DMVAE_CHECK_ARG((((uintptr_t)dy | (uintptr_t)wp_dgrad | (uintptr_t)dx) & 31) == 0, "conv_up2x_dgrad: buffers must be 32-byte aligned");
    if (!dmvae_conv_up2x_supported(B, H, W, Cin, Cout))
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_up2x_dgrad: shape B=%d H=%d W=%d Cin=%d Cout=%d not supported", B, H, W, Cin, Cout);
    TcGeom g;
    up2x_geom(&g, B, H, W, /*K=*/Cout, /*N=*/Cin, 2);
    return launch_conv_tc2h_up<256>(dy, wp_dgrad, nullptr, dx, nullptr, g, (cudaStream_t)stream);
}

// dwp[16][Cout][Cin] (fp32, caller-zeroed or accumulated) += the 16 phase-tap weight gradients of the sub-pixel form:
//   dwp[4*(2py+px) + 2a + b][co][ci] = sum_{i,j} dy[2i+py][2j+px][co] * x[i + a - 1 + py][j + b - 1 + px][ci]
// dmvae_subpixel_fold_wgrad turns them into the 3x3 weight gradient.
DMVAE_API int dmvae_conv_up2x_wgrad(const void* x, const void* dy, float* dwp, int B, int H, int W, int Cin, int Cout, void* stream) {
    DMVAE_CHECK_ARG(x && dy && dwp, "conv_up2x_wgrad: null pointer");
    DMVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dwp) & 15) == 0, "conv_up2x_wgrad: buffers must be 16-byte aligned");
    if (!dmvae_conv_up2x_supported(B, H, W, Cin, Cout))
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_up2x_wgrad: shape B=%d H=%d W=%d Cin=%d Cout=%d not supported", B, H, W, Cin, Cout);
    WgGeom g;
    g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.KH = 2; g.KW = 2; g.pt = 0; g.pl = 0;
    g.stride = 1; g.IH = H; g.IW = W; g.up = 1;
    pick_pixel_tile64(H, W, &g.BW, &g.BH);
    g.tiles_w = W / g.BW; g.tiles_h = H / g.BH;
    g.pix_tiles = B * g.tiles_w * g.tiles_h;
    return Cout % 512 == 0 ? launch_wgrad_tc2<2>(x, dy, dwp, g, (cudaStream_t)stream) : launch_wgrad_tc2<1>(x, dy, dwp, g, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- strided convolution (flux_ae.Downsample, :85-95)
// Stride-2 3x3 with arbitrary top/left padding: the output grid is tiled as usual and the A operand is fetched with TMA
// element strides of 2 over the input image (right/bottom zero padding = TMA out-of-bounds fill).
DMVAE_API int dmvae_conv_tc_strided_supported(int B, int IH, int IW, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride) {
    int bw, bh;
    if (stride != 2 || B <= 0 || Cin % 8 != 0 || Cin < 32 || Cout % 8 != 0 || Cout < 32 || KH != 3 || KW != 3) return 0;
    if (!pick_pixel_tile_n(OH, OW, BM, &bw, &bh)) return 0;
    return bw * stride <= 256 && bh * stride <= 256;
}

DMVAE_API int dmvae_conv_tc_fwd_strided(const void* x, const void* w_packed, const float* bias, void* y, int B, int IH, int IW,
                                        int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad_top, int pad_left,
                                        void* stream) {
    DMVAE_CHECK_ARG(x && w_packed && y, "conv_tc_fwd_strided: null pointer");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)y & 15) == 0,
                    "conv_tc_fwd_strided: buffers must be 16-byte aligned");
    if (!dmvae_conv_tc_strided_supported(B, IH, IW, Cin, OH, OW, Cout, KH, KW, stride))
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_tc_fwd_strided: shape B=%d %dx%d->%dx%d Cin=%d Cout=%d s=%d not supported", B, IH, IW, OH, OW, Cin, Cout, stride);
    TcGeom g;
    g.B = B; g.H = OH; g.W = OW; g.Cin = Cin; g.Cout = Cout; g.KH = KH; g.KW = KW;
    g.pt = pad_top; g.pl = pad_left; g.stride = stride; g.IH = IH; g.IW = IW; g.cpg = 0; g.flags = 0; g.up = 0;
    g.k_chunks = (Cin + BK - 1) / BK;
    pick_pixel_tile_n(OH, OW, BM, &g.BW, &g.BH);
    g.tiles_w = OW / g.BW; g.tiles_h = OH / g.BH;
    g.m_tiles = B * g.tiles_w * g.tiles_h;
    cudaStream_t st = (cudaStream_t)stream;
    const int bn = (Cout % 256 == 0 || Cout > 256) ? 256 : 128;
    g.n_tiles = (Cout + bn - 1) / bn;
    if (bn == 256 && Cout % 256 == 0 && g.m_tiles % 2 == 0 && g_force_mt != 1)
        return launch_conv_tc2<256>(x, w_packed, bias, nullptr, y, nullptr, g, st);
    return bn == 256 ? launch_conv_tc<256, 1>(x, w_packed, bias, nullptr, y, nullptr, g, st)
                     : launch_conv_tc<128, 1>(x, w_packed, bias, nullptr, y, nullptr, g, st);
}

// strided weight gradient: dy lives on the OHxOW grid, x is sampled with TMA element strides
DMVAE_API int dmvae_conv_tc_wgrad_strided(const void* x, const void* dy, float* dw_tap_major, int B, int IH, int IW, int Cin,
                                          int OH, int OW, int Cout, int KH, int KW, int stride, int pad_top, int pad_left,
                                          void* stream) {
    DMVAE_CHECK_ARG(x && dy && dw_tap_major, "conv_tc_wgrad_strided: null pointer");
    DMVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dw_tap_major & 15) == 0,
                    "conv_tc_wgrad_strided: buffers must be 16-byte aligned");
    int bw, bh;
    if (stride != 2 || KH != 3 || KW != 3 || Cin % 8 != 0 || Cout % 8 != 0 || Cin < 32 || Cout < 32 || !pick_pixel_tile64(OH, OW, &bw, &bh) ||
        bw * stride > 256 || bh * stride > 256)
        return dmvae_set_error(DMVAE_EUNSUPPORTED, "conv_tc_wgrad_strided: shape B=%d %dx%d->%dx%d Cin=%d Cout=%d s=%d not supported", B, IH, IW, OH, OW, Cin, Cout, stride);
    WgGeom g;
    g.B = B; g.H = OH; g.W = OW; g.Cin = Cin; g.Cout = Cout; g.KH = KH; g.KW = KW;
    g.pt = pad_top; g.pl = pad_left; g.stride = stride; g.IH = IH; g.IW = IW; g.up = 0;
    g.BW = bw; g.BH = bh;
    g.tiles_w = OW / bw; g.tiles_h = OH / bh;
    g.pix_tiles = B * g.tiles_w * g.tiles_h;
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide_m = Cout >= 256 && g_force_mt != 1;
    if (const int pmt = wgrad_pair_mt(Cin, Cout))
        return pmt == 2 ? launch_wgrad_tc2<2>(x, dy, dw_tap_major, g, st) : launch_wgrad_tc2<1>(x, dy, dw_tap_major, g, st);
    if (Cin >= 256) return wide_m ? launch_wgrad_tc<256, 2, 1>(x, dy, dw_tap_major, g, st) : launch_wgrad_tc<256, 1, 1>(x, dy, dw_tap_major, g, st);
    if (wide_m) return launch_wgrad_tc<128, 2, 2>(x, dy, dw_tap_major, g, st);
    return launch_wgrad_tc<128, 1, 3>(x, dy, dw_tap_major, g, st);
}
