// Probe: can a K-major SWIZZLE_128B UMMA A-operand be read from a TMA-written halo tile at an arbitrary 128-byte row offset
// with a stride-byte-offset that is not a multiple of 1024 B?  (Needed to reuse one (BH+2)x(BW+2) activation halo for all nine
// filter taps of a 3x3 convolution instead of re-fetching a shifted tile per tap.)
//   halo tile: HR x HP pixels (rows of 128 B = 64 bf16 channels), written by one TMA box, smem base 1024-aligned
//   A rows m = 0..127 <-> halo pixel ((m / 8) + dy) * HP + (m % 8) + dx       => start = base + (dy*HP+dx)*128, SBO = HP*128
// Prints, per (dy, dx, base_offset mode), the max abs error against the host result.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int HP = 10, HR = 18, NROWS = HP * HR;     // halo pitch (pixels), halo rows
constexpr int N = 64;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out, int off_rows, int sbo_bytes,
      int bo_mode) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                          // NROWS * 128 B = 23040 -> pad to 23552 (1024 multiple)
    uint8_t* sb = smem + 23552;                  // 64 * 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 23552 + 8192);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(NROWS * 128 + N * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sa)), "l"(&map_a), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sb)), "l"(&map_b), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW0:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D0;\n\tbra W0;\n\tD0:\n\t}"
                     ::"r"(smem_u32(bar)) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(sa) + off_rows * 128;
        uint64_t ad = 0;
        ad |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
        ad |= (uint64_t)1 << 16;
        ad |= (uint64_t)(sbo_bytes >> 4) << 32;
        ad |= (uint64_t)1 << 46;
        if (bo_mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
        ad |= (uint64_t)2 << 61;
        uint64_t bd = 0;
        bd |= (uint64_t)((smem_u32(sb) & 0x3FFFF) >> 4);
        bd |= (uint64_t)1 << 16;
        bd |= (uint64_t)(1024 >> 4) << 32;
        bd |= (uint64_t)1 << 46;
        bd |= (uint64_t)2 << 61;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int kk = 0; kk < 4; ++kk) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad + 2 * kk), "l"(bd + 2 * kk), "r"(idesc), "r"(kk) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
    }
    __syncwarp();
    asm volatile("{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}"
                 ::"r"(smem_u32(done)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) out[(warp * 32 + lane) * N + c0 + e] = __uint_as_float(v[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    EncodeFn encode = (EncodeFn)fn;
    std::vector<__nv_bfloat16> ha(NROWS * 64), hb(N * 64);
    std::vector<float> fa(NROWS * 64), fb(N * 64);
    srand(1);
    for (int i = 0; i < NROWS * 64; ++i) { fa[i] = (float)(rand() % 9 - 4); ha[i] = __float2bfloat16(fa[i]); }
    for (int i = 0; i < N * 64; ++i) { fb[i] = (float)(rand() % 5 - 2); hb[i] = __float2bfloat16(fb[i]); }
    __nv_bfloat16 *da, *db; float* dout;
    CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dout, 128 * N * 4));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)NROWS}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)NROWS}; cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
    }
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)N}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)N}; cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    const int smem_bytes = 23552 + 8192 + 64 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    std::vector<float> hout(128 * N);
    // pitch variants: HP (the halo case, SBO = 1280) and 8 (the ordinary dense tile, SBO = 1024)
    for (int pitch : {8, HP}) {
        for (int bo = 0; bo < 2; ++bo) {
            for (int dy = 0; dy < 3; ++dy) for (int dx = 0; dx < 3; ++dx) {
                const int off = dy * pitch + dx;
                CK(cudaMemset(dout, 0xff, 128 * N * 4));
                probe<<<1, 128, smem_bytes>>>(ma, mb, dout, off, pitch * 128, bo);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("pitch %d bo %d dy %d dx %d: launch error %s\n", pitch, bo, dy, dx, cudaGetErrorString(e)); return 2; }
                CK(cudaMemcpy(hout.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0; int bad_rows = 0;
                for (int m = 0; m < 128; ++m) {
                    const int row = (m / 8) * pitch + (m % 8) + off;
                    double rowerr = 0;
                    for (int n = 0; n < N; ++n) {
                        float acc = 0;
                        if (row < NROWS) for (int k = 0; k < 64; ++k) acc += fa[row * 64 + k] * fb[n * 64 + k];
                        else continue;
                        double d = fabs((double)hout[m * N + n] - acc);
                        if (d > rowerr) rowerr = d;
                    }
                    if (rowerr > 0) ++bad_rows;
                    if (rowerr > maxerr) maxerr = rowerr;
                }
                printf("pitch %2d base_offset_mode %d dy %d dx %d (start row %2d): max err %g, bad rows %d/128 -> %s\n", pitch, bo, dy, dx, off,
                       maxerr, bad_rows, maxerr == 0 ? "EXACT" : "MISMATCH");
            }
        }
    }
    return 0;
}
