// Probe: issue pace of back-to-back tcgen05.mma (SS mode, bf16, M=128 per CTA, K=16) as a function of N and of how many MMAs
// share one tcgen05.commit; operands stay resident in shared memory (no TMA in the loop), one CTA per SM.
// Prints cycles per MMA and the implied fraction of the 128*N*16 MAC / (4096 MAC/clk) floor.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int N>
__global__ void __launch_bounds__(128, 1) pace(long long* cycles, int iters, int per_commit, int stages) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    // stages x (A 16 KB + B N*128 B)
    const int stage_bytes = 16384 + N * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    for (int i = threadIdx.x; i < stages * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        long long t0 = clock64();
        int stage = 0, since = 0, grp = 0;
        auto wait_grp = [&](int gidx) {      // group gidx committed to bar[gidx & 1]; its phase parity is (gidx >> 1) & 1
            asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DL;\n\tbra WL;\n\tDL:\n\t}"
                         ::"r"(smem_u32(bar + (gidx & 1))), "r"((uint32_t)((gidx >> 1) & 1)) : "memory");
        };
        for (int i = 0; i < iters; ++i) {
            const uint32_t sa = smem_u32(smem + stage * stage_bytes);
            const uint64_t ad = desc(sa) + 2 * (i & 3), bd = desc(sa + 16384) + 2 * (i & 3);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem + (uint32_t)((i >> 2) & 1) * N), "l"(ad), "l"(bd), "r"(idesc), "r"(i) : "memory");
            if ((i & 3) == 3 && ++stage == stages) stage = 0;
            if (++since == per_commit || i + 1 == iters) {
                since = 0;
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + (grp & 1))) : "memory");
                if (grp >= 1) wait_grp(grp - 1);      // like a producer waiting for a slot released one group ago: the pipe never drains
                ++grp;
            }
        }
        wait_grp(grp - 1);
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int N>
void run(int grid, int iters, int per_commit, int stages) {
    long long* d; CK(cudaMalloc(&d, grid * 8));
    const int smem_bytes = stages * (16384 + N * 128) + 1024 + 64;
    CK(cudaFuncSetAttribute(pace<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    pace<N><<<grid, 128, smem_bytes>>>(d, iters, per_commit, stages);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    pace<N><<<grid, 128, smem_bytes>>>(d, iters, per_commit, stages);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; CK(cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double cyc = avg / iters, floor_c = 128.0 * N * 16 / 4096 / 2;   // 8192 MAC/clk/SM dense bf16 = 2.25 PF / 148 / ~1.86 GHz
    const double tflops = 2.0 * 128 * N * 16 * iters * grid / (ms * 1e-3) / 1e12;
    printf("N %3d grid %3d stages %d per_commit %2d: %.1f clk/MMA (floor %.0f, %.0f%%), %.0f TFLOP/s chip-wide by wall time\n", N, grid, stages,
           per_commit, cyc, floor_c, 100.0 * floor_c / cyc, tflops);
    fflush(stdout);
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        for (int pc : {4, 8, 64}) {
            run<64>(grid, 20000, pc, 4);
            run<128>(grid, 20000, pc, 4);
            run<256>(grid, 20000, pc, 4);
        }
    }
    return 0;
}
