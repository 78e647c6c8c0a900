// Microbenchmark: issue-bound rate of tcgen05.mma.kind::i8 (uint8 x uint8 -> int32, accumulators in TMEM) on sm_100a, with the
// operand shapes of mad_b200/csrc/match_u8.cu: one CTA per SM (M = 128, N = 256, K = 32 per instruction) and CTA pairs
// (cta_group::2, M = 256, N = 256).  Operands sit in shared memory (SWIZZLE_128B K-major tiles, contents irrelevant) and are
// re-read by every instruction; nothing is loaded in the timed loop, so this is the tensor pipe's own ceiling for this
// instruction -- the denominator of the matcher's roofline (bench.py reads profiles/r02_i8_mma_rate.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_mma_rate i8_mma_rate.cu && ./i8_mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {     // K-major, 128-byte rows, SWIZZLE_128B, SBO = 1024 B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int NCTA>
__global__ void __launch_bounds__(128, 1) i8_rate_kernel(int iters, int n_cols) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t a_tile = base;                 // 128 rows x 128 B
    const uint32_t b_tile = base + 16384;         // 256 rows x 128 B (pair: each CTA's own 128 rows are used)
    const uint32_t bar = base + 16384 + 32768;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bar + 16 - raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (NCTA == 2) cluster_sync_all();
    if (warp == 0) {
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1 && lane == 0 && rank == 0) {
        // kind::i8: D = s32, A = B = u8, K-major; N >> 3 at bits 17-22, M >> 4 at bits 24-28
        const uint32_t M = NCTA == 2 ? 256 : 128;
        const uint32_t idesc = (2u << 4) | ((uint32_t)(n_cols >> 3) << 17) | ((M >> 4) << 24);
        const uint64_t da = umma_smem_desc(a_tile), db = umma_smem_desc(b_tile);
        for (int it = 0; it < iters; ++it) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((it & 1) * n_cols);       // two accumulators, as the matcher alternates
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {                                         // the four K = 32 steps of a 128-byte row
                if (NCTA == 2)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d_tmem), "l"(da + 2 * kk), "l"(db + 2 * kk), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d_tmem), "l"(da + 2 * kk), "l"(db + 2 * kk), "r"(idesc), "r"(1u) : "memory");
            }
        }
        if (NCTA == 2)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(bar), "h"((uint16_t)3) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    if (warp == 1 && lane == 0) mbar_wait(bar, 0);          // both CTAs of a pair get the multicast commit
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (NCTA == 2) cluster_sync_all();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

template <int NCTA>
static double run(int sms, int iters, int n_cols) {
    const size_t smem = 1024 + 16384 + 32768 + 64;
    auto kern = i8_rate_kernel<NCTA>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NCTA == 2 ? (sms / 2) * 2 : sms);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, kern, iters / 8, n_cols);     // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, kern, iters, n_cols);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double m = NCTA == 2 ? 256.0 : 128.0;
    const double groups = NCTA == 2 ? sms / 2 : sms;                               // issuing CTAs (pairs)
    const double ops = 2.0 * m * n_cols * 128.0 * iters * groups;                  // 4 instructions of K = 32 per iteration
    return ops / (best * 1e-3) / 1e12;
}

int main() {
    int dev = 0, sms = 0, mhz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
    const int iters = 200000;
    const double one128 = run<1>(sms, iters, 128), one256 = run<1>(sms, iters, 256), pair256 = run<2>(sms, iters, 256);
    cudaError_t e = cudaGetLastError();
    printf("tcgen05.mma.kind::i8 u8 x u8 -> s32, %d SMs, max clock %d MHz\n", sms, mhz / 1000);
    printf("cta_group::1  M=128 N=128 K=32 : %8.1f TOP/s  (%.0f ops/clk/SM at max clock)\n", one128, one128 * 1e12 / sms / (mhz * 1e3));
    printf("cta_group::1  M=128 N=256 K=32 : %8.1f TOP/s  (%.0f ops/clk/SM at max clock)\n", one256, one256 * 1e12 / sms / (mhz * 1e3));
    printf("cta_group::2  M=256 N=256 K=32 : %8.1f TOP/s  (%.0f ops/clk/SM at max clock)\n", pair256, pair256 * 1e12 / sms / (mhz * 1e3));
    printf("JSON {\"tops\": %.1f, \"one_cta_n128\": %.1f, \"one_cta_n256\": %.1f, \"cta_pair_n256\": %.1f, \"sms\": %d, \"max_mhz\": %d}\n",
           pair256 > one256 ? pair256 : one256, one128, one256, pair256, sms, mhz / 1000);
    printf("%s\n", e == cudaSuccess ? "no error" : cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
