// tcgen05.mma kind::i8 M128 N32 K32 issue / completion rate: same accumulator vs rotating accumulators, commit per product
// or once, one or two issuing threads (different warps).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma(uint32_t td, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// nacc: accumulators rotated over (1, 4, 8); commit_each; nissue: issuing threads (lane 0 of warps 0 .. nissue-1), each with
// its own accumulators and slabs
__global__ void __launch_bounds__(128) rate(int reps, int nacc, int commit_each, int nissue, long long* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar[4], bar2[4];
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 16 * 5120; i += 128) sm[i] = (uint8_t)(i * 7);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[i])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && warp < nissue) {
        const uint32_t idesc = (2u << 4) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t base = smem_u32(sm) + warp * 4 * 5120;
        const uint32_t td0 = tmem_base + warp * 128;
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t da = make_desc(base + u * 5120, 128, 512), db = make_desc(base + u * 5120 + 4096, 128, 512);
                mma(td0 + ((r + u) % nacc) * 32 % 128, da, db, idesc);
                if (commit_each) commit(smem_u32(&bar2[warp]));
            }
        }
        long long t1 = clock64();
        commit(smem_u32(&bar[warp]));
        mbar_wait(smem_u32(&bar[warp]), 0);
        long long t2 = clock64();
        out[2 * warp] = t1 - t0;
        out[2 * warp + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 5120);
    const int reps = 4096;
    for (int nissue : {1, 2, 4})
        for (int nacc : {1, 4})
            for (int ce : {0, 1}) {
                rate<<<1, 128, 16 * 5120>>>(reps, nacc, ce, nissue, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
                printf("issuers=%d accumulators=%d commit_each=%d: issue %.1f cyc/mma per issuer, complete %.1f; aggregate %.1f cyc per mma\n", nissue, nacc, ce,
                       (double)h[0] / reps, (double)h[1] / reps, (double)h[2 * (nissue - 1) + 1] / reps / nissue);
            }
    return 0;
}
