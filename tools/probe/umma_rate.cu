// Throughput of tcgen05.mma kind::i8 M128 x N x K32 from shared memory: MN-major vs K-major operands, N = 32 / 64.
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
// mode 0: both MN-major; 1: both K-major; 2: A MN-major, B K-major.  commit_each: a commit after every mma
__global__ void __launch_bounds__(128) rate(int N, int M, int mode, int reps, int commit_each, int nslab, int kind, int accum, long long* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < nslab * 8192; i += 128) sm[i] = (uint8_t)(i * 7);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (threadIdx.x == 0) {
        const uint32_t amn = mode != 1, bmn = mode == 0;
        // kind 0: i8 (u8 x s8 -> s32); 1: f16 (f16 x f16 -> f32); 2: f8f6f4 (e4m3 x e4m3 -> f32)
        const uint32_t idesc = (kind == 0 ? ((2u << 4) | (1u << 10)) : (1u << 4)) | (amn << 15) | (bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t base = smem_u32(sm);
        const uint64_t da = amn ? make_desc(base, 128, 512) : make_desc(base, 128, 256);
        const uint64_t db = bmn ? make_desc(base + 4096, 128, 512) : make_desc(base + 4096, 128, 256);
        const uint32_t td = tmem_base;
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (kind == 0)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
                else
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
                if (commit_each)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192);
    const int reps = 4096;
    for (int kind : {0, 1})
        for (int M : {64, 128})
        for (int N : {16, 32, 64, 128, 256})
            for (int mode : {0, 1})
                for (int ce : {0, 1}) {
                    if (mode == 0 && N > 128) continue;     // (the probe's operand slab holds 8 N-groups)
                    rate<<<1, 128, 16 * 8192>>>(N, M, mode, reps, ce, 4, kind, 1, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("kind=%s M=%d N=%d %s commit_each=%d: issue %.1f cyc/mma, complete %.1f cyc/mma\n",
                           kind == 0 ? "i8" : "f16", M, N, mode == 0 ? "MN-major" : "K-major", ce,
                           (double)h[0] / reps, (double)h[1] / reps);
                }
    return 0;
}
