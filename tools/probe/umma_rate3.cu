// tcgen05.mma kind::i8 M128 N32 K32 rate on a FULL grid (148 CTAs x 512 threads) while the other warps of the CTA are idle,
// read shared memory, write shared memory or do float64 arithmetic: what slows the products inside the train kernel?
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
__global__ void __launch_bounds__(512) rate(int reps, int nissue, int load, int commit_each, int nslab, long long* out, double* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar[4], bar2[4];
    __shared__ uint32_t tmem_base;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < 16 * 5120; i += 512) sm[i] = (uint8_t)(i * 7);
    if (threadIdx.x == 0) {
        stop = 0;
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 16 - nissue) {
        const int qi = warp - (16 - nissue);
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t base = smem_u32(sm);
            const uint32_t td0 = tmem_base + qi * 128;
            long long t0 = clock64();
            int slab = qi;
            for (int r = 0; r < reps; ++r) {
                const uint64_t da = make_desc(base + slab * 5120, 128, 512), db = make_desc(base + slab * 5120 + 4096, 128, 512);
                mma(td0 + (r & 3) * 32, da, db, idesc);
                if (commit_each) commit(smem_u32(&bar2[qi]));
                slab += nissue;
                if (slab >= nslab) slab = qi;
            }
            long long t1 = clock64();
            commit(smem_u32(&bar[qi]));
            mbar_wait(smem_u32(&bar[qi]), 0);
            long long t2 = clock64();
            if (blockIdx.x == 0) {
                out[2 * qi] = t1 - t0;
                out[2 * qi + 1] = t2 - t0;
            }
            atomicAdd((int*)&stop, 1);
        }
    } else if (load) {
        double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
        uint32_t addr = smem_u32(sm) + 12 * 5120 + ((threadIdx.x * 16) & 8191);   // slabs 12..13: away from the operands
        uint32_t x = threadIdx.x;
        while (stop < nissue) {
            if (load == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    uint32_t v0, v1, v2, v3;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr + ((x * 16) & 4095)));
                    x += v0 + i;
                }
            } else if (load == 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" :: "r"(addr + (((x + i * 37) * 16) & 4095)), "r"(x) : "memory");
                }
                x += 17;
            } else if (load == 3) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { a = fma(a, b, 0.25); b = fma(b, a, 0.125); }
            } else {
                // load 4: the operand slabs themselves are rewritten (generic proxy) + fence, as the producers do
#pragma unroll
                for (int i = 0; i < 10; ++i)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" :: "r"(smem_u32(sm) + (warp % nslab) * 5120 + i * 512 + lane * 16), "r"(x & 0x01010101u) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                x += 17;
#pragma unroll
                for (int i = 0; i < 64; ++i) { a = fma(a, b, 0.25); b = fma(b, a, 0.125); }
            }
        }
        if (a + b + x == 1234.5) sink[0] = a;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    double* sink; cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 5120);
    const int reps = 4096;
    const char* names[] = {"idle", "LDS.128", "STS.128", "DFMA", "slab rewrite + fence + DFMA"};
    for (int grid : {1, 148})
        for (int load = 0; load < 5; ++load)
            for (int ni : {1, 2, 4})
                for (int ce : {0, 1}) {
                    rate<<<grid, 512, 16 * 5120>>>(reps, ni, load, ce, 12, d, sink);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
                    printf("grid=%d others=%s issuers=%d commit_each=%d: issue %.1f cyc/mma per issuer, complete %.1f; aggregate %.1f cyc per mma\n",
                           grid, names[load], ni, ce, (double)h[0] / reps, (double)h[1] / reps, (double)h[1] / reps / ni);
                }
    return 0;
}
