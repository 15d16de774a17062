// Probe for the tcgen05 (UMMA) pieces the train kernel relies on: kind::i8, MN-major A and B without swizzle, S32
// accumulators in TMEM, several accumulators side by side.  Prints which shared-memory descriptor convention the
// hardware follows.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe.bin umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;     // descriptor version (Blackwell)
    return d;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}

// A: M=128 x K=32 (u8), B: N x K=32 (s8); variant: 0 = LBO is the K-group stride, SBO the MN-group stride; 1 = swapped
template <int N>
__global__ void __launch_bounds__(128) probe(const uint8_t* A, const int8_t* B, int32_t* D, int variant, int reps, int col0) {
    __shared__ __align__(128) uint8_t sA[128 * 32];
    __shared__ __align__(128) uint8_t sB[N * 32];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    // canonical MN-major, no swizzle: core matrix = 8 K-rows of 16 bytes along MN
    for (int i = tid; i < 128 * 32; i += 128) {
        const int m = i / 32, k = i % 32;
        sA[(m / 16) * 512 + (k / 8) * 128 + (k % 8) * 16 + (m % 16)] = A[m * 32 + k];
    }
    for (int i = tid; i < N * 32; i += 128) {
        const int n = i / 32, k = i % 32;
        sB[(n / 16) * 512 + (k / 8) * 128 + (k % 8) * 16 + (n % 16)] = (uint8_t)B[n * 32 + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t lbo = variant == 0 ? 128 : 512, sbo = variant == 0 ? 512 : 128;
        const uint64_t da = make_desc(smem_u32(sA), lbo, sbo), db = make_desc(smem_u32(sB), lbo, sbo);
        // c S32 (2<<4) | a u8 (0<<7) | b s8 (1<<10) | a MN-major (1<<15) | b MN-major (1<<16) | N>>3 <<17 | M>>4 <<24
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        for (int r = 0; r < reps; ++r) {
            const uint32_t acc = r > 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tb + col0), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // lane = m; 32 columns per load
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + col0 + c;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[tid * N + c + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128));
}

template <int N>
int run(int variant, int reps, int col0) {
    static uint8_t hA[128 * 32];
    static int8_t hB[N * 32];
    static int32_t hD[128 * N], ref[128 * N];
    srand(1234 + N);
    for (int i = 0; i < 128 * 32; ++i) hA[i] = (rand() % 3 == 0) ? (uint8_t)(rand() % 200) : 0;
    for (int i = 0; i < N * 32; ++i) hB[i] = (int8_t)(rand() % 256 - 128);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            int64_t s = 0;
            for (int k = 0; k < 32; ++k) s += (int)hA[m * 32 + k] * (int)hB[n * 32 + k];
            ref[m * N + n] = (int32_t)(s * reps);
        }
    uint8_t* dA; int8_t* dB; int32_t* dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof hD);
    probe<N><<<1, 128>>>(dA, dB, dD, variant, reps, col0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d variant=%d reps=%d col0=%d: CUDA error %s\n", N, variant, reps, col0, cudaGetErrorString(e)); return -1; }
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128 * N; ++i) bad += hD[i] != ref[i];
    printf("N=%d variant=%d reps=%d col0=%d: %s (%d of %d mismatches)\n", N, variant, reps, col0, bad ? "MISMATCH" : "match", bad, 128 * N);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad;
}

int main() {
    // (variant 1, LBO / SBO swapped, reads outside the shared-memory window: an illegal-address fault on B200)
    int bad = 0;
    bad += run<32>(0, 1, 0) != 0;
    bad += run<32>(0, 3, 32) != 0;
    bad += run<32>(0, 2, 96) != 0;
    bad += run<64>(0, 1, 0) != 0;
    bad += run<64>(0, 2, 64) != 0;
    printf("convention LBO = K-group stride, SBO = MN-group stride: %s\n", bad ? "FAILED" : "confirmed");
    return bad;
}
