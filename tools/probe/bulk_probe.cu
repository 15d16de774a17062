// Probe: 1-D bulk async copies (TMA, cp.async.bulk -> UBLKCP) of small planes into shared memory, completion on an
// mbarrier with expect_tx.  Several copies on one barrier, re-armed over phases.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__global__ void k(const uint64_t* km, const uint32_t* col, int64_t stride, int ntiles, uint64_t* out) {
    __shared__ __align__(128) unsigned char buf[2][256 + 5 * 128];
    __shared__ __align__(8) uint64_t bar[2];
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    auto issue = [&](int t) {
        const int s = t & 1;
        const uint32_t b = smem_u32(&bar[s]), d = smem_u32(buf[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(256 + 5 * 128) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(km + t * 32), "r"(256), "r"(b) : "memory");
        for (int c = 0; c < 5; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(d + 256 + c * 128), "l"(col + c * stride + t * 32), "r"(128), "r"(b) : "memory");
    };
    if (lane == 0) issue(0);
    uint64_t acc = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        if (lane == 0 && t + 1 < ntiles) issue(t + 1);
        mbar_wait(smem_u32(&bar[s]), (t >> 1) & 1);
        const uint64_t code = reinterpret_cast<const uint64_t*>(buf[s])[lane];
        acc += code;
        for (int c = 0; c < 5; ++c) acc += (uint64_t)reinterpret_cast<const uint32_t*>(buf[s] + 256)[c * 32 + lane] << (8 * c);
        __syncwarp();      // every lane has read stage s before lane 0 re-arms it in the next iteration
    }
    out[lane] = acc;
}
int main() {
    const int nt = 37; const int64_t n = nt * 32, stride = n + 32;
    uint64_t* hk = new uint64_t[n]; uint32_t* hc = new uint32_t[5 * stride];
    for (int64_t i = 0; i < n; ++i) hk[i] = 0x9E3779B97F4A7C15ull * (i + 1);
    for (int64_t i = 0; i < 5 * stride; ++i) hc[i] = (uint32_t)(i * 2654435761u) >> 20;
    uint64_t ref[32] = {0};
    for (int64_t i = 0; i < n; ++i) { ref[i % 32] += hk[i]; for (int c = 0; c < 5; ++c) ref[i % 32] += (uint64_t)hc[c * stride + i] << (8 * c); }
    uint64_t *dk, *dout; uint32_t* dc;
    cudaMalloc(&dk, n * 8); cudaMalloc(&dc, 5 * stride * 4); cudaMalloc(&dout, 32 * 8);
    cudaMemcpy(dk, hk, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dc, hc, 5 * stride * 4, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(dk, dc, stride, nt, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("bulk probe: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    uint64_t got[32]; cudaMemcpy(got, dout, sizeof got, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 32; ++i) bad += got[i] != ref[i];
    printf("bulk probe: %s\n", bad ? "MISMATCH" : "match");
    return bad;
}
