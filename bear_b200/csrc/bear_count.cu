// K-mer transition counting on the device: sequences -> packed count table.
// The count table BEAR trains on is defined by brute-force counting over '[' * lag + seq + ']'
// (reference tests/test_summarize.py:96-114); the reference produces it with KMC binaries plus a
// heap merge (summarize.py:103-622).  Here every transition of every sequence is one thread: it packs
// its lag-symbol context into the 2-bit code of the packed table and bumps counts[slot][group][next]
// in an open-addressing hash table keyed by that code (atomicCAS on the key, atomicAdd on the count --
// both native 64/32-bit integer atomics in global memory).
#include "bear_b200.h"
#include "bear_common.cuh"
#include "bear_host.h"

namespace {

using namespace bear;

constexpr int THREADS = 256;
constexpr uint64_t EMPTY = ~0ull;

__device__ __forceinline__ int base_code(uint8_t ch) {
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return -1;
    }
}

// symbol m of sequence [b, b+len) read forward, or of its reverse complement
__device__ __forceinline__ int seq_symbol(const uint8_t* __restrict__ seq, int64_t b, int64_t len, int64_t m, bool rc) {
    const int c = base_code(seq[b + (rc ? len - 1 - m : m)]);
    return (rc && c >= 0) ? 3 - c : c;
}

__global__ void count_transitions_kernel(const uint8_t* __restrict__ seq, const int64_t* __restrict__ offsets,
                                         const int64_t* __restrict__ toff, const int32_t* __restrict__ groups, int64_t nseq,
                                         int64_t ntrans, int lag, int G, int strands, uint64_t* __restrict__ keys,
                                         uint32_t* __restrict__ counts, int64_t cap, unsigned long long* __restrict__ stats) {
    const int64_t total = ntrans * strands;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
        const bool rc = t >= ntrans;
        const int64_t tt = rc ? t - ntrans : t;
        // sequence of this transition: last i with toff[i] <= tt
        int64_t lo = 0, hi = nseq - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (toff[mid] <= tt) lo = mid;
            else hi = mid - 1;
        }
        const int64_t b = offsets[lo], len = offsets[lo + 1] - b;
        const int64_t j = tt - toff[lo];                    // 0..len: position of the predicted symbol
        const int nstart = j < lag ? int(lag - j) : 0;
        uint64_t code = 0;
        bool ok = true;
        for (int p = nstart; p < lag; ++p) {
            const int c = seq_symbol(seq, b, len, j - lag + p, rc);
            ok = ok && c >= 0;
            code = (code << 2) | uint64_t(c & 3);
        }
        int next = 4;                                       // stop symbol
        if (j < len) {
            next = seq_symbol(seq, b, len, j, rc);
            ok = ok && next >= 0;
        }
        if (!ok) {                                          // symbols outside ACGT: the transition is skipped
            atomicAdd(stats + 1, 1ull);
            continue;
        }
        code |= uint64_t(nstart) << 58;
        uint64_t slot = mix64(code) & uint64_t(cap - 1);
        for (;;) {
            const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot), EMPTY, code);
            if (old == EMPTY) atomicAdd(stats, 1ull);       // a new distinct k-mer
            if (old == EMPTY || old == code) break;
            slot = (slot + 1) & uint64_t(cap - 1);
        }
        const uint32_t prev = atomicAdd(counts + (slot * G + groups[lo]) * 5 + next, 1u);
        if (prev == 0xffffffffu) atomicAdd(stats + 2, 1ull);   // 32-bit overflow
    }
}

// occupied slots -> rows of a packed table (kmers[n], counts[G][5][stride]); rows[i] = slot of row i
__global__ void gather_table_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                    const int64_t* __restrict__ rows, int64_t n, int G, int64_t stride,
                                    uint64_t* __restrict__ out_kmers, uint32_t* __restrict__ out_counts) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t s = rows[i];
        out_kmers[i] = keys[s];
        for (int p = 0; p < G * 5; ++p) out_counts[int64_t(p) * stride + i] = counts[s * G * 5 + p];
    }
}

inline int blocks_for(int64_t n) {
    int64_t b = (n + THREADS - 1) / THREADS;
    if (b < 1) b = 1;
    return int(b < 148 * 16 ? b : 148 * 16);
}

}  // namespace

extern "C" int bear_count_transitions(const uint8_t* d_seq, const int64_t* d_offsets, const int64_t* d_toff,
                                      const int32_t* d_groups, int64_t nseq, int64_t ntrans, int lag, int G,
                                      int reverse_complement, uint64_t* d_keys, uint32_t* d_counts, int64_t cap,
                                      uint64_t* d_stats, void* stream) {
    const char* fn = "bear_count_transitions";
    BEAR_REQUIRE(nseq >= 0 && ntrans >= 0 && lag >= 1 && lag <= 29 && G >= 1, fn);
    BEAR_REQUIRE(cap >= 2 && (cap & (cap - 1)) == 0, fn);
    if (nseq == 0 || ntrans == 0) return BEAR_OK;
    BEAR_REQUIRE(d_seq && d_offsets && d_toff && d_groups && d_keys && d_counts && d_stats, fn);
    const int strands = reverse_complement ? 2 : 1;
    count_transitions_kernel<<<blocks_for(ntrans * strands), THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        d_seq, d_offsets, d_toff, d_groups, nseq, ntrans, lag, G, strands, d_keys, d_counts, cap,
        reinterpret_cast<unsigned long long*>(d_stats));
    BEAR_LAUNCH_CHECK("count_transitions_kernel");
    return BEAR_OK;
}

extern "C" int bear_gather_table(const uint64_t* d_keys, const uint32_t* d_counts, const int64_t* d_rows, int64_t n, int G,
                                 int64_t stride, uint64_t* d_out_kmers, uint32_t* d_out_counts, void* stream) {
    const char* fn = "bear_gather_table";
    BEAR_REQUIRE(n >= 0 && G >= 1 && stride >= n, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_keys && d_counts && d_rows && d_out_kmers && d_out_counts, fn);
    gather_table_kernel<<<blocks_for(n), THREADS, 0, static_cast<cudaStream_t>(stream)>>>(d_keys, d_counts, d_rows, n, G, stride,
                                                                                       d_out_kmers, d_out_counts);
    BEAR_LAUNCH_CHECK("gather_table_kernel");
    return BEAR_OK;
}
