// 12-bit coding of a row's count vector for the compact transfer format (include/bear_b200.h, wire & 15 == 12).
// A vector (c_0 .. c_{A1-1}) with N = sum c <= nmax is sent as its rank among all such vectors: the vectors with sum n come
// after those with sum n - 1, in lexicographic order within one sum (combinatorial number system: the number of vectors
// of `parts` entries with sum m is C(m + parts - 1, parts - 1)).  A1 = 5: nmax = 10, 3003 vectors; A1 = 21: nmax = 3, 2024
// vectors; code 4095 = "not representable", the row's non-zero counts travel as escape entries.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BEAR_HD __host__ __device__
#else
#define BEAR_HD
#endif

namespace bear_rank {

constexpr uint32_t ESCAPE = 4095u;

// number of vectors of `parts` non-negative entries with sum m
BEAR_HD inline uint32_t compositions(int m, int parts) {
    uint64_t r = 1;                                   // C(m + parts - 1, m), multiplicative form (exact at every step)
    for (int i = 1; i <= m; ++i) r = r * uint64_t(parts - 1 + i) / uint64_t(i);
    return uint32_t(r);
}

// largest N whose vectors (all sums 0..N) fit below the escape code
BEAR_HD inline int nmax(int A1) {
    uint32_t total = 0;
    int n = 0;
    for (;; ++n) {
        const uint32_t c = compositions(n, A1);
        if (total + c > ESCAPE) break;
        total += c;
    }
    return n - 1;
}

// rank of a count vector read with stride `pitch` (elements), or ESCAPE
template <typename T>
BEAR_HD inline uint32_t rank_of(const T* c, int64_t pitch, int A1, int nmax_) {
    uint32_t N = 0;
    for (int b = 0; b < A1; ++b) {
        const uint32_t v = uint32_t(c[int64_t(b) * pitch]);
        if (v > uint32_t(nmax_)) return ESCAPE;
        N += v;
    }
    if (N > uint32_t(nmax_)) return ESCAPE;
    uint32_t r = 0;
    for (uint32_t n = 0; n < N; ++n) r += compositions(int(n), A1);
    int rem = int(N);
    for (int b = 0; b + 1 < A1; ++b) {
        const int cb = int(c[int64_t(b) * pitch]);
        for (int v = 0; v < cb; ++v) r += compositions(rem - v, A1 - 1 - b);
        rem -= cb;
    }
    return r;
}

// inverse: counts of rank r (r below the number of vectors) into out[A1]
BEAR_HD inline void unrank(uint32_t r, int A1, uint8_t* out) {
    int N = 0;
    for (;; ++N) {
        const uint32_t c = compositions(N, A1);
        if (r < c) break;
        r -= c;
    }
    int rem = N;
    for (int b = 0; b + 1 < A1; ++b) {
        int v = 0;
        for (;;) {
            const uint32_t c = compositions(rem - v, A1 - 1 - b);
            if (r < c) break;
            r -= c;
            ++v;
        }
        out[b] = uint8_t(v);
        rem -= v;
    }
    out[A1 - 1] = uint8_t(rem);
}

BEAR_HD inline uint32_t num_vectors(int A1, int nmax_) {
    uint32_t total = 0;
    for (int n = 0; n <= nmax_; ++n) total += compositions(n, A1);
    return total;
}

}  // namespace bear_rank
