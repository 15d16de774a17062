// Dense (reference-shaped float64 tensor) kernels: unpacking the packed table into the reference's
// tensors, the generic distributions of core.py, and the posterior sampler of log_gamma.py.
#include <math.h>

#include "bear_rank.h"
#include "bear_b200.h"
#include "bear_common.cuh"
#include "bear_host.h"

namespace {

using namespace bear;

constexpr int THREADS = 256;
constexpr int MAX_A1 = 32;

inline int blocks_for(int64_t n, int cap = 148 * 16) {
    int64_t b = (n + THREADS - 1) / THREADS;
    if (b < 1) b = 1;
    return int(b < cap ? b : cap);
}

__device__ __forceinline__ int decode_symbol(uint64_t code, int j, int lag, int alphabet) {
    if (alphabet == BEAR_ALPHABET_PROT) {
        const int c = int((code >> (5 * (lag - 1 - j))) & 31u);
        return c <= 20 ? c : 21;               // 21 = unknown -> all-zero row
    }
    const int nstart = int(code >> 58);
    return j < nstart ? 4 : int((code >> (2 * (lag - 1 - j))) & 3u);
}

// core.tf_one_hot (core.py:156-174) from packed codes.  A warp owns tiles of 32 k-mers: one coalesced load of the codes,
// then the tile's 32 * lag * A1 doubles are written as consecutive 16-byte pairs (the output is 40 * lag bytes per DNA
// k-mer, all of it stores).  Index arithmetic is 32-bit inside a tile: row = pair element / (lag * A1) through an exact
// float reciprocal (elements < 2^13), position / letter through division by the compile-time alphabet size.
template <int A1>
__global__ void __launch_bounds__(THREADS)
decode_onehot_kernel(const uint64_t* __restrict__ kmers, int64_t n, int lag, int alphabet, double* __restrict__ out) {
    const int per_row = lag * A1;
    const float inv = 1.0f / float(per_row);
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5, w0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t ntiles = (n + 31) >> 5;
    for (int64_t t = w0; t < ntiles; t += warps) {
        const int64_t i0 = t << 5;
        const int rows = int(n - i0 < 32 ? n - i0 : 32);
        const uint64_t mine = lane < rows ? __ldg(kmers + i0 + lane) : 0ull;
        double* dst = out + i0 * per_row;                       // 256 * per_row bytes per full tile: 16-byte aligned
        const int total = rows * per_row;
        for (int base = 0; base < total; base += 64) {          // uniform trip count: every lane takes part in the shuffles
            const int e = base + 2 * lane;
            double v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int x = e + h;
                const int row = min(__float2int_rz((float(x) + 0.5f) * inv), 31);
                const int r = x - row * per_row;
                const int j = r / A1, b = r - j * A1;
                const uint64_t code = __shfl_sync(0xffffffffu, mine, row);
                v[h] = (x < total && decode_symbol(code, j, lag, alphabet) == b) ? 1.0 : 0.0;
            }
            if (e + 1 < total) *reinterpret_cast<double2*>(dst + e) = make_double2(v[0], v[1]);
            else if (e < total) dst[e] = v[0];
        }
    }
}

__global__ void decode_symbols_kernel(const uint64_t* __restrict__ kmers, int64_t n, int lag, int alphabet,
                                      uint8_t* __restrict__ out) {
    const int64_t total = n * lag;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = idx / lag;
        const int j = int(idx - i * lag);
        out[idx] = uint8_t(decode_symbol(kmers[i], j, lag, alphabet));
    }
}

// group-planar uint32 [G][A1][stride] -> dense float64 [n, G, A1] (the tensor of dataloader.py:44-46).  A warp owns
// tiles of 32 rows: it reads the G * A1 planes with coalesced 128-byte loads into a shared-memory tile [row][plane]
// (row pitch odd: conflict-free) and writes the tile's 32 * G * A1 doubles as one contiguous run.  GROUPWISE (tiles too
// wide for shared memory: many protein groups): one group at a time, runs of A1 doubles.
template <int A1, bool GROUPWISE>
__global__ void __launch_bounds__(THREADS)
unpack_counts_kernel(const uint32_t* __restrict__ counts, int64_t stride, int64_t n, int G, double* __restrict__ out) {
    extern __shared__ uint32_t unpack_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5, w0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t ntiles = (n + 31) >> 5;
    const int GA = G * A1, W = GROUPWISE ? A1 : GA, pitch = W | 1;
    uint32_t* tile = unpack_smem + warp * 32 * pitch;
    for (int64_t t = w0; t < ntiles; t += warps) {
        const int64_t i0 = t << 5;
        const int rows = int(n - i0 < 32 ? n - i0 : 32);
        for (int g = 0; g < (GROUPWISE ? G : 1); ++g) {
            const uint32_t* src = counts + int64_t(g) * A1 * stride + i0 + lane;
            for (int p = 0; p < W; ++p) tile[lane * pitch + p] = lane < rows ? __ldg(src + p * stride) : 0u;
            __syncwarp();
            for (int e = lane; e < rows * W; e += 32) {
                const int r = e / W, p = e - r * W;
                out[(i0 + r) * GA + g * A1 + p] = double(tile[r * pitch + p]);
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ bool is_count(double v) { return v >= 0.0 && v < 9.0e15 && v == floor(v); }

// lgamma(a + v) - lgamma(a): integer fast path when v is a count, else two lgamma calls as in tf.math.lbeta
__device__ __forceinline__ double lg_diff_real(double a, double v) {
    if (v == 0.0) return 0.0;
    if (is_count(v) && a > 0.0) {
        const LgDg t = lgdg_diff<false>(a, v);
        return t.add + (t.mul == 1.0 ? 0.0 : log(t.mul));
    }
    return lgamma(a + v) - lgamma(a);
}

__device__ __forceinline__ double dg_diff_real(double a, double v) {
    if (v == 0.0) return 0.0;
    if (is_count(v) && a > 0.0) return lgdg_diff<true>(a, v).dg;
    return digamma_pos(a + v) - digamma_pos(a);
}

// core.tfpDirichletMultinomialPerm.counts_log_prob (core.py:73-74)
__global__ void dm_logprob_kernel(const double* __restrict__ conc, int64_t conc_rows, const double* __restrict__ value,
                                  int64_t n, int A1, double* __restrict__ out) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double* a = conc + (i % conc_rows) * A1;
        const double* v = value + i * A1;
        double s = 0.0, tot = 0.0, ll = 0.0;
        for (int b = 0; b < A1; ++b) {
            s += a[b];
            tot += v[b];
            ll += lg_diff_real(a[b], v[b]);
        }
        out[i] = ll - lg_diff_real(s, tot);
    }
}

__global__ void dm_logprob_bwd_kernel(const double* __restrict__ conc, int64_t conc_rows, const double* __restrict__ value,
                                      int64_t n, int A1, const double* __restrict__ gout, double* __restrict__ gconc) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double* a = conc + (i % conc_rows) * A1;
        const double* v = value + i * A1;
        double s = 0.0, tot = 0.0;
        for (int b = 0; b < A1; ++b) {
            s += a[b];
            tot += v[b];
        }
        const double dt = dg_diff_real(s, tot);
        const double go = gout ? gout[i] : 1.0;
        for (int b = 0; b < A1; ++b) gconc[i * A1 + b] = go * (dg_diff_real(a[b], v[b]) - dt);
    }
}

// core.tfpMultinomialPerm.counts_log_prob (core.py:138-139): sum multiply_no_nan(log p, c)
__global__ void mn_logprob_kernel(const double* __restrict__ probs, int64_t probs_rows, const double* __restrict__ value,
                                  int64_t n, int A1, double* __restrict__ out) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double* p = probs + (i % probs_rows) * A1;
        const double* v = value + i * A1;
        double ll = 0.0;
        for (int b = 0; b < A1; ++b)
            if (v[b] != 0.0) ll = fma(v[b], log(p[b]), ll);
        out[i] = ll;
    }
}

__global__ void mn_logprob_bwd_kernel(const double* __restrict__ probs, int64_t probs_rows, const double* __restrict__ value,
                                      int64_t n, int A1, const double* __restrict__ gout, double* __restrict__ gprobs) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double* p = probs + (i % probs_rows) * A1;
        const double* v = value + i * A1;
        const double go = gout ? gout[i] : 1.0;
        for (int b = 0; b < A1; ++b) gprobs[i * A1 + b] = v[b] != 0.0 ? go * v[b] / p[b] : 0.0;
    }
}

// ml_output (core.py:69-71,134-136)
__global__ void ml_output_kernel(const double* __restrict__ x, int64_t n, int A1, double sigma, int64_t seed,
                                 double* __restrict__ out) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double* v = x + i * A1;
        int best = 0;
        double top = -INFINITY;
        for (int b = 0; b < A1; ++b) {
            double t = v[b];
            if (seed >= 0) t += sigma * rng_normal(uint64_t(seed), uint64_t(i), uint64_t(b));
            if (t > top) {
                top = t;
                best = b;
            }
        }
        out[i] = double(best);
    }
}

// log X, X ~ Gamma(a, 1).  a >= 1: Marsaglia-Tsang squeeze; a < 1: log Gamma(a+1) + log(U)/a, which
// stays finite where X itself underflows (the reason log_gamma.py exists, log_gamma.py:17-31).
__device__ double loggamma_draw(double a, uint64_t seed, uint64_t stream) {
    double boost = 0.0;
    uint64_t ctr = 0;
    if (a < 1.0) {
        boost = log(u01(rng_u64(seed, stream, ctr++))) / a;
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        const double x = rng_normal(seed, stream, ctr++);
        const double t = 1.0 + c * x;
        if (t <= 0.0) continue;
        const double v = t * t * t;
        const double u = u01(rng_u64(seed, stream, ctr++));
        const double lv = log(v);
        if (log(u) < 0.5 * x * x + d - d * v + d * lv) return log(d) + lv + boost;
    }
}

__global__ void loggamma_sample_kernel(const double* __restrict__ conc, int64_t n, int64_t n_samples, int64_t seed,
                                       double* __restrict__ out) {
    const int64_t total = n * n_samples;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = idx % n;
        out[idx] = loggamma_draw(conc[i], uint64_t(seed), uint64_t(idx));
    }
}

// x -= logsumexp(x) over groups of A1 (get_var_probs.py:175)
__global__ void log_normalize_kernel(double* __restrict__ x, int64_t groups, int A1) {
    for (int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += int64_t(gridDim.x) * blockDim.x) {
        double* v = x + g * A1;
        double m = -INFINITY;
        for (int b = 0; b < A1; ++b) m = fmax(m, v[b]);
        double s = 0.0;
        for (int b = 0; b < A1; ++b) s += exp(v[b] - m);
        const double lse = m + log(s);
        for (int b = 0; b < A1; ++b) v[b] -= lse;
    }
}


// Deterministic synthetic table (bench / large-scale parity): every cell is a pure function of
// (seed, global row index), so any shard of any size regenerates the same rows.
//   k-mer : a hash bijection of the row index on [0, 4^lag) -- distinct k-mers in pseudo-random
//           order (the reference recommends shuffled tables, docs/usage.rst:191-194); a
//           start_permille/1000 slice gets a start-padded prefix.
//   counts: regime 0 "sparse": N = 1 + Poisson(2) transitions, each to the row's dominant letter
//           with probability 0.7 else uniform over the 4 letters, the last one a stop w.p. 1/150;
//           regime 1 "dense": N = round(LogNormal(ln 300, 1.5)) split by a row-specific profile.
//           Columns g > 0 are binomial thinnings (1/4) of column 0, as ysd1's train:test ratio.
__global__ void synth_table_kernel(uint64_t* __restrict__ kmers, uint32_t* __restrict__ counts, int64_t stride,
                                   int64_t row_begin, int64_t n, int lag, int G, int64_t seed, int regime,
                                   int start_permille) {
    const uint64_t mask = lag >= 29 ? ((1ull << 58) - 1) : ((1ull << (2 * lag)) - 1);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const uint64_t row = uint64_t(row_begin + i);
        // bijection of the row index on [0, 4^lag): odd multiply / xor-shift rounds, each invertible mod 2^(2 lag)
        const int bits = lag >= 29 ? 58 : 2 * lag;
        uint64_t code = (row + 0x632BE59BD9B4E019ull * uint64_t(seed + 1)) & mask;
        for (int round = 0; round < 3; ++round) {
            code = (code * 0x9E3779B97F4A7C15ull) & mask;
            code ^= code >> (bits / 2 + 1);
            code = (code * 0xD1B54A32D192ED03ull) & mask;
            code ^= code >> (bits / 3 + 1);
        }
        if (regime & 2) {
            // sorted order (what KMC / summarize.py emit): strictly increasing distinct codes, one random
            // k-mer out of every block of `gap` consecutive ones; assumes < 2^31 rows per table
            const uint64_t gap = bits > 31 ? (1ull << (bits - 31)) : 1ull;
            code = (row * gap + (gap > 1 ? code % gap : 0ull)) & mask;
        }
        const uint64_t hs = rng_u64(uint64_t(seed), row, 1);
        if (int(hs % 1000) < start_permille) {
            const uint64_t ns = 1 + (hs >> 20) % uint64_t(lag);
            code &= ns >= uint64_t(lag) ? 0ull : ((1ull << (2 * (lag - int(ns)))) - 1);
            code |= ns << 58;
        }
        kmers[i] = code;
        uint32_t c[5] = {0, 0, 0, 0, 0};
        const uint64_t h0 = rng_u64(uint64_t(seed), row, 2);
        const int dom = int(h0 & 3);
        if ((regime & 1) == 0) {
            double u = u01(rng_u64(uint64_t(seed), row, 3));
            int N = 1;
            double p = 0.1353352832366127, cdf = p;     // Poisson(2)
            while (u > cdf && N < 40) { p *= 2.0 / double(N); cdf += p; ++N; }
            for (int t = 0; t < N; ++t) {
                const uint64_t ht = rng_u64(uint64_t(seed), row, 16 + uint64_t(t));
                const int letter = (ht % 10) < 7 ? dom : int((ht >> 8) & 3);
                if (t == N - 1 && (ht >> 16) % 150 == 0) c[4]++;
                else c[letter]++;
            }
        } else {
            const double z = rng_normal(uint64_t(seed), row, 3);
            const double N = rint(exp(5.703782474656201 + 1.5 * z));
            double w[5], ws = 0.0;
            for (int b = 0; b < 5; ++b) {
                w[b] = -log(u01(rng_u64(uint64_t(seed), row, 8 + uint64_t(b)))) * (b == dom ? 3.0 : (b == 4 ? 0.03 : 1.0));
                ws += w[b];
            }
            for (int b = 0; b < 5; ++b) c[b] = uint32_t(fmin(rint(N * w[b] / ws), 4.0e9));
        }
        for (int b = 0; b < 5; ++b) counts[int64_t(b) * stride + i] = c[b];
        for (int g = 1; g < G; ++g) {
            for (int b = 0; b < 5; ++b) {
                uint32_t t = 0;
                if (c[b] <= 64) {
                    for (uint32_t k = 0; k < c[b]; ++k)
                        t += (rng_u64(uint64_t(seed), row, 1000 * uint64_t(g) + 64 * uint64_t(b) + k) & 3) == 0;
                } else {
                    const double m = 0.25 * double(c[b]), sd = sqrt(0.1875 * double(c[b]));
                    const double x = rint(m + sd * rng_normal(uint64_t(seed), row, 1000 * uint64_t(g) + uint64_t(b)));
                    t = uint32_t(fmin(fmax(x, 0.0), double(c[b])));
                }
                counts[(int64_t(g) * 5 + b) * stride + i] = t;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// compact transfer format -> packed table (include/bear_b200.h).  Four rows per thread: one 32-bit load per byte
// plane, 128-bit stores of the count planes.
// ---------------------------------------------------------------------------------------------
// rank -> count vector, 4 bits (A1 = 5, counts <= 10) or 2 bits (A1 = 21, counts <= 3) per letter; rebuilt before every
// expansion that needs it (a few microseconds, stateless, identical values whoever writes them)
__device__ uint64_t g_rank_lut[4096];

__global__ void rank_lut_kernel(int A1, uint32_t nvec) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nvec) return;
    uint8_t c[21];
    bear_rank::unrank(r, A1, c);
    const int cb = A1 == 5 ? 4 : 2;
    uint64_t p = 0;
    for (int b = 0; b < A1; ++b) p |= uint64_t(c[b]) << (cb * b);
    g_rank_lut[r] = p;
}

__global__ void expand_table_kernel(const uint8_t* __restrict__ comp, int64_t pitch, int64_t n, int kb, int kbits_dna_lag,
                                    int nplanes, int A1r, int count_bits, uint64_t* __restrict__ kmers,
                                    uint32_t* __restrict__ counts, int64_t stride, bool vec) {
    const int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;       // rows 4q .. 4q+3
    const int64_t i0 = q * 4;
    if (i0 >= n) return;
    uint64_t v[4] = {0, 0, 0, 0};
    for (int b = 0; b < kb; ++b) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(comp + int64_t(b) * pitch) + q);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] |= uint64_t((w >> (8 * j)) & 0xffu) << (8 * b);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (kbits_dna_lag > 0) {        // DNA / RNA: n_start sits above the 2*lag payload bits (absent when it travels as escapes)
            const uint64_t pay = v[j] & ((uint64_t(1) << (2 * kbits_dna_lag)) - 1);
            v[j] = pay | ((v[j] >> (2 * kbits_dna_lag)) << 58);
        }
        if (i0 + j < n) kmers[i0 + j] = v[j];
    }
    const uint8_t* cplanes = comp + int64_t(kb) * pitch;
    if (count_bits == 12) {             // 12-bit rank of the count vector of a (row, group): byte plane + nibble plane per group
        const int G = nplanes / A1r;
        for (int g = 0; g < G; ++g) {
            const uint8_t* lo8 = cplanes + int64_t(g) * (pitch + pitch / 2);
            const uint32_t wl = __ldg(reinterpret_cast<const uint32_t*>(lo8) + q);
            const uint32_t wh = __ldg(reinterpret_cast<const uint16_t*>(lo8 + pitch) + q);
            uint64_t packed[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t r = ((wl >> (8 * j)) & 0xffu) | (((wh >> (4 * j)) & 0xfu) << 8);
                packed[j] = r == bear_rank::ESCAPE ? 0ull : g_rank_lut[r];      // escaped rows: zeros, patched by the escapes
            }
            const int cb = A1r == 5 ? 4 : 2;                                    // bits per count in a table entry
            for (int b = 0; b < A1r; ++b) {
                uint32_t c[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) c[j] = uint32_t(packed[j] >> (cb * b)) & ((1u << cb) - 1u);
                uint32_t* dst = counts + int64_t(g * A1r + b) * stride + i0;
                if (vec && i0 + 3 < n) {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(c[0], c[1], c[2], c[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + j < n) dst[j] = c[j];
                }
            }
        }
        return;
    }
    const int64_t cpitch = pitch * count_bits / 8;
    for (int pl = 0; pl < nplanes; ++pl) {
        uint32_t c[4];
        if (count_bits == 4) {          // two rows per byte, low nibble = even row: one 16-bit load per four rows
            const uint32_t w = __ldg(reinterpret_cast<const uint16_t*>(cplanes + int64_t(pl) * cpitch) + q);
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = (w >> (4 * j)) & 0xfu;
        } else {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(cplanes + int64_t(pl) * cpitch) + q);
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = (w >> (8 * j)) & 0xffu;
        }
        uint32_t* dst = counts + int64_t(pl) * stride + i0;
        if (vec && i0 + 3 < n) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(c[0], c[1], c[2], c[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i0 + j < n) dst[j] = c[j];
        }
    }
}

__global__ void expand_escapes_kernel(const uint32_t* __restrict__ esc, int64_t n_esc, uint64_t* __restrict__ kmers,
                                      uint32_t* __restrict__ counts, int64_t stride) {
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_esc) return;
    const uint32_t plane = esc[3 * e], row = esc[3 * e + 1], val = esc[3 * e + 2];
    if (plane == 0xffffffffu) kmers[row] |= uint64_t(val) << 58;        // n_start of a start-padded row
    else counts[int64_t(plane) * stride + row] = val;
}

}  // namespace

#define ST(stream) static_cast<cudaStream_t>(stream)

extern "C" int bear_decode_onehot(const uint64_t* d_kmers, int64_t n, int lag, int alphabet, double* d_onehot, void* stream) {
    const char* fn = "bear_decode_onehot";
    BEAR_REQUIRE(bear_alphabet_size(alphabet) > 0 && lag >= 1 && lag <= bear_max_lag(alphabet) && n >= 0, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_kmers && d_onehot, fn);
    const int grid = blocks_for(n, 148 * 8);                     // one warp per 32 k-mers
    if (alphabet == BEAR_ALPHABET_PROT)
        decode_onehot_kernel<21><<<grid, THREADS, 0, ST(stream)>>>(d_kmers, n, lag, alphabet, d_onehot);
    else
        decode_onehot_kernel<5><<<grid, THREADS, 0, ST(stream)>>>(d_kmers, n, lag, alphabet, d_onehot);
    BEAR_LAUNCH_CHECK("decode_onehot_kernel");
    return BEAR_OK;
}

extern "C" int bear_decode_symbols(const uint64_t* d_kmers, int64_t n, int lag, int alphabet, uint8_t* d_sym, void* stream) {
    const char* fn = "bear_decode_symbols";
    BEAR_REQUIRE(bear_alphabet_size(alphabet) > 0 && lag >= 1 && lag <= bear_max_lag(alphabet) && n >= 0, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_kmers && d_sym, fn);
    decode_symbols_kernel<<<blocks_for(n * lag), THREADS, 0, ST(stream)>>>(d_kmers, n, lag, alphabet, d_sym);
    BEAR_LAUNCH_CHECK("decode_symbols_kernel");
    return BEAR_OK;
}

extern "C" int bear_unpack_counts(const uint32_t* d_counts, int64_t stride, int64_t row0, int64_t n, int G, int A1,
                                  double* d_out, void* stream) {
    const char* fn = "bear_unpack_counts";
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n && G >= 1 && (A1 == 5 || A1 == 21), fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_counts && d_out, fn);
    const int grid = blocks_for(n, 148 * 8);                     // one warp per 32 rows
    const bool wide = G * A1 > 80;                               // whole-row tiles up to 8 DNA groups / 3 protein groups
    const size_t smem = size_t(THREADS / 32) * 32 * ((wide ? A1 : G * A1) | 1) * sizeof(uint32_t);
    const uint32_t* src = d_counts + row0;
    if (A1 == 5 && !wide) {
        BEAR_CUDA_CHECK(cudaFuncSetAttribute(unpack_counts_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        unpack_counts_kernel<5, false><<<grid, THREADS, smem, ST(stream)>>>(src, stride, n, G, d_out);
    } else if (A1 == 5) {
        unpack_counts_kernel<5, true><<<grid, THREADS, smem, ST(stream)>>>(src, stride, n, G, d_out);
    } else if (!wide) {
        BEAR_CUDA_CHECK(cudaFuncSetAttribute(unpack_counts_kernel<21, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        unpack_counts_kernel<21, false><<<grid, THREADS, smem, ST(stream)>>>(src, stride, n, G, d_out);
    } else {
        unpack_counts_kernel<21, true><<<grid, THREADS, smem, ST(stream)>>>(src, stride, n, G, d_out);
    }
    BEAR_LAUNCH_CHECK("unpack_counts_kernel");
    return BEAR_OK;
}

#define BEAR_DENSE_ARGS(fn, a, rows, v, n, A1)                                                    \
    BEAR_REQUIRE(n >= 0 && rows >= 1 && A1 >= 1 && A1 <= MAX_A1 && (n % rows) == 0, fn);          \
    if (n == 0) return BEAR_OK;                                                                   \
    BEAR_REQUIRE(a && v, fn)

extern "C" int bear_dm_logprob(const double* d_conc, int64_t conc_rows, const double* d_value, int64_t n, int A1,
                               double* d_out, void* stream) {
    BEAR_DENSE_ARGS("bear_dm_logprob", d_conc, conc_rows, d_value, n, A1);
    BEAR_REQUIRE(d_out != nullptr, "bear_dm_logprob");
    dm_logprob_kernel<<<blocks_for(n), THREADS, 0, ST(stream)>>>(d_conc, conc_rows, d_value, n, A1, d_out);
    BEAR_LAUNCH_CHECK("dm_logprob_kernel");
    return BEAR_OK;
}

extern "C" int bear_dm_logprob_bwd(const double* d_conc, int64_t conc_rows, const double* d_value, int64_t n, int A1,
                                   const double* d_gout, double* d_gconc, void* stream) {
    BEAR_DENSE_ARGS("bear_dm_logprob_bwd", d_conc, conc_rows, d_value, n, A1);
    BEAR_REQUIRE(d_gconc != nullptr, "bear_dm_logprob_bwd");
    dm_logprob_bwd_kernel<<<blocks_for(n), THREADS, 0, ST(stream)>>>(d_conc, conc_rows, d_value, n, A1, d_gout, d_gconc);
    BEAR_LAUNCH_CHECK("dm_logprob_bwd_kernel");
    return BEAR_OK;
}

extern "C" int bear_mn_logprob(const double* d_probs, int64_t probs_rows, const double* d_value, int64_t n, int A1,
                               double* d_out, void* stream) {
    BEAR_DENSE_ARGS("bear_mn_logprob", d_probs, probs_rows, d_value, n, A1);
    BEAR_REQUIRE(d_out != nullptr, "bear_mn_logprob");
    mn_logprob_kernel<<<blocks_for(n), THREADS, 0, ST(stream)>>>(d_probs, probs_rows, d_value, n, A1, d_out);
    BEAR_LAUNCH_CHECK("mn_logprob_kernel");
    return BEAR_OK;
}

extern "C" int bear_mn_logprob_bwd(const double* d_probs, int64_t probs_rows, const double* d_value, int64_t n, int A1,
                                   const double* d_gout, double* d_gprobs, void* stream) {
    BEAR_DENSE_ARGS("bear_mn_logprob_bwd", d_probs, probs_rows, d_value, n, A1);
    BEAR_REQUIRE(d_gprobs != nullptr, "bear_mn_logprob_bwd");
    mn_logprob_bwd_kernel<<<blocks_for(n), THREADS, 0, ST(stream)>>>(d_probs, probs_rows, d_value, n, A1, d_gout, d_gprobs);
    BEAR_LAUNCH_CHECK("mn_logprob_bwd_kernel");
    return BEAR_OK;
}

extern "C" int bear_ml_output(const double* d_x, int64_t n, int A1, double sigma, int64_t seed, double* d_out, void* stream) {
    const char* fn = "bear_ml_output";
    BEAR_REQUIRE(n >= 0 && A1 >= 1, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_x && d_out, fn);
    ml_output_kernel<<<blocks_for(n), THREADS, 0, ST(stream)>>>(d_x, n, A1, sigma, seed, d_out);
    BEAR_LAUNCH_CHECK("ml_output_kernel");
    return BEAR_OK;
}

extern "C" int bear_loggamma_sample(const double* d_conc, int64_t n, int64_t n_samples, int64_t seed, double* d_out, void* stream) {
    const char* fn = "bear_loggamma_sample";
    BEAR_REQUIRE(n >= 0 && n_samples >= 0, fn);
    if (n == 0 || n_samples == 0) return BEAR_OK;
    BEAR_REQUIRE(d_conc && d_out, fn);
    loggamma_sample_kernel<<<blocks_for(n * n_samples), THREADS, 0, ST(stream)>>>(d_conc, n, n_samples, seed, d_out);
    BEAR_LAUNCH_CHECK("loggamma_sample_kernel");
    return BEAR_OK;
}

extern "C" int bear_log_normalize(double* d_x, int64_t n_groups, int A1, void* stream) {
    const char* fn = "bear_log_normalize";
    BEAR_REQUIRE(n_groups >= 0 && A1 >= 1, fn);
    if (n_groups == 0) return BEAR_OK;
    BEAR_REQUIRE(d_x != nullptr, fn);
    log_normalize_kernel<<<blocks_for(n_groups), THREADS, 0, ST(stream)>>>(d_x, n_groups, A1);
    BEAR_LAUNCH_CHECK("log_normalize_kernel");
    return BEAR_OK;
}

extern "C" int bear_synth_table(uint64_t* d_kmers, uint32_t* d_counts, int64_t stride, int64_t row_begin, int64_t n,
                                int lag, int G, int64_t seed, int regime, int start_permille, void* stream) {
    const char* fn = "bear_synth_table";
    BEAR_REQUIRE(n >= 0 && stride >= n && lag >= 1 && lag <= 29 && G >= 1 && row_begin >= 0, fn);
    BEAR_REQUIRE(regime >= 0 && regime <= 3, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_kmers && d_counts, fn);
    synth_table_kernel<<<blocks_for(n), THREADS, 0, ST(stream)>>>(d_kmers, d_counts, stride, row_begin, n, lag, G, seed,
                                                                  regime, start_permille);
    BEAR_LAUNCH_CHECK("synth_table_kernel");
    return BEAR_OK;
}

extern "C" int bear_expand_table(const uint8_t* d_compact, const uint32_t* d_esc, int64_t n_esc, int64_t n, int lag,
                                 int alphabet, int G, int wire, uint64_t* d_kmers, uint32_t* d_counts,
                                 int64_t stride, int64_t dst_row0, void* stream) {
    const char* fn = "bear_expand_table";
    const int a = bear_alphabet_size(alphabet);
    const int count_bits = wire & 15;
    const bool start_esc = (wire & BEAR_WIRE_START_ESC) != 0;
    BEAR_REQUIRE(a > 0 && lag >= 1 && lag <= bear_max_lag(alphabet) && G >= 1 && (count_bits == 4 || count_bits == 8 || count_bits == 12), fn);
    BEAR_REQUIRE((wire & ~(15 | BEAR_WIRE_START_ESC)) == 0 && !(start_esc && alphabet == BEAR_ALPHABET_PROT), fn);
    BEAR_REQUIRE(n >= 0 && n_esc >= 0 && dst_row0 >= 0 && stride >= dst_row0 + n, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_compact && d_kmers && d_counts && (d_esc || n_esc == 0), fn);
    const bool dna = alphabet != BEAR_ALPHABET_PROT;
    const int kb = ((dna ? 2 * lag + (start_esc ? 0 : 6) : 5 * lag) + 7) / 8;
    const int64_t pitch = (n + 15) / 16 * 16;
    const int64_t quads = (n + 3) / 4;
    const bool vec = (dst_row0 & 3) == 0 && (stride & 3) == 0 && (reinterpret_cast<uintptr_t>(d_counts) & 15) == 0;
    if (count_bits == 12) {
        const uint32_t nvec = bear_rank::num_vectors(a + 1, bear_rank::nmax(a + 1));
        rank_lut_kernel<<<(nvec + THREADS - 1) / THREADS, THREADS, 0, ST(stream)>>>(a + 1, nvec);
        BEAR_LAUNCH_CHECK("rank_lut_kernel");
    }
    expand_table_kernel<<<unsigned((quads + THREADS - 1) / THREADS), THREADS, 0, ST(stream)>>>(
        d_compact, pitch, n, kb, dna ? lag : 0, G * (a + 1), a + 1, count_bits, d_kmers + dst_row0, d_counts + dst_row0, stride, vec);
    BEAR_LAUNCH_CHECK("expand_table_kernel");
    if (n_esc > 0) {
        expand_escapes_kernel<<<unsigned((n_esc + THREADS - 1) / THREADS), THREADS, 0, ST(stream)>>>(
            d_esc, n_esc, d_kmers + dst_row0, d_counts + dst_row0, stride);
        BEAR_LAUNCH_CHECK("expand_escapes_kernel");
    }
    return BEAR_OK;
}
