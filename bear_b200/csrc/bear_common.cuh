// Device-side math shared by the BEAR kernels: differences of lgamma / digamma at integer offsets,
// block reductions, counter-based RNG.  float64 throughout (reference default precision).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BEAR_EPS 1e-7          // tf.keras.backend.epsilon(), core.py:8 / bear_net.py:4
#define BEAR_XMIN 10.0         // asymptotic series are used for arguments >= BEAR_XMIN
#define BEAR_NSMALL 10         // rising-factorial path for integer offsets <= BEAR_NSMALL

namespace bear {

// lgamma(x) = (x-1/2) ln x - x + ln(2 pi)/2 + stirling_corr(1/x);  |err| < 1e-16 for x >= 10
__device__ __forceinline__ double stirling_corr(double rx) {
    const double r2 = rx * rx;
    double p = 1.0 / 156.0;
    p = fma(p, r2, -691.0 / 360360.0);
    p = fma(p, r2, 1.0 / 1188.0);
    p = fma(p, r2, -1.0 / 1680.0);
    p = fma(p, r2, 1.0 / 1260.0);
    p = fma(p, r2, -1.0 / 360.0);
    p = fma(p, r2, 1.0 / 12.0);
    return p * rx;
}

// digamma(x) = ln x - 1/(2x) - digamma_tail(1/x);  |err| < 1e-15 for x >= 10
__device__ __forceinline__ double digamma_tail(double rx) {
    const double r2 = rx * rx;
    double p = 1.0 / 12.0;
    p = fma(p, r2, -691.0 / 32760.0);
    p = fma(p, r2, 1.0 / 132.0);
    p = fma(p, r2, -1.0 / 240.0);
    p = fma(p, r2, 1.0 / 252.0);
    p = fma(p, r2, -1.0 / 120.0);
    p = fma(p, r2, 1.0 / 12.0);
    return p * r2;
}

// lgamma(a + c) - lgamma(a) = add + log(mul)   and   digamma(a + c) - digamma(a) = dg
// for a > 0 and an integer-valued offset c >= 0 (held in a double).  The log of the rising
// factorial is left to the caller (`mul`) so that several terms can share one log().
//   c == 0            : exactly 0 (the reference's lgamma(a+0)-lgamma(a) is exactly 0 too)
//   c <= BEAR_NSMALL  : rising factorial  P = prod_{i<c}(a+i),  P'/P = sum 1/(a+i)
//   otherwise         : shift a up to x1 = a+m >= BEAR_XMIN with a rising factorial, then the
//                       cancellation-free Stirling difference
//                         (x1-1/2) log1p(c'/x1) + c' (ln x2 - 1) + S(x2) - S(x1),  c' = c - m
// The reference evaluates the same difference as two separate lgamma calls inside tf.math.lbeta
// (core.py:60-62,74 via TFP); this form has less rounding noise, not more.
struct LgDg {
    double add, mul, dg;
};

template <bool GRAD>
__device__ __noinline__ LgDg lgdg_diff(double a, double c) {
    LgDg r;
    r.add = 0.0;
    r.mul = 1.0;
    r.dg = 0.0;
    if (c == 0.0) return r;
    if (c <= double(BEAR_NSMALL)) {
        double P = a, D = 1.0;
        const int n = int(c);
        for (int i = 1; i < n; ++i) {
            const double t = a + double(i);
            if (GRAD) D = fma(D, t, P);
            P *= t;
        }
        r.mul = P;
        if (GRAD) r.dg = D / P;
        return r;
    }
    double x1 = a, P = 1.0, D = 0.0;
    int m = 0;
    if (a < BEAR_XMIN) {
        m = int(ceil(BEAR_XMIN - a));   // 1..10, and m < c
        P = a;
        D = 1.0;
        for (int i = 1; i < m; ++i) {
            const double t = a + double(i);
            if (GRAD) D = fma(D, t, P);
            P *= t;
        }
        x1 = a + double(m);
    }
    const double cp = c - double(m);
    const double x2 = a + c;
    const double r1 = 1.0 / x1, r2 = 1.0 / x2;
    const double l1p = log1p(cp * r1);          // ln(x2 / x1)
    const double lx2 = log(x2);
    r.add = fma(x1 - 0.5, l1p, cp * (lx2 - 1.0)) + (stirling_corr(r2) - stirling_corr(r1));
    r.mul = P;
    if (GRAD) {
        r.dg = l1p + 0.5 * cp * r1 * r2 - (digamma_tail(r2) - digamma_tail(r1));
        if (m) r.dg += D / P;
    }
    return r;
}

// lgamma(a + c) - lgamma(a) for a ROW-INDEPENDENT a (a BMM prior, the concentration sum 1/h + 5 eps) and a large
// count: x = a + c >= 64, so three terms of the Stirling series are exact to 1e-16, and lgamma(a) enters as the
// precomputed constant K = ln(2 pi)/2 - lgamma(a).  Only for a < 64 (then lgamma(a) is small next to the result and
// nothing cancels); larger a go through lgdg_diff.
#define BEAR_LARGE_C 64.0
__device__ __forceinline__ double lg_shift_large(double a, double c, double K) {
    const double x = a + c, rx = 1.0 / x, r2 = rx * rx;
    return fma(x - 0.5, log(x), K - x) + rx * fma(r2, fma(r2, 1.0 / 1260.0, -1.0 / 360.0), 1.0 / 12.0);
}
__device__ __forceinline__ double lg_shift_const(double a) { return 0.91893853320467274178 - lgamma(a); }

// log() for paths that are rarely taken inside the hot loops of the fused kernels (running-product rescues, fall-backs for
// large counts): out of line, because those loops are large enough for instruction fetch to show up in the stall profile.
static __device__ __noinline__ double log_cold(double x) { return log(x); }

// Accumulates sum_i (add_i + log mul_i) with as few log() calls as possible.
struct LogProd {
    double add = 0.0, mul = 1.0;
    __device__ __forceinline__ void push(double a, double m) {
        add += a;
        if (mul > 1e40 || mul < 1e-40) {
            add += log_cold(mul);
            mul = 1.0;
        }
        mul *= m;
    }
    __device__ __forceinline__ void push(const LgDg& t) { push(t.add, t.mul); }
};

// Long-running variant for per-thread accumulators spanning many rows: when the running product leaves
// [1e-40, 1e40] its binary exponent is moved into an integer (a handful of integer instructions) instead
// of calling log(); value() = add + ex ln 2 + log(mul).
struct LogProdLong {
    double add = 0.0, mul = 1.0;
    int ex = 0;
    __device__ __forceinline__ void push(double a, double m) {
        add += a;
        if (mul > 1e40 || mul < 1e-40) {
            if (mul > 1e-300 && mul < 1e300) {
                const int hi = __double2hiint(mul);
                ex += ((hi >> 20) & 0x7ff) - 1023;
                mul = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(mul));
            } else {                       // zero / subnormal / inf: keep the reference's -inf / inf
                add += log_cold(mul);
                mul = 1.0;
            }
        }
        mul *= m;
    }
    __device__ __forceinline__ double value() const {
        return add + (double(ex) * 0.69314718055994530942 + (mul == 1.0 ? 0.0 : log(mul)));
    }
};

// numerator / denominator pair: value = num - den
__device__ __forceinline__ double logprod_diff(const LogProd& num, const LogProd& den) {
    // both |log10 mul| <= 140 by construction, so the ratio cannot overflow
    const double ratio = num.mul / den.mul;
    return (num.add - den.add) + (ratio == 1.0 ? 0.0 : log_cold(ratio));
}

// digamma(x) for x > 0 (generic dense path: gradients of lgamma at real-valued arguments)
__device__ __forceinline__ double digamma_pos(double x) {
    double shift = 0.0;
    while (x < BEAR_XMIN) {
        shift += 1.0 / x;
        x += 1.0;
    }
    const double rx = 1.0 / x;
    return log(x) - 0.5 * rx - digamma_tail(rx) - shift;
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in thread 0.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// counter-based RNG (splitmix64 finaliser chain); every draw is a pure function of its counter,
// so results do not depend on the launch geometry.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t rng_u64(uint64_t seed, uint64_t a, uint64_t b) {
    return mix64(mix64(mix64(seed) ^ a) ^ (b * 0xD1B54A32D192ED03ull));
}

// uniform in (0, 1)
__device__ __forceinline__ double u01(uint64_t bits) {
    return (double(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

static __device__ __noinline__ double rng_normal(uint64_t seed, uint64_t a, uint64_t b) {
    const uint64_t x = rng_u64(seed, a, b);
    const double u1 = u01(x), u2 = u01(mix64(x));
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// standard normal for the argmax tie-breaking noise only (Box-Muller on two 24-bit uniforms in float32 with the fast
// intrinsics: the noise decides near-ties, its law needs no more than that; |n| <= 5.9).  Samplers use rng_normal.
static __device__ __noinline__ double rng_normal_fast(uint64_t seed, uint64_t a, uint64_t b) {
    const uint64_t x = rng_u64(seed, a, b);
    const float u1 = (float(uint32_t(x >> 40)) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (float(uint32_t(x >> 8) & 0xffffffu) + 0.5f) * (1.0f / 16777216.0f);
    return double(sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853071795865f * u2));
}

}  // namespace bear
