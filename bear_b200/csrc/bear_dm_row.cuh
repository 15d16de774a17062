// Per-row Dirichlet-multinomial / multinomial terms shared by the fused kernels (bear_fused.cu, bear_cnn.cu):
// count loading, rising factorials at small integer offsets, the shared-division reciprocal.
// core.tfpDirichletMultinomialPerm.counts_log_prob (core.py:60-74), core.tfpMultinomialPerm (core.py:125-139).
#pragma once
#include "bear_common.cuh"

namespace bear {

constexpr int A1 = 5;            // DNA/RNA letters + stop

// ------------------------------------------------------------------------------------------------
// per-row pieces
// ------------------------------------------------------------------------------------------------
constexpr uint32_t SMALLC = 8;   // counts up to SMALLC take the polynomial rising-factorial path
constexpr int STIR_N = (SMALLC + 1) * (SMALLC + 1) + 1;   // padded to keep 16-byte alignment after it

struct Counts {
    uint32_t c[A1];
    uint32_t cmax;
    double n;
};

__device__ __forceinline__ Counts load_counts(const uint32_t* __restrict__ col, int64_t stride, int64_t i, bool in_range) {
    Counts r;
#pragma unroll
    for (int b = 0; b < A1; ++b) r.c[b] = in_range ? __ldg(col + b * stride + i) : 0u;
    r.cmax = max(max(max(r.c[0], r.c[1]), max(r.c[2], r.c[3])), r.c[4]);
    if (r.cmax < (1u << 29))                   // the common case: one integer sum, one conversion
        r.n = double((r.c[0] + r.c[1]) + (r.c[2] + r.c[3]) + r.c[4]);
    else
        r.n = (double(r.c[0]) + double(r.c[1])) + (double(r.c[2]) + double(r.c[3])) + double(r.c[4]);
    return r;
}

// Warp-uniform trip count of the small-count loops: the largest count (capped) among the live lanes.
__device__ __forceinline__ uint32_t warp_steps(bool live, uint32_t cmax) {
    return __reduce_max_sync(0xffffffffu, live ? (cmax < SMALLC ? cmax : SMALLC) : 0u);
}

// Rising factorials P_b = prod_{i<c_b}(a_b + i) and derivatives D_b = dP_b/da for the five letters of a
// row with counts <= SMALLC, as polynomials in a: P_c(a) = sum_k S(c,k) a^k with the unsigned Stirling
// numbers of the first kind (all terms positive for a > 0, so Horner is well conditioned).  `steps` is
// the warp-wide largest count, so the degree loop is divergence-free; the coefficient row is picked per
// lane and letter from a shared-memory copy of the triangle.
static __constant__ double kStirling[(SMALLC + 1) * (SMALLC + 1)] = {
    1, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 1, 0, 0, 0, 0, 0, 0, 0,
    0, 1, 1, 0, 0, 0, 0, 0, 0,
    0, 2, 3, 1, 0, 0, 0, 0, 0,
    0, 6, 11, 6, 1, 0, 0, 0, 0,
    0, 24, 50, 35, 10, 1, 0, 0, 0,
    0, 120, 274, 225, 85, 15, 1, 0, 0,
    0, 720, 1764, 1624, 735, 175, 21, 1, 0,
    0, 5040, 13068, 13132, 6769, 1960, 322, 28, 1};

// Rising factorial P = prod_{i < c} (a + i) and its derivative D by a Horner-like recurrence that runs from the TOP:
//     for k = steps-1 .. 0:   D = D (a + k) + P;   P = P (a + k) + [k == c]
// starting from P = [c == steps], D = 0.  A lane's product starts when k reaches its own count (the injected 1.0), stays 0
// before that, and every lane finishes together at k = 0 -- so the warp-uniform loop needs NO select of operands or
// results: per step and letter one DADD, two DFMAs, one compare and one 32-bit select for the high word of the injected
// constant.  (The `if (c > k) { D = D t + P; P *= t; }` form -- also when written as predicated PTX -- is compiled to both
// products plus FOUR selects per step and letter, half of the loop's instructions.)  D is the exact derivative of the
// same recurrence (d/da of (a + k) is 1, the injected constant does not depend on a).
__device__ __forceinline__ double rf_inject(bool on) { return __hiloint2double(on ? 0x3ff00000 : 0, 0); }

template <int NL>
__device__ __forceinline__ void rf_horner(const double (&a)[A1], const uint32_t (&c)[A1], uint32_t steps, double (&P)[A1], double (&D)[A1]) {
    uint32_t ce[NL];
#pragma unroll
    for (int b = 0; b < NL; ++b) {
        ce[b] = min(c[b], steps);                 // (counts above the loop bound are redone by the general routine: keep them finite)
        P[b] = rf_inject(ce[b] == steps);
        D[b] = 0.0;
    }
    double kd = double(steps);
    for (uint32_t k = steps; k-- > 0;) {
        kd -= 1.0;
#pragma unroll
        for (int b = 0; b < NL; ++b) {
            const double t = a[b] + kd;
            D[b] = fma(D[b], t, P[b]);
            P[b] = fma(P[b], t, rf_inject(k == ce[b]));
        }
    }
}

template <bool GRAD, typename TS>
__device__ __forceinline__ void rf_letters(const TS* __restrict__ stir, const double (&a)[A1], const uint32_t (&c)[A1],
                                           uint32_t steps, double (&P)[A1], double (&D)[A1]) {
    if (GRAD) {
        // with derivatives (training): the recurrence above -- no shared-memory traffic, no float -> double conversions
        (void)stir;
        // The last letter is the stop symbol: about one transition per sequence, so its count is 0 or 1 in all but a
        // few rows.  When that holds for the whole warp (one vote) it is set directly and the loop runs over the other
        // letters only: a fifth fewer float64 slots.
        if (!__any_sync(0xffffffffu, c[A1 - 1] > 1u)) {
            rf_horner<A1 - 1>(a, c, steps, P, D);
            P[A1 - 1] = c[A1 - 1] != 0u ? a[A1 - 1] : 1.0;
            D[A1 - 1] = c[A1 - 1] != 0u ? 1.0 : 0.0;
        } else {
            rf_horner<A1>(a, c, steps, P, D);
        }
        return;
    }
    const TS* row[A1];     // TS = float (exact: coefficients <= 13132) halves the shared-memory traffic, double saves the conversion
#pragma unroll
    for (int b = 0; b < A1; ++b) {
        row[b] = stir + (c[b] <= SMALLC ? c[b] : 0u) * (SMALLC + 1);
        P[b] = double(row[b][steps]);
        D[b] = 0.0;
    }
    for (int k = int(steps) - 1; k >= 0; --k) {
#pragma unroll
        for (int b = 0; b < A1; ++b) {
            if (GRAD) D[b] = fma(D[b], a[b], P[b]);
            P[b] = fma(P[b], a[b], double(row[b][k]));
        }
    }
}

// single rising factorial for c <= SMALLC (per-lane loop; used for the "total" term off the table)
template <bool GRAD>
__device__ __forceinline__ void rf_one(double a, uint32_t c, double& P, double& D) {
    P = c >= 1 ? a : 1.0;
    D = c >= 1 ? 1.0 : 0.0;
    for (uint32_t t = 1; t < c; ++t) {
        const double x = a + double(t);
        if (GRAD) D = fma(D, x, P);
        P *= x;
    }
}

// r[b] = 1 / d[b] with a single division; returns prod d
__device__ __forceinline__ double inv5(const double (&d)[A1], double (&r)[A1]) {
    const double p01 = d[0] * d[1], p012 = p01 * d[2], p0123 = p012 * d[3], p = p0123 * d[4];
    double t = 1.0 / p;
    r[4] = t * p0123;
    t *= d[4];
    r[3] = t * p012;
    t *= d[3];
    r[2] = t * p01;
    t *= d[2];
    r[1] = t * d[0];
    r[0] = t * d[1];
    return p;
}

// sum_b [lgamma(conc_b + c_b) - lgamma(conc_b)] = add + log(prod) and, with GRAD, w_b = the digamma
// differences.  Small counts: predicated rising factorials and one shared division; a lane with a
// count above SMALLC redoes its row with the general routine (divergent, rare in sparse tables).
template <bool GRAD, typename TS>
__device__ __forceinline__ void letters_term(const TS* __restrict__ stir, const double (&conc)[A1], const Counts& r,
                                             uint32_t steps, double& add, double& prod, double (&w)[A1]) {
    double P[A1], D[A1];
    rf_letters<GRAD, TS>(stir, conc, r.c, steps, P, D);
    add = 0.0;
    if (GRAD) {
        double ri[A1];
        prod = inv5(P, ri);
#pragma unroll
        for (int b = 0; b < A1; ++b) w[b] = D[b] * ri[b];
    } else {
        prod = ((P[0] * P[1]) * (P[2] * P[3])) * P[4];
    }
    if (r.cmax > SMALLC) {
        LogProd acc;
#pragma unroll
        for (int b = 0; b < A1; ++b) {
            const LgDg t = lgdg_diff<GRAD>(conc[b], double(r.c[b]));
            acc.push(t);
            if (GRAD) w[b] = t.dg;
        }
        add = acc.add;
        prod = acc.mul;
    }
}

// lgamma(s + n) - lgamma(s) = add + log(prod), digamma difference dg
template <bool GRAD>
__device__ __forceinline__ void total_term(double s, const Counts& r, double& add, double& prod, double& dg) {
    if (r.n <= double(SMALLC)) {
        double D;
        rf_one<GRAD>(s, uint32_t(r.n), prod, D);
        add = 0.0;
        if (GRAD) dg = D / prod;
    } else {
        const LgDg t = lgdg_diff<GRAD>(s, r.n);
        add = t.add;
        prod = t.mul;
        if (GRAD) dg = t.dg;
    }
}

// sum_b c_b log p_b = add + log(prod)  (multiply_no_nan semantics, core.py:138-139)
__device__ __forceinline__ void mn_term(const double (&p)[A1], const Counts& r, double& add, double& prod) {
    double pw[A1];
#pragma unroll
    for (int b = 0; b < A1; ++b) {           // p^c for c <= 8 by squaring
        const double p2 = p[b] * p[b], p4 = p2 * p2;
        const uint32_t c = r.c[b];
        double x = (c & 1u) ? p[b] : 1.0;
        x *= (c & 2u) ? p2 : 1.0;
        x *= (c & 4u) ? p4 : 1.0;
        pw[b] = (c & 8u) ? p4 * p4 : x;
    }
    add = 0.0;
    prod = ((pw[0] * pw[1]) * (pw[2] * pw[3])) * pw[4];
    if (r.cmax > SMALLC) {
        prod = 1.0;
#pragma unroll
        for (int b = 0; b < A1; ++b)
            if (r.c[b] != 0) add = fma(double(r.c[b]), log_cold(p[b]), add);
    }
}

}  // namespace bear
