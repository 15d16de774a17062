// Fused packed-path kernels (DNA/RNA, A1 = 5): the count-streaming hot path next to the training step
// (bear_train.cu: linear_train_tc_kernel).
//   explicit_train_kernel the training loss for a caller-evaluated head f (bear_ref / plugins)
//   eval_kernel           bear_net._evaluation_step (bear_net.py:323-371), h_scan (bear_net.py:516-531)
//   bmm_kernel            dataloader._marginal_step (dataloader.py:111-113)
// Every kernel streams the packed table once (8 B k-mer + 20 B per count column per row) and keeps all
// per-row temporaries in registers.  What makes the sparse-count regime cheap:
//   * the linear head is a gather from per-chunk tables of exp-ratios (4 positions per lookup); the tables
//     also hold the start-padded patterns, so every k-mer takes the same path;
//   * lgamma / digamma differences at small integer offsets are rising factorials evaluated with
//     predicated straight-line code; terms that do not depend on the row's k-mer (the "total" term of
//     the Dirichlet-multinomial, BMM priors) come from per-CTA tables indexed by the count;
//   * log() is taken of running products spanning many rows, not once per row; the five reciprocals
//     of a row share one division.
// Reductions are two-stage and deterministic across CTAs: per-CTA partials in the caller's workspace,
// then a fixed-order sum.
#include <math.h>
#include <stdlib.h>

#include "bear_b200.h"
#include "bear_host.h"
#include "bear_linear_head.cuh"

namespace {

using namespace bear;

#ifndef BEAR_EVAL_CTAS
#define BEAR_EVAL_CTAS 2
#endif
#define EVAL_MIN_CTAS(NH, NV) (((NH) <= 1 && (NV) <= 4) ? BEAR_EVAL_CTAS : 2)

// count of letter idx; the evaluation sums these as integers (exact, and no int -> double conversion per row)
__device__ __forceinline__ uint32_t pick5(const uint32_t (&c)[A1], int idx) {
    return idx == 0 ? c[0] : idx == 1 ? c[1] : idx == 2 ? c[2] : idx == 3 ? c[3] : c[4];
}

// argmax of v + sigma * N(0,1) (core.py:69-71,134-136).  Candidates are the entries within 16 sigma of
// the maximum (anything further cannot win, P < 1e-28).  One candidate: no randomness needed.  All
// candidates exactly tied: a uniform pick, which is what iid noise gives.  Otherwise Gaussian noise on
// the candidates only.  seed < 0: no noise, first maximum wins.
struct V5 {
    double v[A1];
};

__device__ __forceinline__ uint32_t tie_hash(int64_t seed, uint64_t row, uint64_t model) {
    return uint32_t(mix64(uint64_t(seed) ^ (row * 0x9E3779B97F4A7C15ull) ^ (model * 0xD1B54A32D192ED03ull)) >> 32);
}

// the randomised part, out of line: ties are rare except for the unconditioned BMM (handled separately)
__device__ __noinline__ int argmax_tiebreak(V5 x, double top, double thr, bool all_exact, int exact, double sigma,
                                            int64_t seed, uint64_t row, uint64_t model) {
    int best = 0;
    if (all_exact) {
        int k = int(tie_hash(seed, row, model) % uint32_t(exact));
        for (int b = 0; b < A1; ++b)
            if (x.v[b] == top) {
                if (k == 0) best = b;
                --k;
            }
        return best;
    }
    double nb = -INFINITY;
    for (int b = 0; b < A1; ++b)
        if (x.v[b] > thr) {
            const double y = x.v[b] + sigma * rng_normal(uint64_t(seed), row, model * 64 + uint64_t(b));
            if (y > nb) {
                nb = y;
                best = b;
            }
        }
    return best;
}

__device__ __forceinline__ int noisy_argmax5(const double (&v)[A1], double sigma, int64_t seed, uint64_t row,
                                             uint64_t model) {
    int best = 0;
    double top = v[0];
#pragma unroll
    for (int b = 1; b < A1; ++b)
        if (v[b] > top) {
            top = v[b];
            best = b;
        }
    if (seed < 0) return best;
    const double thr = top - 16.0 * sigma;
    int near = 0, exact = 0;
#pragma unroll
    for (int b = 0; b < A1; ++b) {
        near += v[b] > thr;
        exact += v[b] == top;
    }
    if (near == 1) return best;
    V5 x;
#pragma unroll
    for (int b = 0; b < A1; ++b) x.v[b] = v[b];
    return argmax_tiebreak(x, top, thr, near == exact, exact, sigma, seed, row, model);
}

// ------------------------------------------------------------------------------------------------
// explicit head: f in, d loss / d f out
// ------------------------------------------------------------------------------------------------
template <bool TRAIN_AR>
__global__ void __launch_bounds__(THREADS)
explicit_train_kernel(const uint32_t* __restrict__ col, int64_t stride, int64_t n,
                      const double* __restrict__ f_in, const double* __restrict__ h_signed, double scale,
                      double* __restrict__ gf, double* __restrict__ ll_out, double* __restrict__ partials) {
    __shared__ double red[32];
    __shared__ double stir[STIR_N];
    for (int i = threadIdx.x; i < STIR_N; i += blockDim.x) stir[i] = kStirling[i];
    __syncthreads();
    const double hinv = exp(-h_signed[0]);
    double ll_sum = 0.0, dh_sum = 0.0;
    // warp-tile loop: every lane of a warp runs the same number of iterations (collectives inside)
    for (int64_t base = int64_t(blockIdx.x) * blockDim.x; base < n; base += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const bool in_range = i < n;
        const Counts r = load_counts(col, stride, i, in_range);
        const bool live = r.cmax != 0;
        const uint32_t steps = warp_steps(live, r.cmax);
        double f[A1], df[A1], ll = 0.0;
#pragma unroll
        for (int b = 0; b < A1; ++b) {
            f[b] = in_range ? f_in[i * A1 + b] : 0.2;
            df[b] = 0.0;
        }
        double add, prod, w[A1];
        if (TRAIN_AR) {
            double p[A1];
#pragma unroll
            for (int b = 0; b < A1; ++b) p[b] = f[b] + BEAR_EPS;
            mn_term(p, r, add, prod);
#pragma unroll
            for (int b = 0; b < A1; ++b) df[b] = r.c[b] == 0 ? 0.0 : double(r.c[b]) / p[b];
        } else {
            double conc[A1], tadd, tprod, tdg;
#pragma unroll
            for (int b = 0; b < A1; ++b) conc[b] = fma(f[b], hinv, BEAR_EPS);
            letters_term<true>(stir, conc, r, steps, add, prod, w);
            const double s = ((conc[0] + conc[1]) + (conc[2] + conc[3])) + conc[4];
            total_term<true>(s, r, tadd, tprod, tdg);
            add -= tadd;
            prod /= tprod;
            if (live) {
#pragma unroll
                for (int b = 0; b < A1; ++b) {
                    df[b] = (w[b] - tdg) * hinv;
                    dh_sum -= f[b] * df[b];
                }
            }
        }
        if (live) ll = add + log(prod);
        ll_sum += ll;
        if (in_range) {
            if (ll_out) ll_out[i] = ll;
            if (gf) {
#pragma unroll
                for (int b = 0; b < A1; ++b) gf[i * A1 + b] = -scale * df[b];
            }
        }
    }
    const double ll_blk = block_sum(ll_sum, red);
    const double dh_blk = block_sum(dh_sum, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x * 2 + 0] = ll_blk;
        partials[blockIdx.x * 2 + 1] = dh_blk;
    }
}

// ------------------------------------------------------------------------------------------------
// dense counts against row-independent concentrations (BMM priors)
// ------------------------------------------------------------------------------------------------
template <int NA1>
struct CountVec {
    uint32_t c[NA1];
};

// lgamma(a + c) - lgamma(a) summed over the letters of one row (minus `total` handled by the caller) for a
// row-independent a: small counts from the count table, large ones by the constant-a Stirling form
template <int NA1>
__device__ __forceinline__ double dense_letters(const CountVec<NA1>& cv, double a, const double* __restrict__ tab_k, double Ka) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < NA1; ++b)
        s += cv.c[b] < uint32_t(TABN) ? tab_k[cv.c[b]] : lg_shift_large(a, double(cv.c[b]), Ka);
    return s;
}

// the vanilla-BMM term of one evaluation row without a conditioning column (prior vk = van_k + eps)
__device__ __noinline__ double van_dense_row(CountVec<A1> cv, double rn, double vk, const double* __restrict__ tv,
                                             double Kv, double Kvt) {
    return dense_letters<A1>(cv, vk, tv, Kv) - lg_shift_large(double(A1) * vk, rn, Kvt);
}

// ------------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------------
// NH / NV bound the number of h values (H) and of BMM priors (V) of one launch.
template <int HEAD, int NH, int NV, bool HAS_TRAIN>
__global__ void __launch_bounds__(THREADS, EVAL_MIN_CTAS(NH, NV))
eval_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ test_col,
            const uint32_t* __restrict__ train_col, int64_t stride, int64_t row0, int64_t n, int lag,
            const ChunkKeys ck, const double* __restrict__ head, const double* __restrict__ d_h, int H,
            const double* __restrict__ d_van, int V, int64_t seed, double* __restrict__ partials) {
    extern __shared__ __align__(16) double smem[];
    const int nch = num_chunks(lag);
    constexpr bool LIN = HEAD == BEAR_HEAD_LINEAR;
    // the sum of a row's BEAR concentrations is row-independent when there is no conditioning column
    // and the head is normalised (or absent)
    constexpr bool TOT_TAB = !HAS_TRAIN && HEAD != BEAR_HEAD_EXPLICIT;
    double* R = smem;                                               // [nch][ENT][4] extended ratio tables
    uint16_t* symtab = reinterpret_cast<uint16_t*>(R + (LIN ? nch * ENT * 4 : 0));
    double* red = R + (LIN ? nch * ENT * 4 + (2 * ENT * 2) / 8 : 0);
    constexpr int NM = NH > NV ? NH : NV;
    double* tab_ear = red + 32;                    // [NM][TABN]  lgamma(S0_k + N) - lgamma(S0_k)
    double* tab_van = tab_ear + NM * TABN;         // [NM][TABN]  lgamma(van_k + eps + c) - lgamma(van_k + eps)
    double* tab_vtot = tab_van + NM * TABN;        // [NM][TABN]  lgamma(5 (van_k + eps) + N) - lgamma(5 (van_k + eps))
    double* kconst = tab_vtot + NM * TABN;         // [3][NM] ln(2 pi)/2 - lgamma(a) of the three table families
    double* stir = kconst + 3 * NM;                // Stirling triangle
    for (int i = threadIdx.x; i < STIR_N; i += blockDim.x) stir[i] = kStirling[i];
    if (LIN) build_ext_tables(head, R, symtab, lag, ck);
    double hinv[NH], van[NV];
#pragma unroll
    for (int k = 0; k < NH; ++k) hinv[k] = k < H ? 1.0 / d_h[k] : 1.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) van[k] = k < V ? d_van[k] : 1.0;
    // constants of the row-independent terms for counts past the tables (lg_shift_large); kept in shared memory
    bool fast_van = true, fast_ear = true;
#pragma unroll
    for (int k = 0; k < NV; ++k) fast_van = fast_van && double(A1) * (van[k] + BEAR_EPS) < BEAR_LARGE_C;
#pragma unroll
    for (int k = 0; k < NH; ++k) fast_ear = fast_ear && (HEAD == BEAR_HEAD_NONE ? 0.0 : hinv[k]) + A1 * BEAR_EPS < BEAR_LARGE_C;
    if (!HAS_TRAIN && threadIdx.x < NM) {
        const int k = threadIdx.x;
        const double hk = k < H ? 1.0 / d_h[k] : 1.0, vk = (k < V ? d_van[k] : 1.0) + BEAR_EPS;
        kconst[k] = lg_shift_const((HEAD == BEAR_HEAD_NONE ? 0.0 : hk) + A1 * BEAR_EPS);
        kconst[NM + k] = lg_shift_const(vk);
        kconst[2 * NM + k] = lg_shift_const(double(A1) * vk);
    }
    if (!HAS_TRAIN) {
        for (int idx = threadIdx.x; idx < NM * TABN; idx += blockDim.x) {
            const int k = idx / TABN;
            const double c = double(idx % TABN);
            const double hk = k < H ? 1.0 / d_h[k] : 1.0, vk = (k < V ? d_van[k] : 1.0) + BEAR_EPS;
            const double s0 = (HEAD == BEAR_HEAD_NONE ? 0.0 : hk) + A1 * BEAR_EPS;
            LgDg t = lgdg_diff<false>(s0, c);
            tab_ear[idx] = t.add + (t.mul == 1.0 ? 0.0 : log(t.mul));
            t = lgdg_diff<false>(vk, c);
            tab_van[idx] = t.add + (t.mul == 1.0 ? 0.0 : log(t.mul));
            t = lgdg_diff<false>(double(A1) * vk, c);
            tab_vtot[idx] = t.add + (t.mul == 1.0 ? 0.0 : log(t.mul));
        }
    }
    __syncthreads();

    double ear_add[NH], van_add[NV];
    unsigned long long cor_ear[NH], cor_van[NV], cor_arm = 0ull;      // test counts at the predicted letters
    LogProdLong ear_prod[NH];
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        ear_add[k] = 0.0;
        cor_ear[k] = 0ull;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        van_add[k] = 0.0;
        cor_van[k] = 0ull;
    }
    double arm_add = 0.0, total = 0.0;
    LogProdLong arm_prod;

    // the k-mer and test counts of the next row are fetched while the current row is evaluated
    const int64_t gstep = int64_t(gridDim.x) * blockDim.x;
    RowIn nxt = load_row(LIN ? kmers : nullptr, test_col, stride, int64_t(blockIdx.x) * blockDim.x + threadIdx.x, n);
    uint32_t tnx[A1] = {0, 0, 0, 0, 0};
    if (HAS_TRAIN) {
        const int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int b = 0; b < A1; ++b) tnx[b] = i0 < n ? __ldg(train_col + b * stride + i0) : 0u;
    }
    for (int64_t base = int64_t(blockIdx.x) * blockDim.x; base < n; base += gstep) {
        const int64_t i = base + threadIdx.x;
        const bool in_range = i < n;
        const RowIn cur = nxt;
        nxt = load_row(LIN ? kmers : nullptr, test_col, stride, i + gstep, n);
        uint32_t tc[A1] = {0, 0, 0, 0, 0};
        if (HAS_TRAIN) {
#pragma unroll
            for (int b = 0; b < A1; ++b) tc[b] = tnx[b];
#pragma unroll
            for (int b = 0; b < A1; ++b) tnx[b] = i + gstep < n ? __ldg(train_col + b * stride + i + gstep) : 0u;
        }
        Counts r;
#pragma unroll
        for (int b = 0; b < A1; ++b) r.c[b] = cur.c[b];
        r.cmax = max(max(max(r.c[0], r.c[1]), max(r.c[2], r.c[3])), r.c[4]);
        if (r.cmax < (1u << 29))
            r.n = double((r.c[0] + r.c[1]) + (r.c[2] + r.c[3]) + r.c[4]);
        else
            r.n = (double(r.c[0]) + double(r.c[1])) + (double(r.c[2]) + double(r.c[3])) + double(r.c[4]);
        const bool live = r.cmax != 0;             // no test transitions: contributes 0 to every output
        const uint32_t steps = warp_steps(live, r.cmax);
        double t[A1] = {0, 0, 0, 0, 0};
        if (HAS_TRAIN && live) {
#pragma unroll
            for (int b = 0; b < A1; ++b) t[b] = double(tc[b]);
        }
        double f[A1];
        if (LIN) {
#pragma unroll
            for (int b = 0; b < A1; ++b) f[b] = 0.2;
            if (live) linear_head_ext(R, head, cur.code, lag, ck, nch, f);
        } else {
#pragma unroll
            for (int b = 0; b < A1; ++b)
                f[b] = HEAD == BEAR_HEAD_EXPLICIT ? (in_range ? head[i * A1 + b] : 0.2)
                                                  : ((HEAD == BEAR_HEAD_STOP && b == A1 - 1) ? 1.0 : 0.0);
        }
        total += r.n;
        const uint64_t grow = uint64_t(row0 + i);
        const bool use_tab = !HAS_TRAIN && r.n < double(TABN);
        double dummy[A1];
        uint64_t van_hash = 0;
        // BEAR: conc = f / h + train + eps   (bear_net.py:43, 335-337)
#pragma unroll
        for (int k = 0; k < NH; ++k) {
            if (k < H) {
                double conc[A1], add, prod;
#pragma unroll
                for (int b = 0; b < A1; ++b) conc[b] = fma(f[b], hinv[k], t[b]) + BEAR_EPS;
                letters_term<false>(stir, conc, r, steps, add, prod, dummy);
                if (TOT_TAB && use_tab) {
                    add -= tab_ear[k * TABN + int(r.n)];
                } else if (TOT_TAB && fast_ear) {
                    add -= lg_shift_large((HEAD == BEAR_HEAD_NONE ? 0.0 : hinv[k]) + A1 * BEAR_EPS, r.n, kconst[k]);
                } else {
                    double tadd, tprod, tdg;
                    const double s = ((conc[0] + conc[1]) + (conc[2] + conc[3])) + conc[4];
                    total_term<false>(s, r, tadd, tprod, tdg);
                    add -= tadd;
                    prod /= tprod;
                }
                if (live) {
                    ear_add[k] += add;
                    ear_prod[k].push(0.0, prod);
                    cor_ear[k] += pick5(r.c, noisy_argmax5(conc, 100.0 * BEAR_EPS, seed, grow, uint64_t(k)));
                }
            }
        }
        // AR: p = f + eps   (bear_net.py:68, 338)
        {
            double p[A1], add, prod;
#pragma unroll
            for (int b = 0; b < A1; ++b) p[b] = f[b] + BEAR_EPS;
            mn_term(p, r, add, prod);
            if (live) {
                arm_add += add;
                arm_prod.push(0.0, prod);
                cor_arm += pick5(r.c, noisy_argmax5(p, BEAR_EPS, seed, grow, 100));
            }
        }
        // vanilla BMM: conc = train + van + eps   (bear_net.py:328-331, 339-340)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (k < V) {
                double conc[A1];
#pragma unroll
                for (int b = 0; b < A1; ++b) conc[b] = (t[b] + van[k]) + BEAR_EPS;
                if (!HAS_TRAIN) {
                    if (use_tab) {
                        const double* tv = tab_van + k * TABN;
                        van_add[k] += (((tv[r.c[0]] + tv[r.c[1]]) + (tv[r.c[2]] + tv[r.c[3]])) + tv[r.c[4]]) -
                                      tab_vtot[k * TABN + int(r.n)];
                    } else if (fast_van) {
                        CountVec<A1> cv;
#pragma unroll
                        for (int b = 0; b < A1; ++b) cv.c[b] = r.c[b];
                        van_add[k] += van_dense_row(cv, r.n, van[k] + BEAR_EPS, tab_van + k * TABN, kconst[NM + k],
                                                    kconst[2 * NM + k]);
                    } else {
                        LogProd num, den;
                        den.push(lgdg_diff<false>(((conc[0] + conc[1]) + (conc[2] + conc[3])) + conc[4], r.n));
#pragma unroll
                        for (int b = 0; b < A1; ++b) num.push(lgdg_diff<false>(conc[b], double(r.c[b])));
                        van_add[k] += logprod_diff(num, den);
                    }
                } else {
                    double add, prod, tadd, tprod, tdg;
                    letters_term<false>(stir, conc, r, steps, add, prod, dummy);
                    const double s = ((conc[0] + conc[1]) + (conc[2] + conc[3])) + conc[4];
                    total_term<false>(s, r, tadd, tprod, tdg);
                    if (live) van_add[k] += (add - tadd) + log(prod / tprod);
                }
                if (live) {
                    // no conditioning column: the five concentrations are equal and the noisy argmax is a uniform
                    // pick; one hash per row serves up to four priors (16-bit fields, multiply-shift to 0..4)
                    int best;
                    if (HAS_TRAIN) {
                        best = noisy_argmax5(conc, 100.0 * BEAR_EPS, seed, grow, 200 + uint64_t(k));
                    } else if (seed < 0) {
                        best = 0;
                    } else {
                        if ((k & 3) == 0) van_hash = mix64(uint64_t(seed) ^ (grow * 0x9E3779B97F4A7C15ull) ^ (uint64_t(200 + k) * 0xD1B54A32D192ED03ull));
                        best = int((uint32_t(van_hash >> (16 * (k & 3))) & 0xffffu) * 5u >> 16);
                    }
                    cor_van[k] += pick5(r.c, best);
                }
            }
        }
    }
    // layout: [ll_ear[H], ll_arm, ll_van[V], cor_ear[H], cor_arm, cor_van[V], total]
    const int P = 2 * H + 2 * V + 3;
    double* out = partials + int64_t(blockIdx.x) * P;
    int o = 0;
#pragma unroll
    for (int k = 0; k < NH; ++k)
        if (k < H) {
            const double v = ear_add[k] + ear_prod[k].value();
            const double s = block_sum(v, red);
            if (threadIdx.x == 0) out[o] = s;
            ++o;
        }
    {
        const double v = arm_add + arm_prod.value();
        const double s = block_sum(v, red);
        if (threadIdx.x == 0) out[o] = s;
        ++o;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k)
        if (k < V) { const double s = block_sum(van_add[k], red); if (threadIdx.x == 0) out[o] = s; ++o; }
#pragma unroll
    for (int k = 0; k < NH; ++k)
        if (k < H) { const double s = block_sum(double(cor_ear[k]), red); if (threadIdx.x == 0) out[o] = s; ++o; }
    { const double s = block_sum(double(cor_arm), red); if (threadIdx.x == 0) out[o] = s; ++o; }
#pragma unroll
    for (int k = 0; k < NV; ++k)
        if (k < V) { const double s = block_sum(double(cor_van[k]), red); if (threadIdx.x == 0) out[o] = s; ++o; }
    { const double s = block_sum(total, red); if (threadIdx.x == 0) out[o] = s; }
}

// ------------------------------------------------------------------------------------------------
// BMM marginal likelihood, every group and alpha
// ------------------------------------------------------------------------------------------------
template <int NA1, int NV>
__global__ void __launch_bounds__(THREADS, NA1 == 5 ? 4 : 1)
bmm_kernel(const uint32_t* __restrict__ counts, int64_t stride, int64_t n, int G,
           const double* __restrict__ d_alpha, int V, double* __restrict__ partials) {
    __shared__ double red[32];
    __shared__ double tab[NV][TABN];      // lgamma(a + c) - lgamma(a)
    __shared__ double tab_tot[NV][TABN];  // lgamma(A1 a + N) - lgamma(A1 a)
    const int g = blockIdx.y;
    const uint32_t* col = counts + int64_t(g) * NA1 * stride;
    double alpha[NV], acc[NV], Ka[NV], Kt[NV];   // K = ln(2 pi)/2 - lgamma(a): constants of lg_shift_large
    bool fast = true;               // every prior small enough for the constant-a Stirling form
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        alpha[k] = k < V ? d_alpha[k] : 1.0;
        acc[k] = 0.0;
        Ka[k] = lg_shift_const(alpha[k]);
        Kt[k] = lg_shift_const(double(NA1) * alpha[k]);
        fast = fast && double(NA1) * alpha[k] < BEAR_LARGE_C;
    }
    for (int idx = threadIdx.x; idx < NV * TABN; idx += blockDim.x) {
        const int k = idx / TABN;
        const double a = k < V ? d_alpha[k] : 1.0, c = double(idx % TABN);
        LgDg t = lgdg_diff<false>(a, c);
        tab[k][idx % TABN] = t.add + (t.mul == 1.0 ? 0.0 : log(t.mul));
        t = lgdg_diff<false>(double(NA1) * a, c);
        tab_tot[k][idx % TABN] = t.add + (t.mul == 1.0 ? 0.0 : log(t.mul));
    }
    __syncthreads();
    auto row_term = [&](const uint32_t (&c)[NA1]) {
        uint32_t cmax = 0, toti = 0;
#pragma unroll
        for (int b = 0; b < NA1; ++b) {
            cmax = max(cmax, c[b]);
            toti += c[b] < uint32_t(TABN) ? c[b] : uint32_t(TABN);
        }
        if (cmax == 0) return;
        if (toti < uint32_t(TABN)) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                if (k < V) {
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < NA1; ++b) s += tab[k][c[b]];
                    acc[k] += s - tab_tot[k][toti];
                }
            }
        } else if (fast) {
            // dense counts: small letters from the table, large ones (and the total) by the constant-a Stirling form
            double tot = 0.0;
#pragma unroll
            for (int b = 0; b < NA1; ++b) tot += double(c[b]);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                if (k < V) {
                    double s = -lg_shift_large(double(NA1) * alpha[k], tot, Kt[k]);
#pragma unroll
                    for (int b = 0; b < NA1; ++b)
                        s += c[b] < uint32_t(TABN) ? tab[k][c[b]] : lg_shift_large(alpha[k], double(c[b]), Ka[k]);
                    acc[k] += s;
                }
            }
        } else {
            double tot = 0.0;
#pragma unroll
            for (int b = 0; b < NA1; ++b) tot += double(c[b]);
            for (int k = 0; k < V; ++k) {
                LogProd num, den;
                den.push(lgdg_diff<false>(double(NA1) * alpha[k], tot));
#pragma unroll
                for (int b = 0; b < NA1; ++b) num.push(lgdg_diff<false>(alpha[k], double(c[b])));
                acc[k] += logprod_diff(num, den);
            }
        }
    };
    // four rows per thread with 128-bit loads when the column is 16-byte aligned (row0 % 4 == 0)
    const bool vec = (reinterpret_cast<uintptr_t>(col) & 15) == 0 && (stride & 3) == 0;
    const int64_t nq = vec ? n / 4 : 0;
    for (int64_t qd = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; qd < nq; qd += int64_t(gridDim.x) * blockDim.x) {
        uint4 v[NA1];
#pragma unroll
        for (int b = 0; b < NA1; ++b) v[b] = __ldg(reinterpret_cast<const uint4*>(col + b * stride) + qd);
        uint32_t c[NA1];
#pragma unroll
        for (int b = 0; b < NA1; ++b) c[b] = v[b].x;
        row_term(c);
#pragma unroll
        for (int b = 0; b < NA1; ++b) c[b] = v[b].y;
        row_term(c);
#pragma unroll
        for (int b = 0; b < NA1; ++b) c[b] = v[b].z;
        row_term(c);
#pragma unroll
        for (int b = 0; b < NA1; ++b) c[b] = v[b].w;
        row_term(c);
    }
    for (int64_t i = nq * 4 + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        uint32_t c[NA1];
#pragma unroll
        for (int b = 0; b < NA1; ++b) c[b] = __ldg(col + b * stride + i);
        row_term(c);
    }
    for (int k = 0; k < V; ++k) {
        const double s = block_sum(acc[k], red);
        if (threadIdx.x == 0) partials[(int64_t(blockIdx.x) * G + g) * V + k] = s;
    }
}


int grid_for(int64_t n, int cap = MAX_GRID) {
    int64_t blocks = (n + THREADS - 1) / THREADS;
    if (blocks < 1) blocks = 1;
    return int(blocks < cap ? blocks : cap);
}


size_t eval_smem_bytes(int head, int lag, int nm) {
    size_t d = 32 + size_t(3) * nm * TABN + 3 * nm + STIR_N;
    if (head == BEAR_HEAD_LINEAR) d += size_t(num_chunks(lag)) * ENT * 4 + (2 * ENT * 2) / 8;
    return sizeof(double) * d;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        BEAR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    }
    return 0;
}

template <int HEAD, int NH, int NV, bool HAS_TRAIN>
int launch_eval(int grid, size_t smem, cudaStream_t st, const uint64_t* km, const uint32_t* te, const uint32_t* tr,
                int64_t stride, int64_t row0, int64_t n, int lag, const double* head, const double* d_h, int H,
                const double* d_van, int V, int64_t seed, double* ws) {
    if (set_smem(eval_kernel<HEAD, NH, NV, HAS_TRAIN>, smem)) return BEAR_ERR_CUDA;
    eval_kernel<HEAD, NH, NV, HAS_TRAIN><<<grid, THREADS, smem, st>>>(km, te, tr, stride, row0, n, lag,
                                                                      make_chunk_keys(lag > 0 && lag <= 29 ? lag : 1), head,
                                                                      d_h, H, d_van, V, seed, ws);
    return 0;
}

template <int HEAD, int NH, int NV>
int launch_eval_t(bool has_train, int grid, size_t smem, cudaStream_t st, const uint64_t* km, const uint32_t* te,
                  const uint32_t* tr, int64_t stride, int64_t row0, int64_t n, int lag, const double* head,
                  const double* d_h, int H, const double* d_van, int V, int64_t seed, double* ws) {
    return has_train ? launch_eval<HEAD, NH, NV, true>(grid, smem, st, km, te, tr, stride, row0, n, lag, head, d_h, H, d_van, V, seed, ws)
                     : launch_eval<HEAD, NH, NV, false>(grid, smem, st, km, te, tr, stride, row0, n, lag, head, d_h, H, d_van, V, seed, ws);
}

// the common call is one h value with up to four priors (evaluation); h_scan uses up to eight h values
template <int HEAD>
int launch_eval_nm(bool small, bool has_train, int grid, size_t smem, cudaStream_t st, const uint64_t* km, const uint32_t* te,
                   const uint32_t* tr, int64_t stride, int64_t row0, int64_t n, int lag, const double* head,
                   const double* d_h, int H, const double* d_van, int V, int64_t seed, double* ws) {
    if (small) return launch_eval_t<HEAD, 1, 4>(has_train, grid, smem, st, km, te, tr, stride, row0, n, lag, head, d_h, H, d_van, V, seed, ws);
    return launch_eval_t<HEAD, 8, 8>(has_train, grid, smem, st, km, te, tr, stride, row0, n, lag, head, d_h, H, d_van, V, seed, ws);
}

}  // namespace

extern "C" int bear_dm_train_step_explicit(const uint32_t* d_col, int64_t stride, int64_t row0, int64_t n,
                                           const double* d_f, const double* d_h_signed, double scale,
                                           int train_ar, double* d_flat, double* d_gf, double* d_ll_out,
                                           double* d_workspace, void* stream) {
    const char* fn = "bear_dm_train_step_explicit";
    BEAR_REQUIRE(d_col && d_f && d_h_signed && d_flat && d_workspace, fn);
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    if (n == 0) return BEAR_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n);
    if (train_ar)
        explicit_train_kernel<true><<<grid, THREADS, 0, st>>>(d_col + row0, stride, n, d_f, d_h_signed, scale, d_gf, d_ll_out, d_workspace);
    else
        explicit_train_kernel<false><<<grid, THREADS, 0, st>>>(d_col + row0, stride, n, d_f, d_h_signed, scale, d_gf, d_ll_out, d_workspace);
    BEAR_LAUNCH_CHECK("explicit_train_kernel");
    reduce_partials_kernel<<<1, 32, 0, st>>>(d_workspace, grid, 2, -scale, d_flat);
    BEAR_LAUNCH_CHECK("reduce_partials_kernel");
    return BEAR_OK;
}

extern "C" int bear_eval_step(const uint64_t* d_kmers, const uint32_t* d_test_col, const uint32_t* d_train_col,
                              int64_t stride, int64_t row0, int64_t n, int lag, int head, const double* d_head,
                              const double* d_h, int H, const double* d_van, int V, int64_t seed, int64_t row_id0,
                              double* d_acc, double* d_workspace, void* stream) {
    const char* fn = "bear_eval_step";
    BEAR_REQUIRE(d_test_col && d_acc && d_workspace, fn);
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    BEAR_REQUIRE(H >= 0 && H <= BEAR_MAX_MODELS && V >= 0 && V <= BEAR_MAX_MODELS, fn);
    BEAR_REQUIRE((H == 0 || d_h) && (V == 0 || d_van), fn);
    BEAR_REQUIRE(head >= BEAR_HEAD_NONE && head <= BEAR_HEAD_STOP, fn);
    if (head == BEAR_HEAD_LINEAR) BEAR_REQUIRE(d_kmers && d_head && lag >= 1 && lag <= 29, fn);
    if (head == BEAR_HEAD_EXPLICIT) BEAR_REQUIRE(d_head != nullptr, fn);
    if (n == 0) return BEAR_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool small = H <= 1 && V <= 4;
    const int grid = grid_for(n, 148 * (small ? BEAR_EVAL_CTAS : 2));
    const size_t smem = eval_smem_bytes(head, lag, small ? 4 : 8);
    const uint64_t* km = d_kmers ? d_kmers + row0 : nullptr;
    const uint32_t* tr = d_train_col ? d_train_col + row0 : nullptr;
    const uint32_t* te = d_test_col + row0;
    const bool ht = tr != nullptr;
    int rc;
    switch (head) {
        case BEAR_HEAD_LINEAR: rc = launch_eval_nm<BEAR_HEAD_LINEAR>(small, ht, grid, smem, st, km, te, tr, stride, row_id0, n, lag, d_head, d_h, H, d_van, V, seed, d_workspace); break;
        case BEAR_HEAD_EXPLICIT: rc = launch_eval_nm<BEAR_HEAD_EXPLICIT>(small, ht, grid, smem, st, km, te, tr, stride, row_id0, n, lag, d_head, d_h, H, d_van, V, seed, d_workspace); break;
        case BEAR_HEAD_STOP: rc = launch_eval_nm<BEAR_HEAD_STOP>(small, ht, grid, smem, st, km, te, tr, stride, row_id0, n, lag, d_head, d_h, H, d_van, V, seed, d_workspace); break;
        default: rc = launch_eval_nm<BEAR_HEAD_NONE>(small, ht, grid, smem, st, km, te, tr, stride, row_id0, n, lag, d_head, d_h, H, d_van, V, seed, d_workspace); break;
    }
    if (rc) return rc;
    BEAR_LAUNCH_CHECK("eval_kernel");
    const int P = 2 * H + 2 * V + 3;
    reduce_partials_kernel<<<1, 64, 0, st>>>(d_workspace, grid, P, 1.0, d_acc);
    BEAR_LAUNCH_CHECK("reduce_partials_kernel");
    return BEAR_OK;
}

extern "C" int bear_bmm_likelihood(const uint32_t* d_counts, int64_t stride, int64_t row0, int64_t n, int G,
                                   int A1v, const double* d_alpha, int V, double* d_out, double* d_workspace,
                                   void* stream) {
    const char* fn = "bear_bmm_likelihood";
    BEAR_REQUIRE(d_counts && d_alpha && d_out && d_workspace, fn);
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    BEAR_REQUIRE(G >= 1 && V >= 1 && V <= BEAR_MAX_MODELS && G * V <= 64, fn);
    BEAR_REQUIRE(A1v == 5 || A1v == 21, fn);
    if (n == 0) return BEAR_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid(grid_for(n), G);
    if (A1v == 5 && V <= 4)
        bmm_kernel<5, 4><<<grid, THREADS, 0, st>>>(d_counts + row0, stride, n, G, d_alpha, V, d_workspace);
    else if (A1v == 5)
        bmm_kernel<5, 8><<<grid, THREADS, 0, st>>>(d_counts + row0, stride, n, G, d_alpha, V, d_workspace);
    else
        bmm_kernel<21, 8><<<grid, THREADS, 0, st>>>(d_counts + row0, stride, n, G, d_alpha, V, d_workspace);
    BEAR_LAUNCH_CHECK("bmm_kernel");
    reduce_partials_kernel<<<1, 64, 0, st>>>(d_workspace, int(grid.x), G * V, 1.0, d_out);
    BEAR_LAUNCH_CHECK("reduce_partials_kernel");
    return BEAR_OK;
}
