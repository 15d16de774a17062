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
#include "bear_eval.cuh"
#include "bear_linear_head.cuh"
#include "bear_sm100.cuh"

namespace {

using namespace bear;
using namespace bear::sm100;

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
// bear_ref with the stop net: head, loss and the three scalar gradients in one pass
// ------------------------------------------------------------------------------------------------
// f = (nw stop + JC(ref, tau)) / (nw + 1) (bear_ref.py:9-69) needs no k-mer: the kernel streams the data column and the
// reference column (40 B per row) and returns [loss, d h_signed, d tau_signed, d net_weight_signed].
template <bool TRAIN_AR>
__global__ void __launch_bounds__(THREADS)
ref_stop_train_kernel(const uint32_t* __restrict__ col, const uint32_t* __restrict__ ref, int64_t stride, int64_t n,
                      const double* __restrict__ h_signed, const double* __restrict__ tau_signed,
                      const double* __restrict__ nw_signed, double* __restrict__ ll_out, double* __restrict__ partials) {
    __shared__ double red[32];
    __shared__ double stir[STIR_N];
    for (int i = threadIdx.x; i < STIR_N; i += blockDim.x) stir[i] = kStirling[i];
    __syncthreads();
    const double hinv = exp(-h_signed[0]);
    const double tau = exp(tau_signed[0]), etau = exp(-tau), nw = exp(nw_signed[0]), inv = 1.0 / (nw + 1.0);
    double ll_sum = 0.0, dh_sum = 0.0, dtau_sum = 0.0, dnw_sum = 0.0;
    for (int64_t base = int64_t(blockIdx.x) * blockDim.x; base < n; base += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const bool in_range = i < n;
        const Counts r = load_counts(col, stride, i, in_range);
        const bool live = r.cmax != 0;
        const uint32_t steps = warp_steps(live, r.cmax);
        // Jukes-Cantor mix of the reference counts (bear_ref.py:9-33 after the map of bear_ref.py:332-337)
        double p[4], s = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            p[b] = double(in_range ? __ldg(ref + b * stride + i) : 0u) + BEAR_EPS;
            s += p[b];
        }
        const double si = 1.0 / s;
        double f[A1], df[A1], ll = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            p[b] = p[b] * si - 0.25;
            f[b] = (0.25 + etau * p[b]) * inv;
        }
        f[4] = nw * inv;
        double add, prod, w[A1];
        if (TRAIN_AR) {
            double q[A1];
#pragma unroll
            for (int b = 0; b < A1; ++b) q[b] = f[b] + BEAR_EPS;
            mn_term(q, r, add, prod);
#pragma unroll
            for (int b = 0; b < A1; ++b) df[b] = r.c[b] == 0 ? 0.0 : double(r.c[b]) / q[b];
        } else {
            double conc[A1], tadd, tprod, tdg;
#pragma unroll
            for (int b = 0; b < A1; ++b) conc[b] = fma(f[b], hinv, BEAR_EPS);
            letters_term<true>(stir, conc, r, steps, add, prod, w);
            const double sc = ((conc[0] + conc[1]) + (conc[2] + conc[3])) + conc[4];
            total_term<true>(sc, r, tadd, tprod, tdg);
            add -= tadd;
            prod /= tprod;
#pragma unroll
            for (int b = 0; b < A1; ++b) df[b] = (w[b] - tdg) * hinv;
        }
        if (live) {
            ll = add + log(prod);
#pragma unroll
            for (int b = 0; b < A1; ++b) {
                if (!TRAIN_AR) dh_sum -= f[b] * df[b];
                if (b < 4) dtau_sum -= df[b] * tau * etau * p[b] * inv;               // d f_b / d tau_signed
                dnw_sum += df[b] * nw * inv * ((b == 4 ? 1.0 : 0.0) - f[b]);           // d f_b / d net_weight_signed
            }
        }
        ll_sum += ll;
        if (in_range && ll_out) ll_out[i] = ll;
    }
    const double a0 = block_sum(ll_sum, red), a1 = block_sum(dh_sum, red), a2 = block_sum(dtau_sum, red), a3 = block_sum(dnw_sum, red);
    if (threadIdx.x == 0) {
        double* out = partials + int64_t(blockIdx.x) * 4;
        out[0] = a0;
        out[1] = a1;
        out[2] = a2;
        out[3] = a3;
    }
}

// ------------------------------------------------------------------------------------------------
// BMM marginal likelihood, every group and alpha
// ------------------------------------------------------------------------------------------------
constexpr int BMM_WARPS = 16, BMM_STAGES = 3;

template <int NA1, int NV>
__global__ void __launch_bounds__(32 * BMM_WARPS, 1)
bmm_kernel(const uint32_t* __restrict__ counts, int64_t stride, int64_t row_lo, int64_t row_hi, int G, int use_tma,
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
    // Sparse rows (every count <= 8, total <= 16; the rule in k-mer tables of whole genomes): the likelihood is a linear
    // function of the HISTOGRAM of the counts -- sum_rows sum_b T_k[c_b] = sum_c hist[c] T_k[c] -- whatever the priors, so
    // a row only bumps integer histogram bins: 7- / 8-bit fields of 64-bit registers (counts 0..8, totals 1..16), spilled
    // into per-lane 32-bit bins before a field can overflow.  No table look-ups, no float arithmetic per row; the bins
    // meet the tables once, at the end.  Integer bins also make the result independent of the order of the rows.
    uint64_t hl = 0, ht0 = 0, ht1 = 0;                     // packed fields: letters with count f + 1, totals f + 1 / f + 9
    uint32_t bin_l[8], bin_t[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) bin_l[j] = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) bin_t[j] = 0u;
    int pending = 0;                                       // rows since the last spill (uniform over the warp)
    auto spill_fields = [&]() {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            bin_l[j] += uint32_t(hl >> (7 * (j + 1))) & 0x7fu;     // (field 0 counts the zero letters: never read)
            bin_t[j] += uint32_t(ht0 >> (8 * j)) & 0xffu;
            bin_t[8 + j] += uint32_t(ht1 >> (8 * j)) & 0xffu;
        }
        hl = ht0 = ht1 = 0;
        pending = 0;
    };
    auto row_term = [&](const uint32_t (&c)[NA1]) {
        uint32_t cmax = 0, toti = 0;
#pragma unroll
        for (int b = 0; b < NA1; ++b) {
            cmax = max(cmax, c[b]);
            toti += c[b] < uint32_t(TABN) ? c[b] : uint32_t(TABN);
        }
        if (cmax == 0) return;
        if (NA1 == 5 && cmax <= 8u && toti <= 16u) {
            // nine 7-bit fields, field c = letters with count c (c = 0 included: no test, no predicate per letter)
#pragma unroll
            for (int b = 0; b < NA1; ++b) hl += 1ull << (7u * c[b]);
            if (toti <= 8u) ht0 += 1ull << (8u * (toti - 1u));
            else ht1 += 1ull << (8u * (toti - 9u));
        } else if (toti < uint32_t(TABN)) {
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
    // Tiles of BT rows aligned to absolute multiples of BT (planes of BT * 4 bytes, 16-byte aligned); every warp is an
    // independent pipeline over its tiles with a ring of BMM_STAGES shared-memory stages that its first NA1 lanes fill
    // with 1-D bulk async copies (TMA), completion on an mbarrier.  NA1 = 5: BT = 128, a lane evaluates four rows from
    // one 128-bit shared-memory load per plane.  Rows outside [row_lo, row_hi) are dead; the tail tile uses guarded loads.
    constexpr int RPL = NA1 == 5 ? 4 : 1;                 // rows per lane
    constexpr int BT = 32 * RPL, PLANE = BT * 4, STAGE = NA1 * PLANE;
    extern __shared__ __align__(128) unsigned char bmm_smem[];
    __shared__ __align__(8) uint64_t bars[BMM_WARPS * BMM_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t my_in = smem_u32(bars) + 8 * warp * BMM_STAGES;
    const uint32_t ring = smem_u32(bmm_smem) + warp * BMM_STAGES * STAGE;
    if (lane == 0) {
        for (int s = 0; s < BMM_STAGES; ++s) mbar_init(my_in + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int64_t a0 = row_lo - row_lo % BT;
    const uint32_t ntiles = uint32_t((row_hi - a0 + BT - 1) / BT);
    const uint32_t t_full = use_tma ? uint32_t((row_hi - a0) / BT) : 0u;
    const uint32_t tstep = gridDim.x * BMM_WARPS;
    uint32_t t = blockIdx.x * BMM_WARPS + warp;
    const uint32_t n_my = t < ntiles ? (ntiles - t + tstep - 1) / tstep : 0u;
    auto issue_tile = [&](uint32_t ti, int stg) {          // lanes 0 .. NA1-1: one plane each
        const uint32_t bar = my_in + 8 * stg;
        if (lane == 0) mbar_arrive_expect_tx(bar, STAGE);
        bulk_g2s(ring + stg * STAGE + lane * PLANE, col + lane * stride + a0 + int64_t(ti) * BT, PLANE, bar);
    };
    if (lane < NA1) {
        for (int s = 0; s < BMM_STAGES; ++s)
            if (uint32_t(s) < n_my && t + s * tstep < t_full) issue_tile(t + s * tstep, s);
    }
    int stg = 0;
    uint32_t in_par = 0;
    for (uint32_t it = 0; it < n_my; ++it, t += tstep) {
        const int64_t r0 = a0 + int64_t(t) * BT + lane * RPL;          // first row of this lane
        uint32_t cc[RPL][NA1];
        if (t < t_full) {
            mbar_wait(my_in + 8 * stg, in_par);
            const uint32_t st = ring + stg * STAGE + lane * (4 * RPL);
#pragma unroll
            for (int b = 0; b < NA1; ++b) {
                if (RPL == 4) {
                    uint32_t x, y, z, w;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(st + b * PLANE));
                    cc[0][b] = x;
                    cc[1 % RPL][b] = y;
                    cc[2 % RPL][b] = z;
                    cc[3 % RPL][b] = w;
                } else {
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cc[0][b]) : "r"(st + b * PLANE));
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < RPL; ++j)
#pragma unroll
                for (int b = 0; b < NA1; ++b) cc[j][b] = r0 + j < row_hi ? __ldg(col + b * stride + r0 + j) : 0u;
        }
        __syncwarp();
        if (lane < NA1 && it + BMM_STAGES < n_my && t + BMM_STAGES * tstep < t_full) issue_tile(t + BMM_STAGES * tstep, stg);
        if (++stg == BMM_STAGES) {
            stg = 0;
            in_par ^= 1u;
        }
        const bool edge = t == 0 || t >= t_full;
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            if (edge && (r0 + j < row_lo || r0 + j >= row_hi)) continue;
            row_term(cc[j]);
        }
        pending += RPL;
        if (pending > 127 / NA1 - RPL) spill_fields();     // a (7-bit) field grows by at most NA1 per row
    }
    spill_fields();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (k < V) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += double(bin_l[j]) * tab[k][j + 1];
#pragma unroll
            for (int j = 0; j < 16; ++j) s -= double(bin_t[j]) * tab_tot[k][j + 1];
            acc[k] += s;
        }
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

extern "C" int bear_ref_train_step(const uint32_t* d_col, const uint32_t* d_ref_col, int64_t stride, int64_t row0, int64_t n,
                                   const double* d_h_signed, const double* d_tau_signed, const double* d_nw_signed, double scale,
                                   int train_ar, double* d_flat, double* d_ll_out, double* d_workspace, void* stream) {
    const char* fn = "bear_ref_train_step";
    BEAR_REQUIRE(d_col && d_ref_col && d_h_signed && d_tau_signed && d_nw_signed && d_flat && d_workspace, fn);
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    if (n == 0) return BEAR_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n);
    if (train_ar)
        ref_stop_train_kernel<true><<<grid, THREADS, 0, st>>>(d_col + row0, d_ref_col + row0, stride, n, d_h_signed, d_tau_signed,
                                                              d_nw_signed, d_ll_out, d_workspace);
    else
        ref_stop_train_kernel<false><<<grid, THREADS, 0, st>>>(d_col + row0, d_ref_col + row0, stride, n, d_h_signed, d_tau_signed,
                                                               d_nw_signed, d_ll_out, d_workspace);
    BEAR_LAUNCH_CHECK("ref_stop_train_kernel");
    reduce_partials_kernel<<<1, 32, 0, st>>>(d_workspace, grid, 4, -scale, d_flat);
    BEAR_LAUNCH_CHECK("reduce_partials_kernel");
    return BEAR_OK;
}

namespace {

int run_eval(const char* fn, bear_eval::EvalArgs& a, double* d_acc) {
    BEAR_REQUIRE(a.test_col && d_acc && a.ws, fn);
    BEAR_REQUIRE(a.n >= 0 && a.row0 >= 0 && a.stride >= a.row0 + a.n, fn);
    BEAR_REQUIRE(a.H >= 0 && a.H <= BEAR_MAX_MODELS && a.V >= 0 && a.V <= BEAR_MAX_MODELS, fn);
    BEAR_REQUIRE((a.H == 0 || a.d_h) && (a.V == 0 || a.d_van), fn);
    if (a.head == BEAR_HEAD_LINEAR || a.head == BEAR_HEAD_REF_LINEAR)
        BEAR_REQUIRE(a.kmers && a.head_ptr && a.lag >= 1 && a.lag <= 29, fn);
    if (a.head == BEAR_HEAD_EXPLICIT) BEAR_REQUIRE(a.head_ptr != nullptr, fn);
    if (a.n == 0) return BEAR_OK;
    int grid;
    switch (a.head) {
        case BEAR_HEAD_LINEAR: grid = bear_eval::launch_linear(a); break;
        case BEAR_HEAD_REF_STOP:
        case BEAR_HEAD_REF_LINEAR: grid = bear_eval::launch_ref(a); break;
        default: grid = bear_eval::launch_misc(a); break;
    }
    if (grid < 0) return grid;
    const int P = 2 * a.H + 2 * a.V + 3;
    reduce_partials_kernel<<<1, 64, 0, a.stream>>>(a.ws, grid, P, 1.0, d_acc);
    BEAR_LAUNCH_CHECK("reduce_partials_kernel");
    return BEAR_OK;
}

}  // namespace

extern "C" int bear_eval_step(const uint64_t* d_kmers, const uint32_t* d_test_col, const uint32_t* d_train_col,
                              int64_t stride, int64_t row0, int64_t n, int lag, int head, const double* d_head,
                              const double* d_h, int H, const double* d_van, int V, int64_t seed, int64_t row_id0,
                              double* d_acc, double* d_workspace, void* stream) {
    const char* fn = "bear_eval_step";
    BEAR_REQUIRE(head >= BEAR_HEAD_NONE && head <= BEAR_HEAD_STOP, fn);
    bear_eval::EvalArgs a{d_kmers, d_test_col, d_train_col, nullptr, stride, row0, n, row_id0, lag, head, d_head, nullptr, nullptr,
                          d_h, d_van, H, V, seed, d_workspace, static_cast<cudaStream_t>(stream)};
    return run_eval(fn, a, d_acc);
}

extern "C" int bear_ref_eval_step(const uint64_t* d_kmers, const uint32_t* d_test_col, const uint32_t* d_train_col,
                                  const uint32_t* d_ref_col, int64_t stride, int64_t row0, int64_t n, int lag, int net,
                                  const double* d_mat, const double* d_tau_signed, const double* d_nw_signed, const double* d_h,
                                  int H, const double* d_van, int V, int64_t seed, int64_t row_id0, double* d_acc,
                                  double* d_workspace, void* stream) {
    const char* fn = "bear_ref_eval_step";
    BEAR_REQUIRE(net == BEAR_HEAD_STOP || net == BEAR_HEAD_LINEAR, fn);
    BEAR_REQUIRE(d_ref_col && d_tau_signed && d_nw_signed, fn);
    bear_eval::EvalArgs a{d_kmers, d_test_col, d_train_col, d_ref_col, stride, row0, n, row_id0, lag,
                          net == BEAR_HEAD_STOP ? BEAR_HEAD_REF_STOP : BEAR_HEAD_REF_LINEAR, d_mat, d_tau_signed, d_nw_signed,
                          d_h, d_van, H, V, seed, d_workspace, static_cast<cudaStream_t>(stream)};
    return run_eval(fn, a, d_acc);
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
    const int bt = A1v == 5 ? 128 : 32;                       // rows per tile (bmm_kernel)
    const int64_t a0 = row0 - row0 % bt, ntiles = (row0 + n - a0 + bt - 1) / bt;
    const int64_t want = (ntiles + BMM_WARPS - 1) / BMM_WARPS;
    const int gx = int(want < 148 ? want : 148);
    const dim3 grid(gx, G);
    const int use_tma = (reinterpret_cast<uintptr_t>(d_counts) & 15) == 0 && (stride & 3) == 0;
    const size_t smem = size_t(BMM_WARPS) * BMM_STAGES * A1v * bt * 4;
    if (A1v == 5 && V <= 4) {
        BEAR_CUDA_CHECK(cudaFuncSetAttribute(bmm_kernel<5, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        bmm_kernel<5, 4><<<grid, 32 * BMM_WARPS, smem, st>>>(d_counts, stride, row0, row0 + n, G, use_tma, d_alpha, V, d_workspace);
    } else if (A1v == 5) {
        BEAR_CUDA_CHECK(cudaFuncSetAttribute(bmm_kernel<5, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        bmm_kernel<5, 8><<<grid, 32 * BMM_WARPS, smem, st>>>(d_counts, stride, row0, row0 + n, G, use_tma, d_alpha, V, d_workspace);
    } else {
        BEAR_CUDA_CHECK(cudaFuncSetAttribute(bmm_kernel<21, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        bmm_kernel<21, 8><<<grid, 32 * BMM_WARPS, smem, st>>>(d_counts, stride, row0, row0 + n, G, use_tma, d_alpha, V, d_workspace);
    }
    BEAR_LAUNCH_CHECK("bmm_kernel");
    reduce_partials_kernel<<<1, 64, 0, st>>>(d_workspace, int(grid.x), G * V, 1.0, d_out);
    BEAR_LAUNCH_CHECK("reduce_partials_kernel");
    return BEAR_OK;
}
