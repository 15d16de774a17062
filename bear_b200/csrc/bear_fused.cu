// Fused packed-path kernels (DNA/RNA, A1 = 5): the count-streaming hot path.
//   linear_train2_kernel  bear_net._train_step (bear_net.py:146-197) + ar_funcs.make_ar_func_linear
//                         (ar_funcs.py:23-46) + core.*.counts_log_prob (core.py:73-74,138-139), fwd+bwd
//   explicit_train_kernel the same loss for a caller-evaluated head f (bear_ref / plugins)
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
//     of a row share one division;
//   * the weight-table gradient is scattered without atomics: producer warps stage rows in shared memory,
//     each gradient chunk table is owned by one consumer warp, rows with equal keys are ranked by the
//     producer (match.any) and applied in separate read-modify-write rounds.
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
// linear head, fused forward + backward, v2: warp-specialised, barrier-light
// ------------------------------------------------------------------------------------------------
// One CTA of 16 warps per SM.  Warps 0..nch-1 are CONSUMERS: warp ch owns the gradient chunk table G[ch] and
// does nothing but scatter staged rows into it.  The other warps are PRODUCERS: each computes TPW tiles of
// 32 rows per iteration (decode -> table-gather head -> Dirichlet-multinomial forward/backward) and stages,
// per row, the four logit gradients and the row's key into every chunk table.  The stage is double
// buffered, so one __syncthreads per iteration is enough: consumers scatter iteration i-1 while producers
// compute iteration i.  The loads of a producer's next tile are issued before the current tile is computed.
//
// Start symbols only occur as a prefix run (summarize.py:441-443), so the chunk tables carry, next to the 4^r
// plain symbol combinations of a chunk of r positions, the sum_{s=1..r} 4^(r-s) patterns with s leading
// starts (<= 85): every k-mer takes the table path, there is no per-position side path and no atomics.
constexpr int T2_THREADS = 512;
constexpr int T2_NW = T2_THREADS / 32;
// G[q] += (s0, s1, s2, s3): plain read-modify-write of one 32-byte table row (two 128-bit accesses)
__device__ __forceinline__ void rmw_row(double* Gc, int q, double s0, double s1, double s2, double s3) {
    const int sw = half_swizzle(q);
    double2* lo = reinterpret_cast<double2*>(Gc + q * 4 + sw);
    double2* hi = reinterpret_cast<double2*>(Gc + q * 4 + (sw ^ 2));
    double2 a = *lo, b = *hi;
    a.x += s0;
    a.y += s1;
    b.x += s2;
    b.y += s3;
    *lo = a;
    *hi = b;
}

// Kernel experiments, compiled in with BEAR_NVCC_EXTRA=-DBEAR_TRAIN_EXPERIMENTS (bear_b200/build.py): with
// BEAR_TRAIN_DEBUG=1 in the environment the consumers skip the scatter -- wrong results, timing of the producer side
// only (tools/ab_train.py; profiles/README.md "Experiments")
#ifdef BEAR_TRAIN_EXPERIMENTS
__constant__ int g_train_debug = 0;
// [0] cycles the producer warps spent computing (summed over warps and CTAs), [1] the same for the consumer warps'
// scatter, [2] cycles between the first and the last barrier summed over CTAs, [3] producer warps, [4] consumer warps:
// busy fraction of a role = [0 or 1] / ([2] * warps of that role per CTA)
__device__ unsigned long long g_train_cycles[5];
#define BEAR_TRAIN_SKIP_SCATTER (g_train_debug != 0)
#define BEAR_TRAIN_CLOCK() clock64()
#else
#define BEAR_TRAIN_SKIP_SCATTER false
#define BEAR_TRAIN_CLOCK() 0ll
#endif

struct Train2Layout {                // offsets in bytes from the start of dynamic shared memory
    int R, G, stage_g, stage_q, stage_m, tags, symtab, tab_lg, tab_dg, stir, red, total;
};

__host__ __device__ inline Train2Layout train2_layout(int nch, int tpw) {
    const int tiles = (T2_NW - nch) * tpw;
    Train2Layout L;
    int o = 0;
    L.R = o;        o += nch * ENT * 4 * 8;
    L.G = o;        o += nch * ENT * 4 * 8;
    L.stage_g = o;  o += 2 * tiles * 4 * 32 * 8;           // [buf][tile][letter][lane] double
    L.stage_q = o;  o += 2 * tiles * nch * 32 * 2;         // [buf][tile][chunk][lane] uint16
    L.stage_m = o;  o += ((2 * tiles * 4 + 15) / 16) * 16; // [buf][tile] uint32 live-row masks
    L.tags = o;     o += ((2 * tiles * nch + 15) / 16) * 16; // [buf][tile][chunk] uint8: largest rank of a row among its key's rows
    L.symtab = o;   o += 2 * ENT * 2;                      // [size class][entry] packed symbols
    L.tab_lg = o;   o += TABN * 8;
    L.tab_dg = o;   o += TABN * 8;
    L.stir = o;     o += STIR_N * 8;
    L.red = o;      o += 32 * 8;
    L.total = o;
    return L;
}

template <bool TRAIN_AR>
__global__ void __launch_bounds__(T2_THREADS, 1)
linear_train2_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ col, int64_t stride,
                     int64_t n, int lag, const ChunkKeys ck, int tpw, const double* __restrict__ mat,
                     const double* __restrict__ h_signed, double* __restrict__ ll_out, double* __restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nch = num_chunks(lag);
    const Train2Layout L = train2_layout(nch, tpw);
    double* R = reinterpret_cast<double*>(smem_raw + L.R);                // [nch][ENT][4] forward ratios
    double* G = reinterpret_cast<double*>(smem_raw + L.G);                // [nch][ENT][4] d ll / d chunk logits
    double* stage_g = reinterpret_cast<double*>(smem_raw + L.stage_g);
    uint16_t* stage_q = reinterpret_cast<uint16_t*>(smem_raw + L.stage_q);
    uint32_t* stage_m = reinterpret_cast<uint32_t*>(smem_raw + L.stage_m);
    uint8_t* stage_r = smem_raw + L.tags;
    uint16_t* symtab = reinterpret_cast<uint16_t*>(smem_raw + L.symtab);
    double* tab_lg = reinterpret_cast<double*>(smem_raw + L.tab_lg);
    double* tab_dg = reinterpret_cast<double*>(smem_raw + L.tab_dg);
    float* stir = reinterpret_cast<float*>(smem_raw + L.stir);
    double* red = reinterpret_cast<double*>(smem_raw + L.red);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = (T2_NW - nch) * tpw;                  // tiles staged per iteration
    const double hinv = exp(-h_signed[0]);                  // 1 / h,  h = exp(h_signed)  (bear_net.py:186)

    // ---------------- tables ----------------
    for (int i = threadIdx.x; i < nch * ENT * 4; i += blockDim.x) G[i] = 0.0;
    for (int i = threadIdx.x; i < STIR_N; i += blockDim.x) stir[i] = float(kStirling[i]);
    for (int i = threadIdx.x; i < 2 * tiles; i += blockDim.x) stage_m[i] = 0u;
    if (!TRAIN_AR && threadIdx.x < TABN) {
        // the concentrations of a row sum to 1/h + 5 eps whatever its k-mer (softmax sums to 1)
        const LgDg t = lgdg_diff<true>(hinv + A1 * BEAR_EPS, double(threadIdx.x));
        tab_lg[threadIdx.x] = t.add + (t.mul == 1.0 ? 0.0 : log(t.mul));
        tab_dg[threadIdx.x] = t.dg;
    }
    build_ext_tables(mat, R, symtab, lag, ck);
    __syncthreads();

    double acc_add = 0.0, dh_sum = 0.0;
    LogProdLong acc_prod;
    const int64_t ntiles = (n + 31) >> 5;
    const int64_t per_iter = int64_t(gridDim.x) * tiles;
    const int64_t niter = (ntiles + per_iter - 1) / per_iter;
    const bool producer = warp >= nch;
    const int slot0 = (warp - nch) * tpw;                   // this producer's first stage slot

    // tile of (iteration it, slot s): consecutive CTAs take consecutive blocks of `tiles` tiles
    auto tile_of = [&](int64_t it, int s) { return (it * gridDim.x + blockIdx.x) * tiles + s; };

    RowIn nxt;
    if (producer) nxt = load_row(kmers, col, stride, (tile_of(0, slot0) << 5) + lane, niter > 0 ? n : 0);

    long long busy = 0;
    const long long t_begin = BEAR_TRAIN_CLOCK();
    for (int64_t it = 0; it <= niter; ++it) {
        const int buf = int(it & 1);
        const long long t_it = BEAR_TRAIN_CLOCK();
        if (producer) {
            if (it < niter) {
                for (int k = 0; k < tpw; ++k) {
                    const int slot = slot0 + k;
                    const int64_t i = (tile_of(it, slot) << 5) + lane;
                    const RowIn cur = nxt;
                    {   // prefetch the next tile of this warp (next slot, or the first slot of the next iteration)
                        const bool last = k + 1 == tpw;
                        const int64_t ti = last ? tile_of(it + 1, slot0) : tile_of(it, slot + 1);
                        nxt = load_row(kmers, col, stride, (ti << 5) + lane, (last && it + 1 >= niter) ? 0 : n);
                    }
                    Counts r;
#pragma unroll
                    for (int b = 0; b < A1; ++b) r.c[b] = cur.c[b];
                    r.cmax = max(max(max(r.c[0], r.c[1]), max(r.c[2], r.c[3])), r.c[4]);
                    if (r.cmax < (1u << 29))
                        r.n = double((r.c[0] + r.c[1]) + (r.c[2] + r.c[3]) + r.c[4]);
                    else
                        r.n = (double(r.c[0]) + double(r.c[1])) + (double(r.c[2]) + double(r.c[3])) + double(r.c[4]);
                    const bool in_range = i < n;
                    const bool live = r.cmax != 0;          // zero-count row: ll = 0 and every gradient is 0
                    const uint32_t steps = warp_steps(live, r.cmax);
                    // ---- head: product of chunk-table rows; the keys are staged for the consumers ----
                    double f[A1];
                    {
                        const int ns = int(cur.code >> 58);
                        const uint64_t v = cur.code & PAYLOAD_MASK;
                        uint16_t* sq = stage_q + ((buf * tiles + slot) * nch) * 32 + lane;
                        double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
                        int sh = 2 * lag, c0 = 0;
                        for (int ch = 0; ch < nch; ++ch) {
                            const int rr = ck.base + (ch < ck.extra ? 1 : 0);
                            sh -= 2 * rr;
                            int q = int(uint32_t(v >> sh) & ((1u << (2 * rr)) - 1u));
                            if (ns > c0) q = ext_key(uint32_t(q), rr, ns - c0);
                            c0 += rr;
                            // rank of this row among the tile's rows with the same key: consumers apply rank 0 rows,
                            // then rank 1 rows, ... so equal keys never meet in one read-modify-write round
                            // (match.any is slow: the table gather below is issued before its result is used)
                            const unsigned grp = __match_any_sync(0xffffffffu, live ? q : 0x10000 + lane);
                            const int sw = half_swizzle(q);
                            const double2 a = *reinterpret_cast<const double2*>(R + (ch * ENT + q) * 4 + sw);
                            const double2 b = *reinterpret_cast<const double2*>(R + (ch * ENT + q) * 4 + (sw ^ 2));
                            p0 *= a.x;
                            p1 *= a.y;
                            p2 *= b.x;
                            p3 *= b.y;
                            const int rank = __popc(grp & ((1u << lane) - 1u));
                            const int maxr = __reduce_max_sync(0xffffffffu, live ? rank : 0);
                            sq[ch * 32] = uint16_t(q | (rank << 10));
                            if (lane == 0) stage_r[(buf * tiles + slot) * nch + ch] = uint8_t(maxr);
                        }
                        const double z = 1.0 + ((p0 + p1) + (p2 + p3));
                        if (z < 1e300 && z > 1e-300) {
                            const double zi = 1.0 / z;
                            f[0] = p0 * zi;
                            f[1] = p1 * zi;
                            f[2] = p2 * zi;
                            f[3] = p3 * zi;
                            f[4] = zi;
                        } else {
                            linear_head_exact(mat, cur.code, lag, f);
                        }
                    }
                    double g[A1] = {0, 0, 0, 0, 0}, ll_row = 0.0;
                    {
                        double add, prod, w[A1];
                        if (TRAIN_AR) {
                            double p[A1], ri[A1];
#pragma unroll
                            for (int b = 0; b < A1; ++b) p[b] = f[b] + BEAR_EPS;                 // bear_net.py:68
                            mn_term(p, r, add, prod);
                            inv5(p, ri);
                            double u = 0.0;
#pragma unroll
                            for (int b = 0; b < A1; ++b) {
                                w[b] = double(r.c[b]) * ri[b];                                   // d ll / d f_b
                                u = fma(f[b], w[b], u);
                            }
#pragma unroll
                            for (int b = 0; b < A1; ++b) g[b] = f[b] * (w[b] - u);
                        } else {
                            double conc[A1];
#pragma unroll
                            for (int b = 0; b < A1; ++b) conc[b] = fma(f[b], hinv, BEAR_EPS);       // bear_net.py:43
                            letters_term<true>(stir, conc, r, steps, add, prod, w);
                            double tadd, tdg;
                            if (r.n < double(TABN)) {
                                tadd = tab_lg[int(r.n)];
                                tdg = tab_dg[int(r.n)];
                            } else {
                                double tprod;
                                const double s = ((conc[0] + conc[1]) + (conc[2] + conc[3])) + conc[4];
                                total_term<true>(s, r, tadd, tprod, tdg);
                                tadd += log(tprod);
                            }
                            add -= tadd;
                            // d ll/d conc_b = w_b - tdg; d ll/d f_b = that / h; softmax backward:
                            // g_b = f_b (d ll/d f_b - sum_j f_j d ll/d f_j) = f_b (w_b - W) / h,  W = sum_j f_j w_j
                            double W = 0.0;
#pragma unroll
                            for (int b = 0; b < A1; ++b) W = fma(f[b], w[b], W);
                            if (live) dh_sum -= (W - tdg) * hinv;   // d ll / d h_signed = -sum_b f_b d ll/d f_b
#pragma unroll
                            for (int b = 0; b < A1; ++b) g[b] = f[b] * hinv * (w[b] - W);
                        }
                        if (live) {
                            if (ll_out) {
                                ll_row = add + log(prod);
                                acc_add += ll_row;
                            } else {
                                acc_add += add;
                                acc_prod.push(0.0, prod);
                            }
                        }
                    }
                    if (ll_out && in_range) ll_out[i] = ll_row;
                    double* sg = stage_g + ((buf * tiles + slot) * 4) * 32 + lane;
#pragma unroll
                    for (int b = 0; b < 4; ++b) sg[b * 32] = g[b];
                    const unsigned m = __ballot_sync(0xffffffffu, live);
                    if (lane == 0) stage_m[buf * tiles + slot] = m;
                }
            }
        } else if (it > 0) {
            // ---------------- consumer: warp ch scatters chunk ch of every tile staged in iteration it-1 --------
            // Rows with equal keys would collide on one table entry.  The producer ranked every row among the rows of
            // its tile with the same key, so round r applies the rank-r rows with plain read-modify-writes: distinct
            // keys within a round, no atomics, no retry loop (random 256-key chunks need 2 rounds 6 times out of 7).
            // Keys with more than five rows in a tile (the leading chunks of a table sorted by k-mer) have their
            // remaining rows summed with a butterfly first.
            const int pb = buf ^ 1;
            const int ch = warp;
            double* Gc = G + ch * ENT * 4;
            const uint32_t* sm = stage_m + pb * tiles;
            const uint8_t* sr = stage_r + pb * tiles * nch + ch;
            const uint16_t* sq = stage_q + (pb * tiles * nch + ch) * 32 + lane;
            const double* sg = stage_g + pb * tiles * 4 * 32;
            // the staged row of the next tile is fetched while the current one is scattered
            unsigned m_n = sm[0];
            int q_n = int(*sq), r_n = int(*sr);
            double n0 = sg[lane], n1 = sg[32 + lane], n2 = sg[64 + lane], n3 = sg[96 + lane];
            for (int t = 0; t < tiles; ++t, sq += nch * 32, sg += 4 * 32, sr += nch) {
                const unsigned m = m_n;
                const bool pend = (m >> lane) & 1u;
                const int q = q_n & 1023, rank = q_n >> 10, rounds = r_n;
                double s0 = n0, s1 = n1, s2 = n2, s3 = n3;
                if (t + 1 < tiles) {
                    m_n = sm[t + 1];
                    q_n = int(sq[nch * 32]);
                    r_n = int(sr[nch]);
                    n0 = sg[128 + lane];
                    n1 = sg[160 + lane];
                    n2 = sg[192 + lane];
                    n3 = sg[224 + lane];
                }
                if (m == 0u || BEAR_TRAIN_SKIP_SCATTER) continue;
                bool todo = pend;
                int plain = rounds;
                if (rounds > 4) {
                    // keys with more than five rows in the tile (at most five such keys; the rule for the leading
                    // chunks of a table sorted by k-mer): all rows of such a key are summed with a butterfly and
                    // applied by the key's rank-5 lane
                    unsigned heads = __ballot_sync(0xffffffffu, pend && rank == 5);
                    while (heads) {
                        const int L = __ffs(heads) - 1;
                        heads &= heads - 1;
                        const int qL = __shfl_sync(0xffffffffu, q, L);       // every lane takes part in the shuffle
                        const bool mem = pend && q == qL;
                        const double t0 = warp_sum(mem ? s0 : 0.0), t1 = warp_sum(mem ? s1 : 0.0);
                        const double t2 = warp_sum(mem ? s2 : 0.0), t3 = warp_sum(mem ? s3 : 0.0);
                        if (lane == L) rmw_row(Gc, q, t0, t1, t2, t3);
                        todo = todo && !mem;
                    }
                    __syncwarp();
                    plain = __any_sync(0xffffffffu, todo) ? 4 : -1;      // the other keys have at most five rows each
                }
                for (int r = 0; r <= plain; ++r) {
                    if (todo && rank == r) rmw_row(Gc, q, s0, s1, s2, s3);
                    __syncwarp();
                }
            }
        }
        busy += BEAR_TRAIN_CLOCK() - t_it;
        __syncthreads();
    }
#ifdef BEAR_TRAIN_EXPERIMENTS
    if (lane == 0) {
        atomicAdd(&g_train_cycles[producer ? 0 : 1], (unsigned long long)busy);
        atomicAdd(&g_train_cycles[producer ? 3 : 4], 1ull);
        if (threadIdx.x == 0) atomicAdd(&g_train_cycles[2], (unsigned long long)(clock64() - t_begin));
    }
#else
    (void)busy;
    (void)t_begin;
#endif

    const int P = 2 + lag * A1 * A1;
    double* out = partials + int64_t(blockIdx.x) * P;
    const double ll_thread = acc_add + acc_prod.value();
    const double ll_blk = block_sum(ll_thread, red);
    const double dh_blk = block_sum(dh_sum, red);
    if (threadIdx.x == 0) {
        out[0] = ll_blk;
        out[1] = dh_blk;
    }
    __syncthreads();
    // marginalise the chunk-table gradients back onto mat[j, s, b] (s = 4: the start symbol)
    for (int idx = threadIdx.x; idx < lag * A1 * A1; idx += blockDim.x) {
        const int b = idx % A1, s = (idx / A1) % A1, j = idx / (A1 * A1);
        int ch = 0;
        ChunkGeom cg = chunk_geom(lag, nch, 0);
        while (j >= cg.start + cg.size) cg = chunk_geom(lag, nch, ++ch);
        const int p = j - cg.start;
        const uint16_t* st = symtab + (ch < ck.extra ? 0 : ENT);
        const int ne = ext_entries(cg.size);
        double acc = 0.0;
        for (int e = 0; e < ne; ++e) {
            if (int((st[e] >> (3 * p)) & 7u) != s) continue;
            const double* src = G + (ch * ENT + e) * 4;
            if (b < 4) acc += src[((b & 2) ^ half_swizzle(e)) + (b & 1)];
            else acc -= (src[0] + src[1]) + (src[2] + src[3]);   // the 5 logit gradients sum to 0
        }
        out[2 + idx] = acc;
    }
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

#ifdef BEAR_TRAIN_EXPERIMENTS
// reads and clears the cycle counters of linear_train2_kernel (synchronises the device)
extern "C" int bear_debug_train_cycles(unsigned long long* out5) {
    BEAR_CUDA_CHECK(cudaDeviceSynchronize());
    BEAR_CUDA_CHECK(cudaMemcpyFromSymbol(out5, g_train_cycles, 5 * sizeof(unsigned long long)));
    const unsigned long long zero[5] = {0, 0, 0, 0, 0};
    BEAR_CUDA_CHECK(cudaMemcpyToSymbol(g_train_cycles, zero, sizeof(zero)));
    return BEAR_OK;
}
#endif

extern "C" int bear_linear_train_step_legacy(const uint64_t* d_kmers, const uint32_t* d_col, int64_t stride,
                                      int64_t row0, int64_t n, int lag, const double* d_mat,
                                      const double* d_h_signed, double scale, int train_ar,
                                      double* d_flat, double* d_ll_out, double* d_workspace, void* stream) {
    const char* fn = "bear_linear_train_step";
    BEAR_REQUIRE(d_kmers && d_col && d_mat && d_h_signed && d_flat && d_workspace, fn);
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    BEAR_REQUIRE(lag >= 1 && lag <= 29, fn);
    if (n == 0) return BEAR_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int P = 2 + lag * A1 * A1;
    const ChunkKeys ck = make_chunk_keys(lag);
    const int nch = num_chunks(lag);
    const int tpw = size_t(train2_layout(nch, 2).total) <= size_t(227 * 1024) ? 2 : 1;
    const size_t smem2 = size_t(train2_layout(nch, tpw).total);
    const int64_t ntiles = (n + 31) / 32, per_cta = int64_t(T2_NW - nch) * tpw;
    const int64_t want = (ntiles + per_cta - 1) / per_cta;
    const int grid2 = int(want < 148 ? want : 148);
#ifdef BEAR_TRAIN_EXPERIMENTS
    static const int debug = getenv("BEAR_TRAIN_DEBUG") ? atoi(getenv("BEAR_TRAIN_DEBUG")) : 0;
    if (debug) BEAR_CUDA_CHECK(cudaMemcpyToSymbolAsync(g_train_debug, &debug, sizeof(int), 0, cudaMemcpyHostToDevice, st));
#endif
    if (train_ar) {
        if (set_smem(linear_train2_kernel<true>, smem2)) return BEAR_ERR_CUDA;
        linear_train2_kernel<true><<<grid2, T2_THREADS, smem2, st>>>(d_kmers + row0, d_col + row0, stride, n, lag, ck, tpw,
                                                                    d_mat, d_h_signed, d_ll_out, d_workspace);
    } else {
        if (set_smem(linear_train2_kernel<false>, smem2)) return BEAR_ERR_CUDA;
        linear_train2_kernel<false><<<grid2, T2_THREADS, smem2, st>>>(d_kmers + row0, d_col + row0, stride, n, lag, ck, tpw,
                                                                     d_mat, d_h_signed, d_ll_out, d_workspace);
    }
    BEAR_LAUNCH_CHECK("linear_train2_kernel");
    reduce_partials_kernel<<<(P + 127) / 128, 128, 0, st>>>(d_workspace, grid2, P, -scale, d_flat);
    BEAR_LAUNCH_CHECK("reduce_partials_kernel");
    return BEAR_OK;
}

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
                              const double* d_h, int H, const double* d_van, int V, int64_t seed,
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
        case BEAR_HEAD_LINEAR: rc = launch_eval_nm<BEAR_HEAD_LINEAR>(small, ht, grid, smem, st, km, te, tr, stride, row0, n, lag, d_head, d_h, H, d_van, V, seed, d_workspace); break;
        case BEAR_HEAD_EXPLICIT: rc = launch_eval_nm<BEAR_HEAD_EXPLICIT>(small, ht, grid, smem, st, km, te, tr, stride, row0, n, lag, d_head, d_h, H, d_van, V, seed, d_workspace); break;
        case BEAR_HEAD_STOP: rc = launch_eval_nm<BEAR_HEAD_STOP>(small, ht, grid, smem, st, km, te, tr, stride, row0, n, lag, d_head, d_h, H, d_van, V, seed, d_workspace); break;
        default: rc = launch_eval_nm<BEAR_HEAD_NONE>(small, ht, grid, smem, st, km, te, tr, stride, row0, n, lag, d_head, d_h, H, d_van, V, seed, d_workspace); break;
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
