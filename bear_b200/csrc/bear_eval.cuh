// Evaluation of a BEAR / AR / BMM model over a packed DNA/RNA table (A1 = 5):
//   bear_net._evaluation_step / evaluation (bear_net.py:323-463), h_scan (bear_net.py:516-531),
//   bear_ref._evaluation_step / evaluation (bear_ref.py:391-539) with the Jukes-Cantor reference head (bear_ref.py:9-69).
// eval_tile_kernel: one persistent CTA of 16 warps per SM; every warp is an independent row pipeline over tiles of 32
// rows.  A warp owns a ring of shared-memory stages that its first lanes fill with 1-D bulk async copies (TMA): the
// k-mer plane (heads that read the k-mer) and the five count planes of the test column, of the conditioning column and
// of the reference column, as far as the launch has them; completion on an mbarrier.  One pass gives the seven
// accumulators of the evaluation: BEAR log-likelihood for up to 8 h values, AR log-likelihood, BMM log-likelihood for
// up to 8 priors, the test counts at the three models' predicted letters, and the total test count.
// The kernel template lives here; the head variants are instantiated in bear_eval_{misc,lin,ref}.cu (compiled in
// parallel) and dispatched from bear_eval_step / bear_ref_eval_step (bear_fused.cu).
#pragma once
#include <math.h>

#include "bear_b200.h"
#include "bear_host.h"
#include "bear_linear_head.cuh"
#include "bear_sm100.cuh"

// net of the reference head of bear_ref (bear_ref.py:63-68): BEAR_HEAD_STOP or BEAR_HEAD_LINEAR, shifted past the plain heads
#define BEAR_HEAD_REF_STOP 4
#define BEAR_HEAD_REF_LINEAR 5

namespace bear_eval {

struct EvalArgs {
    const uint64_t* kmers;           // table base pointers (not offset by row0)
    const uint32_t *test_col, *train_col, *ref_col;
    int64_t stride, row0, n, row_id0;
    int lag, head;
    const double *head_ptr, *tau_signed, *nw_signed, *d_h, *d_van;
    int H, V;
    int64_t seed;
    double* ws;
    cudaStream_t stream;
};

// one function per translation unit: launches eval_tile_kernel for its head variants; returns the grid size (> 0) or a
// negative bear_status
int launch_misc(const EvalArgs& a);
int launch_linear(const EvalArgs& a);
int launch_ref(const EvalArgs& a);

}  // namespace bear_eval

#ifdef BEAR_EVAL_IMPL

namespace {

using namespace bear;
using namespace bear::sm100;

// count of letter idx; the evaluation sums these as integers (exact, and no int -> double conversion per row)
__device__ __forceinline__ uint32_t pick5(const uint32_t (&c)[A1], int idx) {
    return idx == 0 ? c[0] : idx == 1 ? c[1] : idx == 2 ? c[2] : idx == 3 ? c[3] : c[4];
}

// argmax of v + sigma * N(0,1) (core.py:69-71,134-136).  Candidates are the entries within 8 sigma of the maximum (the
// noise lifts anything further over the maximum with probability Phi(-8 / sqrt 2) < 1e-8, and only the few per cent of
// rows whose runner-up lies between 8 and 16 sigma would even be exposed to that).  One candidate: no randomness
// needed.  All candidates exactly tied: a uniform pick, which is what iid noise gives.  Two candidates: the difference
// of their noises is one N(0, 2 sigma^2) draw.  Otherwise Gaussian noise on the candidates only.  seed < 0: no noise,
// first maximum wins.  (With small weights a few per cent of rows have a near-tie, i.e. most warps visit the slow path
// once per tile: its cost is visible -- 10 % of the evaluation kernel on the bench's table before these two cuts.)
struct V5 {
    double v[A1];
};

__device__ __forceinline__ uint32_t tie_hash(int64_t seed, uint64_t row, uint64_t model) {
    return uint32_t(mix64(uint64_t(seed) ^ (row * 0x9E3779B97F4A7C15ull) ^ (model * 0xD1B54A32D192ED03ull)) >> 32);
}

// the randomised part, out of line: ties are rare except for the unconditioned BMM (handled separately)
__device__ __noinline__ int argmax_tiebreak(V5 x, double top, double thr, int near, double sigma,
                                            int64_t seed, uint64_t row, uint64_t model) {
    int best = 0, exact = 0;
    for (int b = 0; b < A1; ++b) exact += x.v[b] == top;
    const bool all_exact = near == exact;
    if (all_exact) {
        int k = int(tie_hash(seed, row, model) % uint32_t(exact));
        for (int b = 0; b < A1; ++b)
            if (x.v[b] == top) {
                if (k == 0) best = b;
                --k;
            }
        return best;
    }
    int ia = -1, ib = -1, ncand = 0;
    for (int b = 0; b < A1; ++b)
        if (x.v[b] > thr) {
            if (ncand == 0) ia = b; else ib = b;
            ++ncand;
        }
    if (ncand == 2) {        // v_a + s n_a > v_b + s n_b  <=>  (v_a - v_b) + s sqrt(2) n > 0 with a single standard normal n
        const double n = rng_normal_fast(uint64_t(seed), row, model * 64);
        return (x.v[ia] - x.v[ib]) + 1.4142135623730951 * sigma * n > 0.0 ? ia : ib;
    }
    double nb = -INFINITY;
    for (int b = 0; b < A1; ++b)
        if (x.v[b] > thr) {
            const double y = x.v[b] + sigma * rng_normal_fast(uint64_t(seed), row, model * 64 + uint64_t(b));
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
    const double thr = top - 8.0 * sigma;
    int near = 0;                       // candidates (the exact ties among them are only counted on the slow path)
#pragma unroll
    for (int b = 0; b < A1; ++b) near += v[b] > thr;
    if (near == 1) return best;
    V5 x;
#pragma unroll
    for (int b = 0; b < A1; ++b) x.v[b] = v[b];
    return argmax_tiebreak(x, top, thr, near, sigma, seed, row, model);
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


// the same for a prior too large for the constant-a form (rare: priors >= 64 / 5): general differences, out of line
__device__ __noinline__ double van_general_row(CountVec<A1> cv, double rn, double vk) {
    LogProd num, den;
    den.push(lgdg_diff<false>(double(A1) * vk, rn));
    for (int b = 0; b < A1; ++b) num.push(lgdg_diff<false>(vk, double(cv.c[b])));
    return logprod_diff(num, den);
}

// row-independent total term past the table, out of line (the sparse regime never gets here)
__device__ __noinline__ double lg_shift_large_cold(double a, double c, double K) { return lg_shift_large(a, c, K); }

// warps per CTA: 16 for the common call (one h, up to four priors); the eight-model variants (h_scan, many priors) keep
// 16 + 16 running accumulators per thread and run 8 warps with up to 255 registers instead of spilling
#ifndef BEAR_EV_WARPS
#define BEAR_EV_WARPS 16
#endif
__host__ __device__ constexpr int ev_warps(int NH, int NV) { return (NH <= 1 && NV <= 4) ? BEAR_EV_WARPS : 8; }
__host__ __device__ constexpr int ev_ctas(int NH, int NV) { return (NH <= 1 && NV <= 4 && BEAR_EV_WARPS <= 16) ? 16 / BEAR_EV_WARPS : 1; }
constexpr int EV_MAX_NW = 32;
constexpr int EV_MAX_STAGES = 4;

struct EvalLayout {                  // offsets in bytes from the start of dynamic shared memory
    int R, symtab, red, tab_ear, tab_van, tab_vtot, kconst, stir, bars, ring, total;
};

__host__ __device__ constexpr EvalLayout eval_layout(bool lin, int nch, int nm, int stage_bytes, int nstage, int nw) {
    EvalLayout L{};
    int o = 0;
    L.R = o;         o += lin ? nch * ENT * 4 * 8 : 0;
    L.symtab = o;    o += lin ? ((2 * ENT * 2 + 15) / 16) * 16 : 0;
    L.red = o;       o += 32 * 8;
    L.tab_ear = o;   o += nm * TABN * 8;
    L.tab_van = o;   o += nm * TABN * 8;
    L.tab_vtot = o;  o += nm * TABN * 8;
    L.kconst = o;    o += 3 * nm * 8;
    L.stir = o;      o += ((STIR_N * 8 + 15) / 16) * 16;
    L.bars = o;      o += EV_MAX_NW * EV_MAX_STAGES * 8;
    o = (o + 127) & ~127;
    L.ring = o;      o += nw * nstage * stage_bytes;
    L.total = o;
    return L;
}

// NH / NV bound the number of h values (H) and of BMM priors (V) of one launch.
template <int HEAD, int NH, int NV, bool HAS_TRAIN>
__global__ void __launch_bounds__(32 * ev_warps(NH, NV), ev_ctas(NH, NV))
eval_tile_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ test_col, const uint32_t* __restrict__ train_col,
                 const uint32_t* __restrict__ ref_col, int64_t stride, int64_t row_lo, int64_t row_hi, int64_t row_id0, int lag,
                 const ChunkKeys ck, const HeadGeom hg, int nstage, int use_tma, const double* __restrict__ head, const double* __restrict__ tau_signed,
                 const double* __restrict__ nw_signed, const double* __restrict__ d_h, int H, const double* __restrict__ d_van, int V,
                 int64_t seed, double* __restrict__ partials) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool REF = HEAD == BEAR_HEAD_REF_STOP || HEAD == BEAR_HEAD_REF_LINEAR;
    constexpr bool LIN = HEAD == BEAR_HEAD_LINEAR || HEAD == BEAR_HEAD_REF_LINEAR;      // the net is the chunk-table head
    constexpr bool STOP = HEAD == BEAR_HEAD_STOP || HEAD == BEAR_HEAD_REF_STOP;
    // the sum of a row's BEAR concentrations is row-independent when there is no conditioning column
    // and the head is normalised (or absent)
    constexpr bool TOT_TAB = !HAS_TRAIN && HEAD != BEAR_HEAD_EXPLICIT;
    constexpr int NM = NH > NV ? NH : NV;
    constexpr int EV_NW = ev_warps(NH, NV);
    // planes of a stage: [k-mers 256 B] [test 5 x 128 B] [train 5 x 128 B] [ref 5 x 128 B]
    constexpr int OFF_TEST = LIN ? 256 : 0, OFF_TRAIN = OFF_TEST + 640, OFF_REF = OFF_TRAIN + (HAS_TRAIN ? 640 : 0);
    constexpr int STAGE = OFF_REF + (REF ? 640 : 0);
    constexpr int NPLANES = (LIN ? 1 : 0) + 5 + (HAS_TRAIN ? 5 : 0) + (REF ? 5 : 0);
    const int nch = num_chunks(lag);
    const EvalLayout L = eval_layout(LIN, nch, NM, STAGE, nstage, EV_NW);
    double* R = reinterpret_cast<double*>(smem_raw + L.R);                 // [nch][ENT][4] extended ratio tables
    uint16_t* symtab = reinterpret_cast<uint16_t*>(smem_raw + L.symtab);
    double* red = reinterpret_cast<double*>(smem_raw + L.red);
    double* tab_ear = reinterpret_cast<double*>(smem_raw + L.tab_ear);     // [NM][TABN]  lgamma(S0_k + N) - lgamma(S0_k)
    double* tab_van = reinterpret_cast<double*>(smem_raw + L.tab_van);     // [NM][TABN]  lgamma(van_k + eps + c) - lgamma(van_k + eps)
    double* tab_vtot = reinterpret_cast<double*>(smem_raw + L.tab_vtot);   // [NM][TABN]  lgamma(5 (van_k + eps) + N) - lgamma(5 (van_k + eps))
    double* kconst = reinterpret_cast<double*>(smem_raw + L.kconst);       // [3][NM] ln(2 pi)/2 - lgamma(a) of the three table families
    double* stir = reinterpret_cast<double*>(smem_raw + L.stir);           // Stirling triangle
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t my_in = smem_u32(smem_raw + L.bars) + 8 * warp * EV_MAX_STAGES;
    const uint32_t ring = smem_u32(smem_raw + L.ring) + warp * nstage * STAGE;

    for (int i = threadIdx.x; i < STIR_N; i += blockDim.x) stir[i] = kStirling[i];
    if (lane == 0) {
        for (int s = 0; s < EV_MAX_STAGES; ++s) mbar_init(my_in + 8 * s, 1);
        mbar_fence_init();
    }
    if (LIN) build_ext_tables(head, R, symtab, lag, ck);
    double hinv[NH], van[NV];
#pragma unroll
    for (int k = 0; k < NH; ++k) hinv[k] = k < H ? 1.0 / d_h[k] : 1.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) van[k] = k < V ? d_van[k] : 1.0;
    // reference head (bear_ref.py:63-68): f = (nw g + jc) / (nw + 1)
    double etau = 0.0, nwt = 0.0, nwi = 1.0;
    if (REF) {
        etau = exp(-exp(tau_signed[0]));
        nwt = exp(nw_signed[0]);
        nwi = 1.0 / (nwt + 1.0);
    }
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

    // Tiles are aligned to absolute multiples of 32 rows (128-byte aligned planes); rows outside [row_lo, row_hi) are dead.
    // Iteration i of warp w of CTA c works on tile (i gridDim + c) EV_NW + w: consecutive warps stream consecutive tiles.
    const int64_t a0 = row_lo & ~int64_t(31);
    const uint32_t ntiles = uint32_t((row_hi - a0 + 31) >> 5);
    const uint32_t tstep = gridDim.x * EV_NW;
    // tiles below t_full lie entirely below row_hi and are fetched by the TMA engine; the tail tile uses guarded loads
    const uint32_t t_full = use_tma ? uint32_t((row_hi - a0) >> 5) : 0u;
    uint32_t t = blockIdx.x * EV_NW + warp;
    const uint32_t n_my = t < ntiles ? (ntiles - t + tstep - 1) / tstep : 0u;       // tiles of this warp
    // lane p < NPLANES copies plane p of a tile
    auto issue_tile = [&](uint32_t ti, int stg) {
        const uint32_t bar = my_in + 8 * stg, dst = ring + stg * STAGE;
        const int64_t r0 = a0 + (int64_t(ti) << 5);
        if (lane == 0) mbar_arrive_expect_tx(bar, STAGE);
        int p = lane;
        if (LIN) {
            if (p == 0) {
                bulk_g2s(dst, kmers + r0, 256, bar);
                return;
            }
            --p;
        }
        const uint32_t* src = p < 5 ? test_col : (HAS_TRAIN && p < 10) ? train_col : ref_col;
        const int off = p < 5 ? OFF_TEST : (HAS_TRAIN && p < 10) ? OFF_TRAIN : OFF_REF;
        const int b = p % 5;
        bulk_g2s(dst + off + b * 128, src + b * stride + r0, 128, bar);
    };
    if (lane < NPLANES) {
        for (int s = 0; s < nstage; ++s)
            if (uint32_t(s) < n_my && t + s * tstep < t_full) issue_tile(t + s * tstep, s);
    }
    int stg = 0;
    uint32_t in_par = 0;
    for (uint32_t it = 0; it < n_my; ++it, t += tstep) {
        const int64_t arow = a0 + (int64_t(t) << 5) + lane;
        uint64_t code = 0ull;
        Counts r;
        uint32_t tc[A1] = {0, 0, 0, 0, 0}, rc[A1] = {0, 0, 0, 0, 0};
        if (t < t_full) {
            mbar_wait(my_in + 8 * stg, in_par);
            const uint32_t st = ring + stg * STAGE;
            if (LIN) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(code) : "r"(st + lane * 8));
#pragma unroll
            for (int b = 0; b < A1; ++b) {
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r.c[b]) : "r"(st + OFF_TEST + b * 128 + lane * 4));
                if (HAS_TRAIN) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tc[b]) : "r"(st + OFF_TRAIN + b * 128 + lane * 4));
                if (REF) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rc[b]) : "r"(st + OFF_REF + b * 128 + lane * 4));
            }
        } else {
            const bool ok = arow < row_hi;
            if (LIN) code = ok ? __ldg(kmers + arow) : 0ull;
#pragma unroll
            for (int b = 0; b < A1; ++b) {
                r.c[b] = ok ? __ldg(test_col + b * stride + arow) : 0u;
                if (HAS_TRAIN) tc[b] = ok ? __ldg(train_col + b * stride + arow) : 0u;
                if (REF) rc[b] = ok ? __ldg(ref_col + b * stride + arow) : 0u;
            }
        }
        bool in_range = true;
        if (t == 0 || t >= t_full) {                // only the first and the tail tile can hold rows outside the batch
            in_range = arow >= row_lo && arow < row_hi;
            if (!in_range) {
#pragma unroll
                for (int b = 0; b < A1; ++b) r.c[b] = 0u;
            }
        }
        r.cmax = max(max(max(r.c[0], r.c[1]), max(r.c[2], r.c[3])), r.c[4]);
        if (r.cmax < (1u << 29))
            r.n = double((r.c[0] + r.c[1]) + (r.c[2] + r.c[3]) + r.c[4]);
        else
            r.n = (double(r.c[0]) + double(r.c[1])) + (double(r.c[2]) + double(r.c[3])) + double(r.c[4]);
        const bool live = r.cmax != 0;             // no test transitions: contributes 0 to every output
        const uint32_t steps = warp_steps(live, r.cmax);
        // The test counts stay in the stage until the end of the iteration: the count at a predicted letter is one
        // shared-memory load by index (count_at) instead of a chain of selects.  Tail tiles (guarded loads) select.
        const bool staged = t < t_full;
        const uint32_t cnt_addr = ring + stg * STAGE + OFF_TEST + lane * 4;
        auto count_at = [&](int idx) -> uint32_t {
            if (!staged) return pick5(r.c, idx);
            uint32_t v;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(cnt_addr + uint32_t(idx) * 128u));
            return v;
        };
        double tr[A1] = {0, 0, 0, 0, 0};
        if (HAS_TRAIN && live) {
#pragma unroll
            for (int b = 0; b < A1; ++b) tr[b] = double(tc[b]);
        }
        // ---- head ----
        double f[A1];
        const bool any_start = LIN && __any_sync(0xffffffffu, (code >> 58) != 0ull);
        if (LIN) {
#pragma unroll
            for (int b = 0; b < A1; ++b) f[b] = 0.2;
            if (live) linear_head_geom<0>(R, head, code, lag, hg, nch, f, any_start);
        } else {
#pragma unroll
            for (int b = 0; b < A1; ++b)
                f[b] = HEAD == BEAR_HEAD_EXPLICIT ? (in_range ? head[(arow - row_lo) * A1 + b] : 0.2)
                                                  : ((STOP && b == A1 - 1) ? 1.0 : 0.0);
        }
        if (REF) {
            // Jukes-Cantor mix of the reference counts (bear_ref.py:9-33 after the map of bear_ref.py:332-337):
            // rt = (ref + eps) * not_stop; p = rt / sum |rt|; jc = u + exp(-tau) (p - u), u = [1/4, 1/4, 1/4, 1/4, 0]
            double p[4], s = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                p[b] = double(rc[b]) + BEAR_EPS;
                s += p[b];
            }
            const double si = 1.0 / s;
#pragma unroll
            for (int b = 0; b < 4; ++b) f[b] = (nwt * f[b] + (0.25 + etau * (p[b] * si - 0.25))) * nwi;
            f[4] = nwt * f[4] * nwi;
        }
        total += r.n;
        const uint64_t grow = uint64_t(row_id0 + (arow - row_lo));
        const bool use_tab = !HAS_TRAIN && r.n < double(TABN);
        double dummy[A1];
        uint64_t van_hash = 0;
        // BEAR: conc = f / h + train + eps   (bear_net.py:43, 335-337)
#pragma unroll
        for (int k = 0; k < NH; ++k) {
            if (k < H) {
                double conc[A1], add, prod;
#pragma unroll
                for (int b = 0; b < A1; ++b) conc[b] = fma(f[b], hinv[k], tr[b]) + BEAR_EPS;
                letters_term<false>(stir, conc, r, steps, add, prod, dummy);
                if (TOT_TAB && use_tab) {
                    add -= tab_ear[k * TABN + int(r.n)];
                } else if (TOT_TAB && fast_ear) {
                    add -= lg_shift_large_cold((HEAD == BEAR_HEAD_NONE ? 0.0 : hinv[k]) + A1 * BEAR_EPS, r.n, kconst[k]);
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
                    cor_ear[k] += count_at(noisy_argmax5(conc, 100.0 * BEAR_EPS, seed, grow, uint64_t(k)));
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
                cor_arm += count_at(noisy_argmax5(p, BEAR_EPS, seed, grow, 100));
            }
        }
        // vanilla BMM: conc = train + van + eps   (bear_net.py:328-331, 339-340)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (k < V) {
                double conc[A1];
#pragma unroll
                for (int b = 0; b < A1; ++b) conc[b] = (tr[b] + van[k]) + BEAR_EPS;
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
                        CountVec<A1> cv;
#pragma unroll
                        for (int b = 0; b < A1; ++b) cv.c[b] = r.c[b];
                        van_add[k] += van_general_row(cv, r.n, van[k] + BEAR_EPS);
                    }
                } else {
                    double add, prod, tadd, tprod, tdg;
                    letters_term<false>(stir, conc, r, steps, add, prod, dummy);
                    const double s = ((conc[0] + conc[1]) + (conc[2] + conc[3])) + conc[4];
                    total_term<false>(s, r, tadd, tprod, tdg);
                    if (live) van_add[k] += (add - tadd) + log(prod / tprod);        // (heldout: hot, stays inline)
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
                    cor_van[k] += count_at(best);
                }
            }
        }
        __syncwarp();                               // every lane is done with the stage
        if (lane < NPLANES && it + nstage < n_my && t + nstage * tstep < t_full)
            issue_tile(t + nstage * tstep, stg);    // refill this stage with the tile `nstage` iterations ahead
        if (++stg == nstage) {
            stg = 0;
            in_par ^= 1u;
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

template <int HEAD, int NH, int NV, bool HAS_TRAIN>
int launch_one(const bear_eval::EvalArgs& a) {
    constexpr bool REF = HEAD == BEAR_HEAD_REF_STOP || HEAD == BEAR_HEAD_REF_LINEAR;
    constexpr bool LIN = HEAD == BEAR_HEAD_LINEAR || HEAD == BEAR_HEAD_REF_LINEAR;
    constexpr int NM = NH > NV ? NH : NV;
    constexpr int EV_NW = ev_warps(NH, NV);
    constexpr int STAGE = (LIN ? 256 : 0) + 640 + (HAS_TRAIN ? 640 : 0) + (REF ? 640 : 0);
    const int lag = LIN ? a.lag : 1;
    const int nch = num_chunks(lag);
    int nstage = EV_MAX_STAGES;
    while (nstage > 2 && size_t(eval_layout(LIN, nch, NM, STAGE, nstage, EV_NW).total) > size_t(227 * 1024) / ev_ctas(NH, NV)) --nstage;
    const size_t smem = size_t(eval_layout(LIN, nch, NM, STAGE, nstage, EV_NW).total);
    // bulk copies need 16-byte aligned planes (tiles start at absolute multiples of 32 rows)
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const int use_tma = (a.stride & 3) == 0 && al(a.test_col) && (!LIN || al(a.kmers)) && (!HAS_TRAIN || al(a.train_col)) &&
                        (!REF || al(a.ref_col));
    const int64_t a0 = a.row0 & ~int64_t(31);
    const int64_t ntiles = (a.row0 + a.n - a0 + 31) / 32;
    const int64_t want = (ntiles + EV_NW - 1) / EV_NW;
    const int cap = 148 * ev_ctas(NH, NV);
    const int grid = int(want < cap ? want : cap);
    BEAR_CUDA_CHECK(cudaFuncSetAttribute(eval_tile_kernel<HEAD, NH, NV, HAS_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    eval_tile_kernel<HEAD, NH, NV, HAS_TRAIN><<<grid, 32 * EV_NW, smem, a.stream>>>(
        a.kmers, a.test_col, a.train_col, a.ref_col, a.stride, a.row0, a.row0 + a.n, a.row_id0, lag, make_chunk_keys(lag), make_head_geom(lag), nstage,
        use_tma, a.head_ptr, a.tau_signed, a.nw_signed, a.d_h, a.H, a.d_van, a.V, a.seed, a.ws);
    BEAR_LAUNCH_CHECK("eval_tile_kernel");
    return grid;
}

// the common call is one h value with up to four priors (evaluation); h_scan uses up to eight h values
template <int HEAD>
int launch_head(const bear_eval::EvalArgs& a) {
    const bool small = a.H <= 1 && a.V <= 4, ht = a.train_col != nullptr;
    if (small) return ht ? launch_one<HEAD, 1, 4, true>(a) : launch_one<HEAD, 1, 4, false>(a);
    return ht ? launch_one<HEAD, 8, 8, true>(a) : launch_one<HEAD, 8, 8, false>(a);
}

}  // namespace

#endif  // BEAR_EVAL_IMPL
