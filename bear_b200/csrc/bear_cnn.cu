// Fused CNN autoregressive head (ar_funcs.make_ar_func_cnn, ar_funcs.py:49-99) on packed DNA/RNA k-mers:
//   x0 = scale0 * LN_F(conv1d_VALID(onehot, filters[W,5,F])) + int0          [P = lag-W+1, F]
//   x1 = scale1 * LN(elu(x0) . W1[P,F,H1]) + int1                            [H1]
//   f  = softmax(elu(x1) . W2[H1,5] + int2)                                  [5]
// followed (training) by the Dirichlet-multinomial / multinomial loss of bear_net._train_step
// (bear_net.py:146-197) and the whole backward pass, all in ONE kernel: no activation ever leaves the SM
// (the reference round-trips [B,P,F] float64 activations through memory for every op).
//
// A CTA owns a tile of TB rows per iteration:
//   1. conv + layer norm + elu: one warp per (position, row) pair, lane = filter.  The convolution over
//      a one-hot input is a gather of W filter rows, not a contraction.  Result E0[TB, P*F] in shared memory.
//   2. dense layer 1 on the FP64 tensor cores (mma.sync m8n8k4 f64, "DMMA"):  Y = E0 . W1.
//   3. per row (16 lanes = the H1 units): layer norm, elu, dense layer 2, softmax, the loss gradient
//      d ll / d f, and the backward pass down to dY, written over Y.
//   4. dW1 += E0^T . dY on the tensor cores, accumulators live in registers for the whole kernel.
//   5. dE0 = dY . W1^T on the tensor cores, elu' applied in the epilogue, written over E0.
//   6. layer-norm / conv backward with the partition of step 1 (statistics recomputed, not stored);
//      filter gradients go to a warp-private table (lane = filter => no conflicts, no atomics).
// Per-CTA partial sums land in the caller's workspace; a fixed-order second stage adds them to the flat
// [loss, d h_signed, d params...] buffer.
#include <math.h>

#include "bear_b200.h"
#include "bear_common.cuh"
#include "bear_dm_row.cuh"
#include "bear_host.h"

namespace {

using namespace bear;

// A CTA has NW = 16 warps when its tile fits shared memory (the kernel is latency-bound: more warps hide more), else 8.
__host__ __device__ constexpr int max_tiles(int nw) { return nw == 16 ? 7 : 13; }   // 8x8 dW1 accumulator tiles per warp
constexpr int MAX_SMEM = 227 * 1024;
constexpr uint64_t PAYLOAD_MASK = (1ull << 58) - 1;

enum { MODE_FWD = 0, MODE_TRAIN_BEAR = 1, MODE_TRAIN_AR = 2, MODE_BWD = 3 };

struct CnnDims {
    int lag, W, F, H1, P, PF;
    int PFp, Hp;         // PF, H1 rounded up to the 8-wide tensor-core tiles
    int es, ws;          // row strides (doubles) of E0 and of W1s / Y: = 4 or 12 (mod 16) => conflict-free fragments
    int nfil;            // W * 5 * F
    int nparams;
    int o_fil, o_int0, o_w1, o_int1, o_w2, o_int2, o_sc0, o_sc1;   // offsets in the flat parameter block
};

__host__ __device__ inline int even(int x) { return (x + 1) & ~1; }

inline CnnDims make_dims(int lag, int W, int F, int H1) {
    CnnDims d;
    d.lag = lag; d.W = W; d.F = F; d.H1 = H1;
    d.P = lag - W + 1;
    d.PF = d.P * F;
    d.PFp = (d.PF + 7) & ~7;
    d.Hp = (H1 + 7) & ~7;
    d.es = d.PFp + 4;
    d.ws = d.Hp + 4;
    d.nfil = W * A1 * F;
    // reference order (ar_funcs.py:98-99): filters, int0, W1, int1, W2, int2, scale0, scale1
    int o = 0;
    d.o_fil = o;  o += d.nfil;
    d.o_int0 = o; o += d.PF;
    d.o_w1 = o;   o += d.PF * H1;
    d.o_int1 = o; o += H1;
    d.o_w2 = o;   o += H1 * A1;
    d.o_int2 = o; o += A1;
    d.o_sc0 = o;  o += d.PF;
    d.o_sc1 = o;  o += H1;
    d.nparams = o;
    return d;
}

// shared-memory carve-up (in doubles); every segment has an even length (16-byte alignment)
struct Layout {
    int e0, w1s, y, fil, sc0, in0, small, stir, red, codes, stats, soff, fbuf, tbuf, dfil, dsc0, din0, smallg, total;
};
constexpr int SMALL_N = 16 + 16 + 8;             // int1[16], scale1[16], int2[8]
constexpr int SMALLG_N = 16 * A1 + 8 + 16 + 16;  // dW2[16][5], dint2[8], dint1[16], dscale1[16]

__host__ __device__ inline Layout make_layout(const CnnDims& d, int TB, bool train, int nw) {
    Layout L;
    int o = 0;
    L.e0 = o;    o += TB * d.es;
    L.w1s = o;   o += d.PFp * d.ws;
    L.y = o;     o += TB * d.ws;
    L.fil = o;   o += even(d.nfil);
    L.sc0 = o;   o += even(d.PF);
    L.in0 = o;   o += even(d.PF);
    L.small = o; o += SMALL_N;
    L.stir = o;
    L.red = o;   o += 32;
    L.codes = o; o += TB;
    L.stats = o; o += 2 * TB * d.P;                  // (mean, rstd) of every (row, position)
    L.soff = o;  o += even((TB * d.lag + 3) / 4);        // uint16 [TB][lag]: symbol * F
    L.fbuf = o;  o += train ? TB * A1 : 0;               // f, then d objective / d logits, of the tile's rows
    L.tbuf = o;  o += train ? (A1 + 1) * TB * 2 : 0;     // per row and term: lgamma difference, digamma difference
    L.dfil = o;  o += train ? nw * even(d.nfil) : 0;
    L.dsc0 = o;  o += train ? even(d.PF) : 0;
    L.din0 = o;  o += train ? even(d.PF) : 0;
    L.smallg = o; o += train ? SMALLG_N : 0;
    L.total = o;
    return L;
}

// D[8x8] += A[8x4] . B[4x8] on the FP64 tensor cores.  Fragments (PTX ISA, mma.m8n8k4 .f64): A element
// (row lane/4, col lane%4); B element (row lane%4, col lane/4); C/D elements (row lane/4, cols 2*(lane%4)+{0,1}).
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double half_sum(double v) {      // sum over the 16 lanes of a half-warp
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// conv1d output of one row at position p for filter `lane`: the one-hot input makes it a gather of W filter
// rows (ar_funcs.py:92).  so = the row's per-position offsets symbol * F.
__device__ __forceinline__ double conv_at(const double* __restrict__ fil, const uint16_t* __restrict__ so, int p, int lane,
                                          const CnnDims& d) {
    const double* f = fil + lane;
    const int ws = A1 * d.F;
    so += p;
    if (d.W == 3) return (f[so[0]] + f[ws + so[1]]) + f[2 * ws + so[2]];        // config_files/bear_cnn_bear.cfg
    if (d.W == 8)                                                               // the default (ar_funcs.py:50)
        return ((f[so[0]] + f[ws + so[1]]) + (f[2 * ws + so[2]] + f[3 * ws + so[3]])) +
               ((f[4 * ws + so[4]] + f[5 * ws + so[5]]) + (f[6 * ws + so[6]] + f[7 * ws + so[7]]));
    double c = 0.0;
#pragma unroll 1
    for (int w = 0; w < d.W; ++w) c += f[w * ws + so[w]];
    return c;
}

// sums of a and b over the warp with 6 exchange rounds instead of 10: after the first round the lower
// half-warp carries a, the upper one b.
__device__ __forceinline__ void warp_sum2(double& a, double& b) {
    const bool lower = (threadIdx.x & 16) == 0;
    double v = (lower ? a : b) + __shfl_xor_sync(0xffffffffu, lower ? b : a, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const double other = __shfl_xor_sync(0xffffffffu, v, 16);
    a = lower ? v : other;
    b = lower ? other : v;
}

__device__ __forceinline__ double elu(double x) { return x > 0.0 ? x : exp(x) - 1.0; }   // tf.nn.elu: exp(x) - 1

template <int MODE, int TB, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
cnn_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ col, int64_t stride, int64_t n,
           const CnnDims d, const double* __restrict__ params, const double* __restrict__ h_signed,
           const double* __restrict__ gf_in, double* __restrict__ f_out, double* __restrict__ ll_out,
           double* __restrict__ partials) {
    constexpr bool TRAIN = MODE != MODE_FWD;
    constexpr int THREADS = NW * 32, NWARP = NW, MAX_TILES = max_tiles(NW);
    extern __shared__ __align__(16) double smem[];
    const Layout L = make_layout(d, TB, TRAIN, NW);
    double* E0 = smem + L.e0;
    double* W1s = smem + L.w1s;
    double* Y = smem + L.y;
    double* fil = smem + L.fil;
    double* sc0 = smem + L.sc0;
    double* in0 = smem + L.in0;
    double* small = smem + L.small;
    double* red = smem + L.red;
    uint64_t* codes = reinterpret_cast<uint64_t*>(smem + L.codes);
    double2* stats = reinterpret_cast<double2*>(smem + L.stats);
    uint16_t* soff = reinterpret_cast<uint16_t*>(smem + L.soff);
    double* fbuf = smem + L.fbuf;
    double* tbuf = smem + L.tbuf;
    double* dfil = smem + L.dfil;
    double* dsc0 = smem + L.dsc0;
    double* din0 = smem + L.din0;
    double* smallg = smem + L.smallg;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g4 = lane >> 2, t4 = lane & 3;             // tensor-core fragment coordinates
    const int nfil2 = even(d.nfil);

    // ---------------- parameters -> shared memory / registers ----------------
    for (int i = threadIdx.x; i < d.nfil; i += THREADS) fil[i] = params[d.o_fil + i];
    for (int i = threadIdx.x; i < d.PF; i += THREADS) {
        sc0[i] = params[d.o_sc0 + i];
        in0[i] = params[d.o_int0 + i];
    }
    for (int i = threadIdx.x; i < d.PFp * d.ws; i += THREADS) {
        const int pf = i / d.ws, h = i % d.ws;
        W1s[i] = (pf < d.PF && h < d.H1) ? params[d.o_w1 + pf * d.H1 + h] : 0.0;
    }
    for (int i = threadIdx.x; i < TB * d.es; i += THREADS) E0[i] = 0.0;      // incl. the padding columns
    for (int i = threadIdx.x; i < TB * d.ws; i += THREADS) Y[i] = 0.0;
    if (threadIdx.x < 16) {
        small[threadIdx.x] = threadIdx.x < d.H1 ? params[d.o_int1 + threadIdx.x] : 0.0;
        small[16 + threadIdx.x] = threadIdx.x < d.H1 ? params[d.o_sc1 + threadIdx.x] : 0.0;
        if (threadIdx.x < 8) small[32 + threadIdx.x] = threadIdx.x < A1 ? params[d.o_int2 + threadIdx.x] : 0.0;
    }
    if (TRAIN) {
        for (int i = threadIdx.x; i < NWARP * nfil2; i += THREADS) dfil[i] = 0.0;
        for (int i = threadIdx.x; i < d.PF; i += THREADS) {
            dsc0[i] = 0.0;
            din0[i] = 0.0;
        }
        for (int i = threadIdx.x; i < SMALLG_N; i += THREADS) smallg[i] = 0.0;
    }
    const int hl = lane & 15, half = lane >> 4;
    const bool hok = hl < d.H1;
    double w2[A1];
#pragma unroll
    for (int b = 0; b < A1; ++b) w2[b] = hok ? params[d.o_w2 + hl * A1 + b] : 0.0;
    const double in1_l = hok ? params[d.o_int1 + hl] : 0.0;
    const double sc1_l = hok ? params[d.o_sc1 + hl] : 0.0;
    const double hinv = (MODE == MODE_TRAIN_BEAR) ? exp(-h_signed[0]) : 1.0;
    const double invH = 1.0 / double(d.H1), invF = 1.0 / double(d.F);
    __syncthreads();

    // accumulators that live in registers for the whole kernel
    double accW1[MAX_TILES][2];
#pragma unroll
    for (int i = 0; i < MAX_TILES; ++i) accW1[i][0] = accW1[i][1] = 0.0;
    double aW2[A1] = {0, 0, 0, 0, 0}, aI2[A1] = {0, 0, 0, 0, 0}, aS1 = 0.0, aI1 = 0.0;
    double ll_sum = 0.0, dh_sum = 0.0;

    const int npairs = d.P * TB;
    const int ppw = ((npairs + NWARP - 1) / NWARP + 1) & ~1;     // even: the conv stages take two pairs per trip
    const int pair0 = warp * ppw, pair1 = min(npairs, pair0 + ppw);
    const int nt_h = d.Hp >> 3;                       // 8-wide tiles across the H1 units
    const int nt_pf = d.PFp >> 3;                     // ... across the P*F conv features
    const int ntiles_w1 = nt_pf * nt_h;
    const int hshift = nt_h >> 1;                     // H1 <= 16: one or two tiles across
    double* mydfil = dfil + warp * nfil2;

    const int64_t ntile = (n + TB - 1) / TB;
    for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int64_t row0 = tile * TB;
        if (threadIdx.x < TB) codes[threadIdx.x] = row0 + threadIdx.x < n ? __ldg(kmers + row0 + threadIdx.x) : 0ull;
        __syncthreads();
        // ---------------- 0. per tile: symbol offsets and the layer-norm statistics of every (row, position) ----
        for (int i = threadIdx.x; i < TB * d.lag; i += THREADS) {
            const int r = i / d.lag, j = i - r * d.lag;
            const uint64_t code = codes[r];
            const int ns = int(code >> 58);
            const int sym = j < ns ? 4 : int(((code & PAYLOAD_MASK) >> (2 * (d.lag - 1 - j))) & 3u);
            soff[i] = uint16_t(sym * d.F);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < npairs; i += THREADS) {
            // one thread per (row, position): moments over the F filters (ar_funcs.py:18-19), serially
            const int p = i / TB, r = i % TB;
            const uint16_t* so = soff + r * d.lag;
            double s1 = 0.0, s2 = 0.0;
            for (int f = 0; f < d.F; ++f) {
                const double c = conv_at(fil, so, p, f, d);
                s1 += c;
                s2 = fma(c, c, s2);
            }
            const double mean = s1 * invF;
            const double var = fmax(fma(-mean, mean, s2 * invF), 0.0);
            stats[i] = make_double2(mean, rsqrt(var + 1e-5));
        }
        __syncthreads();

        // ---------------- 1. conv -> layer norm -> elu ----------------
        // two (position, row) pairs per trip: independent dependency chains (pair ranges start and end on even indices)
        for (int idx = pair0; idx < pair1; idx += 2) {
            const int p = idx / TB, r = idx % TB;
            if (lane < d.F) {
                const double2 st0 = stats[idx], st1 = stats[idx + 1];
                const double sc = sc0[p * d.F + lane], in = in0[p * d.F + lane];
                const double x0 = fma(sc, (conv_at(fil, soff + r * d.lag, p, lane, d) - st0.x) * st0.y, in);
                const double x1 = fma(sc, (conv_at(fil, soff + (r + 1) * d.lag, p, lane, d) - st1.x) * st1.y, in);
                const double ex0 = exp(fmin(x0, 0.0)), ex1 = exp(fmin(x1, 0.0));     // tf.nn.elu: exp(x) - 1 for x < 0
                E0[r * d.es + p * d.F + lane] = x0 > 0.0 ? x0 : ex0 - 1.0;
                E0[(r + 1) * d.es + p * d.F + lane] = x1 > 0.0 ? x1 : ex1 - 1.0;
            }
        }
        __syncthreads();

        // ---------------- 2. Y = E0 . W1 (tensor cores) ----------------
        for (int t = warp; t < (TB >> 3) * nt_h; t += NWARP) {
            const int m = t / nt_h, nn = t % nt_h;
            const double* a_ptr = E0 + (m * 8 + g4) * d.es + t4;
            const double* b_ptr = W1s + t4 * d.ws + nn * 8 + g4;
            double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
            int k0 = 0;
            for (; k0 + 16 <= d.PFp; k0 += 16) {
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(c[j][0], c[j][1], a_ptr[k0 + 4 * j], b_ptr[(k0 + 4 * j) * d.ws]);
            }
            if (k0 < d.PFp) {                               // PFp is a multiple of 8: one 8-wide remainder at most
                dmma(c[0][0], c[0][1], a_ptr[k0], b_ptr[k0 * d.ws]);
                dmma(c[1][0], c[1][1], a_ptr[k0 + 4], b_ptr[(k0 + 4) * d.ws]);
            }
            double* y = Y + (m * 8 + g4) * d.ws + nn * 8 + 2 * t4;
            y[0] = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
            y[1] = (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
        }
        __syncthreads();

        // ---------------- 3a. per row (half-warp, lane = H1 unit): LN, elu, dense 2, softmax ----------------
        constexpr int RPW = TB / NW;                         // rows per warp in the per-row stages (1, 2 or 4)
        constexpr int NIT = RPW >= 2 ? RPW / 2 : 1;
        double k_yhat[NIT], k_rstd[NIT], k_x1[NIT], k_e1[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const bool act = it * 2 + half < RPW;             // RPW = 1: the upper half-warp has no row
            const int r = act ? warp * RPW + it * 2 + half : 0;
            const int64_t i = row0 + r;
            const double yv = hok ? Y[r * d.ws + hl] : 0.0;
            const double mean = half_sum(yv) * invH;
            const double dl = hok ? yv - mean : 0.0;
            const double var = half_sum(dl * dl) * invH;
            const double rstd = rsqrt(var + 1e-5);
            const double yhat = dl * rstd;
            const double x1 = fma(sc1_l, yhat, in1_l);
            const double e1 = elu(x1);
            k_yhat[it] = yhat; k_rstd[it] = rstd; k_x1[it] = x1; k_e1[it] = e1;
            double f[A1], x2[A1], mx = -INFINITY, z = 0.0;
#pragma unroll
            for (int b = 0; b < A1; ++b) {
                x2[b] = half_sum(e1 * w2[b]) + small[32 + b];
                mx = fmax(mx, x2[b]);
            }
#pragma unroll
            for (int b = 0; b < A1; ++b) {
                f[b] = exp(x2[b] - mx);
                z += f[b];
            }
            const double zi = 1.0 / z;
            if (hl < A1 && act) {
                const double v = (hl == 0 ? f[0] : hl == 1 ? f[1] : hl == 2 ? f[2] : hl == 3 ? f[3] : f[4]) * zi;
                if (MODE == MODE_FWD) {
                    if (i < n) f_out[i * A1 + hl] = v;
                } else {
                    fbuf[r * A1 + hl] = v;
                }
            }
        }
        __syncthreads();
        if (MODE == MODE_FWD) continue;

        // ---------------- 3b. loss and its gradient w.r.t. the logits ----------------
        // BEAR: the six lgamma / digamma differences of a row (five letters + the total) are independent, so six
        // warps evaluate one term each for the tile's rows (lane = row) instead of one warp walking all six.
        if (MODE == MODE_TRAIN_BEAR && warp < A1 + 1) {
            const int r = lane < TB ? lane : TB - 1;
            const int64_t i = row0 + r;
            const bool in_range = lane < TB && i < n;
            double a, c;
            if (warp < A1) {
                a = fma(fbuf[r * A1 + warp], hinv, BEAR_EPS);                                   // bear_net.py:43
                c = in_range ? double(__ldg(col + warp * stride + i)) : 0.0;
            } else {
                a = 0.0;
                c = 0.0;
#pragma unroll
                for (int b2 = 0; b2 < A1; ++b2) {
                    a += fma(fbuf[r * A1 + b2], hinv, BEAR_EPS);
                    c += in_range ? double(__ldg(col + b2 * stride + i)) : 0.0;
                }
            }
            const LgDg t = lgdg_diff<true>(a, c);
            if (lane < TB) {
                tbuf[(warp * TB + r) * 2] = t.add + (t.mul == 1.0 ? 0.0 : log(t.mul));
                tbuf[(warp * TB + r) * 2 + 1] = t.dg;
            }
        }
        if (MODE == MODE_TRAIN_BEAR) __syncthreads();
        if (warp == 0) {
            const int r = lane < TB ? lane : TB - 1;
            const int64_t i = row0 + r;
            const bool in_range = lane < TB && i < n;
            double f[A1], df[A1] = {0, 0, 0, 0, 0};          // df = d (objective) / d f
#pragma unroll
            for (int b = 0; b < A1; ++b) f[b] = fbuf[r * A1 + b];
            if (MODE == MODE_BWD) {
#pragma unroll
                for (int b = 0; b < A1; ++b) df[b] = in_range ? __ldg(gf_in + i * A1 + b) : 0.0;
            } else if (MODE == MODE_TRAIN_AR) {
                const Counts cr = load_counts(col, stride, i, in_range);
                double add, prod, pr[A1];
#pragma unroll
                for (int b = 0; b < A1; ++b) pr[b] = f[b] + BEAR_EPS;                      // bear_net.py:68
                mn_term(pr, cr, add, prod);
#pragma unroll
                for (int b = 0; b < A1; ++b) df[b] = cr.c[b] == 0 ? 0.0 : double(cr.c[b]) / pr[b];
                const double ll = cr.cmax != 0 ? add + log(prod) : 0.0;
                ll_sum += ll;
                if (ll_out && in_range) ll_out[i] = ll;
            } else {
                // ll = sum_b [lgamma(conc_b + c_b) - lgamma(conc_b)] - [lgamma(S + N) - lgamma(S)]; zero-count rows
                // give exact zeros in every term
                const double on = lane < TB ? 1.0 : 0.0;      // 16-row tiles: the upper half-warp has no row
                const double tdg = tbuf[(A1 * TB + r) * 2 + 1];
                double ll = -tbuf[(A1 * TB + r) * 2];
#pragma unroll
                for (int b = 0; b < A1; ++b) {
                    ll += tbuf[(b * TB + r) * 2];
                    df[b] = on * (tbuf[(b * TB + r) * 2 + 1] - tdg) * hinv;
                    dh_sum -= f[b] * df[b];
                }
                ll *= on;
                ll_sum += ll;
                if (ll_out && in_range) ll_out[i] = ll;
            }
            double u = 0.0;                                   // softmax backward
#pragma unroll
            for (int b = 0; b < A1; ++b) u = fma(f[b], df[b], u);
            if (lane < TB) {
#pragma unroll
                for (int b = 0; b < A1; ++b) fbuf[r * A1 + b] = f[b] * (df[b] - u);
            }
        }
        __syncthreads();

        // ---------------- 3c. dense 2 backward, elu', layer-norm backward -> dY over Y ----------------
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const bool act = it * 2 + half < RPW;
            const int r = act ? warp * RPW + it * 2 + half : 0;
            const double yhat = k_yhat[it], e1 = k_e1[it];
            double de1 = 0.0;
#pragma unroll
            for (int b = 0; b < A1; ++b) {
                const double g = act ? fbuf[r * A1 + b] : 0.0;
                de1 = fma(w2[b], g, de1);
                aW2[b] = fma(e1, g, aW2[b]);
                aI2[b] += g;
            }
            const double dx1 = de1 * (k_x1[it] > 0.0 ? 1.0 : e1 + 1.0);
            aS1 = fma(dx1, yhat, aS1);
            aI1 += dx1;
            const double dyh = dx1 * sc1_l;
            const double m1 = half_sum(dyh) * invH;
            const double m2 = half_sum(dyh * yhat) * invH;
            if (hok && act) Y[r * d.ws + hl] = k_rstd[it] * (dyh - m1 - yhat * m2);
        }
        __syncthreads();

        // ---------------- 4. dW1 += E0^T . dY (tensor cores, register accumulators) ----------------
        for (int k0 = 0; k0 < TB; k0 += 4) {
            const double* a_row = E0 + (k0 + t4) * d.es + g4;
            const double* b_row = Y + (k0 + t4) * d.ws + g4;
#pragma unroll
            for (int j = 0; j < MAX_TILES; ++j) {
                const int t = warp + NWARP * j;
                if (t < ntiles_w1) {
                    const int m = t >> hshift, nn = t & hshift;        // nt_h is 1 or 2
                    dmma(accW1[j][0], accW1[j][1], a_row[m * 8], b_row[nn * 8]);
                }
            }
        }
        __syncthreads();

        // ---------------- 5. dX0 = (dY . W1^T) * elu'(x0), over E0 (tensor cores) ----------------
        for (int t = warp; t < (TB >> 3) * nt_pf; t += NWARP) {
            const int m = t % (TB >> 3), nn = t / (TB >> 3);
            const double* a_ptr = Y + (m * 8 + g4) * d.ws + t4;
            const double* b_ptr = W1s + (nn * 8 + g4) * d.ws + t4;
            double c[2][2] = {{0, 0}, {0, 0}};
            for (int k0 = 0; k0 < d.Hp; k0 += 8) {
                dmma(c[0][0], c[0][1], a_ptr[k0], b_ptr[k0]);
                dmma(c[1][0], c[1][1], a_ptr[k0 + 4], b_ptr[k0 + 4]);
            }
            double2* e = reinterpret_cast<double2*>(E0 + (m * 8 + g4) * d.es + nn * 8 + 2 * t4);
            double2 v = *e;
            v.x = (c[0][0] + c[1][0]) * (v.x > 0.0 ? 1.0 : v.x + 1.0);       // elu' = 1 or elu + 1
            v.y = (c[0][1] + c[1][1]) * (v.y > 0.0 ? 1.0 : v.y + 1.0);
            *e = v;
        }
        __syncthreads();

        // ---------------- 6. layer norm 0 / conv backward ----------------
        {
            double aS0 = 0.0, aI0 = 0.0;
            int curp = -1;
            for (int idx = pair0; idx < pair1; idx += 2) {
                const int p = idx / TB, r = idx % TB;            // rows r and r + 1 at position p
                if (p != curp) {
                    if (curp >= 0 && lane < d.F) {
                        atomicAdd(dsc0 + curp * d.F + lane, aS0);
                        atomicAdd(din0 + curp * d.F + lane, aI0);
                    }
                    curp = p;
                    aS0 = aI0 = 0.0;
                }
                const uint16_t* soA = soff + r * d.lag;
                const uint16_t* soB = soA + d.lag;
                const bool fok = lane < d.F;
                const double2 stA = stats[idx], stB = stats[idx + 1];
                const double sc = fok ? sc0[p * d.F + lane] : 0.0;
                const double xhA = fok ? (conv_at(fil, soA, p, lane, d) - stA.x) * stA.y : 0.0;
                const double xhB = fok ? (conv_at(fil, soB, p, lane, d) - stB.x) * stB.y : 0.0;
                const double dxA = fok ? E0[r * d.es + p * d.F + lane] : 0.0;
                const double dxB = fok ? E0[(r + 1) * d.es + p * d.F + lane] : 0.0;
                aS0 = fma(dxA, xhA, fma(dxB, xhB, aS0));
                aI0 += dxA + dxB;
                const double dhA = dxA * sc, dhB = dxB * sc;
                double m1A = dhA, m2A = dhA * xhA, m1B = dhB, m2B = dhB * xhB;
                warp_sum2(m1A, m2A);
                warp_sum2(m1B, m2B);
                const double dcA = stA.y * (dhA - m1A * invF - xhA * (m2A * invF));
                const double dcB = stB.y * (dhB - m1B * invF - xhB * (m2B * invF));
                if (fok) {
                    double* g = mydfil + lane;
                    const int ws = A1 * d.F;
                    if (d.W == 3) {
                        g[soA[p]] += dcA;
                        g[ws + soA[p + 1]] += dcA;
                        g[2 * ws + soA[p + 2]] += dcA;
                        g[soB[p]] += dcB;
                        g[ws + soB[p + 1]] += dcB;
                        g[2 * ws + soB[p + 2]] += dcB;
                    } else {
#pragma unroll 1
                        for (int w = 0; w < d.W; ++w) {
                            g[w * ws + soA[p + w]] += dcA;
                            g[w * ws + soB[p + w]] += dcB;
                        }
                    }
                }
            }
            if (curp >= 0 && lane < d.F) {
                atomicAdd(dsc0 + curp * d.F + lane, aS0);
                atomicAdd(din0 + curp * d.F + lane, aI0);
            }
        }
        __syncthreads();
    }

    if (!TRAIN) return;

    // ---------------- per-CTA partial sums: [ll, d/dh, d params...] ----------------
    if (hok) {
#pragma unroll
        for (int b = 0; b < A1; ++b) atomicAdd(smallg + hl * A1 + b, aW2[b]);
        atomicAdd(smallg + 16 * A1 + 8 + hl, aI1);
        atomicAdd(smallg + 16 * A1 + 8 + 16 + hl, aS1);
        if (hl == 0) {
#pragma unroll
            for (int b = 0; b < A1; ++b) atomicAdd(smallg + 16 * A1 + b, aI2[b]);
        }
    }
    double* out = partials + int64_t(blockIdx.x) * (2 + d.nparams);
    const double ll_blk = block_sum(ll_sum, red);
    const double dh_blk = block_sum(dh_sum, red);
    if (threadIdx.x == 0) {
        out[0] = ll_blk;
        out[1] = dh_blk;
    }
    __syncthreads();
    double* po = out + 2;
    for (int i = threadIdx.x; i < d.nfil; i += THREADS) {
        double s = 0.0;
        for (int w = 0; w < NWARP; ++w) s += dfil[w * nfil2 + i];
        po[d.o_fil + i] = s;
    }
    for (int i = threadIdx.x; i < d.PF; i += THREADS) {
        po[d.o_int0 + i] = din0[i];
        po[d.o_sc0 + i] = dsc0[i];
    }
#pragma unroll
    for (int j = 0; j < MAX_TILES; ++j) {
        const int t = warp + NWARP * j;
        if (t < ntiles_w1) {
            const int pf = (t / nt_h) * 8 + g4, h = (t % nt_h) * 8 + 2 * t4;
            if (pf < d.PF) {
                if (h < d.H1) po[d.o_w1 + pf * d.H1 + h] = accW1[j][0];
                if (h + 1 < d.H1) po[d.o_w1 + pf * d.H1 + h + 1] = accW1[j][1];
            }
        }
    }
    if (threadIdx.x < d.H1) {
        const int h = threadIdx.x;
#pragma unroll
        for (int b = 0; b < A1; ++b) po[d.o_w2 + h * A1 + b] = smallg[h * A1 + b];
        po[d.o_int1 + h] = smallg[16 * A1 + 8 + h];
        po[d.o_sc1 + h] = smallg[16 * A1 + 8 + 16 + h];
    }
    if (threadIdx.x < A1) po[d.o_int2 + threadIdx.x] = smallg[16 * A1 + threadIdx.x];
}

// fixed-order second stage: out[p] += mult * sum_blk partials[blk, p]
__global__ void cnn_reduce_kernel(const double* __restrict__ partials, int nblk, int P, int p_begin, double mult,
                                  double* __restrict__ out) {
    const int p = p_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partials[int64_t(b) * P + p];
    out[p - p_begin] += mult * s;
}

// rows per tile and warps per CTA for these dimensions: the first of (32 rows, 16 warps), (32, 8), (16, 16), (16, 8)
// whose shared-memory carve-up fits; {0, 0} = outside the fused kernel
struct TileCfg {
    int tb, nw;
};
TileCfg pick_cfg(const CnnDims& d, bool train) {
    if (d.W < 1 || d.P < 1 || d.F < 1 || d.F > 32 || d.H1 < 1 || d.H1 > 16 || d.lag > 29) return {0, 0};
    const int ntiles = (d.PFp >> 3) * (d.Hp >> 3);
    for (int tb : {32, 16})
        for (int nw : {16, 8}) {
            if ((ntiles + nw - 1) / nw > max_tiles(nw)) continue;
            if (size_t(make_layout(d, tb, train, nw).total) * 8 <= size_t(MAX_SMEM)) return {tb, nw};
        }
    return {0, 0};
}

template <int MODE, int TB, int NW>
int launch(cudaStream_t st, const uint64_t* kmers, const uint32_t* col, int64_t stride, int64_t n, const CnnDims& d,
           const double* params, const double* h_signed, const double* gf, double* f_out, double* ll_out,
           double* partials, int* grid_out) {
    const size_t smem = size_t(make_layout(d, TB, MODE != MODE_FWD, NW).total) * 8;
    cudaError_t e = cudaFuncSetAttribute(cnn_kernel<MODE, TB, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) {
        bear_set_error("cudaFuncSetAttribute(cnn_kernel, %zu bytes) failed: %s", smem, cudaGetErrorString(e));
        return BEAR_ERR_CUDA;
    }
    const int64_t ntile = (n + TB - 1) / TB;
    const int grid = int(ntile < 148 ? ntile : 148);
    cnn_kernel<MODE, TB, NW><<<grid, NW * 32, smem, st>>>(kmers, col, stride, n, d, params, h_signed, gf, f_out, ll_out, partials);
    BEAR_LAUNCH_CHECK("cnn_kernel");
    *grid_out = grid;
    return BEAR_OK;
}

template <int MODE>
int launch_cfg(TileCfg c, cudaStream_t st, const uint64_t* kmers, const uint32_t* col, int64_t stride, int64_t n,
               const CnnDims& d, const double* params, const double* h_signed, const double* gf, double* f_out,
               double* ll_out, double* partials, int* grid_out) {
#define BEAR_CNN_GO(TB_, NW_) \
    return launch<MODE, TB_, NW_>(st, kmers, col, stride, n, d, params, h_signed, gf, f_out, ll_out, partials, grid_out)
    if (c.tb == 32 && c.nw == 16) BEAR_CNN_GO(32, 16);
    if (c.tb == 32) BEAR_CNN_GO(32, 8);
    if (c.nw == 16) BEAR_CNN_GO(16, 16);
    BEAR_CNN_GO(16, 8);
#undef BEAR_CNN_GO
}

}  // namespace

extern "C" int bear_cnn_supported(int lag, int filter_width, int num_filters, int layer1_width) {
    if (lag < 1 || filter_width < 1 || filter_width > lag) return 0;
    return pick_cfg(make_dims(lag, filter_width, num_filters, layer1_width), true).tb != 0;
}

extern "C" int64_t bear_cnn_num_params(int lag, int filter_width, int num_filters, int layer1_width) {
    if (lag < 1 || filter_width < 1 || filter_width > lag || num_filters < 1 || layer1_width < 1) return -1;
    return make_dims(lag, filter_width, num_filters, layer1_width).nparams;
}

extern "C" int bear_cnn_head_forward(const uint64_t* d_kmers, int64_t row0, int64_t n, int lag, int filter_width,
                                     int num_filters, int layer1_width, const double* d_params, double* d_f,
                                     void* stream) {
    const char* fn = "bear_cnn_head_forward";
    BEAR_REQUIRE(n >= 0 && row0 >= 0, fn);
    BEAR_REQUIRE(lag >= 1 && lag <= 29 && filter_width >= 1 && filter_width <= lag, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_kmers && d_params && d_f, fn);
    const CnnDims d = make_dims(lag, filter_width, num_filters, layer1_width);
    const TileCfg tb = pick_cfg(d, false);
    if (!tb.tb) {
        bear_set_error("%s: dimensions outside the fused kernel (F <= 32, H1 <= 16, shared memory)", fn);
        return BEAR_ERR_RANGE;
    }
    int grid;
    return launch_cfg<MODE_FWD>(tb, static_cast<cudaStream_t>(stream), d_kmers + row0, nullptr, 0, n, d, d_params, nullptr,
                               nullptr, d_f, nullptr, nullptr, &grid);
}

extern "C" int bear_cnn_train_step(const uint64_t* d_kmers, const uint32_t* d_col, int64_t stride, int64_t row0,
                                   int64_t n, int lag, int filter_width, int num_filters, int layer1_width,
                                   const double* d_params, const double* d_h_signed, double scale, int train_ar,
                                   double* d_flat, double* d_ll_out, double* d_workspace, void* stream) {
    const char* fn = "bear_cnn_train_step";
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    BEAR_REQUIRE(lag >= 1 && lag <= 29 && filter_width >= 1 && filter_width <= lag, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_kmers && d_col && d_params && d_h_signed && d_flat && d_workspace, fn);
    const CnnDims d = make_dims(lag, filter_width, num_filters, layer1_width);
    const TileCfg tb = pick_cfg(d, true);
    if (!tb.tb) {
        bear_set_error("%s: dimensions outside the fused kernel (F <= 32, H1 <= 16, shared memory)", fn);
        return BEAR_ERR_RANGE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid, rc;
    if (train_ar)
        rc = launch_cfg<MODE_TRAIN_AR>(tb, st, d_kmers + row0, d_col + row0, stride, n, d, d_params, d_h_signed, nullptr,
                                      nullptr, d_ll_out, d_workspace, &grid);
    else
        rc = launch_cfg<MODE_TRAIN_BEAR>(tb, st, d_kmers + row0, d_col + row0, stride, n, d, d_params, d_h_signed, nullptr,
                                        nullptr, d_ll_out, d_workspace, &grid);
    if (rc) return rc;
    const int P = 2 + d.nparams;
    cnn_reduce_kernel<<<(P + 127) / 128, 128, 0, st>>>(d_workspace, grid, P, 0, -scale, d_flat);
    BEAR_LAUNCH_CHECK("cnn_reduce_kernel");
    return BEAR_OK;
}

extern "C" int bear_cnn_head_backward(const uint64_t* d_kmers, int64_t row0, int64_t n, int lag, int filter_width,
                                      int num_filters, int layer1_width, const double* d_params, const double* d_gf,
                                      double* d_gparams, double* d_workspace, void* stream) {
    const char* fn = "bear_cnn_head_backward";
    BEAR_REQUIRE(n >= 0 && row0 >= 0, fn);
    BEAR_REQUIRE(lag >= 1 && lag <= 29 && filter_width >= 1 && filter_width <= lag, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_kmers && d_params && d_gf && d_gparams && d_workspace, fn);
    const CnnDims d = make_dims(lag, filter_width, num_filters, layer1_width);
    const TileCfg tb = pick_cfg(d, true);
    if (!tb.tb) {
        bear_set_error("%s: dimensions outside the fused kernel (F <= 32, H1 <= 16, shared memory)", fn);
        return BEAR_ERR_RANGE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int grid;
    const int rc = launch_cfg<MODE_BWD>(tb, st, d_kmers + row0, nullptr, 0, n, d, d_params, nullptr, d_gf, nullptr, nullptr,
                                       d_workspace, &grid);
    if (rc) return rc;
    const int P = 2 + d.nparams;
    cnn_reduce_kernel<<<(d.nparams + 127) / 128, 128, 0, st>>>(d_workspace, grid, P, 2, 1.0, d_gparams);
    BEAR_LAUNCH_CHECK("cnn_reduce_kernel");
    return BEAR_OK;
}
