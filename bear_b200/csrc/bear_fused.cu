// Fused packed-path kernels (DNA/RNA, A1 = 5): the count-streaming hot path.
//   linear_train_kernel   bear_net._train_step (bear_net.py:146-197) + ar_funcs.make_ar_func_linear
//                         (ar_funcs.py:23-46) + core.*.counts_log_prob (core.py:73-74,138-139), fwd+bwd
//   explicit_train_kernel the same loss for a caller-evaluated head f (CNN / bear_ref / plugins)
//   eval_kernel           bear_net._evaluation_step (bear_net.py:323-371), h_scan (bear_net.py:516-531)
//   bmm_kernel            dataloader._marginal_step (dataloader.py:111-113)
// Every kernel streams the packed table once (8 B k-mer + 20 B per count column per row), keeps all
// per-row temporaries in registers and reduces in two deterministic stages: per-CTA partials in the
// caller's workspace, then a fixed-order sum.
#include <math.h>

#include "bear_b200.h"
#include "bear_common.cuh"
#include "bear_host.h"

namespace {

using namespace bear;

constexpr int A1 = 5;            // DNA/RNA letters + stop
constexpr int CHUNK = 4;         // positions per chunk table
constexpr int COMBOS = 256;      // 4^CHUNK
constexpr int THREADS = 256;
constexpr int MAX_GRID = 148 * 4;

__host__ __device__ inline int num_chunks(int lag) { return (lag + CHUNK - 1) / CHUNK; }

// ------------------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------------------
struct Row {
    double c[A1];
    double n;
};

__device__ __forceinline__ Row load_row(const uint32_t* __restrict__ col, int64_t stride, int64_t i) {
    Row r;
    uint32_t raw[A1];
#pragma unroll
    for (int b = 0; b < A1; ++b) raw[b] = __ldg(col + b * stride + i);
    r.n = 0.0;
#pragma unroll
    for (int b = 0; b < A1; ++b) {
        r.c[b] = double(raw[b]);
        r.n += r.c[b];
    }
    return r;
}

// Chunk tables of the linear head.  For chunk ch (positions 4ch..4ch+3) and symbol combination q,
// R[ch][q][b] = exp(l_b - l_4), b < 4, with l = sum of the chunk's rows of `mat`.  The softmax of a
// start-free k-mer is then prod_ch R[ch][q_ch][b] / (1 + sum_b prod_ch R[ch][q_ch][b]).
__device__ void build_ratio_tables(const double* smat, double* R, int lag) {
    const int nch = num_chunks(lag);
    for (int idx = threadIdx.x; idx < nch * COMBOS; idx += blockDim.x) {
        const int ch = idx >> 8, q = idx & 255;
        const int r = (ch == nch - 1) ? lag - CHUNK * ch : CHUNK;
        double l[A1] = {0, 0, 0, 0, 0};
        if (q < (1 << (2 * r))) {
            for (int p = 0; p < r; ++p) {
                const int s = (q >> (2 * (r - 1 - p))) & 3;
                const double* row = smat + ((CHUNK * ch + p) * A1 + s) * A1;
#pragma unroll
                for (int b = 0; b < A1; ++b) l[b] += row[b];
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) R[idx * 4 + b] = exp(l[b] - l[4]);
    }
}

__device__ __forceinline__ int chunk_key(uint64_t v, int ch, int nch, int lag) {
    if (ch == nch - 1) return int(v & ((1u << (2 * (lag - CHUNK * ch))) - 1u));
    return int((v >> (2 * (lag - CHUNK * ch - CHUNK))) & 255u);
}

__device__ __forceinline__ int symbol_at(uint64_t v, int j, int lag, int nstart) {
    return j < nstart ? 4 : int((v >> (2 * (lag - 1 - j))) & 3u);
}

// softmax(sum_j mat[j, s_j, :]); returns false if the fast (chunk-table) path could not be used
__device__ __forceinline__ bool linear_head_fast(const double* R, uint64_t v, int lag, int nch, double (&f)[A1]) {
    double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
    for (int ch = 0; ch < nch; ++ch) {
        const int q = chunk_key(v, ch, nch, lag);
        const double2 a = *reinterpret_cast<const double2*>(R + (ch * COMBOS + q) * 4);
        const double2 b = *reinterpret_cast<const double2*>(R + (ch * COMBOS + q) * 4 + 2);
        p0 *= a.x;
        p1 *= a.y;
        p2 *= b.x;
        p3 *= b.y;
    }
    const double z = 1.0 + ((p0 + p1) + (p2 + p3));
    if (!(z < 1e300)) return false;
    const double zi = 1.0 / z;
    f[0] = p0 * zi;
    f[1] = p1 * zi;
    f[2] = p2 * zi;
    f[3] = p3 * zi;
    f[4] = zi;
    return true;
}

__device__ __forceinline__ void linear_head_slow(const double* smat, uint64_t v, int lag, int nstart, double (&f)[A1]) {
    double l[A1] = {0, 0, 0, 0, 0};
    for (int j = 0; j < lag; ++j) {
        const double* row = smat + (j * A1 + symbol_at(v, j, lag, nstart)) * A1;
#pragma unroll
        for (int b = 0; b < A1; ++b) l[b] += row[b];
    }
    double m = l[0];
#pragma unroll
    for (int b = 1; b < A1; ++b) m = fmax(m, l[b]);
    double z = 0.0;
#pragma unroll
    for (int b = 0; b < A1; ++b) {
        f[b] = exp(l[b] - m);
        z += f[b];
    }
    const double zi = 1.0 / z;
#pragma unroll
    for (int b = 0; b < A1; ++b) f[b] *= zi;
}

// Dirichlet-multinomial log-likelihood of one row and d ll / d conc  (core.py:73-74 via TFP's lbeta
// difference; the add-then-subtract log_combinations term is omitted, it cancels analytically).
template <bool GRAD>
__device__ __forceinline__ double dm_row(const double (&conc)[A1], const Row& r, double (&dconc)[A1]) {
    LogProd num, den;
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < A1; ++b) s += conc[b];
    const LgDg t = lgdg_diff<GRAD>(s, r.n);
    den.push(t);
#pragma unroll
    for (int b = 0; b < A1; ++b) {
        const LgDg x = lgdg_diff<GRAD>(conc[b], r.c[b]);
        num.push(x);
        if (GRAD) dconc[b] = x.dg - t.dg;
    }
    return logprod_diff(num, den);
}

// Multinomial log-likelihood sum_b c_b log p_b with multiply_no_nan semantics (core.py:138-139)
__device__ __forceinline__ double mn_row(const double (&p)[A1], const Row& r) {
    double ll = 0.0;
#pragma unroll
    for (int b = 0; b < A1; ++b)
        if (r.c[b] != 0.0) ll = fma(r.c[b], log(p[b]), ll);
    return ll;
}

// Fixed-order second stage: out[p] += mult * sum_blk partials[blk, p]
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblk, int P, double mult,
                                       double* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partials[int64_t(b) * P + p];
    out[p] += mult * s;
}

// ------------------------------------------------------------------------------------------------
// linear head, fused forward + backward
// ------------------------------------------------------------------------------------------------
template <bool TRAIN_AR>
__global__ void __launch_bounds__(THREADS, 2)
linear_train_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ col, int64_t stride,
                    int64_t n, int lag, const double* __restrict__ mat, const double* __restrict__ h_signed,
                    double* __restrict__ ll_out, double* __restrict__ partials) {
    extern __shared__ __align__(16) double smem[];
    const int nch = num_chunks(lag);
    double* R = smem;                          // [nch][256][4] forward ratio tables
    double* G = R + nch * COMBOS * 4;          // [nch][256][4] d ll / d chunk-logits (letters 0..3)
    double* smat = G + nch * COMBOS * 4;       // [lag][5][5]
    double* gmat = smat + lag * A1 * A1;       // [lag][5][5]  gradient from slow-path rows
    double* red = gmat + lag * A1 * A1;        // [32]

    for (int i = threadIdx.x; i < lag * A1 * A1; i += blockDim.x) {
        smat[i] = mat[i];
        gmat[i] = 0.0;
    }
    for (int i = threadIdx.x; i < nch * COMBOS * 4; i += blockDim.x) G[i] = 0.0;
    __syncthreads();
    build_ratio_tables(smat, R, lag);
    __syncthreads();

    const double hinv = exp(-h_signed[0]);     // 1 / h,  h = exp(h_signed)  (bear_net.py:186)
    double ll_sum = 0.0, dh_sum = 0.0;

    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const uint64_t code = __ldg(kmers + i);
        const Row r = load_row(col, stride, i);
        if (r.n == 0.0) {                      // zero-count row: ll = 0 and every gradient is 0
            if (ll_out) ll_out[i] = 0.0;
            continue;
        }
        const int nstart = int(code >> 58);
        const uint64_t v = code & ((1ull << 58) - 1);
        double f[A1];
        bool fast = nstart == 0 && linear_head_fast(R, v, lag, nch, f);
        if (!fast) linear_head_slow(smat, v, lag, nstart, f);

        double ll, df[A1];                     // df = d ll / d f
        if (TRAIN_AR) {
            double p[A1];
#pragma unroll
            for (int b = 0; b < A1; ++b) p[b] = f[b] + BEAR_EPS;           // bear_net.py:68
            ll = mn_row(p, r);
#pragma unroll
            for (int b = 0; b < A1; ++b) df[b] = r.c[b] == 0.0 ? 0.0 : r.c[b] / p[b];
        } else {
            double conc[A1], dconc[A1];
#pragma unroll
            for (int b = 0; b < A1; ++b) conc[b] = fma(f[b], hinv, BEAR_EPS);   // bear_net.py:43
            ll = dm_row<true>(conc, r, dconc);
#pragma unroll
            for (int b = 0; b < A1; ++b) df[b] = dconc[b] * hinv;
        }
        double u = 0.0;
#pragma unroll
        for (int b = 0; b < A1; ++b) u = fma(f[b], df[b], u);
        ll_sum += ll;
        if (!TRAIN_AR) dh_sum -= u;            // d ll / d h_signed = -sum_b f_b d ll/d f_b
        if (ll_out) ll_out[i] = ll;

        double g[A1];                          // softmax backward: d ll / d logits
#pragma unroll
        for (int b = 0; b < A1; ++b) g[b] = f[b] * (df[b] - u);
        if (fast) {
            for (int ch = 0; ch < nch; ++ch) {
                double* dst = G + (ch * COMBOS + chunk_key(v, ch, nch, lag)) * 4;
#pragma unroll
                for (int b = 0; b < 4; ++b) atomicAdd(dst + b, g[b]);
            }
        } else {
            for (int j = 0; j < lag; ++j) {
                double* dst = gmat + (j * A1 + symbol_at(v, j, lag, nstart)) * A1;
#pragma unroll
                for (int b = 0; b < A1; ++b) atomicAdd(dst + b, g[b]);
            }
        }
    }

    const int P = 2 + lag * A1 * A1;
    double* out = partials + int64_t(blockIdx.x) * P;
    const double ll_blk = block_sum(ll_sum, red);
    const double dh_blk = block_sum(dh_sum, red);
    if (threadIdx.x == 0) {
        out[0] = ll_blk;
        out[1] = dh_blk;
    }
    __syncthreads();
    // marginalise the chunk-table gradients back onto mat[j, s, b]
    for (int idx = threadIdx.x; idx < lag * A1 * A1; idx += blockDim.x) {
        const int b = idx % A1, s = (idx / A1) % A1, j = idx / (A1 * A1);
        double val = gmat[idx];
        if (s < 4) {
            const int ch = j / CHUNK, p = j % CHUNK;
            const int r = (ch == nch - 1) ? lag - CHUNK * ch : CHUNK;
            const int shift = 2 * (r - 1 - p);
            double acc = 0.0;
            for (int q = 0; q < (1 << (2 * r)); ++q) {
                if (((q >> shift) & 3) != s) continue;
                const double* src = G + (ch * COMBOS + q) * 4;
                if (b < 4) acc += src[b];
                else acc -= (src[0] + src[1]) + (src[2] + src[3]);   // the 5 logit gradients sum to 0
            }
            val += acc;
        }
        out[2 + idx] = val;
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
    const double hinv = exp(-h_signed[0]);
    double ll_sum = 0.0, dh_sum = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const Row r = load_row(col, stride, i);
        double f[A1], df[A1], ll = 0.0;
#pragma unroll
        for (int b = 0; b < A1; ++b) {
            f[b] = f_in[i * A1 + b];
            df[b] = 0.0;
        }
        if (r.n != 0.0) {
            if (TRAIN_AR) {
                double p[A1];
#pragma unroll
                for (int b = 0; b < A1; ++b) p[b] = f[b] + BEAR_EPS;
                ll = mn_row(p, r);
#pragma unroll
                for (int b = 0; b < A1; ++b) df[b] = r.c[b] == 0.0 ? 0.0 : r.c[b] / p[b];
            } else {
                double conc[A1], dconc[A1];
#pragma unroll
                for (int b = 0; b < A1; ++b) conc[b] = fma(f[b], hinv, BEAR_EPS);
                ll = dm_row<true>(conc, r, dconc);
#pragma unroll
                for (int b = 0; b < A1; ++b) {
                    df[b] = dconc[b] * hinv;
                    dh_sum -= f[b] * df[b];
                }
            }
        }
        ll_sum += ll;
        if (ll_out) ll_out[i] = ll;
        if (gf) {
#pragma unroll
            for (int b = 0; b < A1; ++b) gf[i * A1 + b] = -scale * df[b];
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
// evaluation
// ------------------------------------------------------------------------------------------------
struct EvalParams {
    double h[BEAR_MAX_MODELS];
    double van[BEAR_MAX_MODELS];
    int H, V;
};

template <int HEAD>
__global__ void __launch_bounds__(THREADS, 1)
eval_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ test_col,
            const uint32_t* __restrict__ train_col, int64_t stride, int64_t row0, int64_t n, int lag,
            const double* __restrict__ head, const double* __restrict__ d_h, int H,
            const double* __restrict__ d_van, int V, int64_t seed, double* __restrict__ partials) {
    extern __shared__ __align__(16) double smem[];
    const int nch = num_chunks(lag);
    double* R = smem;
    double* smat = R + (HEAD == BEAR_HEAD_LINEAR ? nch * COMBOS * 4 : 0);
    double* red = smat + (HEAD == BEAR_HEAD_LINEAR ? lag * A1 * A1 : 0);
    if (HEAD == BEAR_HEAD_LINEAR) {
        for (int i = threadIdx.x; i < lag * A1 * A1; i += blockDim.x) smat[i] = head[i];
        __syncthreads();
        build_ratio_tables(smat, R, lag);
        __syncthreads();
    }
    double hinv[BEAR_MAX_MODELS], van[BEAR_MAX_MODELS];
#pragma unroll
    for (int k = 0; k < BEAR_MAX_MODELS; ++k) {
        hinv[k] = k < H ? 1.0 / d_h[k] : 0.0;
        van[k] = k < V ? d_van[k] : 0.0;
    }
    double ll_ear[BEAR_MAX_MODELS], cor_ear[BEAR_MAX_MODELS], ll_van[BEAR_MAX_MODELS], cor_van[BEAR_MAX_MODELS];
#pragma unroll
    for (int k = 0; k < BEAR_MAX_MODELS; ++k) ll_ear[k] = cor_ear[k] = ll_van[k] = cor_van[k] = 0.0;
    double ll_arm = 0.0, cor_arm = 0.0, total = 0.0;

    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const Row r = load_row(test_col, stride, i);
        if (r.n == 0.0) continue;              // no test transitions: contributes 0 to every output
        double t[A1] = {0, 0, 0, 0, 0};
        if (train_col) {
#pragma unroll
            for (int b = 0; b < A1; ++b) t[b] = double(__ldg(train_col + b * stride + i));
        }
        double f[A1] = {0, 0, 0, 0, 0};
        if (HEAD == BEAR_HEAD_LINEAR) {
            const uint64_t code = __ldg(kmers + i);
            const int nstart = int(code >> 58);
            const uint64_t v = code & ((1ull << 58) - 1);
            if (!(nstart == 0 && linear_head_fast(R, v, lag, nch, f))) linear_head_slow(smat, v, lag, nstart, f);
        } else if (HEAD == BEAR_HEAD_EXPLICIT) {
#pragma unroll
            for (int b = 0; b < A1; ++b) f[b] = head[i * A1 + b];
        } else if (HEAD == BEAR_HEAD_STOP) {
            f[A1 - 1] = 1.0;
        }
        total += r.n;
        const uint64_t grow = uint64_t(row0 + i);
        double dummy[A1];
        // BEAR: conc = f / h + train + eps   (bear_net.py:43, 335-337)
        for (int k = 0; k < H; ++k) {
            double conc[A1];
#pragma unroll
            for (int b = 0; b < A1; ++b) conc[b] = fma(f[b], hinv[k], t[b]) + BEAR_EPS;
            ll_ear[k] += dm_row<false>(conc, r, dummy);
            cor_ear[k] += r.c[noisy_argmax<A1>(conc, 100.0 * BEAR_EPS, seed, grow, uint64_t(k))];
        }
        // AR: p = f + eps   (bear_net.py:68, 338)
        {
            double p[A1];
#pragma unroll
            for (int b = 0; b < A1; ++b) p[b] = f[b] + BEAR_EPS;
            ll_arm += mn_row(p, r);
            cor_arm += r.c[noisy_argmax<A1>(p, BEAR_EPS, seed, grow, 100)];
        }
        // vanilla BMM: conc = train + van + eps   (bear_net.py:328-331, 339-340)
        for (int k = 0; k < V; ++k) {
            double conc[A1];
#pragma unroll
            for (int b = 0; b < A1; ++b) conc[b] = (t[b] + van[k]) + BEAR_EPS;
            ll_van[k] += dm_row<false>(conc, r, dummy);
            cor_van[k] += r.c[noisy_argmax<A1>(conc, 100.0 * BEAR_EPS, seed, grow, 200 + uint64_t(k))];
        }
    }
    // layout: [ll_ear[H], ll_arm, ll_van[V], cor_ear[H], cor_arm, cor_van[V], total]
    const int P = 2 * H + 2 * V + 3;
    double* out = partials + int64_t(blockIdx.x) * P;
    int o = 0;
    for (int k = 0; k < H; ++k, ++o) { const double s = block_sum(ll_ear[k], red); if (threadIdx.x == 0) out[o] = s; }
    { const double s = block_sum(ll_arm, red); if (threadIdx.x == 0) out[o] = s; ++o; }
    for (int k = 0; k < V; ++k, ++o) { const double s = block_sum(ll_van[k], red); if (threadIdx.x == 0) out[o] = s; }
    for (int k = 0; k < H; ++k, ++o) { const double s = block_sum(cor_ear[k], red); if (threadIdx.x == 0) out[o] = s; }
    { const double s = block_sum(cor_arm, red); if (threadIdx.x == 0) out[o] = s; ++o; }
    for (int k = 0; k < V; ++k, ++o) { const double s = block_sum(cor_van[k], red); if (threadIdx.x == 0) out[o] = s; }
    { const double s = block_sum(total, red); if (threadIdx.x == 0) out[o] = s; }
}

// ------------------------------------------------------------------------------------------------
// BMM marginal likelihood, every group and alpha
// ------------------------------------------------------------------------------------------------
template <int NA1>
__global__ void __launch_bounds__(THREADS)
bmm_kernel(const uint32_t* __restrict__ counts, int64_t stride, int64_t n, int G,
           const double* __restrict__ d_alpha, int V, double* __restrict__ partials) {
    __shared__ double red[32];
    const int g = blockIdx.y;
    const uint32_t* col = counts + int64_t(g) * NA1 * stride;
    double alpha[BEAR_MAX_MODELS], acc[BEAR_MAX_MODELS];
#pragma unroll
    for (int k = 0; k < BEAR_MAX_MODELS; ++k) {
        alpha[k] = k < V ? d_alpha[k] : 1.0;
        acc[k] = 0.0;
    }
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        double c[NA1], tot = 0.0;
#pragma unroll
        for (int b = 0; b < NA1; ++b) {
            c[b] = double(__ldg(col + b * stride + i));
            tot += c[b];
        }
        if (tot == 0.0) continue;
        for (int k = 0; k < V; ++k) {
            LogProd num, den;
            den.push(lgdg_diff<false>(double(NA1) * alpha[k], tot));
#pragma unroll
            for (int b = 0; b < NA1; ++b) num.push(lgdg_diff<false>(alpha[k], c[b]));
            acc[k] += logprod_diff(num, den);
        }
    }
    for (int k = 0; k < V; ++k) {
        const double s = block_sum(acc[k], red);
        if (threadIdx.x == 0) partials[(int64_t(blockIdx.x) * G + g) * V + k] = s;
    }
}

int grid_for(int64_t n) {
    int64_t blocks = (n + THREADS - 1) / THREADS;
    if (blocks < 1) blocks = 1;
    return int(blocks < MAX_GRID ? blocks : MAX_GRID);
}

size_t train_smem_bytes(int lag) {
    return sizeof(double) * (size_t(num_chunks(lag)) * COMBOS * 4 * 2 + size_t(lag) * A1 * A1 * 2 + 32);
}

size_t eval_smem_bytes(int head, int lag) {
    size_t d = 32;
    if (head == BEAR_HEAD_LINEAR) d += size_t(num_chunks(lag)) * COMBOS * 4 + size_t(lag) * A1 * A1;
    return sizeof(double) * d;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        BEAR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    }
    return 0;
}

}  // namespace

extern "C" int64_t bear_workspace_doubles(int64_t n, int lag, int nparams) {
    (void)n;
    int64_t p = 2 + int64_t(lag) * A1 * A1;
    if (nparams + 2 > p) p = nparams + 2;
    if (p < 64) p = 64;
    return int64_t(MAX_GRID) * p;
}

extern "C" int bear_linear_train_step(const uint64_t* d_kmers, const uint32_t* d_col, int64_t stride,
                                      int64_t row0, int64_t n, int lag, const double* d_mat,
                                      const double* d_h_signed, double scale, int train_ar,
                                      double* d_flat, double* d_ll_out, double* d_workspace, void* stream) {
    const char* fn = "bear_linear_train_step";
    BEAR_REQUIRE(d_kmers && d_col && d_mat && d_h_signed && d_flat && d_workspace, fn);
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    BEAR_REQUIRE(lag >= 1 && lag <= 29, fn);
    if (n == 0) return BEAR_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n);
    const size_t smem = train_smem_bytes(lag);
    const int P = 2 + lag * A1 * A1;
    if (train_ar) {
        if (set_smem(linear_train_kernel<true>, smem)) return BEAR_ERR_CUDA;
        linear_train_kernel<true><<<grid, THREADS, smem, st>>>(d_kmers + row0, d_col + row0, stride, n, lag, d_mat,
                                                              d_h_signed, d_ll_out, d_workspace);
    } else {
        if (set_smem(linear_train_kernel<false>, smem)) return BEAR_ERR_CUDA;
        linear_train_kernel<false><<<grid, THREADS, smem, st>>>(d_kmers + row0, d_col + row0, stride, n, lag, d_mat,
                                                               d_h_signed, d_ll_out, d_workspace);
    }
    BEAR_LAUNCH_CHECK("linear_train_kernel");
    reduce_partials_kernel<<<(P + 127) / 128, 128, 0, st>>>(d_workspace, grid, P, -scale, d_flat);
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
    const int grid = grid_for(n);
    const size_t smem = eval_smem_bytes(head, lag);
    const uint64_t* km = d_kmers ? d_kmers + row0 : nullptr;
    const uint32_t* tr = d_train_col ? d_train_col + row0 : nullptr;
#define BEAR_EVAL_LAUNCH(HEADV, HEADPTR)                                                                       \
    do {                                                                                                       \
        if (set_smem(eval_kernel<HEADV>, smem)) return BEAR_ERR_CUDA;                                          \
        eval_kernel<HEADV><<<grid, THREADS, smem, st>>>(km, d_test_col + row0, tr, stride, row0, n, lag,       \
                                                        HEADPTR, d_h, H, d_van, V, seed, d_workspace);         \
    } while (0)
    switch (head) {
        case BEAR_HEAD_LINEAR: BEAR_EVAL_LAUNCH(BEAR_HEAD_LINEAR, d_head); break;
        case BEAR_HEAD_EXPLICIT: BEAR_EVAL_LAUNCH(BEAR_HEAD_EXPLICIT, d_head); break;
        case BEAR_HEAD_STOP: BEAR_EVAL_LAUNCH(BEAR_HEAD_STOP, d_head); break;
        default: BEAR_EVAL_LAUNCH(BEAR_HEAD_NONE, d_head); break;
    }
#undef BEAR_EVAL_LAUNCH
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
    if (A1v == 5)
        bmm_kernel<5><<<grid, THREADS, 0, st>>>(d_counts + row0, stride, n, G, d_alpha, V, d_workspace);
    else
        bmm_kernel<21><<<grid, THREADS, 0, st>>>(d_counts + row0, stride, n, G, d_alpha, V, d_workspace);
    BEAR_LAUNCH_CHECK("bmm_kernel");
    reduce_partials_kernel<<<1, 64, 0, st>>>(d_workspace, int(grid.x), G * V, 1.0, d_out);
    BEAR_LAUNCH_CHECK("reduce_partials_kernel");
    return BEAR_OK;
}
