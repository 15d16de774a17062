// Small kernels around the fused path: the reference-genome head of bear_ref.py and the Keras-style
// Adam update applied to the flat parameter buffer after the per-step allreduce.
#include <math.h>

#include "bear_b200.h"
#include "bear_common.cuh"
#include "bear_host.h"

namespace {

using namespace bear;

constexpr int A1 = 5;
constexpr int THREADS = 256;
constexpr int MAX_GRID = 148 * 4;

inline int grid_for(int64_t n) {
    int64_t b = (n + THREADS - 1) / THREADS;
    if (b < 1) b = 1;
    return int(b < MAX_GRID ? b : MAX_GRID);
}

// Jukes-Cantor transition probabilities from reference counts (bear_ref.py:9-33) with the map of
// bear_ref.py:332-337 applied first: r = (ref + eps) * not_stop; p = r / sum|r|;
// jc = u + exp(-tau) (p - u), u = [1/A, ..., 1/A, 0].
__device__ __forceinline__ void jukes_cantor(const uint32_t* __restrict__ col, int64_t stride, int64_t i, double etau,
                                             double (&p)[A1], double (&jc)[A1]) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < A1 - 1; ++b) {
        p[b] = double(__ldg(col + b * stride + i)) + BEAR_EPS;
        s += p[b];
    }
    p[A1 - 1] = 0.0;
    const double si = 1.0 / s;
    const double u = 1.0 / double(A1 - 1);
#pragma unroll
    for (int b = 0; b < A1 - 1; ++b) {
        p[b] *= si;
        jc[b] = u + etau * (p[b] - u);
    }
    jc[A1 - 1] = 0.0;
}

// bear_ref._make_ref_ar_func.ar_func (bear_ref.py:63-68): f = (nw g + jc) / (nw + 1)
__global__ void ref_head_kernel(const uint32_t* __restrict__ col, int64_t stride, int64_t n, const double* __restrict__ g,
                                const double* __restrict__ tau_signed, const double* __restrict__ nw_signed,
                                double* __restrict__ f) {
    const double etau = exp(-exp(tau_signed[0]));
    const double nw = exp(nw_signed[0]);
    const double inv = 1.0 / (nw + 1.0);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        double p[A1], jc[A1];
        jukes_cantor(col, stride, i, etau, p, jc);
#pragma unroll
        for (int b = 0; b < A1; ++b) {
            const double gb = g ? g[i * A1 + b] : (b == A1 - 1 ? 1.0 : 0.0);   // NULL = stop head (ar_funcs.py:121-126)
            f[i * A1 + b] = (nw * gb + jc[b]) * inv;
        }
    }
}

__global__ void ref_head_bwd_kernel(const uint32_t* __restrict__ col, int64_t stride, int64_t n, const double* __restrict__ g,
                                    const double* __restrict__ tau_signed, const double* __restrict__ nw_signed,
                                    const double* __restrict__ gf, double* __restrict__ gg, double* __restrict__ partials) {
    __shared__ double red[32];
    const double tau = exp(tau_signed[0]);
    const double etau = exp(-tau);
    const double nw = exp(nw_signed[0]);
    const double inv = 1.0 / (nw + 1.0);
    double dtau = 0.0, dnw = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        double p[A1], jc[A1];
        jukes_cantor(col, stride, i, etau, p, jc);
        const double u = 1.0 / double(A1 - 1);
#pragma unroll
        for (int b = 0; b < A1; ++b) {
            const double gb = g ? g[i * A1 + b] : (b == A1 - 1 ? 1.0 : 0.0);
            const double up = gf[i * A1 + b];
            const double fb = (nw * gb + jc[b]) * inv;
            if (b < A1 - 1) dtau += up * (-tau * etau * (p[b] - u) * inv);   // d f / d tau_signed
            dnw += up * (nw * (gb - fb) * inv);                              // d f / d net_weight_signed
            if (gg) gg[i * A1 + b] = up * nw * inv;
        }
    }
    const double a = block_sum(dtau, red);
    const double b = block_sum(dnw, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x * 2 + 0] = a;
        partials[blockIdx.x * 2 + 1] = b;
    }
}

__global__ void sum2_kernel(const double* __restrict__ partials, int nblk, double* __restrict__ out) {
    const int p = threadIdx.x;
    if (p >= 2) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partials[b * 2 + p];
    out[p] += s;
}

// tf.keras.optimizers.Adam (OptimizerV2) as used by bear_net.py:264-265,277-282
__global__ void adam_kernel(double* __restrict__ p, const double* __restrict__ g, double* __restrict__ m, double* __restrict__ v,
                            int64_t n, double lr, double b1, double b2, double eps, const int64_t* __restrict__ step) {
    const double t = double(step[0] + 1);
    const double lr_t = lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t));
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double gi = g[i];
        const double mi = b1 * m[i] + (1.0 - b1) * gi;
        const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= lr_t * mi / (sqrt(vi) + eps);
    }
}

__global__ void bump_kernel(int64_t* step) { step[0] += 1; }

// One optimizer step of the training loop as ONE launch (small parameter vectors: one CTA): records the loss,
// applies Adam, clears the accumulated gradient buffer and advances the step counter.
__global__ void __launch_bounds__(1024)
adam_step_kernel(double* __restrict__ p, double* __restrict__ flat, double* __restrict__ m, double* __restrict__ v, int64_t n,
                 double lr, double b1, double b2, double eps, int64_t* __restrict__ step, double* __restrict__ loss_out,
                 double loss_scale, int zero_flat) {
    const double t = double(step[0] + 1);
    const double lr_t = lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t));
    const double loss = flat[0];
    __syncthreads();                                     // every thread has read the step counter and the loss
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double gi = flat[1 + i];
        const double mi = b1 * m[i] + (1.0 - b1) * gi;
        const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= lr_t * mi / (sqrt(vi) + eps);
        if (zero_flat) flat[1 + i] = 0.0;
    }
    if (threadIdx.x == 0) {
        if (loss_out) loss_out[0] = loss_scale * loss;
        if (zero_flat) flat[0] = 0.0;
        step[0] += 1;
    }
}

}  // namespace

#define ST(stream) static_cast<cudaStream_t>(stream)

extern "C" int bear_ref_head(const uint32_t* d_ref_col, int64_t stride, int64_t row0, int64_t n, const double* d_g,
                             const double* d_tau_signed, const double* d_nw_signed, double* d_f, void* stream) {
    const char* fn = "bear_ref_head";
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_ref_col && d_tau_signed && d_nw_signed && d_f, fn);
    ref_head_kernel<<<grid_for(n), THREADS, 0, ST(stream)>>>(d_ref_col + row0, stride, n, d_g, d_tau_signed, d_nw_signed, d_f);
    BEAR_LAUNCH_CHECK("ref_head_kernel");
    return BEAR_OK;
}

extern "C" int bear_ref_head_bwd(const uint32_t* d_ref_col, int64_t stride, int64_t row0, int64_t n, const double* d_g,
                                 const double* d_tau_signed, const double* d_nw_signed, const double* d_gf, double* d_gg,
                                 double* d_flat2, double* d_workspace, void* stream) {
    const char* fn = "bear_ref_head_bwd";
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_ref_col && d_tau_signed && d_nw_signed && d_gf && d_flat2 && d_workspace, fn);
    const int grid = grid_for(n);
    ref_head_bwd_kernel<<<grid, THREADS, 0, ST(stream)>>>(d_ref_col + row0, stride, n, d_g, d_tau_signed, d_nw_signed, d_gf,
                                                          d_gg, d_workspace);
    BEAR_LAUNCH_CHECK("ref_head_bwd_kernel");
    sum2_kernel<<<1, 32, 0, ST(stream)>>>(d_workspace, grid, d_flat2);
    BEAR_LAUNCH_CHECK("sum2_kernel");
    return BEAR_OK;
}

extern "C" int bear_adam_step(double* d_params, double* d_flat, double* d_m, double* d_v, int64_t n, double lr, double beta1,
                              double beta2, double eps, int64_t* d_step, double* d_loss_out, double loss_scale, int zero_flat,
                              void* stream) {
    const char* fn = "bear_adam_step";
    BEAR_REQUIRE(n >= 0 && n <= (int64_t(1) << 22), fn);
    BEAR_REQUIRE(d_params && d_flat && d_m && d_v && d_step, fn);
    adam_step_kernel<<<1, 1024, 0, ST(stream)>>>(d_params, d_flat, d_m, d_v, n, lr, beta1, beta2, eps, d_step, d_loss_out, loss_scale,
                                                 zero_flat);
    BEAR_LAUNCH_CHECK("adam_step_kernel");
    return BEAR_OK;
}

extern "C" int bear_adam_update(double* d_params, const double* d_grads, double* d_m, double* d_v, int64_t n, double lr,
                                double beta1, double beta2, double eps, int64_t* d_step, void* stream) {
    const char* fn = "bear_adam_update";
    BEAR_REQUIRE(n >= 0, fn);
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(d_params && d_grads && d_m && d_v && d_step, fn);
    adam_kernel<<<grid_for(n), THREADS, 0, ST(stream)>>>(d_params, d_grads, d_m, d_v, n, lr, beta1, beta2, eps, d_step);
    BEAR_LAUNCH_CHECK("adam_kernel");
    bump_kernel<<<1, 1, 0, ST(stream)>>>(d_step);
    BEAR_LAUNCH_CHECK("bump_kernel");
    return BEAR_OK;
}
