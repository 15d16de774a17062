// Internal host-side helpers shared by the translation units of libbear_b200.so.
#ifndef BEAR_HOST_H
#define BEAR_HOST_H

void bear_set_error(const char* fmt, ...);

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define BEAR_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            bear_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return BEAR_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)
#define BEAR_LAUNCH_CHECK(name)                                                            \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            bear_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));       \
            return BEAR_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)
#endif

#define BEAR_REQUIRE(cond, fn)                                                             \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            bear_set_error("%s: bad argument (%s)", fn, #cond);                            \
            return BEAR_ERR_ARG;                                                           \
        }                                                                                  \
    } while (0)

#endif
