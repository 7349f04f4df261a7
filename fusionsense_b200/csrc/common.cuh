// common.cuh — launch helpers shared by every kernel file in libfsb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define FSB_API extern "C" __attribute__((visibility("default")))

// Every C-ABI entry point returns a cudaError_t-compatible int (0 = success) and
// never throws.  FSB_E_ARG is our own code for an argument the kernels refuse.
#define FSB_E_ARG 10001

// number of kernels this library has launched (bench.py reports it as "gpu_launches")
extern unsigned long long g_fsb_launches;

#define FSB_LAUNCH_CHECK()                         \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
        __atomic_fetch_add(&g_fsb_launches, 1ull, __ATOMIC_RELAXED); \
    } while (0)

#define FSB_CUDA(x)                                \
    do {                                           \
        cudaError_t e__ = (x);                     \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

static inline int fsb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

static inline size_t fsb_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#define FSB_NUM_SMS 148  // B200

// "Static capacity" mode (CUDA-graph capturable, no host read of the intersection count): the host passes the
// CAPACITY of the list buffers as the count and a device pointer to the true count; kernels use min(true, capacity).
#ifdef __CUDACC__
__device__ __forceinline__ int64_t fsb_eff_n(int64_t n_or_cap, const int64_t* __restrict__ n_dev) {
    if (n_dev == nullptr) return n_or_cap;
    const int64_t v = *n_dev;
    return v < n_or_cap ? v : n_or_cap;
}
#endif
