// Shared device/host helpers for the den_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "den_b200.h"

namespace den {

constexpr int kSmCountDefault = 148;   // B200: 2 dies x 74 SMs

// ---- error reporting (thread-local message behind den_last_error()) ---------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);

#define DEN_CHECK_ARG(cond, msg)                                    \
    do {                                                            \
        if (!(cond)) {                                              \
            ::den::set_error("%s: %s", __func__, msg);              \
            return DEN_ERR_INVALID_ARGUMENT;                        \
        }                                                           \
    } while (0)

#define DEN_CHECK_LAUNCH()                                          \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return ::den::cuda_fail(e__, __func__); \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// grid size for a grid-stride kernel: enough CTAs for `work` items, capped at a
// multiple of the SM count (persistent-style residency).
int sm_count();
inline int grid_for(int64_t work_items, int per_cta, int ctas_per_sm) {
    int64_t need = (work_items + per_cta - 1) / per_cta;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- device helpers ----------------------------------------------------------
// Sample count of a per-sample kernel: the host passes the CAPACITY of the buffers as `n`; when the
// true count lives on the device (written by the march / the compaction of the same step, so the
// host never has to read it back) `n_dev` points at it and the kernel works on min(n, *n_dev) rows.
__device__ __forceinline__ int64_t effective_n(int64_t n, const int32_t* __restrict__ n_dev) {
    if (n_dev == nullptr) return n;
    const int64_t m = (int64_t)*n_dev;
    return m < n ? (m > 0 ? m : 0) : n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// inclusive prefix sum across the warp
__device__ __forceinline__ float warp_inclusive_sum(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ float softplus_beta(float x, float beta, float threshold = 20.f) {
    // torch.nn.functional.softplus: x if beta*x > threshold else log1p(exp(beta*x))/beta
    float bx = beta * x;
    return bx > threshold ? x : log1pf(expf(bx)) / beta;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace den
