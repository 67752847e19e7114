// Transmittance / weights / accumulation and the fused front-to-back compositor (sm_100a).
//
// Replaces nerfacc.render_weight_from_density / render_weight_from_alpha /
// accumulate_along_rays and the reference's `rendering()` that strings them together
// (external/vol_rendering.py:16-128); semantics per nerfacc 0.3.1 vol_rendering.py and
// cuda/csrc/render_weight.cu, restated in oracle/nerfacc_ref.py.
//
// One warp owns one ray: its samples are contiguous (ray-major packing), so every
// load/store is a coalesced 128-byte line; the per-ray prefix (sum of sigma*dt, or
// product of 1-alpha) is a shuffle scan with a running carry, and the per-ray outputs
// are warp reductions written by lane 0 — deterministic, no atomics, unlike the
// reference's scatter_add_ (external/vol_rendering.py:111-122).
// HBM-bound by construction: fwd 20 B/sample + 12 B/ray, bwd 28 B/sample + 12 B/ray at C=1.
#include "den_common.cuh"

namespace den {

constexpr int kRayWarpsPerCta = 8;
constexpr int kRayThreads = kRayWarpsPerCta * 32;

#define DEN_FOR_EACH_RAY(r)                                                                   \
    const int lane = threadIdx.x & 31;                                                        \
    const int64_t warp0__ = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;            \
    const int64_t nwarps__ = ((int64_t)gridDim.x * blockDim.x) >> 5;                          \
    for (int64_t r = warp0__; r < n_rays; r += nwarps__)

// ---------------------------------------------------------------- weights ----
__global__ void __launch_bounds__(kRayThreads)
weight_density_fwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ t0,
                          const float* __restrict__ t1, const int32_t* __restrict__ offsets,
                          int64_t n_rays, float* __restrict__ weights) {
    DEN_FOR_EACH_RAY(r) {
        const int beg = offsets[r], end = offsets[r + 1];
        float carry = 0.f;
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const float sdt = i < end ? sigmas[i] * (t1[i] - t0[i]) : 0.f;
            const float inc = warp_inclusive_sum(sdt, lane);
            const float excl = carry + inc - sdt;
            if (i < end) weights[i] = __expf(-excl) * (1.f - __expf(-sdt));
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
}

// dL/dsigma_i = dt_i * ( g_i * T_i * (1 - a_i) - sum_{j>i} g_j w_j ),  g = dL/dw
__global__ void __launch_bounds__(kRayThreads)
weight_density_bwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ t0,
                          const float* __restrict__ t1, const int32_t* __restrict__ offsets,
                          int64_t n_rays, const float* __restrict__ dweights,
                          float* __restrict__ dsigmas) {
    DEN_FOR_EACH_RAY(r) {
        const int beg = offsets[r], end = offsets[r + 1];
        if (beg == end) continue;
        // pass A: total optical depth of the ray
        float tot = 0.f;
        for (int i = beg + lane; i < end; i += 32) tot += sigmas[i] * (t1[i] - t0[i]);
        tot = warp_sum(tot);
        // pass B: back to front, carrying the suffix sums
        float suf_sdt = 0.f, suf_gw = 0.f;
        const int n = end - beg;
        const int n_chunks = (n + 31) / 32;
        for (int c = n_chunks - 1; c >= 0; --c) {
            const int i = beg + c * 32 + lane;
            const bool ok = i < end;
            const float dt = ok ? (t1[i] - t0[i]) : 0.f;
            const float sdt = ok ? sigmas[i] * dt : 0.f;
            const float g = ok ? dweights[i] : 0.f;
            // inclusive suffix of sdt inside the chunk
            float suf = sdt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                float t = __shfl_down_sync(0xffffffffu, suf, d);
                if (lane + d < 32) suf += t;
            }
            const float excl_prefix = tot - (suf_sdt + suf);     // sum_{j<i} sdt_j
            const float T = __expf(-excl_prefix);
            const float keep = __expf(-sdt);                     // 1 - alpha
            const float w = T * (1.f - keep);
            const float gw = g * w;
            float sgw = gw;                                       // inclusive suffix of g*w
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                float t = __shfl_down_sync(0xffffffffu, sgw, d);
                if (lane + d < 32) sgw += t;
            }
            const float after = suf_gw + sgw - gw;               // sum_{j>i} g_j w_j
            if (ok) dsigmas[i] = dt * (g * T * keep - after);
            suf_sdt += __shfl_sync(0xffffffffu, suf, 0);
            suf_gw += __shfl_sync(0xffffffffu, sgw, 0);
        }
    }
}

__global__ void __launch_bounds__(kRayThreads)
weight_alpha_fwd_kernel(const float* __restrict__ alphas, const int32_t* __restrict__ offsets,
                        int64_t n_rays, float* __restrict__ weights) {
    DEN_FOR_EACH_RAY(r) {
        const int beg = offsets[r], end = offsets[r + 1];
        float carry = 1.f;
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const float a = i < end ? alphas[i] : 0.f;
            float inc = 1.f - a;                                  // inclusive product
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                float t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc *= t;
            }
            float excl = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) excl = 1.f;
            if (i < end) weights[i] = carry * excl * a;
            carry *= __shfl_sync(0xffffffffu, inc, 31);
        }
    }
}

// w_i = a_i T_i, T_i = prod_{j<i}(1-a_j):  dL/da_i = g_i T_i - (sum_{j>i} g_j w_j) / (1 - a_i)
__global__ void __launch_bounds__(kRayThreads)
weight_alpha_bwd_kernel(const float* __restrict__ alphas, const int32_t* __restrict__ offsets,
                        int64_t n_rays, const float* __restrict__ dweights,
                        float* __restrict__ dalphas) {
    DEN_FOR_EACH_RAY(r) {
        const int beg = offsets[r], end = offsets[r + 1];
        if (beg == end) continue;
        // pass A: total sum of g*w
        float carry = 1.f, tot = 0.f;
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const float a = i < end ? alphas[i] : 0.f;
            float inc = 1.f - a;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                float t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc *= t;
            }
            float excl = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) excl = 1.f;
            if (i < end) tot += dweights[i] * carry * excl * a;
            carry *= __shfl_sync(0xffffffffu, inc, 31);
        }
        tot = warp_sum(tot);
        // pass B: forward again with the running prefix of g*w
        carry = 1.f;
        float pre_gw = 0.f;
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const float a = i < end ? alphas[i] : 0.f;
            const float g = i < end ? dweights[i] : 0.f;
            float inc = 1.f - a;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                float t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc *= t;
            }
            float excl = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) excl = 1.f;
            const float T = carry * excl;
            const float gw = g * T * a;
            const float inc_gw = warp_inclusive_sum(gw, lane);
            const float after = tot - (pre_gw + inc_gw);
            if (i < end) dalphas[i] = g * T - after / fmaxf(1.f - a, 1e-10f);
            carry *= __shfl_sync(0xffffffffu, inc, 31);
            pre_gw += __shfl_sync(0xffffffffu, inc_gw, 31);
        }
    }
}

// ------------------------------------------------------------- accumulate ----
__global__ void __launch_bounds__(kRayThreads)
accumulate_fwd_kernel(const float* __restrict__ weights, const float* __restrict__ values,
                      const int32_t* __restrict__ offsets, int64_t n_rays, int dim,
                      float* __restrict__ out) {
    DEN_FOR_EACH_RAY(r) {
        const int beg = offsets[r], end = offsets[r + 1];
        for (int k = 0; k < dim; ++k) {
            float acc = 0.f;
            for (int i = beg + lane; i < end; i += 32)
                acc += (weights ? weights[i] : 1.f) * (values ? values[(int64_t)i * dim + k] : 1.f);
            acc = warp_sum(acc);
            if (lane == 0) out[r * dim + k] = acc;
        }
    }
}

__global__ void accumulate_bwd_kernel(const float* __restrict__ weights,
                                      const float* __restrict__ values,
                                      const int32_t* __restrict__ ray_indices,
                                      const float* __restrict__ dout, int64_t n, int dim,
                                      float* __restrict__ dweights, float* __restrict__ dvalues) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = ray_indices[i];
        const float w = weights[i];
        float dw = 0.f;
        for (int k = 0; k < dim; ++k) {
            const float go = dout[r * dim + k];
            if (values) {
                dw += go * values[i * dim + k];
                if (dvalues) dvalues[i * dim + k] = go * w;
            } else {
                dw += go;
            }
        }
        if (dweights) dweights[i] = dw;
    }
}

// ------------------------------------------------------- fused compositor ----
// Memory-level parallelism: a ray is consumed in groups of kGroup chunks of 32 samples; the loads
// of a whole group — sigma, t0, t1, rgb: 4 * kGroup independent coalesced 128-byte lines per
// warp — are issued in one burst before any of the dependent scan arithmetic of the group starts.
// v1 had one load -> scan -> load chain per chunk and sat at 0.23 of the HBM roofline,
// long-scoreboard bound (profiles/r01_ncu_full_misc_kernels.md).
constexpr int kGroup = 8;        // forward
constexpr int kGroupBwd = 4;     // backward (more live registers per chunk)

template <int C>
__global__ void __launch_bounds__(kRayThreads)
composite_fwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                     const float* __restrict__ t0, const float* __restrict__ t1,
                     const int32_t* __restrict__ offsets, int64_t n_rays,
                     const float* __restrict__ bkgd, float* __restrict__ colour,
                     float* __restrict__ opacity, float* __restrict__ depth) {
    DEN_FOR_EACH_RAY(r) {
        const int beg = offsets[r], end = offsets[r + 1];
        float carry = 0.f, acc_o = 0.f, acc_d = 0.f, acc_c[C];
#pragma unroll
        for (int k = 0; k < C; ++k) acc_c[k] = 0.f;
        for (int g0 = beg; g0 < end; g0 += 32 * kGroup) {
            float sg[kGroup], ta[kGroup], tb[kGroup], col[kGroup][C];
#pragma unroll
            for (int c = 0; c < kGroup; ++c) {
                const int i = g0 + 32 * c + lane;
                const bool ok = i < end;
                sg[c] = ok ? __ldg(sigmas + i) : 0.f;
                ta[c] = ok ? __ldg(t0 + i) : 0.f;
                tb[c] = ok ? __ldg(t1 + i) : 0.f;
#pragma unroll
                for (int k = 0; k < C; ++k) col[c][k] = ok ? __ldg(rgbs + (int64_t)i * C + k) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < kGroup; ++c) {
                if (g0 + 32 * c >= end) break;                     // warp-uniform
                const float sdt = sg[c] * (tb[c] - ta[c]);         // 0 beyond the ray's end
                const float inc = warp_inclusive_sum(sdt, lane);
                const float w = __expf(-(carry + inc - sdt)) * (1.f - __expf(-sdt));
                acc_o += w;
                acc_d += w * (ta[c] + tb[c]) * 0.5f;
#pragma unroll
                for (int k = 0; k < C; ++k) acc_c[k] += w * col[c][k];
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        acc_o = warp_sum(acc_o);
        acc_d = warp_sum(acc_d);
#pragma unroll
        for (int k = 0; k < C; ++k) acc_c[k] = warp_sum(acc_c[k]);
        if (lane == 0) {
            opacity[r] = acc_o;
            depth[r] = acc_d;
#pragma unroll
            for (int k = 0; k < C; ++k)
                colour[r * C + k] = acc_c[k] + (bkgd ? bkgd[k] * (1.f - acc_o) : 0.f);
        }
    }
}

// one chunk of the back-to-front sweep (shared by the cached and the streaming path)
template <int C>
__device__ __forceinline__ void composite_bwd_chunk(int lane, bool ok, float a, float b, float sigma,
                                                    const float (&rgb)[C], const float (&gc)[C], float g_op,
                                                    float g_dp, float tot, float& suf_sdt, float& suf_gw,
                                                    float& d_sigma, float (&d_rgb)[C]) {
    const float dt = b - a;
    const float sdt = ok ? sigma * dt : 0.f;
    float g = g_op + g_dp * (a + b) * 0.5f;               // dL/dw_i
#pragma unroll
    for (int k = 0; k < C; ++k) g += gc[k] * rgb[k];
    if (!ok) g = 0.f;
    float suf = sdt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, suf, d);
        if (lane + d < 32) suf += t;
    }
    const float T = __expf(-(tot - (suf_sdt + suf)));
    const float keep = __expf(-sdt);
    const float w = T * (1.f - keep);
    const float gw = g * w;
    float sgw = gw;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, sgw, d);
        if (lane + d < 32) sgw += t;
    }
    d_sigma = dt * (g * T * keep - (suf_gw + sgw - gw));
#pragma unroll
    for (int k = 0; k < C; ++k) d_rgb[k] = gc[k] * w;
    suf_sdt += __shfl_sync(0xffffffffu, suf, 0);
    suf_gw += __shfl_sync(0xffffffffu, sgw, 0);
}

template <int C>
__global__ void __launch_bounds__(kRayThreads)
composite_bwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                     const float* __restrict__ t0, const float* __restrict__ t1,
                     const int32_t* __restrict__ offsets, int64_t n_rays,
                     const float* __restrict__ bkgd, const float* __restrict__ opacity,
                     const float* __restrict__ d_colour, const float* __restrict__ d_opacity,
                     const float* __restrict__ d_depth, float* __restrict__ d_sigmas,
                     float* __restrict__ d_rgbs, float* __restrict__ d_bkgd) {
    float bk_acc[C];
#pragma unroll
    for (int k = 0; k < C; ++k) bk_acc[k] = 0.f;

    DEN_FOR_EACH_RAY(r) {
        const int beg = offsets[r], end = offsets[r + 1];
        const int n = end - beg;
        float gc[C];
        float g_op = d_opacity ? d_opacity[r] : 0.f;
        const float g_dp = d_depth ? d_depth[r] : 0.f;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            gc[k] = d_colour ? d_colour[r * C + k] : 0.f;
            if (bkgd) {
                g_op -= gc[k] * bkgd[k];
                if (lane == 0) bk_acc[k] += gc[k] * (1.f - opacity[r]);
            }
        }
        if (n == 0) continue;
        // pass A: total optical depth of the ray (burst loads, kGroupBwd * 2 chunks in flight)
        float tot = 0.f;
        for (int g0 = beg; g0 < end; g0 += 32 * kGroupBwd) {
            float sg[kGroupBwd], ta[kGroupBwd], tb[kGroupBwd];
#pragma unroll
            for (int c = 0; c < kGroupBwd; ++c) {
                const int i = g0 + 32 * c + lane;
                const bool ok = i < end;
                sg[c] = ok ? __ldg(sigmas + i) : 0.f;
                ta[c] = ok ? __ldg(t0 + i) : 0.f;
                tb[c] = ok ? __ldg(t1 + i) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < kGroupBwd; ++c) tot += sg[c] * (tb[c] - ta[c]);
        }
        tot = warp_sum(tot);
        // pass B: back to front in groups of kGroupBwd chunks (the ray's lines are L1/L2-hot from pass A)
        float suf_sdt = 0.f, suf_gw = 0.f;
        const int n_chunks = (n + 31) / 32;
        for (int c_hi = n_chunks; c_hi > 0; c_hi -= kGroupBwd) {
            float sg[kGroupBwd], ta[kGroupBwd], tb[kGroupBwd], col[kGroupBwd][C];
#pragma unroll
            for (int j = 0; j < kGroupBwd; ++j) {
                const int c = c_hi - 1 - j;
                const int i = beg + 32 * c + lane;
                const bool ok = c >= 0 && i < end;
                sg[j] = ok ? __ldg(sigmas + i) : 0.f;
                ta[j] = ok ? __ldg(t0 + i) : 0.f;
                tb[j] = ok ? __ldg(t1 + i) : 0.f;
#pragma unroll
                for (int k = 0; k < C; ++k) col[j][k] = ok ? __ldg(rgbs + (int64_t)i * C + k) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < kGroupBwd; ++j) {
                const int c = c_hi - 1 - j;
                if (c < 0) break;                                  // warp-uniform
                const int i = beg + 32 * c + lane;
                const bool ok = i < end;
                float ds, drgb[C];
                composite_bwd_chunk<C>(lane, ok, ta[j], tb[j], sg[j], col[j], gc, g_op, g_dp, tot, suf_sdt,
                                       suf_gw, ds, drgb);
                if (ok) {
                    d_sigmas[i] = ds;
#pragma unroll
                    for (int k = 0; k < C; ++k) d_rgbs[(int64_t)i * C + k] = drgb[k];
                }
            }
        }
    }
    if (bkgd && d_bkgd && (threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < C; ++k)
            if (bk_acc[k] != 0.f) atomicAdd(d_bkgd + k, bk_acc[k]);
    }
}

// Single-sweep backward: with the forward outputs at hand the suffix sum_{j>i} g_j w_j is
// total - prefix, where total = sum_i g_i w_i = g_op * opacity + g_depth * depth + sum_k g_c[k] * colour_acc[k]
// needs no pass over the samples — so the ray is read ONCE, front to back, exactly like the forward
// (28 B/sample of real traffic instead of 44).  The cancellation error of total - prefix is
// ~1e-7 * |total| absolute, i.e. only visible on samples whose gradient is itself negligible.
template <int C>
__global__ void __launch_bounds__(kRayThreads)
composite_bwd_sweep_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                           const float* __restrict__ t0, const float* __restrict__ t1,
                           const int32_t* __restrict__ offsets, int64_t n_rays,
                           const float* __restrict__ bkgd, const float* __restrict__ colour,
                           const float* __restrict__ opacity, const float* __restrict__ depth,
                           const float* __restrict__ d_colour, const float* __restrict__ d_opacity,
                           const float* __restrict__ d_depth, float* __restrict__ d_sigmas,
                           float* __restrict__ d_rgbs, float* __restrict__ d_bkgd) {
    float bk_acc[C];
#pragma unroll
    for (int k = 0; k < C; ++k) bk_acc[k] = 0.f;

    DEN_FOR_EACH_RAY(r) {
        const int beg = offsets[r], end = offsets[r + 1];
        const float opa = opacity[r];
        float gc[C];
        float g_op = d_opacity ? d_opacity[r] : 0.f;
        const float g_dp = d_depth ? d_depth[r] : 0.f;
        float total = g_dp * depth[r];
#pragma unroll
        for (int k = 0; k < C; ++k) {
            gc[k] = d_colour ? d_colour[r * C + k] : 0.f;
            float acc = colour[r * C + k];
            if (bkgd) {
                g_op -= gc[k] * bkgd[k];
                acc -= bkgd[k] * (1.f - opa);
                if (lane == 0) bk_acc[k] += gc[k] * (1.f - opa);
            }
            total += gc[k] * acc;
        }
        total += g_op * opa;
        float carry = 0.f, carry_gw = 0.f;
        for (int g0 = beg; g0 < end; g0 += 32 * kGroup) {
            float sg[kGroup], ta[kGroup], tb[kGroup], col[kGroup][C];
#pragma unroll
            for (int c = 0; c < kGroup; ++c) {
                const int i = g0 + 32 * c + lane;
                const bool ok = i < end;
                sg[c] = ok ? __ldg(sigmas + i) : 0.f;
                ta[c] = ok ? __ldg(t0 + i) : 0.f;
                tb[c] = ok ? __ldg(t1 + i) : 0.f;
#pragma unroll
                for (int k = 0; k < C; ++k) col[c][k] = ok ? __ldg(rgbs + (int64_t)i * C + k) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < kGroup; ++c) {
                if (g0 + 32 * c >= end) break;                     // warp-uniform
                const int i = g0 + 32 * c + lane;
                const bool ok = i < end;
                const float dt = tb[c] - ta[c];
                const float sdt = sg[c] * dt;                      // 0 beyond the ray's end
                float g = g_op + g_dp * (ta[c] + tb[c]) * 0.5f;
#pragma unroll
                for (int k = 0; k < C; ++k) g += gc[k] * col[c][k];
                const float inc = warp_inclusive_sum(sdt, lane);
                const float T = __expf(-(carry + inc - sdt));
                const float keep = __expf(-sdt);
                const float w = T * (1.f - keep);
                const float gw = ok ? g * w : 0.f;
                const float ginc = warp_inclusive_sum(gw, lane);
                if (ok) {
                    d_sigmas[i] = dt * (g * T * keep - (total - (carry_gw + ginc)));
#pragma unroll
                    for (int k = 0; k < C; ++k) d_rgbs[(int64_t)i * C + k] = gc[k] * w;
                }
                carry += __shfl_sync(0xffffffffu, inc, 31);
                carry_gw += __shfl_sync(0xffffffffu, ginc, 31);
            }
        }
    }
    if (bkgd && d_bkgd && (threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < C; ++k)
            if (bk_acc[k] != 0.f) atomicAdd(d_bkgd + k, bk_acc[k]);
    }
}

inline int ray_grid(int64_t n_rays) { return grid_for(n_rays, kRayWarpsPerCta, 8); }

}  // namespace den

extern "C" {

int den_weight_from_density_fwd(const float* sigmas, const float* t0, const float* t1,
                                const int32_t* offsets, int64_t n_rays, float* weights,
                                void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "negative ray count");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(offsets, "null offsets");
    weight_density_fwd_kernel<<<ray_grid(n_rays), kRayThreads, 0, as_stream(stream)>>>(
        sigmas, t0, t1, offsets, n_rays, weights);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_weight_from_density_bwd(const float* sigmas, const float* t0, const float* t1,
                                const int32_t* offsets, int64_t n_rays, const float* dweights,
                                float* dsigmas, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "negative ray count");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(offsets, "null offsets");
    weight_density_bwd_kernel<<<ray_grid(n_rays), kRayThreads, 0, as_stream(stream)>>>(
        sigmas, t0, t1, offsets, n_rays, dweights, dsigmas);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_weight_from_alpha_fwd(const float* alphas, const int32_t* offsets, int64_t n_rays,
                              float* weights, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "negative ray count");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(offsets, "null offsets");
    weight_alpha_fwd_kernel<<<ray_grid(n_rays), kRayThreads, 0, as_stream(stream)>>>(
        alphas, offsets, n_rays, weights);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_weight_from_alpha_bwd(const float* alphas, const int32_t* offsets, int64_t n_rays,
                              const float* dweights, float* dalphas, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "negative ray count");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(offsets, "null offsets");
    weight_alpha_bwd_kernel<<<ray_grid(n_rays), kRayThreads, 0, as_stream(stream)>>>(
        alphas, offsets, n_rays, dweights, dalphas);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_accumulate_fwd(const float* weights, const float* values, const int32_t* offsets,
                       int64_t n_rays, int32_t dim, float* out, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0 && dim >= 1, "bad size");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(offsets && out, "null pointer");
    accumulate_fwd_kernel<<<ray_grid(n_rays), kRayThreads, 0, as_stream(stream)>>>(
        weights, values, offsets, n_rays, dim, out);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_accumulate_bwd(const float* weights, const float* values, const int32_t* ray_indices,
                       const float* dout, int64_t n, int32_t dim, float* dweights, float* dvalues,
                       void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0 && dim >= 1, "bad size");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(weights && ray_indices && dout, "null pointer");
    accumulate_bwd_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        weights, values, ray_indices, dout, n, dim, dweights, dvalues);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_composite_fwd(const float* sigmas, const float* rgbs, const float* t0, const float* t1,
                      const int32_t* offsets, int64_t n_rays, int32_t channels, const float* bkgd,
                      float* colour, float* opacity, float* depth, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "negative ray count");
    DEN_CHECK_ARG(channels == 1 || channels == 3, "channels must be 1 or 3");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(offsets && colour && opacity && depth, "null pointer");
    if (channels == 1)
        composite_fwd_kernel<1><<<ray_grid(n_rays), kRayThreads, 0, as_stream(stream)>>>(
            sigmas, rgbs, t0, t1, offsets, n_rays, bkgd, colour, opacity, depth);
    else
        composite_fwd_kernel<3><<<ray_grid(n_rays), kRayThreads, 0, as_stream(stream)>>>(
            sigmas, rgbs, t0, t1, offsets, n_rays, bkgd, colour, opacity, depth);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_composite_bwd(const float* sigmas, const float* rgbs, const float* t0, const float* t1,
                      const int32_t* offsets, int64_t n_rays, int32_t channels, const float* bkgd,
                      const float* colour, const float* opacity, const float* depth,
                      const float* d_colour, const float* d_opacity, const float* d_depth,
                      float* d_sigmas, float* d_rgbs, float* d_bkgd, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "negative ray count");
    DEN_CHECK_ARG(channels == 1 || channels == 3, "channels must be 1 or 3");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(offsets, "null offsets");
    DEN_CHECK_ARG(!bkgd || opacity, "background blend needs the forward opacity");
    const int grid = ray_grid(n_rays);
    if (colour && opacity && depth) {
        // forward outputs available: single sweep
        if (channels == 1)
            composite_bwd_sweep_kernel<1><<<grid, kRayThreads, 0, as_stream(stream)>>>(
                sigmas, rgbs, t0, t1, offsets, n_rays, bkgd, colour, opacity, depth, d_colour, d_opacity,
                d_depth, d_sigmas, d_rgbs, d_bkgd);
        else
            composite_bwd_sweep_kernel<3><<<grid, kRayThreads, 0, as_stream(stream)>>>(
                sigmas, rgbs, t0, t1, offsets, n_rays, bkgd, colour, opacity, depth, d_colour, d_opacity,
                d_depth, d_sigmas, d_rgbs, d_bkgd);
    } else if (channels == 1) {
        composite_bwd_kernel<1><<<grid, kRayThreads, 0, as_stream(stream)>>>(
            sigmas, rgbs, t0, t1, offsets, n_rays, bkgd, opacity, d_colour, d_opacity, d_depth,
            d_sigmas, d_rgbs, d_bkgd);
    } else {
        composite_bwd_kernel<3><<<grid, kRayThreads, 0, as_stream(stream)>>>(
            sigmas, rgbs, t0, t1, offsets, n_rays, bkgd, opacity, d_colour, d_opacity, d_depth,
            d_sigmas, d_rgbs, d_bkgd);
    }
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
