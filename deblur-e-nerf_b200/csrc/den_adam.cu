// Adam over a list of fp32 parameter tensors in one or two launches (sm_100a).
//
// Replaces torch.optim.Adam as configured by DeblurENeRF.configure_optimizers
// (models/deblur_e_nerf.py:1055-1112: param groups with their own lr / weight decay, betas
// (0.9, 0.999), eps 1e-8, L2 weight decay added to the gradient, no amsgrad) for the fp32
// parameters — the 12.6 M-entry hash table is 99.9 % of them.  Semantics of one step t (1-based),
// exactly torch's single-tensor formula:
//     g  = grad_scale * grad + weight_decay * p        (grad_scale: 1 / world size folded in under
//                                                       data parallelism — the all-reduce sums)
//     m  = beta1 * m + (1 - beta1) * g
//     v  = beta2 * v + (1 - beta2) * g * g
//     p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// HBM-bound: 16 B read + 12 B written per parameter.
#include "den_common.cuh"

namespace den {

constexpr int kAdamMaxTensors = 24;

struct AdamBatch {
    float* p[kAdamMaxTensors];
    const float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    int64_t n[kAdamMaxTensors];
    float lr[kAdamMaxTensors];
    float wd[kAdamMaxTensors];
};

// omb1 = 1 - beta1, omb2 = 1 - beta2 come from the host in double precision (1.f - 0.999f is off by 1.3e-5)
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float wd, float b1, float b2,
                                            float omb1, float omb2, float step_size, float inv_sqrt_bias2,
                                            float eps, float gs) {
    g = fmaf(wd, p, g * gs);
    m = fmaf(b1, m, omb1 * g);
    v = fmaf(b2, v, omb2 * g * g);
    const float denom = sqrtf(v) * inv_sqrt_bias2 + eps;
    p -= step_size * (m / denom);
}

// blockIdx.y = tensor of the batch, grid-stride over its elements (float4 when all four arrays are
// 16-byte aligned — every torch allocation is)
__global__ void __launch_bounds__(256)
adam_kernel(const __grid_constant__ AdamBatch b, float beta1, float beta2, float omb1, float omb2, float eps,
            float bias1, float inv_sqrt_bias2, float gs, const int64_t* __restrict__ step_dev, double beta1_d,
            double beta2_d, const int32_t* __restrict__ skip_flag) {
    // a step whose sample buffers overflowed (den_clamp_offsets set the flag) lost samples: its
    // gradients are not applied — the update is skipped on the device, nobody waits for the flag
    if (skip_flag != nullptr && *skip_flag != 0) return;
    if (step_dev != nullptr) {
        // the step number lives on the device (a captured CUDA graph replays this launch with the same
        // kernel arguments every step): the bias corrections are derived here
        const double t_d = (double)*step_dev;
        bias1 = (float)(1.0 - pow(beta1_d, t_d));
        inv_sqrt_bias2 = (float)(1.0 / sqrt(1.0 - pow(beta2_d, t_d)));
    }
    const int t = blockIdx.y;
    float* __restrict__ p = b.p[t];
    const float* __restrict__ g = b.g[t];
    float* __restrict__ m = b.m[t];
    float* __restrict__ v = b.v[t];
    const int64_t n = b.n[t];
    const float wd = b.wd[t];
    const float step_size = b.lr[t] / bias1;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                           reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const int64_t n4 = aligned ? n / 4 : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        float4 p4 = reinterpret_cast<float4*>(p)[i];
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 m4 = reinterpret_cast<float4*>(m)[i];
        float4 v4 = reinterpret_cast<float4*>(v)[i];
        adam_update(p4.x, g4.x, m4.x, v4.x, wd, beta1, beta2, omb1, omb2, step_size, inv_sqrt_bias2, eps, gs);
        adam_update(p4.y, g4.y, m4.y, v4.y, wd, beta1, beta2, omb1, omb2, step_size, inv_sqrt_bias2, eps, gs);
        adam_update(p4.z, g4.z, m4.z, v4.z, wd, beta1, beta2, omb1, omb2, step_size, inv_sqrt_bias2, eps, gs);
        adam_update(p4.w, g4.w, m4.w, v4.w, wd, beta1, beta2, omb1, omb2, step_size, inv_sqrt_bias2, eps, gs);
        reinterpret_cast<float4*>(p)[i] = p4;
        reinterpret_cast<float4*>(m)[i] = m4;
        reinterpret_cast<float4*>(v)[i] = v4;
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += stride) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_update(pi, g[i], mi, vi, wd, beta1, beta2, omb1, omb2, step_size, inv_sqrt_bias2, eps, gs);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
    }
}

}  // namespace den

extern "C" int den_adam_step(const den_adam_tensor* tensors, int32_t n_tensors, double beta1_d, double beta2_d,
                             double eps_d, int64_t step, const int64_t* step_dev, double grad_scale,
                             const int32_t* skip_flag, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_tensors >= 0 && (step >= 1 || step_dev != nullptr), "bad tensor count / step");
    if (step < 1) step = 1;
    DEN_CHECK_ARG(n_tensors == 0 || tensors != nullptr, "null tensor list");
    const float beta1 = (float)beta1_d, beta2 = (float)beta2_d, eps = (float)eps_d;
    const double bias1 = 1.0 - pow(beta1_d, (double)step);
    const double bias2 = 1.0 - pow(beta2_d, (double)step);
    const float inv_sqrt_bias2 = (float)(1.0 / sqrt(bias2));
    // big tensors get their own grid-stride launch, the small ones share launches (blockIdx.y)
    for (int pass = 0; pass < 2; ++pass) {
        AdamBatch b;
        int count = 0;
        int64_t largest = 0;
        auto flush = [&]() -> int {
            if (count == 0) return DEN_OK;
            const int64_t blocks = (largest / 4 + 255) / 256;
            dim3 grid((unsigned)grid_for(blocks > 0 ? blocks : 1, 1, 8), (unsigned)count);
            adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(b, beta1, beta2, (float)(1.0 - (double)beta1_d),
                                                             (float)(1.0 - (double)beta2_d), eps, (float)bias1,
                                                             inv_sqrt_bias2, (float)grad_scale, step_dev, beta1_d,
                                                             beta2_d, skip_flag);
            cudaError_t e = cudaGetLastError();
            count = 0;
            largest = 0;
            return e == cudaSuccess ? DEN_OK : cuda_fail(e, "den_adam_step");
        };
        for (int i = 0; i < n_tensors; ++i) {
            const den_adam_tensor& t = tensors[i];
            if (t.n <= 0) continue;
            if (!(t.param && t.grad && t.exp_avg && t.exp_avg_sq)) {
                set_error("den_adam_step: null pointer in tensor %d", i);
                return DEN_ERR_INVALID_ARGUMENT;
            }
            const bool big = t.n >= (1 << 16);
            if (big != (pass == 0)) continue;
            b.p[count] = t.param; b.g[count] = t.grad; b.m[count] = t.exp_avg; b.v[count] = t.exp_avg_sq;
            b.n[count] = t.n; b.lr[count] = t.lr; b.wd[count] = t.weight_decay;
            largest = t.n > largest ? t.n : largest;
            ++count;
            if (big || count == kAdamMaxTensors) {
                int rc = flush();
                if (rc) return rc;
            }
        }
        int rc = flush();
        if (rc) return rc;
    }
    return DEN_OK;
}
