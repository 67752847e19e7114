// Evaluation post-processing on the device — replaces the CPU part of
// DeblurENeRF.evaluation_epoch_end (models/deblur_e_nerf.py:705-969): the reference moves every
// predicted / target image to the host and runs the float64 affine least squares in log space
// (torch.linalg.lstsq over B*H*W x 2, :789-797), the Levenberg-Marquardt refinement of the
// offset-gamma correction (models/offset_gamma_correction.py, external/optimizer.py:60-111: the normal
// equations J^T J of a B*H*W x 3 Jacobian) and the per-image L1 / PSNR terms
// (loss_metric/metric.py:57-72) there.  All three are sums over pixels: here each is ONE pass over the
// images with fp64 accumulation (per-thread -> warp shuffle -> one fp64 atomic per CTA and moment), and
// the 2x2 / 3x3 solves run on the handful of moments.
//
//   den_eval_affine_moments   per channel c: n, sum x, sum y, sum xx, sum xy   with x = log pred,
//                             y = log target - log g_b          (g_b: normalised gain-exposure product)
//   den_eval_lm_moments       per channel c: J^T J (6), J^T r (3), sum r^2 of the model
//                             f(x) = g_b (s x^gamma - o),  r = f(x) - target,  J = d f / d (s, gamma, o),
//                             x = exp(a log pred + b): the affinely corrected prediction, formed in fp64
//                             from the raw prediction (never stored)
//   den_eval_apply            corrected prediction f(x) (written) + per-image sum |d|, sum d^2
// params (C, 5) fp64 on the device: a, b (log-space affine), s, gamma, o.
#include "den_common.cuh"

namespace den {

constexpr int kEvalThreads = 256;

template <int K>
__device__ __forceinline__ void block_accumulate(double (&v)[K], double* __restrict__ out) {
    __shared__ double s_red[K];
    if (threadIdx.x < K) s_red[threadIdx.x] = 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = v[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_red[k], x);
    }
    __syncthreads();
    if (threadIdx.x < K) atomicAdd(out + threadIdx.x, s_red[threadIdx.x]);
}

// images (B, C, HW) fp32; blockIdx.y = b * C + c
__global__ void __launch_bounds__(kEvalThreads)
eval_affine_moments_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                           const double* __restrict__ log_gain, int C, int64_t HW,
                           double* __restrict__ moments /* (C, 5) */) {
    const int bc = blockIdx.y, b = bc / C, c = bc - b * C;
    const float* p = pred + (int64_t)bc * HW;
    const float* t = target + (int64_t)bc * HW;
    const double lg = log_gain[b];
    double v[5] = {0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = (double)logf(p[i]);          // the reference takes the logs in fp32 (:733-734)
        const double y = (double)(logf(t[i]) - (float)lg);
        v[0] += 1.0; v[1] += x; v[2] += y; v[3] += x * x; v[4] += x * y;
    }
    block_accumulate<5>(v, moments + 5 * c);
}

__device__ __forceinline__ double corrected_input(float p, const double* __restrict__ prm) {
    return exp(prm[0] * (double)logf(p) + prm[1]);
}

__global__ void __launch_bounds__(kEvalThreads)
eval_lm_moments_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                       const double* __restrict__ gain, const double* __restrict__ params, int C,
                       int64_t HW, double* __restrict__ moments /* (C, 10) */) {
    const int bc = blockIdx.y, b = bc / C, c = bc - b * C;
    const float* p = pred + (int64_t)bc * HW;
    const float* t = target + (int64_t)bc * HW;
    const double* prm = params + 5 * c;
    const double g = gain[b], s = prm[2], gamma = prm[3], o = prm[4];
    double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = corrected_input(p[i], prm);
        const double xg = pow(x, gamma);
        const double js = g * xg;                     // d f / d scale
        const double jg = s * log(x) * js;            // d f / d gamma
        const double jo = -g;                         // d f / d offset
        const double r = g * (s * xg - o) - (double)t[i];
        v[0] += js * js; v[1] += js * jg; v[2] += js * jo; v[3] += jg * jg; v[4] += jg * jo; v[5] += jo * jo;
        v[6] += js * r;  v[7] += jg * r;  v[8] += jo * r;  v[9] += r * r;
    }
    block_accumulate<10>(v, moments + 10 * c);
}

// out = g_b (s x^gamma - o) (fp32), per-image sums of |out - target| and (out - target)^2 over C, HW
__global__ void __launch_bounds__(kEvalThreads)
eval_apply_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                  const double* __restrict__ gain, const double* __restrict__ params, int C, int64_t HW,
                  float* __restrict__ out, double* __restrict__ image_sums /* (B, 2) */) {
    const int bc = blockIdx.y, b = bc / C, c = bc - b * C;
    const float* p = pred + (int64_t)bc * HW;
    const float* t = target + (int64_t)bc * HW;
    float* q = out + (int64_t)bc * HW;
    const double* prm = params + 5 * c;
    const double g = gain[b], s = prm[2], gamma = prm[3], o = prm[4];
    double v[2] = {0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        const float y = (float)(g * (s * pow(corrected_input(p[i], prm), gamma) - o));
        q[i] = y;
        const double d = (double)y - (double)t[i];
        v[0] += fabs(d);
        v[1] += d * d;
    }
    block_accumulate<2>(v, image_sums + 2 * b);
}

static int eval_grid(int64_t HW) {
    const int64_t need = (HW + kEvalThreads * 4 - 1) / (kEvalThreads * 4);
    return (int)(need < 1 ? 1 : (need > 64 ? 64 : need));
}

}  // namespace den

extern "C" {

int den_eval_affine_moments(const float* pred, const float* target, const double* log_gain, int32_t B,
                            int32_t C, int64_t HW, double* moments, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(B >= 1 && C >= 1 && C <= 3 && HW >= 1, "bad image batch shape");
    DEN_CHECK_ARG(pred && target && log_gain && moments, "null pointer");
    dim3 grid(eval_grid(HW), B * C);
    eval_affine_moments_kernel<<<grid, kEvalThreads, 0, as_stream(stream)>>>(pred, target, log_gain, C, HW, moments);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_eval_lm_moments(const float* pred, const float* target, const double* gain, const double* params,
                        int32_t B, int32_t C, int64_t HW, double* moments, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(B >= 1 && C >= 1 && C <= 3 && HW >= 1, "bad image batch shape");
    DEN_CHECK_ARG(pred && target && gain && params && moments, "null pointer");
    dim3 grid(eval_grid(HW), B * C);
    eval_lm_moments_kernel<<<grid, kEvalThreads, 0, as_stream(stream)>>>(pred, target, gain, params, C, HW, moments);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_eval_apply(const float* pred, const float* target, const double* gain, const double* params,
                   int32_t B, int32_t C, int64_t HW, float* out, double* image_sums, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(B >= 1 && C >= 1 && C <= 3 && HW >= 1, "bad image batch shape");
    DEN_CHECK_ARG(pred && target && gain && params && out && image_sums, "null pointer");
    dim3 grid(eval_grid(HW), B * C);
    eval_apply_kernel<<<grid, kEvalThreads, 0, as_stream(stream)>>>(pred, target, gain, params, C, HW, out, image_sums);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
