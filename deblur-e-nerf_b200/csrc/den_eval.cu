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

// ------------------------------------------------------------------------------------ SSIM --
// Metric.compute's SSIM term (loss_metric/metric.py:74-81): torchmetrics 0.6.2
// `functional.ssim(preds, target, data_range = max_target_val)` = an 11 x 11 Gaussian window
// (sigma 1.5, weights gauss / sum(gauss), 2-D window = outer product), windowed means of p, t, pp, tt,
// pt, the SSIM index per pixel, and — after its reflect padding — a crop of (k - 1) / 2 pixels per side
// before the mean: only pixels whose whole window lies inside the image contribute, so the padding never
// reaches the result and this kernel needs none.  One CTA = a 32 x 8 tile of window centres; the
// (8 + k - 1) x (32 + k - 1) patches of both images are staged in shared memory once and every
// thread runs the k x k window from there (fp32 like the conv2d it replaces, fp64 sum of the indices).
constexpr int kSsimTileW = 32, kSsimTileH = 8, kSsimMaxK = 15;

__global__ void __launch_bounds__(kEvalThreads)
eval_ssim_kernel(const float* __restrict__ pred, const float* __restrict__ target, int H, int W, int K,
                 float sigma, float c1, float c2, int C, double* __restrict__ image_sums /* (B) */) {
    extern __shared__ float s_ssim[];
    __shared__ float s_w[kSsimMaxK];
    const int PW = kSsimTileW + K - 1, PH = kSsimTileH + K - 1;
    float* sp = s_ssim;
    float* st = s_ssim + PW * PH;
    const int bc = blockIdx.z, b = bc / C;
    const float* p = pred + (int64_t)bc * H * W;
    const float* t = target + (int64_t)bc * H * W;
    const int x0 = blockIdx.x * kSsimTileW, y0 = blockIdx.y * kSsimTileH;   // top-left of the patch
    const int tid = threadIdx.x, tx = tid % kSsimTileW, ty = tid / kSsimTileW;   // 1-D CTA (block_accumulate)
    if (tid < K) {
        const float dist = (float)tid + 0.5f * (float)(1 - K);               // arange((1-K)/2, (1+K)/2)
        const float q = dist / sigma;
        s_w[tid] = expf(-(q * q) / 2.f);
    }
    for (int i = tid; i < PW * PH; i += kSsimTileW * kSsimTileH) {
        const int py = i / PW, px = i - py * PW;
        const int y = y0 + py, x = x0 + px;
        const bool in = y < H && x < W;
        sp[i] = in ? __ldg(p + (int64_t)y * W + x) : 0.f;
        st[i] = in ? __ldg(t + (int64_t)y * W + x) : 0.f;
    }
    __syncthreads();
    float wsum = 0.f;
    for (int k = 0; k < K; ++k) wsum += s_w[k];
    double v[1] = {0.0};
    // window centre (y0 + ty + r, x0 + tx + r), r = (K - 1) / 2: inside the cropped region iff the window fits
    if (y0 + ty + K <= H && x0 + tx + K <= W) {
        float mp = 0.f, mt = 0.f, spp = 0.f, stt = 0.f, spt = 0.f;
        for (int i = 0; i < K; ++i) {
            const float wi = s_w[i] / wsum;
            const float* rp = sp + (ty + i) * PW + tx;
            const float* rt = st + (ty + i) * PW + tx;
            for (int j = 0; j < K; ++j) {
                const float w = wi * (s_w[j] / wsum);
                const float a = rp[j], c = rt[j];
                mp = fmaf(w, a, mp);
                mt = fmaf(w, c, mt);
                spp = fmaf(w, a * a, spp);
                stt = fmaf(w, c * c, stt);
                spt = fmaf(w, a * c, spt);
            }
        }
        const float mpp = mp * mp, mtt = mt * mt, mpt = mp * mt;
        const float upper = 2.f * (spt - mpt) + c2;
        const float lower = (spp - mpp) + (stt - mtt) + c2;
        v[0] = (double)(((2.f * mpt + c1) * upper) / ((mpp + mtt + c1) * lower));
    }
    block_accumulate<1>(v, image_sums + b);
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

int den_eval_ssim(const float* pred, const float* target, int32_t B, int32_t C, int32_t H, int32_t W,
                  int32_t kernel_size, double sigma, double c1, double c2, double* image_sums, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(B >= 1 && C >= 1 && C <= 3, "bad image batch shape");
    DEN_CHECK_ARG(kernel_size >= 1 && kernel_size <= kSsimMaxK && (kernel_size & 1), "ssim: the window size must be odd and <= 15");
    DEN_CHECK_ARG(sigma > 0.0, "ssim: sigma must be positive");
    DEN_CHECK_ARG(H >= kernel_size && W >= kernel_size, "ssim: the image is smaller than the window");
    DEN_CHECK_ARG(pred && target && image_sums, "null pointer");
    const int nx = (W - kernel_size + 1 + kSsimTileW - 1) / kSsimTileW;
    const int ny = (H - kernel_size + 1 + kSsimTileH - 1) / kSsimTileH;
    DEN_CHECK_ARG(ny <= 65535 && (int64_t)B * C <= 65535, "ssim: image batch too large for one launch");
    static_assert(kSsimTileW * kSsimTileH == kEvalThreads, "one thread per window centre");
    dim3 grid(nx, ny, B * C);
    const size_t smem = 2 * (size_t)(kSsimTileW + kernel_size - 1) * (kSsimTileH + kernel_size - 1) * sizeof(float);
    eval_ssim_kernel<<<grid, kEvalThreads, smem, as_stream(stream)>>>(pred, target, H, W, kernel_size, (float)sigma,
                                                               (float)c1, (float)c2, C, image_sums);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
