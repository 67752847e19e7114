// Pixel-bandwidth low-pass filter: intensity samples -> band-limited log-intensity, forward and
// reverse mode, all internal arithmetic in fp64 (sm_100a).  S <= 32: one warp per event, one lane per
// interval (den_lpf_loss.cu); 32 < S <= 64: one thread per event (this file).
//
// Replaces PixelBandwidth.intensity_sample_to_weight / linearize_sys / linearized_sys_params /
// discretized_sys_to_weight / the normalisation half of weighted_it_sample_to_output_log_it
// (models/pixel_bandwidth.py:181-228,260-296,369-415) and control.foh_cont2discrete
// (utils/control.py:29-123).  The reference issues 7k-15k ATen ops per call (batched
// matrix_exp, two LU solves, a 28-iteration Python loop of bmm); here one launch
// per direction.
//
// Model.  Per event and per interval k between input samples k and k+1 the 4th-order system is
// linearised at I = I[k+1]:   x' = A x + B u,  state (y', y, s, d),
//     A = [[-a, -b, 0, 0], [1, 0, 0, 0], [0, w_sf, -w_sf, 0], [0, 0, w_d, -w_d]],  B = (b,0,0,0)^T
//     a = alpha0 + alpha1 I   (2 zeta w_n),        b = beta I   (w_n^2)
// (the reference's tau_in = P_in / I etc. collapse to these affine forms; the five
// coefficients alpha0, alpha1, beta, w_sf, w_d are computed by the host module from the six
// softplus-parametrised parameters, so their gradients flow back through torch autograd).
// FOH discretisation with the state preserved (utils/control.py:87-93,109-113):
//     Phi = expm(A dt), G1 = (Phi - I) A^-1 B, G2 = (A dt)^-1 G1 - A^-1 B, Bd = G1 - G2, Bt = G2.
// Every stage has unity DC gain, so A^-1 B = -(0,1,1,1)^T =: -e exactly, hence
//     G1 = e - Phi e,   G2 = A^-1 G1 / dt + e                       (no LU, two closed-form solves).
// The matrix is balanced with x0 -> x0 / w_n (entries become the four rates a, w_n, w_sf, w_d;
// ||A dt|| <= ~2e3 instead of 4e7) and exponentiated by scaling-and-squaring with a degree-12
// Taylor polynomial.  The sample weights follow the reference's recursion
//     w[i] = r_{i+1} Bd[i] + r_i Bt[i-1],   r_j = C Phi[S-2] ... Phi[j]
// in one descending sweep, then out = sum (w / sum w) log I.
// Reverse mode: the adjoint of expm is the Frechet derivative at the transposed argument,
// L(M^T, Gbar), evaluated with the block-triangular pair recurrence (no stored intermediates).
#include "den_lpf.cuh"

namespace den {

// one descending sweep over the intervals; r_hist (optional) records r_{j+1} per interval
template <bool kRecord>
__device__ void sweep_weights(const float* __restrict__ intensity, const float* __restrict__ dt_ns,
                              const double* coef, int S, int64_t N, int64_t n, int nc,
                              double* w /*[S][2]*/, double* r_hist /*[S][2][4]*/) {
    for (int i = 0; i < S * 2; ++i) w[i] = 0.0;
    double r[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    if (nc == 2) { r[0][2] = 1.0; r[1][3] = 1.0; } else { r[0][3] = 1.0; }
    Interval iv;
    for (int j = S - 2; j >= 0; --j) {
        discretize((double)intensity[(int64_t)(j + 1) * N + n], 1e-9 * (double)dt_ns[(int64_t)j * N + n],
                   coef, iv);
        for (int c = 0; c < nc; ++c) {
            if (kRecord)
                for (int k = 0; k < 4; ++k) r_hist[(j * 2 + c) * 4 + k] = r[c][k];
            double rp[4] = {r[c][0] * iv.wn, r[c][1], r[c][2], r[c][3]};
            double wt = 0.0, wd = 0.0, t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                wt = fma(rp[k], iv.bt[k], wt);
                wd = fma(rp[k], iv.bd[k], wd);
            }
            w[(j + 1) * 2 + c] += wt;
            w[j * 2 + c] += wd;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                t[k] = rp[0] * iv.phi[k] + rp[1] * iv.phi[4 + k] + rp[2] * iv.phi[8 + k] + rp[3] * iv.phi[12 + k];
            r[c][0] = t[0] / iv.wn;
            r[c][1] = t[1];
            r[c][2] = t[2];
            r[c][3] = t[3];
        }
    }
}

__global__ void __launch_bounds__(kLpfThreads)
lpf_fwd_kernel(const float* __restrict__ intensity, const float* __restrict__ dt_ns,
               const double* __restrict__ coef_dev, int S, int64_t N, int nc, float* __restrict__ out) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= N) return;
    double coef[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) coef[i] = coef_dev[i];
    double w[kLpfMaxS * 2];
    sweep_weights<false>(intensity, dt_ns, coef, S, N, n, nc, w, nullptr);
    for (int c = 0; c < nc; ++c) {
        double W = 0.0, acc = 0.0;
        for (int i = 0; i < S; ++i) {
            W += w[i * 2 + c];
            acc = fma(w[i * 2 + c], log((double)intensity[(int64_t)i * N + n]), acc);
        }
        out[n * nc + c] = (float)(acc / W);
    }
}

__global__ void __launch_bounds__(kLpfThreads)
lpf_bwd_kernel(const float* __restrict__ intensity, const float* __restrict__ dt_ns,
               const double* __restrict__ coef_dev, int S, int64_t N, int nc,
               const float* __restrict__ d_out, float* __restrict__ d_intensity,
               double* __restrict__ d_coef) {
    __shared__ double s_red[5];
    if (threadIdx.x < 5) s_red[threadIdx.x] = 0.0;
    __syncthreads();
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double dcoef[5] = {0, 0, 0, 0, 0};
    if (n < N) {
        double coef[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) coef[i] = coef_dev[i];
        double w[kLpfMaxS * 2], wbar[kLpfMaxS * 2], r_hist[kLpfMaxS * 8], dI[kLpfMaxS];
        sweep_weights<true>(intensity, dt_ns, coef, S, N, n, nc, w, r_hist);
        for (int i = 0; i < S; ++i) dI[i] = 0.0;
        for (int i = 0; i < S * 2; ++i) wbar[i] = 0.0;
        // out_c = sum_i (w_ic / W_c) L_i
        for (int c = 0; c < nc; ++c) {
            const double go = (double)d_out[n * nc + c];
            double W = 0.0, acc = 0.0;
            for (int i = 0; i < S; ++i) {
                W += w[i * 2 + c];
                acc = fma(w[i * 2 + c], log((double)intensity[(int64_t)i * N + n]), acc);
            }
            const double o = acc / W;
            for (int i = 0; i < S; ++i) {
                const double Ii = (double)intensity[(int64_t)i * N + n];
                wbar[i * 2 + c] = go * (log(Ii) - o) / W;
                dI[i] += go * (w[i * 2 + c] / W) / Ii;
            }
        }
        // adjoint sweep, ascending j: rbar holds the adjoint of r_j
        double rbar[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        Interval iv;
        for (int j = 0; j <= S - 2; ++j) {
            const double Ij = (double)intensity[(int64_t)(j + 1) * N + n];
            discretize(Ij, 1e-9 * (double)dt_ns[(int64_t)j * N + n], coef, iv);
            double phibar[16], bdbar[4] = {0, 0, 0, 0}, btbar[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 16; ++i) phibar[i] = 0.0;
            double wnbar = 0.0;
            for (int c = 0; c < nc; ++c) {
                const double* r = &r_hist[(j * 2 + c) * 4];           // r_{j+1}
                const double rp[4] = {r[0] * iv.wn, r[1], r[2], r[3]};
                double t0 = rp[0] * iv.phi[0] + rp[1] * iv.phi[4] + rp[2] * iv.phi[8] + rp[3] * iv.phi[12];
                // r_j = (t0 / wn, t1, t2, t3)
                const double tb[4] = {rbar[c][0] / iv.wn, rbar[c][1], rbar[c][2], rbar[c][3]};
                wnbar -= rbar[c][0] * t0 / (iv.wn * iv.wn);
                const double gw_t = wbar[(j + 1) * 2 + c], gw_d = wbar[j * 2 + c];
                double rpbar[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    rpbar[k] = iv.phi[4 * k] * tb[0] + iv.phi[4 * k + 1] * tb[1] + iv.phi[4 * k + 2] * tb[2] +
                               iv.phi[4 * k + 3] * tb[3] + gw_t * iv.bt[k] + gw_d * iv.bd[k];
                    btbar[k] += gw_t * rp[k];
                    bdbar[k] += gw_d * rp[k];
#pragma unroll
                    for (int q = 0; q < 4; ++q) phibar[4 * k + q] += rp[k] * tb[q];
                }
                wnbar += rpbar[0] * r[0];
                rbar[c][0] = rpbar[0] * iv.wn;
                rbar[c][1] = rpbar[1];
                rbar[c][2] = rpbar[2];
                rbar[c][3] = rpbar[3];
            }
            double abar, bbar, wsfbar, wdbar;
            interval_adjoint(iv, phibar, bdbar, btbar, wnbar, abar, bbar, wsfbar, wdbar);
            dI[j + 1] += abar * coef[1] + bbar * coef[2];
            dcoef[0] += abar;
            dcoef[1] += abar * Ij;
            dcoef[2] += bbar * Ij;
            dcoef[3] += wsfbar;
            dcoef[4] += wdbar;
        }
        for (int i = 0; i < S; ++i) d_intensity[(int64_t)i * N + n] = (float)dI[i];
    }
    if (d_coef) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double v = dcoef[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if ((threadIdx.x & 31) == 0) atomicAdd(&s_red[k], v);
        }
        __syncthreads();
        if (threadIdx.x < 5) atomicAdd(d_coef + threadIdx.x, s_red[threadIdx.x]);
    }
}

}  // namespace den

extern "C" {

int den_lpf_fwd(const float* intensity, const float* sample_dt_ns, const double* coef, int32_t S,
                int64_t N, int32_t n_channels, float* out, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(S >= 2 && S <= kLpfMaxS, "it_sample_size must be in [2, 64]");
    DEN_CHECK_ARG(n_channels == 1 || n_channels == 2, "n_channels must be 1 or 2");
    DEN_CHECK_ARG(N >= 0, "negative event count");
    if (N == 0) return DEN_OK;
    DEN_CHECK_ARG(intensity && sample_dt_ns && coef && out, "null pointer");
    if (S <= 32) {          // one warp per event, one lane per interval
        lpf_warp_fwd(intensity, sample_dt_ns, coef, S, N, n_channels, out, as_stream(stream));
        DEN_CHECK_LAUNCH();
        return DEN_OK;
    }
    const unsigned grid = (unsigned)((N + kLpfThreads - 1) / kLpfThreads);
    lpf_fwd_kernel<<<grid, kLpfThreads, 0, as_stream(stream)>>>(intensity, sample_dt_ns, coef, S, N,
                                                               n_channels, out);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_lpf_bwd(const float* intensity, const float* sample_dt_ns, const double* coef, int32_t S,
                int64_t N, int32_t n_channels, const float* d_out, float* d_intensity, double* d_coef,
                void* stream) {
    using namespace den;
    DEN_CHECK_ARG(S >= 2 && S <= kLpfMaxS, "it_sample_size must be in [2, 64]");
    DEN_CHECK_ARG(n_channels == 1 || n_channels == 2, "n_channels must be 1 or 2");
    DEN_CHECK_ARG(N >= 0, "negative event count");
    if (N == 0) return DEN_OK;
    DEN_CHECK_ARG(intensity && sample_dt_ns && coef && d_out && d_intensity, "null pointer");
    if (S <= 32) {
        lpf_warp_bwd(intensity, sample_dt_ns, coef, S, N, n_channels, d_out, d_intensity, d_coef,
                     as_stream(stream));
        DEN_CHECK_LAUNCH();
        return DEN_OK;
    }
    const unsigned grid = (unsigned)((N + kLpfThreads - 1) / kLpfThreads);
    lpf_bwd_kernel<<<grid, kLpfThreads, 0, as_stream(stream)>>>(intensity, sample_dt_ns, coef, S, N,
                                                               n_channels, d_out, d_intensity, d_coef);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
