// Pixel-bandwidth filter FUSED with the event-loss reduction, all render requests of a training step
// in one launch (SURVEY.md §8(b) family `lpf_loss`, rows A15-A18).
//
// Replaces, per step: the 2K calls of PixelBandwidth.forward's filter half
// (models/pixel_bandwidth.py:369-448: linearise, FOH-discretise, weight recursion, normalised sum,
// differencing-amplifier reset carried from the first request to the others) and Loss.compute
// (loss_metric/loss.py:34-96: per-event Huber / L1 / MSE / MAPE on the normalised log-intensity
// difference of each request pair, masked means), forward and reverse mode.
//
// Mapping.  One CTA per event, one WARP per render request, one LANE per discretisation interval
// (S - 1 <= 31 intervals; it_sample_size is 30 in every shipped config): the expensive, mutually
// independent part — expm of the balanced 4x4 system matrix and, in reverse mode, its Frechet
// adjoint — runs on all lanes at once, and only the short weight recursion r_j = r_{j+1} Phi_j
// (a 1x4 by 4x4 product per interval) is serial, passed from lane to lane with shuffles.  The math
// of an interval is den_lpf.cu's (shared through den_lpf.cuh), fp64 throughout.  The per-event loss
// terms go to a (P, N) buffer and the LAST CTA to finish (ticket counter) sums them in index order:
// the masked means are deterministic and the host never sees an intermediate.
#include "den_lpf.cuh"

namespace den {
namespace lpf {

constexpr int kMaxRequests = 8;
constexpr int kMaxPairs = kMaxRequests / 2;

enum ErrKind { kErrL1 = 0, kErrMse = 1, kErrHuber = 2, kErrMape = 3 };

__device__ __forceinline__ double err_value(int kind, double x, double t) {
    const double d = x - t, ad = fabs(d);
    switch (kind) {
        case kErrL1: return ad;
        case kErrMse: return d * d;
        case kErrHuber: return ad < 1.0 ? 0.5 * d * d : ad - 0.5;
        default: return ad / fmax(fabs(t), 2.220446049250313e-16);
    }
}
__device__ __forceinline__ double sgn(double v) { return (v > 0.0) - (v < 0.0); }
// d err / d x and d err / d t
__device__ __forceinline__ void err_grad(int kind, double x, double t, double& gx, double& gt) {
    const double d = x - t, ad = fabs(d);
    switch (kind) {
        case kErrL1: gx = sgn(d); gt = -gx; return;
        case kErrMse: gx = 2.0 * d; gt = -gx; return;
        case kErrHuber: gx = ad < 1.0 ? d : sgn(d); gt = -gx; return;
        default: {
            const double eps = 2.220446049250313e-16, at = fabs(t);
            const double den = fmax(at, eps);
            gx = sgn(d) / den;
            gt = -gx - (at > eps ? ad * sgn(t) / (den * den) : 0.0);
            return;
        }
    }
}

struct WarpState {
    Interval iv;            // this lane's interval (lane j: samples j and j+1)
    double rh[2][4];        // r_{j+1} per output channel, recorded during the descending sweep
    double w[2];            // weight of sample `lane` per channel
    double W[2], out[2];    // warp-uniform: weight sums and filter outputs
    double I, L;            // intensity / log-intensity of sample `lane`
    double I_next;          // intensity of sample lane + 1 (linearisation point of the interval)
    bool has_interval, has_sample;
};

// forward of one request on one warp: lane i holds sample i and interval i
__device__ __forceinline__ void warp_forward(const float* __restrict__ intensity /* (S,N) of the request */,
                                             const float* __restrict__ dt_ns /* (S-1,N) */,
                                             const double* coef, int S, int64_t N, int64_t n, int nc,
                                             int lane, WarpState& st) {
    const unsigned full = 0xffffffffu;
    st.has_sample = lane < S;
    st.has_interval = lane < S - 1;
    st.I = st.has_sample ? (double)__ldg(intensity + (int64_t)lane * N + n) : 1.0;
    st.L = st.has_sample ? log(st.I) : 0.0;
    st.I_next = __shfl_down_sync(full, st.I, 1);
    if (st.has_interval) {
        discretize(st.I_next, 1e-9 * (double)__ldg(dt_ns + (int64_t)lane * N + n), coef, st.iv);
    } else {
        st.iv.a = st.iv.wn = st.iv.wsf = st.iv.wd = st.iv.dt = 1.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) st.iv.phi[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) st.iv.u[i] = st.iv.bd[i] = st.iv.bt[i] = 0.0;
    }
    double r[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    if (nc == 2) { r[0][2] = 1.0; r[1][3] = 1.0; } else { r[0][3] = 1.0; }
    double wt[2] = {0, 0}, wdv[2] = {0, 0};
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) st.rh[c][k] = 0.0;
    const Interval& iv = st.iv;
    for (int j = S - 2; j >= 0; --j) {
        for (int c = 0; c < nc; ++c) {
            const double rp[4] = {r[c][0] * iv.wn, r[c][1], r[c][2], r[c][3]};
            double a_t = 0.0, a_d = 0.0, t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                a_t = fma(rp[k], iv.bt[k], a_t);
                a_d = fma(rp[k], iv.bd[k], a_d);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                t[k] = rp[0] * iv.phi[k] + rp[1] * iv.phi[4 + k] + rp[2] * iv.phi[8 + k] + rp[3] * iv.phi[12 + k];
            if (lane == j) {
#pragma unroll
                for (int k = 0; k < 4; ++k) st.rh[c][k] = r[c][k];
                wt[c] = a_t;
                wdv[c] = a_d;
            }
            t[0] /= iv.wn;
#pragma unroll
            for (int k = 0; k < 4; ++k) r[c][k] = __shfl_sync(full, t[k], j);
        }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double from_prev = __shfl_up_sync(full, wt[c], 1);      // Bt of interval lane-1 weighs sample lane
        double w = 0.0;
        if (c < nc && st.has_sample) w = (st.has_interval ? wdv[c] : 0.0) + (lane >= 1 ? from_prev : 0.0);
        st.w[c] = w;
        double W = w, acc = w * st.L;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            W += __shfl_xor_sync(full, W, d);
            acc += __shfl_xor_sync(full, acc, d);
        }
        st.W[c] = W;
        st.out[c] = c < nc ? acc / W : 0.0;
    }
}

// reverse mode of warp_forward: go[c] = dL/d out[c].  Returns this lane's dL/dI(sample lane) and adds
// the lane's coefficient gradients to dcoef[5].
__device__ __forceinline__ double warp_backward(const WarpState& st, const double* coef, int S, int nc,
                                                int lane, const double go[2], double dcoef[5]) {
    const unsigned full = 0xffffffffu;
    const Interval& iv = st.iv;
    double wbar[2] = {0, 0}, gw_t[2], dI = 0.0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c < nc && st.has_sample) {
            wbar[c] = go[c] * (st.L - st.out[c]) / st.W[c];
            dI += go[c] * (st.w[c] / st.W[c]) / st.I;
        }
        gw_t[c] = __shfl_down_sync(full, wbar[c], 1);        // adjoint of the weight of sample lane + 1
    }
    double rbar[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    double phibar[16], bdbar[4] = {0, 0, 0, 0}, btbar[4] = {0, 0, 0, 0}, wnbar = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) phibar[i] = 0.0;
    for (int j = 0; j <= S - 2; ++j) {
        for (int c = 0; c < nc; ++c) {
            const double* r = st.rh[c];
            const double rp[4] = {r[0] * iv.wn, r[1], r[2], r[3]};
            const double t0 = rp[0] * iv.phi[0] + rp[1] * iv.phi[4] + rp[2] * iv.phi[8] + rp[3] * iv.phi[12];
            const double tb[4] = {rbar[c][0] / iv.wn, rbar[c][1], rbar[c][2], rbar[c][3]};
            double rpbar[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                rpbar[k] = iv.phi[4 * k] * tb[0] + iv.phi[4 * k + 1] * tb[1] + iv.phi[4 * k + 2] * tb[2] +
                           iv.phi[4 * k + 3] * tb[3] + gw_t[c] * iv.bt[k] + wbar[c] * iv.bd[k];
            if (lane == j) {
                wnbar += rpbar[0] * r[0] - rbar[c][0] * t0 / (iv.wn * iv.wn);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    btbar[k] += gw_t[c] * rp[k];
                    bdbar[k] += wbar[c] * rp[k];
#pragma unroll
                    for (int q = 0; q < 4; ++q) phibar[4 * k + q] += rp[k] * tb[q];
                }
            }
            rpbar[0] *= iv.wn;
#pragma unroll
            for (int k = 0; k < 4; ++k) rbar[c][k] = __shfl_sync(full, rpbar[k], j);
        }
    }
    double to_next = 0.0;       // dL/dI of sample lane + 1 through this lane's linearisation
    if (st.has_interval) {
        double abar, bbar, wsfbar, wdbar;
        interval_adjoint(iv, phibar, bdbar, btbar, wnbar, abar, bbar, wsfbar, wdbar);
        to_next = abar * coef[1] + bbar * coef[2];
        dcoef[0] += abar;
        dcoef[1] += abar * st.I_next;
        dcoef[2] += bbar * st.I_next;
        dcoef[3] += wsfbar;
        dcoef[4] += wdbar;
    }
    const double from_prev = __shfl_up_sync(full, to_next, 1);
    if (lane >= 1 && st.has_sample) dI += from_prev;
    return dI;
}

struct LossArgs {
    const float* intensity;         // (K, S, N)
    const float* dt_ns;             // (K, S-1, N)
    const double* coef;             // 5, device
    const double* reset_dt_ns;      // (K, N) output_ts[k] - output_ts[0]; row 0 unused; NULL without reset
    const float* target;            // (P, N) or NULL (all zero)
    const uint8_t* target_present;  // host-side flags folded into has_target[p]
    const float* inv_k;             // (P) device: 1 / normaliser of pair p
    const uint8_t* valid;           // (P, N)
    int S, K, P, has_reset;
    int64_t N;
    int kind[kMaxPairs];
    int has_target[kMaxPairs];
};

struct Epilogue {                   // per-event scalars every thread of the CTA can derive from smem
    double final_[kMaxRequests];
    double decay[kMaxRequests];
    double delta;
};

__device__ __forceinline__ void event_epilogue(const LossArgs& a, const double (*outs)[2], int64_t n, Epilogue& e) {
    e.delta = a.has_reset ? outs[0][1] - outs[0][0] : 0.0;
    e.final_[0] = outs[0][0];
    e.decay[0] = 0.0;
    for (int k = 1; k < a.K; ++k) {
        e.decay[k] = a.has_reset ? exp(-a.coef[4] * (1e-9 * a.reset_dt_ns[(int64_t)k * a.N + n])) : 0.0;
        e.final_[k] = outs[k][0] - e.delta * e.decay[k];
    }
}

// workspace layout (doubles): err (P*N) | ticket (1, as uint32)
// kMaxK: the largest number of render requests (warps) a launch of this instance takes.  A training step
// has K <= 4: with 128 threads per CTA the reverse pass is compiled for THREE CTAs per SM (168 registers,
// ~0.5 KB of spills per thread) instead of the two its 250 registers allow — it is bound by the latency of
// its fp64 chains at 8 resident warps per SM, not by the fp64 pipe.
template <bool kBackward, int kMaxK>
__global__ void __launch_bounds__(32 * kMaxK, (kBackward && kMaxK <= 4) ? 3 : 1)
lpf_loss_kernel(const LossArgs a, double* __restrict__ err_buf, unsigned* __restrict__ ticket,
                float* __restrict__ terms /* (P) */, int32_t* __restrict__ counts /* (P) */,
                float* __restrict__ log_intensity /* (K, N) filter outputs after the reset, may be NULL */,
                // backward only
                const float* __restrict__ d_terms /* (P) */, float* __restrict__ d_intensity /* (K,S,N) */,
                double* __restrict__ d_coef /* 5 */, double* __restrict__ d_reset_dt /* (K,N) */,
                float* __restrict__ d_target /* (P,N) */, double* __restrict__ d_inv_k /* (P) */) {
    __shared__ double s_out[kMaxRequests][2];
    __shared__ double s_red[5 + kMaxPairs];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = blockIdx.x;
    const int nc = (warp == 0 && a.has_reset) ? 2 : 1;
    double coef[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) coef[i] = a.coef[i];
    if (threadIdx.x < 5 + kMaxPairs) s_red[threadIdx.x] = 0.0;

    WarpState st;
    warp_forward(a.intensity + (int64_t)warp * a.S * a.N, a.dt_ns + (int64_t)warp * (a.S - 1) * a.N, coef, a.S,
                 a.N, n, nc, lane, st);
    if (lane == 0) {
        s_out[warp][0] = st.out[0];
        s_out[warp][1] = st.out[1];
    }
    __syncthreads();
    Epilogue e;
    event_epilogue(a, s_out, n, e);

    if (!kBackward) {
        if (threadIdx.x < a.K && log_intensity) log_intensity[(int64_t)threadIdx.x * a.N + n] = (float)e.final_[threadIdx.x];
        if (threadIdx.x < a.P) {
            const int p = threadIdx.x;
            const double x = (e.final_[2 * p + 1] - e.final_[2 * p]) * (double)a.inv_k[p];
            const double t = a.has_target[p] ? (double)a.target[(int64_t)p * a.N + n] : 0.0;
            const bool ok = a.valid[(int64_t)p * a.N + n] != 0;
            // masked-out events contribute exactly zero whatever their (possibly non-finite) error is
            err_buf[(int64_t)p * a.N + n] = ok ? err_value(a.kind[p], x, t) : 0.0;
        }
        // the last CTA to arrive reduces the per-event terms in index order (deterministic means)
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        __shared__ double s_sum[32 * kMaxRequests];
        __shared__ int s_cnt[32 * kMaxRequests];
        for (int p = 0; p < a.P; ++p) {
            double acc = 0.0;
            int cnt = 0;
            for (int64_t i = threadIdx.x; i < a.N; i += blockDim.x) {
                acc += __ldcg(err_buf + (int64_t)p * a.N + i);
                cnt += a.valid[(int64_t)p * a.N + i] != 0;
            }
            s_sum[threadIdx.x] = acc;
            s_cnt[threadIdx.x] = cnt;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                int c = 0;
                for (int i = 0; i < (int)blockDim.x; ++i) { tot += s_sum[i]; c += s_cnt[i]; }
                terms[p] = (float)(tot / (double)c);        // 0 / 0 = NaN like the mean of an empty selection
                counts[p] = c;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) *ticket = 0;                  // ready for the next launch
        return;
    } else {
        // ---- d terms -> d final -> d filter outputs ------------------------------------------
        double dfinal[kMaxRequests];
        for (int k = 0; k < a.K; ++k) dfinal[k] = 0.0;
        for (int p = 0; p < a.P; ++p) {
            const bool ok = a.valid[(int64_t)p * a.N + n] != 0;
            double gx = 0.0, gt = 0.0;
            const double pred = e.final_[2 * p + 1] - e.final_[2 * p];
            if (ok) {
                const double t = a.has_target[p] ? (double)a.target[(int64_t)p * a.N + n] : 0.0;
                err_grad(a.kind[p], pred * (double)a.inv_k[p], t, gx, gt);
                const double scale = (double)d_terms[p] / (double)counts[p];
                gx *= scale;
                gt *= scale;
            }
            dfinal[2 * p + 1] += gx * (double)a.inv_k[p];
            dfinal[2 * p] -= gx * (double)a.inv_k[p];
            if (threadIdx.x == 0) {
                if (d_target && a.has_target[p]) d_target[(int64_t)p * a.N + n] = (float)gt;
                if (d_inv_k) atomicAdd(&s_red[5 + p], gx * pred);
            }
        }
        // final_k = out_k - delta decay_k (k >= 1); final_0 = sf; delta = before_0 - sf_0
        double ddelta = 0.0, dwd_reset = 0.0;
        for (int k = 1; k < a.K; ++k) {
            if (!a.has_reset) break;
            ddelta -= dfinal[k] * e.decay[k];
            const double ddecay = -dfinal[k] * e.delta;
            const double rdt = 1e-9 * a.reset_dt_ns[(int64_t)k * a.N + n];
            dwd_reset += ddecay * e.decay[k] * (-rdt);
            if (threadIdx.x == 0 && d_reset_dt)
                d_reset_dt[(int64_t)k * a.N + n] = ddecay * e.decay[k] * (-coef[4] * 1e-9);
        }
        if (threadIdx.x == 0 && d_reset_dt) d_reset_dt[n] = 0.0;
        double go[2];
        if (warp == 0) {
            go[0] = dfinal[0] - ddelta;         // source-follower output (or the only output without reset)
            go[1] = ddelta;                     // differencing-amplifier output before the reset
        } else {
            go[0] = dfinal[warp];
            go[1] = 0.0;
        }
        double dcoef[5] = {0, 0, 0, 0, 0};
        const double dI = warp_backward(st, coef, a.S, nc, lane, go, dcoef);
        if (st.has_sample) d_intensity[((int64_t)warp * a.S + lane) * a.N + n] = (float)dI;
        if (d_coef) {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                double v = dcoef[k];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                if (lane == 0) atomicAdd(&s_red[k], v);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && d_coef) atomicAdd(&s_red[4], dwd_reset);
        __syncthreads();
        if (threadIdx.x < 5 && d_coef) atomicAdd(d_coef + threadIdx.x, s_red[threadIdx.x]);
        if (threadIdx.x >= 5 && threadIdx.x < 5 + a.P && d_inv_k)
            atomicAdd(d_inv_k + (threadIdx.x - 5), s_red[threadIdx.x]);
    }
}

// ---- warp-per-event form of the plain filter (den_lpf_fwd / den_lpf_bwd for S <= 32) ---------------
constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kWarpsPerCta)
lpf_warp_fwd_kernel(const float* __restrict__ intensity, const float* __restrict__ dt_ns,
                    const double* __restrict__ coef_dev, int S, int64_t N, int nc, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t n = blockIdx.x * (int64_t)kWarpsPerCta + (threadIdx.x >> 5);
    if (n >= N) return;
    double coef[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) coef[i] = coef_dev[i];
    WarpState st;
    warp_forward(intensity, dt_ns, coef, S, N, n, nc, lane, st);
    if (lane < nc) out[n * nc + lane] = (float)(lane == 0 ? st.out[0] : st.out[1]);
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
lpf_warp_bwd_kernel(const float* __restrict__ intensity, const float* __restrict__ dt_ns,
                    const double* __restrict__ coef_dev, int S, int64_t N, int nc,
                    const float* __restrict__ d_out, float* __restrict__ d_intensity,
                    double* __restrict__ d_coef) {
    __shared__ double s_red[5];
    if (threadIdx.x < 5) s_red[threadIdx.x] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t n = blockIdx.x * (int64_t)kWarpsPerCta + (threadIdx.x >> 5);
    double dcoef[5] = {0, 0, 0, 0, 0};
    if (n < N) {
        double coef[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) coef[i] = coef_dev[i];
        WarpState st;
        warp_forward(intensity, dt_ns, coef, S, N, n, nc, lane, st);
        const double go[2] = {(double)d_out[n * nc], nc == 2 ? (double)d_out[n * nc + 1] : 0.0};
        const double dI = warp_backward(st, coef, S, nc, lane, go, dcoef);
        if (st.has_sample) d_intensity[(int64_t)lane * N + n] = (float)dI;
    }
    if (d_coef) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double v = dcoef[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) atomicAdd(&s_red[k], v);
        }
        __syncthreads();
        if (threadIdx.x < 5) atomicAdd(d_coef + threadIdx.x, s_red[threadIdx.x]);
    }
}

}  // namespace lpf

int lpf_warp_fwd(const float* intensity, const float* dt_ns, const double* coef, int S, int64_t N, int nc,
                 float* out, cudaStream_t stream) {
    const unsigned grid = (unsigned)((N + lpf::kWarpsPerCta - 1) / lpf::kWarpsPerCta);
    lpf::lpf_warp_fwd_kernel<<<grid, 32 * lpf::kWarpsPerCta, 0, stream>>>(intensity, dt_ns, coef, S, N, nc, out);
    return 0;
}
int lpf_warp_bwd(const float* intensity, const float* dt_ns, const double* coef, int S, int64_t N, int nc,
                 const float* d_out, float* d_intensity, double* d_coef, cudaStream_t stream) {
    const unsigned grid = (unsigned)((N + lpf::kWarpsPerCta - 1) / lpf::kWarpsPerCta);
    lpf::lpf_warp_bwd_kernel<<<grid, 32 * lpf::kWarpsPerCta, 0, stream>>>(intensity, dt_ns, coef, S, N, nc, d_out,
                                                                           d_intensity, d_coef);
    return 0;
}

}  // namespace den

extern "C" {

size_t den_lpf_loss_workspace_bytes(int32_t n_pairs, int64_t N) {
    return (size_t)(n_pairs > 0 ? n_pairs : 0) * (size_t)(N > 0 ? N : 0) * sizeof(double) + 16;
}

static int fill_args(den::lpf::LossArgs& a, const den_lpf_loss_desc* d, const float* intensity,
                     const float* dt, const double* coef, const double* reset_dt, const float* target,
                     const float* inv_k, const uint8_t* valid, int64_t N) {
    using namespace den;
    DEN_CHECK_ARG(d != nullptr, "null descriptor");
    DEN_CHECK_ARG(d->it_sample_size >= 2 && d->it_sample_size <= 32, "it_sample_size must be in [2, 32]");
    DEN_CHECK_ARG(d->n_requests >= 2 && d->n_requests <= lpf::kMaxRequests && d->n_requests % 2 == 0,
                  "n_requests must be an even number in [2, 8]");
    DEN_CHECK_ARG(N >= 1, "no events");
    DEN_CHECK_ARG(intensity && dt && coef && inv_k && valid, "null pointer");
    DEN_CHECK_ARG(!d->has_reset || reset_dt, "the reset needs reset_dt_ns");
    a.intensity = intensity;
    a.dt_ns = dt;
    a.coef = coef;
    a.reset_dt_ns = reset_dt;
    a.target = target;
    a.target_present = nullptr;
    a.inv_k = inv_k;
    a.valid = valid;
    a.S = d->it_sample_size;
    a.K = d->n_requests;
    a.P = d->n_requests / 2;
    a.has_reset = d->has_reset != 0;
    a.N = N;
    for (int p = 0; p < a.P; ++p) {
        DEN_CHECK_ARG(d->error_kind[p] >= 0 && d->error_kind[p] <= 3, "error_kind must be 0..3");
        a.kind[p] = d->error_kind[p];
        a.has_target[p] = d->has_target[p] != 0;
        DEN_CHECK_ARG(!a.has_target[p] || target, "a pair has a target but `target` is NULL");
    }
    return DEN_OK;
}

int den_lpf_loss_fwd(const den_lpf_loss_desc* d, const float* intensity, const float* sample_dt_ns,
                     const double* coef, const double* reset_dt_ns, const float* target,
                     const float* inv_k, const uint8_t* valid, int64_t N, float* terms, int32_t* counts,
                     float* log_intensity, void* workspace, void* stream) {
    using namespace den;
    lpf::LossArgs a;
    int rc = fill_args(a, d, intensity, sample_dt_ns, coef, reset_dt_ns, target, inv_k, valid, N);
    if (rc) return rc;
    DEN_CHECK_ARG(terms && counts && workspace, "null pointer");
    double* err = reinterpret_cast<double*>(workspace);
    unsigned* ticket = reinterpret_cast<unsigned*>(err + (size_t)a.P * N);
    if (a.K <= 4)
        lpf::lpf_loss_kernel<false, 4><<<(unsigned)N, 32 * a.K, 0, as_stream(stream)>>>(
            a, err, ticket, terms, counts, log_intensity, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    else
        lpf::lpf_loss_kernel<false, lpf::kMaxRequests><<<(unsigned)N, 32 * a.K, 0, as_stream(stream)>>>(
            a, err, ticket, terms, counts, log_intensity, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_lpf_loss_bwd(const den_lpf_loss_desc* d, const float* intensity, const float* sample_dt_ns,
                     const double* coef, const double* reset_dt_ns, const float* target,
                     const float* inv_k, const uint8_t* valid, int64_t N, const int32_t* counts,
                     const float* d_terms, float* d_intensity, double* d_coef, double* d_reset_dt_ns,
                     float* d_target, double* d_inv_k, void* stream) {
    using namespace den;
    lpf::LossArgs a;
    int rc = fill_args(a, d, intensity, sample_dt_ns, coef, reset_dt_ns, target, inv_k, valid, N);
    if (rc) return rc;
    DEN_CHECK_ARG(counts && d_terms && d_intensity, "null pointer");
    if (a.K <= 4)
        lpf::lpf_loss_kernel<true, 4><<<(unsigned)N, 32 * a.K, 0, as_stream(stream)>>>(
            a, nullptr, nullptr, nullptr, const_cast<int32_t*>(counts), nullptr, d_terms, d_intensity, d_coef,
            d_reset_dt_ns, d_target, d_inv_k);
    else
        lpf::lpf_loss_kernel<true, lpf::kMaxRequests><<<(unsigned)N, 32 * a.K, 0, as_stream(stream)>>>(
            a, nullptr, nullptr, nullptr, const_cast<int32_t*>(counts), nullptr, d_terms, d_intensity, d_coef,
            d_reset_dt_ns, d_target, d_inv_k);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
