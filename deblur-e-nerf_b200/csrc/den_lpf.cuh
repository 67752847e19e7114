// Interval mathematics of the pixel-bandwidth low-pass filter shared by den_lpf.cu (one thread per
// event, any S <= 64) and den_lpf_loss.cu (one lane per interval, S <= 32, fused with the loss).
// See den_lpf.cu for the model and the derivation.
#pragma once
#include "den_common.cuh"

namespace den {


constexpr int kLpfMaxS = 64;
constexpr int kLpfThreads = 64;
constexpr int kTaylor = 12;

struct Mat4 {
    double m[16];
};

__device__ __forceinline__ void mat_mul(const double* __restrict__ a, const double* __restrict__ b,
                                        double* __restrict__ c) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = fma(a[4 * i + k], b[4 * k + j], s);
            c[4 * i + j] = s;
        }
}

__device__ __forceinline__ int scaling_power(const double* m) {
    double norm = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double col = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) col += fabs(m[4 * i + j]);
        norm = fmax(norm, col);
    }
    if (!(norm > 0.5)) return 0;
    int e;
    frexp(norm, &e);            // norm = f * 2^e, f in [0.5, 1)
    return min(e + 1, 60);      // ||m / 2^s|| <= 0.5
}

// E = expm(M)
__device__ inline void expm4(const double* M, double* E) {
    const int s = scaling_power(M);
    const double sc = ldexp(1.0, -s);
    double X[16], T[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) X[i] = M[i] * sc;
#pragma unroll
    for (int i = 0; i < 16; ++i) E[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int k = kTaylor; k >= 1; --k) {
        mat_mul(X, E, T);
        const double inv = 1.0 / k;
#pragma unroll
        for (int i = 0; i < 16; ++i) E[i] = ((i % 5 == 0) ? 1.0 : 0.0) + T[i] * inv;
    }
    for (int q = 0; q < s; ++q) {
        mat_mul(E, E, T);
#pragma unroll
        for (int i = 0; i < 16; ++i) E[i] = T[i];
    }
}

// Lout = L(M^T, G): adjoint of expm at M applied to the output adjoint G
__device__ inline void expm4_adjoint(const double* M, const double* G, double* Lout) {
    const int s = scaling_power(M);
    const double sc = ldexp(1.0, -s);
    double X[16], Y[16], E[16], F[16], T1[16], T2[16], T3[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) X[4 * i + j] = M[4 * j + i] * sc;      // transpose
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        Y[i] = G[i] * sc;
        E[i] = (i % 5 == 0) ? 1.0 : 0.0;
        F[i] = 0.0;
    }
    for (int k = kTaylor; k >= 1; --k) {
        mat_mul(X, E, T1);
        mat_mul(X, F, T2);
        mat_mul(Y, E, T3);
        const double inv = 1.0 / k;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            E[i] = ((i % 5 == 0) ? 1.0 : 0.0) + T1[i] * inv;
            F[i] = (T2[i] + T3[i]) * inv;
        }
    }
    for (int q = 0; q < s; ++q) {
        mat_mul(E, F, T1);
        mat_mul(F, E, T2);
        mat_mul(E, E, T3);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            F[i] = T1[i] + T2[i];
            E[i] = T3[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) Lout[i] = F[i];
}

struct Interval {
    double a, wn, wsf, wd, dt;      // balanced rates and the step in seconds
    double phi[16];                 // expm(A' dt)
    double u[4];                    // A'^-1 G1
    double bd[4], bt[4];            // Bd', Bt' (balanced coordinates)
};

__device__ __forceinline__ void build_balanced(double a, double wn, double wsf, double wd, double dt,
                                               double* M) {
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = 0.0;
    M[0] = -a * dt;
    M[1] = -wn * dt;
    M[4] = wn * dt;
    M[9] = wsf * dt;
    M[10] = -wsf * dt;
    M[14] = wd * dt;
    M[15] = -wd * dt;
}

__device__ inline void discretize(double I, double dt_s, const double* coef, Interval& iv) {
    iv.a = coef[0] + coef[1] * I;
    iv.wn = sqrt(coef[2] * I);
    iv.wsf = coef[3];
    iv.wd = coef[4];
    iv.dt = dt_s;
    double M[16];
    build_balanced(iv.a, iv.wn, iv.wsf, iv.wd, dt_s, M);
    expm4(M, iv.phi);
    // G1 = e - Phi e, e = (0,1,1,1)
    double g1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        g1[r] = (r > 0 ? 1.0 : 0.0) - (iv.phi[4 * r + 1] + iv.phi[4 * r + 2] + iv.phi[4 * r + 3]);
    // u = A'^-1 G1
    iv.u[0] = g1[1] / iv.wn;
    iv.u[1] = -(g1[0] + iv.a * iv.u[0]) / iv.wn;
    iv.u[2] = iv.u[1] - g1[2] / iv.wsf;
    iv.u[3] = iv.u[2] - g1[3] / iv.wd;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        iv.bt[r] = iv.u[r] / dt_s + (r > 0 ? 1.0 : 0.0);      // G2
        iv.bd[r] = g1[r] - iv.bt[r];                          // G1 - G2
    }
}


// Reverse mode of discretize() for one interval: given the adjoints of Phi, Bd, Bt (balanced
// coordinates) and the direct adjoint of wn collected by the weight recursion, returns the adjoints
// of a = alpha0 + alpha1 I, b = beta I, w_sf and w_d.
__device__ inline void interval_adjoint(const Interval& iv, double* phibar /* modified */, const double* bdbar,
                                        const double* btbar, double wnbar, double& abar, double& bbar,
                                        double& wsfbar, double& wdbar) {
    // Bd = G1 - G2, Bt = G2
    double g1bar[4], g2bar[4], ub[4], v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { g1bar[k] = bdbar[k]; g2bar[k] = btbar[k] - bdbar[k]; }
    // G2 = u / dt + e, u = A'^-1 G1
#pragma unroll
    for (int k = 0; k < 4; ++k) ub[k] = g2bar[k] / iv.dt;
    v[3] = -ub[3] / iv.wd;                                  // v = A'^-T ub
    v[2] = (iv.wd * v[3] - ub[2]) / iv.wsf;
    v[0] = (iv.wsf * v[2] - ub[1]) / iv.wn;
    v[1] = (ub[0] + iv.a * v[0]) / iv.wn;
    double Abar[16];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        g1bar[r] += v[r];
#pragma unroll
        for (int q = 0; q < 4; ++q) Abar[4 * r + q] = -v[r] * iv.u[q];
    }
    // G1 = e - Phi e
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        phibar[4 * r + 1] -= g1bar[r];
        phibar[4 * r + 2] -= g1bar[r];
        phibar[4 * r + 3] -= g1bar[r];
    }
    double M[16], Mbar[16];
    build_balanced(iv.a, iv.wn, iv.wsf, iv.wd, iv.dt, M);
    expm4_adjoint(M, phibar, Mbar);
#pragma unroll
    for (int i = 0; i < 16; ++i) Abar[i] += iv.dt * Mbar[i];
    abar = -Abar[0];
    wnbar += -Abar[1] + Abar[4];
    wsfbar = Abar[9] - Abar[10];
    wdbar = Abar[14] - Abar[15];
    bbar = wnbar / (2.0 * iv.wn);                           // wn = sqrt(b)
}

// warp-cooperative kernels (den_lpf_loss.cu): one warp per event, one lane per interval
int lpf_warp_fwd(const float* intensity, const float* dt_ns, const double* coef, int S, int64_t N, int nc,
                 float* out, cudaStream_t stream);
int lpf_warp_bwd(const float* intensity, const float* dt_ns, const double* coef, int S, int64_t N, int nc,
                 const float* d_out, float* d_intensity, double* d_coef, cudaStream_t stream);

}  // namespace den
