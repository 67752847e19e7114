// Device helpers shared by the fused field kernels (forward, backward, tensor-core).
#pragma once
#include "den_common.cuh"
#include "den_hashgrid.cuh"

namespace den {

// The architecture every shipped config uses (configs/train/*.yaml:81-103):
constexpr int kEncDim = 32;    // L*F, zero-padded when L < 16
constexpr int kWidth = 64;     // n_neurons
constexpr int kBaseOut = 16;   // 1 density + 15 geo features
constexpr int kGeo = 15;
constexpr int kShDim = 16;     // SH degree 4
constexpr int kHeadIn = 32;    // 16 SH + 15 geo + 1 zero pad
constexpr int kOutPad = 4;     // radiance channels padded to 4

// activation ids (den_b200.h)
constexpr int kActRelu = 0, kActSoftplus100 = 1;
constexpr int kDensTruncExp = 0, kDensSoftplus = 1, kDensShiftedSoftplus = 2;
constexpr int kRadSoftplus = 0, kRadSigmoid = 1;

struct FieldSmem {
    float *wb1, *bb1, *wb2, *bb2, *w1, *b1, *w2, *b2, *w3, *b3;
};

__host__ __device__ constexpr int field_smem_floats() {
    return kEncDim * kWidth + kWidth + kWidth * kBaseOut + kBaseOut + kHeadIn * kWidth + kWidth +
           kWidth * kWidth + kWidth + kWidth * kOutPad + kOutPad;
}
inline size_t field_smem_bytes() { return (size_t)field_smem_floats() * sizeof(float); }

__device__ __forceinline__ FieldSmem carve_field_smem(float* base) {
    FieldSmem s;
    s.wb1 = base;                      base += kEncDim * kWidth;
    s.bb1 = base;                      base += kWidth;
    s.wb2 = base;                      base += kWidth * kBaseOut;
    s.bb2 = base;                      base += kBaseOut;
    s.w1 = base;                       base += kHeadIn * kWidth;
    s.b1 = base;                       base += kWidth;
    s.w2 = base;                       base += kWidth * kWidth;
    s.b2 = base;                       base += kWidth;
    s.w3 = base;                       base += kWidth * kOutPad;
    s.b3 = base;
    return s;
}

// (out, in) row-major global -> [in_pad][out_pad] shared, zero padded
__device__ __forceinline__ void load_transposed(float* dst, const float* __restrict__ src, int n_out,
                                                int n_in, int out_pad, int in_pad) {
    for (int i = threadIdx.x; i < in_pad * out_pad; i += blockDim.x) {
        const int k = i / out_pad, j = i - k * out_pad;
        dst[i] = (k < n_in && j < n_out) ? __ldg(src + j * n_in + k) : 0.f;
    }
}
__device__ __forceinline__ void load_padded(float* dst, const float* __restrict__ src, int n, int pad) {
    for (int i = threadIdx.x; i < pad; i += blockDim.x) dst[i] = i < n ? __ldg(src + i) : 0.f;
}

__device__ __forceinline__ void load_field_weights(const FieldSmem& s, const den_field_desc& f,
                                                   const den_field_params& p, bool full) {
    const int enc = f.grid.n_levels * 2;
    load_transposed(s.wb1, p.wb1, kWidth, enc, kWidth, kEncDim);
    load_padded(s.bb1, p.bb1, kWidth, kWidth);
    load_transposed(s.wb2, p.wb2, kBaseOut, kWidth, kBaseOut, kWidth);
    load_padded(s.bb2, p.bb2, kBaseOut, kBaseOut);
    if (full) {
        load_transposed(s.w1, p.w1, kWidth, kShDim + kGeo, kWidth, kHeadIn);
        load_padded(s.b1, p.b1, kWidth, kWidth);
        load_transposed(s.w2, p.w2, kWidth, kWidth, kWidth, kWidth);
        load_padded(s.b2, p.b2, kWidth, kWidth);
        load_transposed(s.w3, p.w3, f.channels, kWidth, kOutPad, kWidth);
        load_padded(s.b3, p.b3, f.channels, kOutPad);
    }
}

// Field-side contraction (external/ngp.py:68-106,230-238).  Returns the selector
// all(0 < u < 1) that gates the density (ngp.py:238,247-250).
__device__ __forceinline__ bool contract_position(const den_field_desc& f, const float pos[3],
                                                  float u[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = __fdiv_rn(pos[d] - f.aabb[d], f.aabb[d + 3] - f.aabb[d]);
    if (f.contraction == DEN_CONTRACT_SPHERE) {
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = u[d] * 2.f - 1.f;
        const float mag = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (mag > 1.f) {
            const float k = (2.f - 1.f / mag) / mag;
#pragma unroll
            for (int d = 0; d < 3; ++d) u[d] *= k;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = u[d] * 0.25f + 0.5f;
    } else if (f.contraction == DEN_CONTRACT_TANH) {
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = (tanhf(u[d] - 0.5f) + 1.f) * 0.5f;
    }
    return u[0] > 0.f && u[0] < 1.f && u[1] > 0.f && u[1] < 1.f && u[2] > 0.f && u[2] < 1.f;
}

// all levels of one sample into registers (enc[2l], enc[2l+1]); unused tail zeroed
__device__ __forceinline__ void encode_sample(const den_hashgrid_desc& g,
                                              const float2* __restrict__ table, const float u[3],
                                              float (&enc)[kEncDim]) {
#pragma unroll
    for (int level = 0; level < kEncDim / 2; ++level) {
        if (level < g.n_levels) {
            const LevelInfo li = make_level(g, level);
            const float2* __restrict__ base = table + g.offset[level];
            const CellFrac cf = locate(li.scale, u[0], u[1], u[2]);
            float2 v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = __ldg(base + corner_index(li, cf, c));
            float ax = 0.f, ay = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = corner_weight(cf, c);
                ax = fmaf(w, v[c].x, ax);
                ay = fmaf(w, v[c].y, ay);
            }
            enc[2 * level] = ax;
            enc[2 * level + 1] = ay;
        } else {
            enc[2 * level] = 0.f;
            enc[2 * level + 1] = 0.f;
        }
    }
}

// Real spherical harmonics, degree 4 (external/sh_encoder.py:42-77)
__device__ __forceinline__ void sh_degree4(const float d[3], float* out) {
    const float x = d[0], y = d[1], z = d[2];
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    out[0] = 0.28209479177387814f;
    out[1] = -0.48860251190291987f * y;
    out[2] = 0.48860251190291987f * z;
    out[3] = -0.48860251190291987f * x;
    out[4] = 1.0925484305920792f * xy;
    out[5] = -1.0925484305920792f * yz;
    out[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    out[7] = -1.0925484305920792f * xz;
    out[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    out[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    out[10] = 2.8906114426405538f * xy * z;
    out[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    out[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    out[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    out[14] = 1.4453057213202769f * z * (x2 - y2);
    out[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// dL/d(dir) = J_SH(dir)^T dsh  (reverse mode of sh_degree4; the tau-gradient path)
__device__ __forceinline__ void sh_degree4_grad(const float d[3], const float* g, float out[3]) {
    const float x = d[0], y = d[1], z = d[2];
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    const float c1 = 0.48860251190291987f, c2 = 1.0925484305920792f, c3 = 0.94617469575755997f,
                c5 = 0.54627421529603959f, c6 = 0.59004358992664352f, c7 = 2.8906114426405538f,
                c8 = 0.45704579946446572f, c9 = 0.3731763325901154f, c10 = 1.4453057213202769f;
    out[0] = -c1 * g[3] + c2 * y * g[4] - c2 * z * g[7] + 2.f * c5 * x * g[8] - 6.f * c6 * x * y * g[9] +
             c7 * y * z * g[10] + c8 * (1.f - 5.f * z2) * g[13] + 2.f * c10 * x * z * g[14] +
             c6 * (3.f * y2 - 3.f * x2) * g[15];
    out[1] = -c1 * g[1] + c2 * x * g[4] - c2 * z * g[5] - 2.f * c5 * y * g[8] +
             c6 * (3.f * y2 - 3.f * x2) * g[9] + c7 * x * z * g[10] + c8 * (1.f - 5.f * z2) * g[11] -
             2.f * c10 * y * z * g[14] + 6.f * c6 * x * y * g[15];
    out[2] = c1 * g[2] - c2 * y * g[5] + 2.f * c3 * z * g[6] - c2 * x * g[7] + c7 * x * y * g[10] -
             10.f * c8 * y * z * g[11] + c9 * (15.f * z2 - 3.f) * g[12] - 10.f * c8 * x * z * g[13] +
             c10 * (x2 - y2) * g[14];
}

// dL/dpos = J_contract(pos)^T du  (reverse mode of contract_position)
__device__ __forceinline__ void contract_position_grad(const den_field_desc& f, const float pos[3],
                                                       const float du[3], float dpos[3]) {
    float inv_ext[3], u[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        inv_ext[d] = 1.f / (f.aabb[d + 3] - f.aabb[d]);
        u[d] = (pos[d] - f.aabb[d]) * inv_ext[d];
    }
    if (f.contraction == DEN_CONTRACT_SPHERE) {
        float v[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = u[d] * 2.f - 1.f;
        const float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        float g[3] = {du[0] * 0.25f, du[1] * 0.25f, du[2] * 0.25f};
        if (n > 1.f) {
            const float s = 2.f / n - 1.f / (n * n);
            const float k = (-2.f / (n * n * n) + 2.f / (n * n * n * n)) * (v[0] * g[0] + v[1] * g[1] + v[2] * g[2]);
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] = s * g[d] + k * v[d];
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) dpos[d] = g[d] * 2.f * inv_ext[d];
    } else if (f.contraction == DEN_CONTRACT_TANH) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float t = tanhf(u[d] - 0.5f);
            dpos[d] = du[d] * 0.5f * (1.f - t * t) * inv_ext[d];
        }
    } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) dpos[d] = du[d] * inv_ext[d];
    }
}

// ---- activations (models/nerf.py:17-29, external/ngp.py:45-65) -----------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus(beta=100): max(z,0) + log1p(exp(-100|z|))/100.  torch's threshold branch (z > 0.2 ->
// identity) only guards exp overflow; here exp(-100|z|) <= 1 and the correction is < 2.1e-11
// beyond the threshold, i.e. below fp32 resolution of z, so the branch-free form is exact to
// rounding.  6 instructions: FMUL, EX2, FADD, LG2, FMNMX, FFMA.
__device__ __forceinline__ float softplus100(float z) {
    const float e = ex2_approx(-144.26950408889634f * fabsf(z));          // exp(-100 |z|)
    return fmaf(lg2_approx(1.f + e), 0.0069314718055994531f, fmaxf(z, 0.f));   // ln2 / 100
}
__device__ __forceinline__ float hidden_act(int id, float z) {
    return id == kActSoftplus100 ? softplus100(z) : fmaxf(z, 0.f);
}
// derivative expressed from the post-activation value h: sigmoid(100 z) = 1 - exp(-100 h)
__device__ __forceinline__ float hidden_act_grad_from_out(int id, float h) {
    if (id == kActSoftplus100) return 1.f - ex2_approx(-144.26950408889634f * h);
    return h > 0.f ? 1.f : 0.f;
}
// vector forms with the activation switch hoisted out of the element loop
template <int N>
__device__ __forceinline__ void bias_hidden_act(int id, float (&h)[N], const float* __restrict__ bias) {
    if (id == kActSoftplus100) {
#pragma unroll
        for (int j = 0; j < N; ++j) h[j] = softplus100(h[j] + bias[j]);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) h[j] = fmaxf(h[j] + bias[j], 0.f);
    }
}
template <int N>
__device__ __forceinline__ void mul_hidden_act_grad(int id, float (&d)[N], const float (&h)[N]) {
    if (id == kActSoftplus100) {
#pragma unroll
        for (int j = 0; j < N; ++j) d[j] = fmaf(-d[j], ex2_approx(-144.26950408889634f * h[j]), d[j]);   // d (1 - e)
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) d[j] = h[j] > 0.f ? d[j] : 0.f;
    }
}
__device__ __forceinline__ float softplus1(float z) {
    return z > 20.f ? z : fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
}
__device__ __forceinline__ float density_act(int id, float raw) {
    if (id == kDensTruncExp) return expf(raw - 1.f);
    if (id == kDensSoftplus) return softplus1(raw);
    return softplus1(raw - 1.f);
}
// d density / d raw  (trunc_exp backward clamps the exponent at 15, ngp.py:55-58)
__device__ __forceinline__ float density_act_grad(int id, float raw) {
    if (id == kDensTruncExp) return expf(fminf(raw - 1.f, 15.f));
    const float z = id == kDensSoftplus ? raw : raw - 1.f;
    return z > 20.f ? 1.f : 1.f / (1.f + expf(-z));
}
__device__ __forceinline__ float radiance_act(int id, float z) {
    return id == kRadSigmoid ? 1.f / (1.f + expf(-z)) : softplus1(z);
}
__device__ __forceinline__ float radiance_act_grad(int id, float z) {
    const float s = 1.f / (1.f + expf(-z));
    if (id == kRadSigmoid) return s * (1.f - s);
    return z > 20.f ? 1.f : s;
}

int check_field(const den_field_desc* f, const den_field_params* p, bool full);

}  // namespace den
