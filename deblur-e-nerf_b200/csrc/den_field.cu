// Fused radiance-field evaluation: contraction -> hash-grid gather -> base MLP -> density
// activation -> (SH || geo) -> head MLP -> radiance activation, one thread per sample, no
// intermediate ever written to HBM (SIMT fp32 path; sm_100a).
//
// Replaces NGPradianceField.query_density / forward (external/ngp.py:230-280), the
// tcnn.Encoding call inside it (ngp.py:240), MLP.forward (external/mlp.py:99-113),
// SHEncoder.forward (external/sh_encoder.py:28-77, degree 4) and the position
// computation of the sigma_fn / rgb_sigma_fn closures (external/utils.py:68-96).
// The reference round-trips (M,32), (M,64) x3 and (M,16) activations through HBM between
// cuBLAS SGEMMs; here the only per-sample traffic is 16 B in (ray index, t0, t1 + the
// ray's o/d from cache), the table gathers (L2-resident, 1 KiB/sample) and 4(1+C) B out.
//
// Layout: nn.Linear weights (out,in) are transposed into shared memory as [in][out] at
// CTA start so that a layer is `out[j] += W[k][j] * x[k]` with x[k] in a register and the
// 64 weights of row k fetched by broadcast LDS.128 (all lanes read the same address).
#include "den_common.cuh"
#include "den_field.cuh"

namespace den {

constexpr int kFieldThreads = 128;

template <int K, int N>
__device__ __forceinline__ void dense(const float* __restrict__ sW, const float* __restrict__ sb,
                                      const float (&x)[K], float (&y)[N]) {
    static_assert(N % 4 == 0, "N must be a multiple of 4");
#pragma unroll
    for (int j = 0; j < N; ++j) y[j] = sb[j];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float xk = x[k];
#pragma unroll
        for (int j4 = 0; j4 < N / 4; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(sW + k * N + 4 * j4);
            y[4 * j4 + 0] = fmaf(w.x, xk, y[4 * j4 + 0]);
            y[4 * j4 + 1] = fmaf(w.y, xk, y[4 * j4 + 1]);
            y[4 * j4 + 2] = fmaf(w.z, xk, y[4 * j4 + 2]);
            y[4 * j4 + 3] = fmaf(w.w, xk, y[4 * j4 + 3]);
        }
    }
}

template <bool kFromRays, bool kFull>
__global__ void __launch_bounds__(kFieldThreads, 3)
field_fwd_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                 const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                 const int32_t* __restrict__ ray_indices, const float* __restrict__ t_starts,
                 const float* __restrict__ t_ends, const float* __restrict__ positions,
                 int64_t n_host, const int32_t* __restrict__ n_dev, float* __restrict__ sigmas,
                 float* __restrict__ rgbs) {
    extern __shared__ __align__(16) float smem[];
    FieldSmem s = carve_field_smem(smem);
    load_field_weights(s, f, p, kFull);
    __syncthreads();

    const int64_t n = n_dev ? (int64_t)min(*n_dev, (int32_t)min(n_host, (int64_t)INT32_MAX)) : n_host;
    const int C = f.channels;

    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float pos[3], dir[3] = {0.f, 0.f, 1.f};
        if (kFromRays) {
            const int64_t r = ray_indices[i];
            const float tm = (t_starts[i] + t_ends[i]);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                dir[d] = __ldg(rays_d + 3 * r + d);
                pos[d] = __ldg(rays_o + 3 * r + d) + (dir[d] * tm) * 0.5f;
            }
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) pos[d] = positions[3 * i + d];
        }
        float u[3];
        const bool inside = contract_position(f, pos, u);

        float enc[kEncDim];
        encode_sample(f.grid, reinterpret_cast<const float2*>(p.table), u, enc);

        float hb[kWidth];
        dense<kEncDim, kWidth>(s.wb1, s.bb1, enc, hb);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) hb[j] = hidden_act(f.hidden_act, hb[j]);

        if (!kFull) {
            float raw = s.bb2[0];
#pragma unroll
            for (int k = 0; k < kWidth; ++k) raw = fmaf(s.wb2[k * kBaseOut], hb[k], raw);
            sigmas[i] = inside ? density_act(f.density_act, raw) : 0.f;
            continue;
        }
        float y[kBaseOut];
        dense<kWidth, kBaseOut>(s.wb2, s.bb2, hb, y);
        sigmas[i] = inside ? density_act(f.density_act, y[0]) : 0.f;

        float in1[kHeadIn];
        sh_degree4(dir, in1);
#pragma unroll
        for (int j = 0; j < kGeo; ++j) in1[kShDim + j] = y[1 + j];
        in1[kHeadIn - 1] = 0.f;

        float h1[kWidth];
        dense<kHeadIn, kWidth>(s.w1, s.b1, in1, h1);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) h1[j] = hidden_act(f.hidden_act, h1[j]);
        float h2[kWidth];
        dense<kWidth, kWidth>(s.w2, s.b2, h1, h2);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) h2[j] = hidden_act(f.hidden_act, h2[j]);
        float out[kOutPad];
        dense<kWidth, kOutPad>(s.w3, s.b3, h2, out);
        for (int c = 0; c < C; ++c) rgbs[i * C + c] = radiance_act(f.radiance_act, out[c]);
    }
}

int check_field(const den_field_desc* f, const den_field_params* p, bool full) {
    if (!f || !p) { set_error("field: null descriptor"); return DEN_ERR_INVALID_ARGUMENT; }
    if (f->grid.n_features != 2 || f->grid.n_levels < 1 || f->grid.n_levels * 2 > kEncDim) {
        set_error("field: hash grid must have F=2 and at most %d levels (got L=%d, F=%d)",
                  kEncDim / 2, f->grid.n_levels, f->grid.n_features);
        return DEN_ERR_UNSUPPORTED;
    }
    if (f->channels != 1 && f->channels != 3) {
        set_error("field: radiance channels must be 1 or 3 (got %d)", f->channels);
        return DEN_ERR_UNSUPPORTED;
    }
    if (f->width != kWidth || f->geo_feat_dim != kGeo || f->sh_degree != 4 ||
        f->n_hidden_base != 1 || f->n_hidden_head != 2) {
        set_error("field: only the shipped architecture is built (base 1x64 -> 1+15, SH degree 4, "
                  "head 2x64); got width=%d geo=%d sh=%d base=%d head=%d",
                  f->width, f->geo_feat_dim, f->sh_degree, f->n_hidden_base, f->n_hidden_head);
        return DEN_ERR_UNSUPPORTED;
    }
    if (!p->table || !p->wb1 || !p->bb1 || !p->wb2 || !p->bb2) {
        set_error("field: null base parameters");
        return DEN_ERR_INVALID_ARGUMENT;
    }
    if (full && (!p->w1 || !p->b1 || !p->w2 || !p->b2 || !p->w3 || !p->b3)) {
        set_error("field: null head parameters");
        return DEN_ERR_INVALID_ARGUMENT;
    }
    return DEN_OK;
}

}  // namespace den

extern "C" {

int den_field_fwd(const den_field_desc* f, const den_field_params* p, const float* rays_o,
                  const float* rays_d, const int32_t* ray_indices, const float* t_starts,
                  const float* t_ends, int64_t n, const int32_t* n_dev, float* sigmas, float* rgbs,
                  void* stream) {
    using namespace den;
    const bool full = rgbs != nullptr;
    int rc = check_field(f, p, full);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(rays_o && rays_d && ray_indices && t_starts && t_ends && sigmas, "null pointer");
    const size_t smem = field_smem_bytes();
    const int grid = grid_for(n, kFieldThreads, 3);
    if (full) {
        cudaFuncSetAttribute(field_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        field_fwd_kernel<true, true><<<grid, kFieldThreads, smem, as_stream(stream)>>>(
            *f, *p, rays_o, rays_d, ray_indices, t_starts, t_ends, nullptr, n, n_dev, sigmas, rgbs);
    } else {
        cudaFuncSetAttribute(field_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        field_fwd_kernel<true, false><<<grid, kFieldThreads, smem, as_stream(stream)>>>(
            *f, *p, rays_o, rays_d, ray_indices, t_starts, t_ends, nullptr, n, n_dev, sigmas, nullptr);
    }
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_field_density_at(const den_field_desc* f, const den_field_params* p, const float* positions,
                         int64_t n, float* sigmas, void* stream) {
    using namespace den;
    int rc = check_field(f, p, false);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(positions && sigmas, "null pointer");
    const size_t smem = field_smem_bytes();
    cudaFuncSetAttribute(field_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    field_fwd_kernel<false, false><<<grid_for(n, kFieldThreads, 3), kFieldThreads, smem, as_stream(stream)>>>(
        *f, *p, nullptr, nullptr, nullptr, nullptr, nullptr, positions, n, nullptr, sigmas, nullptr);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
