// Density / colour MLP fused per 128-sample tile on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM) — the one genuine dense contraction of the hot path.
//
// Replaces MLP.forward (external/mlp.py:99-113) for mlp_base[1] and mlp_head as called from
// NGPradianceField.query_density / _query_rgb (external/ngp.py:239-267), the SH direction
// encoding (external/sh_encoder.py:28-77), the density / radiance activations
// (models/nerf.py:17-29, external/ngp.py:45-65).  The reference runs five cuBLAS SGEMMs with
// TF32 off and round-trips every (M,64) activation through HBM; here activations go
// registers -> bf16 hi/lo operand tiles in shared memory -> tcgen05.mma -> TMEM -> registers,
// and only enc (128 B), sigma and rgb touch HBM.
//
// fp32 fidelity: every operand is split x = hi + lo into two bf16 tiles and each layer issues
// three MMA passes (hi*hi + lo*hi + hi*lo) accumulated in fp32 TMEM: ~2^-17 relative per
// product, measured against the fp32 evaluation in tests/test_gpu_mlp_tc.py (single-pass TF32
// measured 2e-3 on the head-weight gradients, outside the 1e-3 parity bound).
//
// Thread layout: see den_mlp_tc.cuh (S = 2: 8 epilogue warps + 1 MMA warp, 2 CTAs per SM).
#include "den_mlp_tc.cuh"

namespace den {

using namespace mlp;

constexpr int kFwdS = 2;
constexpr int kFwdCols = kWidth / kFwdS;                 // 32 accumulator columns per thread
constexpr int kFwdEpiThreads = 128 * kFwdS;
constexpr int kFwdThreads = kFwdEpiThreads + 32;
constexpr uint32_t kFwdTmemCols = 64;

struct FwdSmem {
    static constexpr int a32 = (Weights::end + 127) / 128 * 128;      // (128, 32): enc, then [SH|geo]
    static constexpr int a64 = a32 + Tile<kTile, kEncDim>::bytes;     // (128, 64): hb, h1, h2
    static constexpr int bar = a64 + Tile<kTile, kWidth>::bytes;
    static constexpr int tmem_ptr = bar + 8;
    static constexpr int total = tmem_ptr + 8;
};

template <bool kFull>
__global__ void __launch_bounds__(kFwdThreads, 2)
mlp_fwd_tc_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                  const float* __restrict__ enc, const float* __restrict__ rays_o,
                  const float* __restrict__ rays_d, const int32_t* __restrict__ ray_indices,
                  const float* __restrict__ t_starts, const float* __restrict__ t_ends, int64_t n,
                  float* __restrict__ sigmas, float* __restrict__ rgbs) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + FwdSmem::bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FwdSmem::tmem_ptr);
    const float* s_bb1 = reinterpret_cast<const float*>(smem + Weights::bias);
    const float* s_bb2 = s_bb1 + kWidth;
    const float* s_b1 = s_bb2 + kBaseOut;
    const float* s_b2 = s_b1 + kWidth;
    const float* s_b3 = s_b2 + kWidth;

    const Tile<kWidth, kEncDim> Wb1(smem + Weights::wb1);
    const Tile<kBaseOut, kWidth> Wb2(smem + Weights::wb2);
    const Tile<kWidth, kHeadIn> W1(smem + Weights::w1);
    const Tile<kWidth, kWidth> W2(smem + Weights::w2);
    const Tile<kOutN, kWidth> W3(smem + Weights::w3);
    const Tile<kTile, kEncDim> T32(smem + FwdSmem::a32);
    const Tile<kTile, kWidth> T64(smem + FwdSmem::a64);

    load_all_weights(smem, f, p, kFull);
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kFwdTmemCols);
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int enc_dim = f.grid.n_levels * 2;
    const int C = f.channels;

    if (warp == kFwdEpiThreads / 32) {
        // ===================== MMA warp: one issue per round =====================
        // hand-off in: named barrier 1 (producers arrive, this warp syncs); hand-off out: tcgen05.commit
        // on the mbarrier; one elected lane issues the whole round
#define DEN_ROUND(...)                                     \
    handoff_sync(kFwdThreads);                             \
    if (tc::elect_one()) { __VA_ARGS__; tc::mma_commit_1t(bar); } \
    __syncwarp();
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            DEN_ROUND(mma_fwd_1t<kWidth, kEncDim>(tmem_base, T32, Wb1))
            DEN_ROUND(mma_fwd_1t<kBaseOut, kWidth>(tmem_base, T64, Wb2))
            if (kFull) {
                DEN_ROUND(mma_fwd_1t<kWidth, kHeadIn>(tmem_base, T32, W1))
                DEN_ROUND(mma_fwd_1t<kWidth, kWidth>(tmem_base, T64, W2))
                DEN_ROUND(mma_fwd_1t<kOutN, kWidth>(tmem_base, T64, W3))
            }
        }
#undef DEN_ROUND
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3, cg = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t i = tile * kTile + row;
            const bool valid = i < n;
            // ---- stage 0: 16 of the 32 encoding features of this row -> T32 ------------------
            float x16[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) x16[k] = 0.f;
            float dir[3] = {0.f, 0.f, 1.f};
            bool inside = false;
            if (valid) {
                const float4* src = reinterpret_cast<const float4*>(enc + i * enc_dim + 16 * cg);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4)
                    if (16 * cg + 4 * v4 < enc_dim) {
                        const float4 v = __ldg(src + v4);
                        x16[4 * v4] = v.x; x16[4 * v4 + 1] = v.y; x16[4 * v4 + 2] = v.z; x16[4 * v4 + 3] = v.w;
                    }
                const int64_t r = ray_indices[i];
                const float tm = t_starts[i] + t_ends[i];
                float pos[3], u[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    dir[d] = __ldg(rays_d + 3 * r + d);
                    pos[d] = __ldg(rays_o + 3 * r + d) + (dir[d] * tm) * 0.5f;
                }
                if (cg == 0) inside = contract_position(f, pos, u);
            }
            store_cols<16, kEncDim>(T32, row, 16 * cg, x16);
            publish_arrive(kFwdThreads);

            // ---- base layer 1 -------------------------------------------------------------------
            await(bar, phase);
            float h[kFwdCols];
            tmem_ld_cols<kFwdCols>(tmem_lane + kFwdCols * cg, h);
bias_hidden_act<kFwdCols>(f.hidden_act, h, s_bb1 + kFwdCols * cg);
            store_cols<kFwdCols, kWidth>(T64, row, kFwdCols * cg, h);
            publish_arrive(kFwdThreads);

            // ---- base layer 2: density + geo features; [SH | geo | 0] -> T32 --------------------
            await(bar, phase);
            if (cg == 0) {
                float y[kBaseOut];
                tmem_ld_cols<kBaseOut>(tmem_lane, y);
#pragma unroll
                for (int j = 0; j < kBaseOut; ++j) y[j] += s_bb2[j];
                if (valid) sigmas[i] = inside ? density_act(f.density_act, y[0]) : 0.f;
                if (kFull) {
#pragma unroll
                    for (int j = 0; j < kGeo; ++j) x16[j] = y[1 + j];
                    x16[15] = 0.f;
                    store_cols<16, kHeadIn>(T32, row, 16, x16);
                }
            } else if (kFull) {
                sh_degree4(dir, x16);
                store_cols<16, kHeadIn>(T32, row, 0, x16);
            }
            if (!kFull) {
                tc::tc_fence_before_sync();
                continue;
            }
            publish_arrive(kFwdThreads);

            // ---- head layers ----------------------------------------------------------------------
            await(bar, phase);
            tmem_ld_cols<kFwdCols>(tmem_lane + kFwdCols * cg, h);
bias_hidden_act<kFwdCols>(f.hidden_act, h, s_b1 + kFwdCols * cg);
            store_cols<kFwdCols, kWidth>(T64, row, kFwdCols * cg, h);
            publish_arrive(kFwdThreads);

            await(bar, phase);
            tmem_ld_cols<kFwdCols>(tmem_lane + kFwdCols * cg, h);
bias_hidden_act<kFwdCols>(f.hidden_act, h, s_b2 + kFwdCols * cg);
            store_cols<kFwdCols, kWidth>(T64, row, kFwdCols * cg, h);
            publish_arrive(kFwdThreads);

            await(bar, phase);
            if (cg == 0) {
                float out[kOutN];
                tmem_ld_cols<kOutN>(tmem_lane, out);
                if (valid)
                    for (int c = 0; c < C; ++c) rgbs[i * C + c] = radiance_act(f.radiance_act, out[c] + s_b3[c]);
            }
            tc::tc_fence_before_sync();
        }
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kFwdTmemCols);
}

// positions of marched samples in the field's unit cube (input of the hash-grid kernels)
__global__ void contract_samples_kernel(const __grid_constant__ den_field_desc f,
                                        const float* __restrict__ rays_o,
                                        const float* __restrict__ rays_d,
                                        const int32_t* __restrict__ ray_indices,
                                        const float* __restrict__ t_starts,
                                        const float* __restrict__ t_ends, int64_t n,
                                        float* __restrict__ unit_pos) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = ray_indices[i];
        const float tm = t_starts[i] + t_ends[i];
        float pos[3], u[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)
            pos[d] = __ldg(rays_o + 3 * r + d) + (__ldg(rays_d + 3 * r + d) * tm) * 0.5f;
        contract_position(f, pos, u);
        unit_pos[3 * i + 0] = u[0];
        unit_pos[3 * i + 1] = u[1];
        unit_pos[3 * i + 2] = u[2];
    }
}

// reverse mode of contract_samples_kernel w.r.t. the ray: per sample dL/dpos and dL/dpos * tm/2
// (pos = o + d * tm / 2); the per-ray sums are taken by den_accumulate_fwd
__global__ void contract_samples_bwd_kernel(const __grid_constant__ den_field_desc f,
                                            const float* __restrict__ rays_o,
                                            const float* __restrict__ rays_d,
                                            const int32_t* __restrict__ ray_indices,
                                            const float* __restrict__ t_starts,
                                            const float* __restrict__ t_ends,
                                            const float* __restrict__ d_unit, int64_t n,
                                            float* __restrict__ d_pos, float* __restrict__ d_pos_t) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = ray_indices[i];
        const float tm = t_starts[i] + t_ends[i];
        float pos[3], g[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)
            pos[d] = __ldg(rays_o + 3 * r + d) + (__ldg(rays_d + 3 * r + d) * tm) * 0.5f;
        const float du[3] = {d_unit[3 * i], d_unit[3 * i + 1], d_unit[3 * i + 2]};
        contract_position_grad(f, pos, du, g);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            d_pos[3 * i + d] = g[d];
            d_pos_t[3 * i + d] = g[d] * tm * 0.5f;
        }
    }
}

}  // namespace den

extern "C" {

int den_contract_samples_bwd(const den_field_desc* f, const float* rays_o, const float* rays_d,
                             const int32_t* ray_indices, const float* t_starts, const float* t_ends,
                             const float* d_unit, int64_t n, float* d_pos, float* d_pos_t,
                             void* stream) {
    using namespace den;
    DEN_CHECK_ARG(f != nullptr, "null descriptor");
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(rays_o && rays_d && ray_indices && t_starts && t_ends && d_unit && d_pos && d_pos_t,
                  "null pointer");
    contract_samples_bwd_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        *f, rays_o, rays_d, ray_indices, t_starts, t_ends, d_unit, n, d_pos, d_pos_t);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_contract_samples(const den_field_desc* f, const float* rays_o, const float* rays_d,
                         const int32_t* ray_indices, const float* t_starts, const float* t_ends,
                         int64_t n, float* unit_pos, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(f != nullptr, "null descriptor");
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(rays_o && rays_d && ray_indices && t_starts && t_ends && unit_pos, "null pointer");
    contract_samples_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        *f, rays_o, rays_d, ray_indices, t_starts, t_ends, n, unit_pos);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_mlp_fwd(const den_field_desc* f, const den_field_params* p, const float* enc,
                const float* rays_o, const float* rays_d, const int32_t* ray_indices,
                const float* t_starts, const float* t_ends, int64_t n, float* sigmas, float* rgbs,
                void* stream) {
    using namespace den;
    const bool full = rgbs != nullptr;
    int rc = check_field(f, p, full);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(enc && rays_o && rays_d && ray_indices && t_starts && t_ends && sigmas,
                  "null pointer");
    DEN_CHECK_ARG((f->grid.n_levels * 2) % 4 == 0, "encoding width must be a multiple of 4");
    const int grid = grid_for((n + kTile - 1) / kTile, 1, 2);
    const size_t smem = FwdSmem::total;
    if (full) {
        cudaFuncSetAttribute(mlp_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mlp_fwd_tc_kernel<true><<<grid, kFwdThreads, smem, as_stream(stream)>>>(
            *f, *p, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, n, sigmas, rgbs);
    } else {
        cudaFuncSetAttribute(mlp_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mlp_fwd_tc_kernel<false><<<grid, kFwdThreads, smem, as_stream(stream)>>>(
            *f, *p, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, n, sigmas, nullptr);
    }
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
