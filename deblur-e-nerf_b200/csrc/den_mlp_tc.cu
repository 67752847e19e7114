// Density / colour MLP fused per 128-sample tile on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM) — the one genuine dense contraction of the hot path.
//
// Replaces MLP.forward (external/mlp.py:99-113) for mlp_base[1] and mlp_head as called from
// NGPradianceField.query_density / _query_rgb (external/ngp.py:239-267), the SH direction
// encoding (external/sh_encoder.py:28-77), the density / radiance activations
// (models/nerf.py:17-29, external/ngp.py:45-65) and their autograd.  The reference runs five
// cuBLAS SGEMMs with TF32 off and round-trips every (M,64) activation through HBM; here a
// CTA owns 128 samples (thread t <-> sample row t <-> TMEM lane t), activations go
// registers -> bf16 hi/lo operand tiles in shared memory -> tcgen05.mma -> TMEM ->
// registers, and only enc (128 B), sigma and rgb touch HBM.
//
// fp32 fidelity: every operand is split x = hi + lo into two bf16 tiles and each layer issues
// three MMAs (hi*hi + lo*hi + hi*lo) accumulated in fp32 TMEM: ~2^-17 relative per product,
// measured against the fp32 reference in tests/test_gpu_mlp_tc.py (single-pass TF32 was
// measured at 2e-3 on the head-weight gradients, outside the 1e-3 parity bound).
#include "den_common.cuh"
#include "den_field.cuh"
#include "den_tc.cuh"

namespace den {

constexpr int kTcThreads = 128;
constexpr int kTcTile = 128;
constexpr int kOutN = 16;            // radiance GEMM N (C padded to the UMMA minimum for M=128)
constexpr uint32_t kTmemCols = 64;

// shared-memory plan (bytes)
struct TcSmem {
    // weights, bf16 hi / lo, K-major UMMA tiles
    static constexpr int wb1 = 0;                                  // (64, 32)
    static constexpr int wb2 = wb1 + 2 * kWidth * kEncDim * 2;     // (16, 64)
    static constexpr int w1 = wb2 + 2 * kBaseOut * kWidth * 2;     // (64, 32)
    static constexpr int w2 = w1 + 2 * kWidth * kHeadIn * 2;       // (64, 64)
    static constexpr int w3 = w2 + 2 * kWidth * kWidth * 2;        // (16, 64)
    static constexpr int bias = w3 + 2 * kOutN * kWidth * 2;       // fp32: 64 + 16 + 64 + 64 + 16
    static constexpr int a32 = bias + (kWidth * 3 + kBaseOut + kOutN) * 4;   // (128, 32) hi / lo
    static constexpr int a64 = a32 + 2 * kTcTile * kEncDim * 2;              // (128, 64) hi / lo
    static constexpr int bar = a64 + 2 * kTcTile * kWidth * 2;
    static constexpr int tmem_ptr = bar + 8;
    static constexpr int total = tmem_ptr + 8;
};
static_assert(TcSmem::a32 % 128 == 0 && TcSmem::a64 % 128 == 0, "operand tiles must be 128 B aligned");

template <int N, int K>
__device__ __forceinline__ void issue_layer(uint8_t* smem, uint32_t tmem_d, int a_off, int w_off,
                                            uint64_t* bar) {
    // hi tile first, lo tile right after it (same size)
    constexpr int a_bytes = kTcTile * K * 2;
    constexpr int w_bytes = N * K * 2;
    tc::gemm_split_kmajor<N, K>(tmem_d, smem + a_off, smem + a_off + a_bytes, smem + w_off,
                                smem + w_off + w_bytes, bar);
}

template <bool kFull>
__global__ void __launch_bounds__(kTcThreads, 2)
mlp_fwd_tc_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                  const float* __restrict__ enc, const float* __restrict__ rays_o,
                  const float* __restrict__ rays_d, const int32_t* __restrict__ ray_indices,
                  const float* __restrict__ t_starts, const float* __restrict__ t_ends, int64_t n,
                  float* __restrict__ sigmas, float* __restrict__ rgbs) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TcSmem::bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TcSmem::tmem_ptr);
    float* s_bias = reinterpret_cast<float*>(smem + TcSmem::bias);
    float* s_bb1 = s_bias;
    float* s_bb2 = s_bb1 + kWidth;
    float* s_b1 = s_bb2 + kBaseOut;
    float* s_b2 = s_b1 + kWidth;
    float* s_b3 = s_b2 + kWidth;

    // ---- one-time setup: weights -> bf16 hi/lo operand tiles, barrier, TMEM -------------
    const int enc_dim = f.grid.n_levels * 2;
    const int C = f.channels;
    tc::load_weight_split(smem + TcSmem::wb1, smem + TcSmem::wb1 + kWidth * kEncDim * 2, p.wb1,
                          kWidth, enc_dim, kWidth, kEncDim);
    tc::load_weight_split(smem + TcSmem::wb2, smem + TcSmem::wb2 + kBaseOut * kWidth * 2, p.wb2,
                          kBaseOut, kWidth, kBaseOut, kWidth);
    load_padded(s_bb1, p.bb1, kWidth, kWidth);
    load_padded(s_bb2, p.bb2, kBaseOut, kBaseOut);
    if (kFull) {
        tc::load_weight_split(smem + TcSmem::w1, smem + TcSmem::w1 + kWidth * kHeadIn * 2, p.w1,
                              kWidth, kShDim + kGeo, kWidth, kHeadIn);
        tc::load_weight_split(smem + TcSmem::w2, smem + TcSmem::w2 + kWidth * kWidth * 2, p.w2,
                              kWidth, kWidth, kWidth, kWidth);
        tc::load_weight_split(smem + TcSmem::w3, smem + TcSmem::w3 + kOutN * kWidth * 2, p.w3, C,
                              kWidth, kOutN, kWidth);
        load_padded(s_b1, p.b1, kWidth, kWidth);
        load_padded(s_b2, p.b2, kWidth, kWidth);
        load_padded(s_b3, p.b3, C, kOutN);
    }
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemCols);
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;

    uint8_t* a32_hi = smem + TcSmem::a32;
    uint8_t* a32_lo = a32_hi + kTcTile * kEncDim * 2;
    uint8_t* a64_hi = smem + TcSmem::a64;
    uint8_t* a64_lo = a64_hi + kTcTile * kWidth * 2;

    const int64_t n_tiles = (n + kTcTile - 1) / kTcTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * kTcTile + tid;
        const bool valid = i < n;

        // ---- stage 0: this thread's encoding row -> A32 ------------------------------------
        float x32[kEncDim];
#pragma unroll
        for (int k = 0; k < kEncDim; ++k) x32[k] = 0.f;
        float dir[3] = {0.f, 0.f, 1.f};
        bool inside = false;
        if (valid) {
            const float4* row = reinterpret_cast<const float4*>(enc + i * enc_dim);
#pragma unroll
            for (int q = 0; q < kEncDim / 4; ++q) {
                if (4 * q < enc_dim) {
                    const float4 v = __ldg(row + q);
                    x32[4 * q] = v.x; x32[4 * q + 1] = v.y; x32[4 * q + 2] = v.z; x32[4 * q + 3] = v.w;
                }
            }
            const int64_t r = ray_indices[i];
            const float tm = t_starts[i] + t_ends[i];
            float pos[3], u[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                dir[d] = __ldg(rays_d + 3 * r + d);
                pos[d] = __ldg(rays_o + 3 * r + d) + (dir[d] * tm) * 0.5f;
            }
            inside = contract_position(f, pos, u);
        }
        tc::store_row_split<kEncDim>(a32_hi, a32_lo, tid, x32);
        tc::fence_smem_to_async_proxy();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- base layer 1: (128,32) x (64,32)^T ---------------------------------------------
        if (tid == 0) {
            tc::tc_fence_after_sync();
            issue_layer<kWidth, kEncDim>(smem, tmem_base, TcSmem::a32, TcSmem::wb1, bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        float h[kWidth];
        tc::tmem_ld<kWidth>(tmem_lane, h);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) h[j] = hidden_act(f.hidden_act, h[j] + s_bb1[j]);
        tc::store_row_split<kWidth>(a64_hi, a64_lo, tid, h);
        tc::fence_smem_to_async_proxy();
        tc::tc_fence_before_sync();
        __syncthreads();

        // ---- base layer 2: (128,64) x (16,64)^T ---------------------------------------------
        if (tid == 0) {
            tc::tc_fence_after_sync();
            issue_layer<kBaseOut, kWidth>(smem, tmem_base, TcSmem::a64, TcSmem::wb2, bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        float y[kBaseOut];
        tc::tmem_ld<kBaseOut>(tmem_lane, y);
#pragma unroll
        for (int j = 0; j < kBaseOut; ++j) y[j] += s_bb2[j];
        if (valid) sigmas[i] = inside ? density_act(f.density_act, y[0]) : 0.f;
        if (!kFull) {
            tc::tc_fence_before_sync();
            __syncthreads();
            continue;
        }

        // ---- head: [SH(dir) | geo | 0] -> 64 -> 64 -> C ---------------------------------------
        sh_degree4(dir, x32);
#pragma unroll
        for (int j = 0; j < kGeo; ++j) x32[kShDim + j] = y[1 + j];
        x32[kHeadIn - 1] = 0.f;
        tc::store_row_split<kHeadIn>(a32_hi, a32_lo, tid, x32);
        tc::fence_smem_to_async_proxy();
        tc::tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after_sync();
            issue_layer<kWidth, kHeadIn>(smem, tmem_base, TcSmem::a32, TcSmem::w1, bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        tc::tmem_ld<kWidth>(tmem_lane, h);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) h[j] = hidden_act(f.hidden_act, h[j] + s_b1[j]);
        tc::store_row_split<kWidth>(a64_hi, a64_lo, tid, h);
        tc::fence_smem_to_async_proxy();
        tc::tc_fence_before_sync();
        __syncthreads();

        if (tid == 0) {
            tc::tc_fence_after_sync();
            issue_layer<kWidth, kWidth>(smem, tmem_base, TcSmem::a64, TcSmem::w2, bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        tc::tmem_ld<kWidth>(tmem_lane, h);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) h[j] = hidden_act(f.hidden_act, h[j] + s_b2[j]);
        // the previous MMA has finished reading A64: safe to overwrite it
        tc::store_row_split<kWidth>(a64_hi, a64_lo, tid, h);
        tc::fence_smem_to_async_proxy();
        tc::tc_fence_before_sync();
        __syncthreads();

        if (tid == 0) {
            tc::tc_fence_after_sync();
            issue_layer<kOutN, kWidth>(smem, tmem_base, TcSmem::a64, TcSmem::w3, bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
        float out[kOutN];
        tc::tmem_ld<kOutN>(tmem_lane, out);
        if (valid) {
            for (int c = 0; c < C; ++c) rgbs[i * C + c] = radiance_act(f.radiance_act, out[c] + s_b3[c]);
        }
        tc::tc_fence_before_sync();
        __syncthreads();
    }

    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kTmemCols);
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

}  // namespace den

extern "C" {

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
    const int grid = grid_for((n + kTcTile - 1) / kTcTile, 1, 2);
    const size_t smem = TcSmem::total;
    if (full) {
        cudaFuncSetAttribute(mlp_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mlp_fwd_tc_kernel<true><<<grid, kTcThreads, smem, as_stream(stream)>>>(
            *f, *p, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, n, sigmas, rgbs);
    } else {
        cudaFuncSetAttribute(mlp_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mlp_fwd_tc_kernel<false><<<grid, kTcThreads, smem, as_stream(stream)>>>(
            *f, *p, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, n, sigmas, nullptr);
    }
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
