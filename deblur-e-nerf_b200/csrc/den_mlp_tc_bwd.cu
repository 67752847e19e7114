// Backward of the tensor-core MLP: forward recompute + input gradient + weight gradients,
// fused per 128-sample tile (tcgen05.mma, TMEM accumulators; sm_100a).
//
// Replaces the autograd of MLP.forward / SHEncoder / activations under
// NGPradianceField.forward (external/ngp.py:269-280) — in the reference ~30 cuBLAS and
// elementwise launches that read and write every (M,64) activation twice.  Here nothing but
// enc (128 B/sample in), the two upstream gradients and dL/denc (128 B/sample out) touches
// HBM: the forward activations are recomputed into shared-memory operand tiles, which are
// then consumed three ways without being rewritten —
//   K-major  A operand        : next layer forward,  dX = dY * W
//   MN-major A/B operand      : dW += dY^T * X   (K = the 128 samples of the tile)
// Weight-gradient accumulators (64 x N, fp32) stay in TMEM across all tiles of a CTA and are
// flushed once with atomics; bias gradients of the 64-wide layers come from one extra MMA
// against a tile of ones.
//
// Thread layout: see den_mlp_tc.cuh (S = 4: 16 epilogue warps + 1 MMA warp, 1 CTA per SM —
// the five activation tiles + gradient tiles + weights fill 213 KB of shared memory).
#include "den_mlp_tc.cuh"

namespace den {

using namespace mlp;

constexpr int kBwdS = 4;
constexpr int kBwdCols = kWidth / kBwdS;                  // 16 accumulator columns per thread
constexpr int kBwdEpiThreads = 128 * kBwdS;
constexpr int kBwdThreads = kBwdEpiThreads + 32;
constexpr uint32_t kBwdTmemCols = 256;

// TMEM column plan
constexpr uint32_t kColScratch = 0;      // 64: forward / dX results (M = 128)
constexpr uint32_t kColDW1 = 64;         // 32: dW1    (out 64 x in 32)
constexpr uint32_t kColDW2 = 96;         // 64: dW2    (64 x 64)
constexpr uint32_t kColDWb1 = 160;       // 32: dWb1   (64 x 32)
constexpr uint32_t kColDWb2T = 192;      // 16: dWb2^T (in 64 x out 16)
constexpr uint32_t kColDW3T = 208;       // 16: dW3^T  (in 64 x out 16)
constexpr uint32_t kColDB1 = 224;        // 8
constexpr uint32_t kColDB2 = 232;        // 8
constexpr uint32_t kColDBb1 = 240;       // 8

struct BwdSmem {
    static constexpr int w3f = Weights::end;                               // fp32 W3 rows (4 x 64)
    static constexpr int small_acc = w3f + 4 * kWidth * 4;                 // dbb2[16] | db3[4] | pad
    static constexpr int ones = (small_acc + 32 * 4 + 127) / 128 * 128;    // (8, 128) bf16 ones
    static constexpr int t_enc = ones + 8 * kTile * 2;
    static constexpr int t_hb = t_enc + Tile<kTile, kEncDim>::bytes;
    static constexpr int t_in1 = t_hb + Tile<kTile, kWidth>::bytes;
    static constexpr int t_h1 = t_in1 + Tile<kTile, kHeadIn>::bytes;
    static constexpr int t_h2 = t_h1 + Tile<kTile, kWidth>::bytes;
    static constexpr int t_d64 = t_h2 + Tile<kTile, kWidth>::bytes;
    static constexpr int t_d16 = t_d64 + Tile<kTile, kWidth>::bytes;
    static constexpr int bar = t_d16 + Tile<kTile, kBaseOut>::bytes;
    static constexpr int tmem_ptr = bar + 8;
    static constexpr int total = tmem_ptr + 8;
};
static_assert(BwdSmem::t_enc % 128 == 0, "tile alignment");
static_assert(BwdSmem::total <= 227 * 1024, "shared-memory plan exceeds 227 KB");

__global__ void __launch_bounds__(kBwdThreads, 1)
mlp_bwd_tc_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                  const __grid_constant__ den_field_grads g, const float* __restrict__ enc,
                  const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                  const int32_t* __restrict__ ray_indices, const float* __restrict__ t_starts,
                  const float* __restrict__ t_ends, const float* __restrict__ d_sigmas,
                  const float* __restrict__ d_rgbs, int64_t n, float* __restrict__ d_enc,
                  float* __restrict__ d_dirs) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + BwdSmem::bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BwdSmem::tmem_ptr);
    const float* s_bb1 = reinterpret_cast<const float*>(smem + Weights::bias);
    const float* s_bb2 = s_bb1 + kWidth;
    const float* s_b1 = s_bb2 + kBaseOut;
    const float* s_b2 = s_b1 + kWidth;
    const float* s_b3 = s_b2 + kWidth;
    float* s_w3f = reinterpret_cast<float*>(smem + BwdSmem::w3f);
    float* s_small = reinterpret_cast<float*>(smem + BwdSmem::small_acc);   // [0,16) dbb2, [16,20) db3

    const Tile<kWidth, kEncDim> Wb1(smem + Weights::wb1);
    const Tile<kBaseOut, kWidth> Wb2(smem + Weights::wb2);
    const Tile<kWidth, kHeadIn> W1(smem + Weights::w1);
    const Tile<kWidth, kWidth> W2(smem + Weights::w2);
    const Tile<kOutN, kWidth> W3(smem + Weights::w3);
    const Tile<kTile, kEncDim> Tenc(smem + BwdSmem::t_enc);
    const Tile<kTile, kWidth> Thb(smem + BwdSmem::t_hb);
    const Tile<kTile, kHeadIn> Tin1(smem + BwdSmem::t_in1);
    const Tile<kTile, kWidth> Th1(smem + BwdSmem::t_h1);
    const Tile<kTile, kWidth> Th2(smem + BwdSmem::t_h2);
    const Tile<kTile, kWidth> Td64(smem + BwdSmem::t_d64);
    const Tile<kTile, kBaseOut> Td16(smem + BwdSmem::t_d16);
    uint8_t* ones = smem + BwdSmem::ones;

    const int enc_dim = f.grid.n_levels * 2;
    const int C = f.channels;

    // ---- setup ---------------------------------------------------------------------------
    load_all_weights(smem, f, p, true);
    for (int i = tid; i < 4 * kWidth; i += blockDim.x)
        s_w3f[i] = (i / kWidth) < C ? __ldg(p.w3 + i) : 0.f;
    if (tid < 32) s_small[tid] = 0.f;
    for (int i = tid; i < 8 * kTile; i += blockDim.x)
        reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16_rn(1.0f);
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kBwdTmemCols);
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t D = tmem_base + kColScratch;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const bool have_tiles = (int64_t)blockIdx.x < n_tiles;

    if (warp == kBwdEpiThreads / 32) {
        // ===================== MMA warp: ten issue rounds per tile =====================
        bool acc = false;       // weight-gradient accumulators already hold earlier tiles
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#define DEN_ISSUE(...)                                                        \
    __syncthreads();                                                          \
    { tc::tc_fence_after_sync(); __VA_ARGS__; tc::mma_commit(bar); }
            DEN_ISSUE(mma_fwd<kWidth, kEncDim>(D, Tenc, Wb1))
            DEN_ISSUE(mma_fwd<kBaseOut, kWidth>(D, Thb, Wb2))
            DEN_ISSUE(mma_fwd<kWidth, kHeadIn>(D, Tin1, W1))
            DEN_ISSUE(mma_fwd<kWidth, kWidth>(D, Th1, W2))
            DEN_ISSUE(mma_fwd<kOutN, kWidth>(D, Th2, W3))
            DEN_ISSUE(mma_dw<kBaseOut>(tmem_base + kColDW3T, Th2, Td16, acc);      // dW3^T += h2^T d3
                      mma_dw<kWidth>(tmem_base + kColDW2, Td64, Th1, acc);         // dW2  += d2^T h1
                      mma_colsum(tmem_base + kColDB2, Td64, ones, acc);            // db2
                      mma_dx<kWidth, kWidth>(D, Td64, W2))                         // dh1 = d2 W2
            DEN_ISSUE(mma_dw<kHeadIn>(tmem_base + kColDW1, Td64, Tin1, acc);       // dW1 += d1^T in1
                      mma_colsum(tmem_base + kColDB1, Td64, ones, acc);            // db1
                      mma_dx<kHeadIn, kWidth>(D, Td64, W1))                        // din1 = d1 W1
            DEN_ISSUE(mma_dw<kBaseOut>(tmem_base + kColDWb2T, Thb, Td16, acc);     // dWb2^T += hb^T dy
                      mma_dx<kWidth, kBaseOut>(D, Td16, Wb2))                      // dhb = dy Wb2
            DEN_ISSUE(mma_dw<kEncDim>(tmem_base + kColDWb1, Td64, Tenc, acc);      // dWb1 += db1^T enc
                      mma_colsum(tmem_base + kColDBb1, Td64, ones, acc);           // dbb1
                      mma_dx<kEncDim, kWidth>(D, Td64, Wb1))                       // denc = db1 Wb1
#undef DEN_ISSUE
            acc = true;
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3, cg = warp >> 2;
        const int row = q * 32 + lane;
        const int col0 = kBwdCols * cg;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t i = tile * kTile + row;
            const bool valid = i < n;

            // ---- forward recompute -----------------------------------------------------------
            float x8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x8[k] = 0.f;
            float dir[3] = {0.f, 0.f, 1.f};
            bool inside = false;
            float g_sigma = 0.f, g_rgb[3] = {0.f, 0.f, 0.f};
            if (valid) {
                if (8 * cg < enc_dim) {
                    const float4* src = reinterpret_cast<const float4*>(enc + i * enc_dim + 8 * cg);
                    const float4 a = __ldg(src), b = __ldg(src + 1);
                    x8[0] = a.x; x8[1] = a.y; x8[2] = a.z; x8[3] = a.w;
                    x8[4] = b.x; x8[5] = b.y; x8[6] = b.z; x8[7] = b.w;
                }
                if (cg <= 1) {
                    const int64_t r = ray_indices[i];
                    const float tm = t_starts[i] + t_ends[i];
                    float pos[3], u[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        dir[d] = __ldg(rays_d + 3 * r + d);
                        pos[d] = __ldg(rays_o + 3 * r + d) + (dir[d] * tm) * 0.5f;
                    }
                    if (cg == 0) {
                        inside = contract_position(f, pos, u);
                        g_sigma = d_sigmas[i];
                    }
                }
                for (int c = 0; c < C; ++c) g_rgb[c] = d_rgbs[i * C + c];
            }
            store_cols<8, kEncDim>(Tenc, row, 8 * cg, x8);
            publish();

            float h[kBwdCols];
            await(bar, phase);                                              // hb
            tmem_ld_cols<kBwdCols>(tmem_lane + kColScratch + col0, h);
bias_hidden_act<kBwdCols>(f.hidden_act, h, s_bb1 + col0);
            store_cols<kBwdCols, kWidth>(Thb, row, col0, h);
            publish();

            float raw = 0.f;
            await(bar, phase);                                              // y -> [SH | geo | 0]
            if (cg == 0) {
                float y[kBaseOut];
                tmem_ld_cols<kBaseOut>(tmem_lane + kColScratch, y);
#pragma unroll
                for (int j = 0; j < kBaseOut; ++j) y[j] += s_bb2[j];
                raw = y[0];
#pragma unroll
                for (int j = 0; j < kGeo; ++j) h[j] = y[1 + j];
                h[15] = 0.f;
                store_cols<16, kHeadIn>(Tin1, row, 16, h);
            } else if (cg == 1) {
                sh_degree4(dir, h);
                store_cols<16, kHeadIn>(Tin1, row, 0, h);
            }
            publish();

            await(bar, phase);                                              // h1
            tmem_ld_cols<kBwdCols>(tmem_lane + kColScratch + col0, h);
bias_hidden_act<kBwdCols>(f.hidden_act, h, s_b1 + col0);
            store_cols<kBwdCols, kWidth>(Th1, row, col0, h);
            publish();

            await(bar, phase);                                              // h2 (kept in registers)
            tmem_ld_cols<kBwdCols>(tmem_lane + kColScratch + col0, h);
bias_hidden_act<kBwdCols>(f.hidden_act, h, s_b2 + col0);
            store_cols<kBwdCols, kWidth>(Th2, row, col0, h);
            publish();

            // ---- output layer backward (SIMT: C <= 3 rows) -----------------------------------
            await(bar, phase);
            float d16[kBaseOut];
            {
                float z3[kOutN];
                tmem_ld_cols<kOutN>(tmem_lane + kColScratch, z3);
#pragma unroll
                for (int j = 0; j < kBaseOut; ++j) d16[j] = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (c < C) d16[c] = g_rgb[c] * radiance_act_grad(f.radiance_act, z3[c] + s_b3[c]);
            }
            float dl[kBwdCols];
#pragma unroll
            for (int j = 0; j < kBwdCols; ++j) {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) a = fmaf(d16[c], s_w3f[c * kWidth + col0 + j], a);
                dl[j] = a * hidden_act_grad_from_out(f.hidden_act, h[j]);
            }
            store_cols<kBwdCols, kWidth>(Td64, row, col0, dl);
            if (cg == 0) {
                store_cols<kBaseOut, kBaseOut>(Td16, row, 0, d16);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float s = warp_sum(d16[c]);
                    if (lane == 0 && c < C) atomicAdd(&s_small[16 + c], s);
                }
            }
            publish();

            await(bar, phase);                                              // dh1 -> d1
            tmem_ld_cols<kBwdCols>(tmem_lane + kColScratch + col0, dl);
            load_cols<kBwdCols, kWidth>(Th1, row, col0, h);
mul_hidden_act_grad<kBwdCols>(f.hidden_act, dl, h);
            store_cols<kBwdCols, kWidth>(Td64, row, col0, dl);
            publish();

            await(bar, phase);                                              // din1 -> dy
            if (cg == 0) {
                float dgeo[16];
                tmem_ld_cols<16>(tmem_lane + kColScratch + kShDim, dgeo);
                d16[0] = inside ? g_sigma * density_act_grad(f.density_act, raw) : 0.f;
#pragma unroll
                for (int j = 0; j < kGeo; ++j) d16[1 + j] = dgeo[j];
                store_cols<kBaseOut, kBaseOut>(Td16, row, 0, d16);
#pragma unroll
                for (int j = 0; j < kBaseOut; ++j) {
                    const float s = warp_sum(d16[j]);
                    if (lane == 0) atomicAdd(&s_small[j], s);
                }
            } else if (cg == 1 && d_dirs != nullptr) {
                // dL/d(view direction) through the SH encoding (only the tau path needs it)
                float dsh[16], dd[3];
                tmem_ld_cols<16>(tmem_lane + kColScratch, dsh);
                sh_degree4_grad(dir, dsh, dd);
                if (valid) { d_dirs[3 * i] = dd[0]; d_dirs[3 * i + 1] = dd[1]; d_dirs[3 * i + 2] = dd[2]; }
            }
            publish();

            await(bar, phase);                                              // dhb -> db1
            tmem_ld_cols<kBwdCols>(tmem_lane + kColScratch + col0, dl);
            load_cols<kBwdCols, kWidth>(Thb, row, col0, h);
mul_hidden_act_grad<kBwdCols>(f.hidden_act, dl, h);
            store_cols<kBwdCols, kWidth>(Td64, row, col0, dl);
            publish();

            await(bar, phase);                                              // denc -> HBM
            {
                float de[8];
                tmem_ld_cols<8>(tmem_lane + kColScratch + 8 * cg, de);
                if (valid && 8 * cg < enc_dim) {
                    float4* out = reinterpret_cast<float4*>(d_enc + i * enc_dim + 8 * cg);
                    out[0] = make_float4(de[0], de[1], de[2], de[3]);
                    out[1] = make_float4(de[4], de[5], de[6], de[7]);
                }
            }
            tc::tc_fence_before_sync();
        }
    }

    // ---- flush the weight-gradient accumulators (M = 64: row m in lane m%16 + 32*(m/16)) ------
    tc::tc_fence_before_sync();
    __syncthreads();
    if (have_tiles && warp < 4) {
        tc::tc_fence_after_sync();
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int row = warp * 16 + lane;              // valid for lane < 16
        const bool owner = lane < 16;
        float v[16];
        for (int c0 = 0; c0 < kHeadIn; c0 += 16) {     // dW1 (64, 31)
            tmem_ld_cols<16>(tmem_lane + kColDW1 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < kShDim + kGeo) atomicAdd(g.w1 + row * (kShDim + kGeo) + c0 + j, v[j]);
        }
        for (int c0 = 0; c0 < kWidth; c0 += 16) {      // dW2 (64, 64)
            tmem_ld_cols<16>(tmem_lane + kColDW2 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j) atomicAdd(g.w2 + row * kWidth + c0 + j, v[j]);
        }
        for (int c0 = 0; c0 < kEncDim; c0 += 16) {     // dWb1 (64, enc_dim)
            tmem_ld_cols<16>(tmem_lane + kColDWb1 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < enc_dim) atomicAdd(g.wb1 + row * enc_dim + c0 + j, v[j]);
        }
        tmem_ld_cols<16>(tmem_lane + kColDWb2T, v);    // dWb2^T (in 64, out 16)
        if (owner)
            for (int j = 0; j < kBaseOut; ++j) atomicAdd(g.wb2 + j * kWidth + row, v[j]);
        tmem_ld_cols<16>(tmem_lane + kColDW3T, v);     // dW3^T (in 64, out C)
        if (owner)
            for (int j = 0; j < C; ++j) atomicAdd(g.w3 + j * kWidth + row, v[j]);
        tmem_ld_cols<16>(tmem_lane + kColDB1, v);      // columns 224..239: db1 | db2
        if (owner) {
            atomicAdd(g.b1 + row, v[0]);
            atomicAdd(g.b2 + row, v[8]);
        }
        tmem_ld_cols<16>(tmem_lane + kColDBb1, v);
        if (owner) atomicAdd(g.bb1 + row, v[0]);
        if (tid < kBaseOut) atomicAdd(g.bb2 + tid, s_small[tid]);
        if (tid < C) atomicAdd(g.b3 + tid, s_small[16 + tid]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kBwdTmemCols);
}

}  // namespace den

extern "C" int den_mlp_bwd(const den_field_desc* f, const den_field_params* p,
                           const den_field_grads* g, const float* enc, const float* rays_o,
                           const float* rays_d, const int32_t* ray_indices, const float* t_starts,
                           const float* t_ends, const float* d_sigmas, const float* d_rgbs,
                           int64_t n, float* d_enc, float* d_dirs, void* stream) {
    using namespace den;
    int rc = check_field(f, p, true);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(g && g->wb1 && g->bb1 && g->wb2 && g->bb2 && g->w1 && g->b1 && g->w2 && g->b2 &&
                      g->w3 && g->b3,
                  "null gradient pointer");
    DEN_CHECK_ARG(enc && rays_o && rays_d && ray_indices && t_starts && t_ends && d_sigmas &&
                      d_rgbs && d_enc,
                  "null pointer");
    DEN_CHECK_ARG((f->grid.n_levels * 2) % 8 == 0, "encoding width must be a multiple of 8");
    const int grid = grid_for((n + kTile - 1) / kTile, 1, 1);
    cudaFuncSetAttribute(mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)BwdSmem::total);
    mlp_bwd_tc_kernel<<<grid, kBwdThreads, BwdSmem::total, as_stream(stream)>>>(
        *f, *p, *g, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, d_sigmas, d_rgbs, n, d_enc, d_dirs);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}
