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
#include "den_common.cuh"
#include "den_field.cuh"
#include "den_tc.cuh"

namespace den {

constexpr int kBwdThreads = 128;
constexpr int kBwdTile = 128;
constexpr int kOutNb = 16;
constexpr uint32_t kBwdTmemCols = 256;

// TMEM column plan
constexpr uint32_t kColScratch = 0;      // 64: forward / dX results (M = 128)
constexpr uint32_t kColDW1 = 64;         // 32: dW1   (out 64 x in 32)
constexpr uint32_t kColDW2 = 96;         // 64: dW2   (64 x 64)
constexpr uint32_t kColDWb1 = 160;       // 32: dWb1  (64 x 32)
constexpr uint32_t kColDWb2T = 192;      // 16: dWb2^T (in 64 x out 16)
constexpr uint32_t kColDW3T = 208;       // 16: dW3^T  (in 64 x out 16)
constexpr uint32_t kColDB1 = 224;        // 8
constexpr uint32_t kColDB2 = 232;        // 8
constexpr uint32_t kColDBb1 = 240;       // 8

struct BwdSmem {
    static constexpr int wb1 = 0;
    static constexpr int wb2 = wb1 + 2 * kWidth * kEncDim * 2;
    static constexpr int w1 = wb2 + 2 * kBaseOut * kWidth * 2;
    static constexpr int w2 = w1 + 2 * kWidth * kHeadIn * 2;
    static constexpr int w3 = w2 + 2 * kWidth * kWidth * 2;
    static constexpr int bias = w3 + 2 * kOutNb * kWidth * 2;                      // 224 floats
    static constexpr int w3f = bias + (kWidth * 3 + kBaseOut + kOutNb) * 4;        // fp32 W3 (4 x 64)
    static constexpr int small_acc = w3f + 4 * kWidth * 4;                         // dbb2[16] + db3[4] (+pad)
    static constexpr int ones = small_acc + 32 * 4;                                // (8, 128) bf16
    static constexpr int t_enc = ones + 8 * kBwdTile * 2;
    static constexpr int t_hb = t_enc + 2 * kBwdTile * kEncDim * 2;
    static constexpr int t_in1 = t_hb + 2 * kBwdTile * kWidth * 2;
    static constexpr int t_h1 = t_in1 + 2 * kBwdTile * kHeadIn * 2;
    static constexpr int t_h2 = t_h1 + 2 * kBwdTile * kWidth * 2;
    static constexpr int t_d64 = t_h2 + 2 * kBwdTile * kWidth * 2;
    static constexpr int t_d16 = t_d64 + 2 * kBwdTile * kWidth * 2;
    static constexpr int bar = t_d16 + 2 * kBwdTile * kBaseOut * 2;
    static constexpr int tmem_ptr = bar + 8;
    static constexpr int total = tmem_ptr + 8;
};
static_assert(BwdSmem::ones % 128 == 0 && BwdSmem::t_enc % 128 == 0, "tile alignment");
static_assert(BwdSmem::total <= 227 * 1024, "shared-memory plan exceeds 227 KB");

// tile = hi half followed by lo half
template <int ROWS, int COLS>
struct Tile {
    uint8_t* hi;
    uint8_t* lo;
    __device__ explicit Tile(uint8_t* base) : hi(base), lo(base + ROWS * COLS * 2) {}
    static constexpr uint32_t row_group = (COLS / 8) * 128;      // stride between groups of 8 rows
};

// D[128 x N] = A[128 x K] * W^T, W tile (N, K) K-major — forward
template <int N, int K>
__device__ __forceinline__ void mma_fwd(uint32_t tmem_d, const Tile<kBwdTile, K>& a,
                                        const Tile<N, K>& w) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(128, N, false, false);
    const uint32_t aa[3] = {tc::smem_u32(a.hi), tc::smem_u32(a.lo), tc::smem_u32(a.hi)};
    const uint32_t bb[3] = {tc::smem_u32(w.hi), tc::smem_u32(w.hi), tc::smem_u32(w.lo)};
    bool acc = false;
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
            tc::mma_bf16(tmem_d, tc::smem_desc(aa[t] + ks * 256, 128, Tile<kBwdTile, K>::row_group),
                         tc::smem_desc(bb[t] + ks * 256, 128, Tile<N, K>::row_group), idesc, acc);
            acc = true;
        }
}

// D[128 x N] = dY[128 x K] * W, W tile (K rows = out, N feats = in) read MN-major — dX
template <int N, int K>
__device__ __forceinline__ void mma_dx(uint32_t tmem_d, const Tile<kBwdTile, K>& dy,
                                       const Tile<K, N>& w) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(128, N, false, true);
    constexpr uint32_t wrg = Tile<K, N>::row_group;
    const uint32_t aa[3] = {tc::smem_u32(dy.hi), tc::smem_u32(dy.lo), tc::smem_u32(dy.hi)};
    const uint32_t bb[3] = {tc::smem_u32(w.hi), tc::smem_u32(w.hi), tc::smem_u32(w.lo)};
    bool acc = false;
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
            tc::mma_bf16(tmem_d, tc::smem_desc(aa[t] + ks * 256, 128, Tile<kBwdTile, K>::row_group),
                         tc::smem_desc(bb[t] + ks * 2 * wrg, wrg, 128), idesc, acc);
            acc = true;
        }
}

// D[64 x N] (+)= A^T * B over the 128 rows of the tile: A tile (128, 64), B tile (128, N),
// both read MN-major — dW
template <int N>
__device__ __forceinline__ void mma_dw(uint32_t tmem_d, const Tile<kBwdTile, 64>& a,
                                       const Tile<kBwdTile, N>& b, bool accumulate) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(64, N, true, true);
    constexpr uint32_t arg = Tile<kBwdTile, 64>::row_group;
    constexpr uint32_t brg = Tile<kBwdTile, N>::row_group;
    const uint32_t aa[3] = {tc::smem_u32(a.hi), tc::smem_u32(a.lo), tc::smem_u32(a.hi)};
    const uint32_t bb[3] = {tc::smem_u32(b.hi), tc::smem_u32(b.hi), tc::smem_u32(b.lo)};
    bool acc = accumulate;
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int ks = 0; ks < kBwdTile / 16; ++ks) {
            tc::mma_bf16(tmem_d, tc::smem_desc(aa[t] + ks * 2 * arg, arg, 128),
                         tc::smem_desc(bb[t] + ks * 2 * brg, brg, 128), idesc, acc);
            acc = true;
        }
}

// D[64 x 8] (+)= A^T * ones : column sums of the (128, 64) tile (bias gradient)
__device__ __forceinline__ void mma_colsum(uint32_t tmem_d, const Tile<kBwdTile, 64>& a,
                                           const uint8_t* ones, bool accumulate) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(64, 8, true, false);
    constexpr uint32_t arg = Tile<kBwdTile, 64>::row_group;
    const uint32_t aa[2] = {tc::smem_u32(a.hi), tc::smem_u32(a.lo)};
    const uint32_t ob = tc::smem_u32(ones);
    bool acc = accumulate;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int ks = 0; ks < kBwdTile / 16; ++ks) {
            tc::mma_bf16(tmem_d, tc::smem_desc(aa[t] + ks * 2 * arg, arg, 128),
                         tc::smem_desc(ob + ks * 256, 128, (kBwdTile / 8) * 128), idesc, acc);
            acc = true;
        }
}

// read this thread's row back from a hi/lo tile
template <int K>
__device__ __forceinline__ void load_row(const uint8_t* hi, const uint8_t* lo, int r, float (&x)[K]) {
#pragma unroll
    for (int c = 0; c < K / 8; ++c) {
        const uint32_t off = tc::chunk_offset(r, c, K);
        const uint4 h = *reinterpret_cast<const uint4*>(hi + off);
        const uint4 l = *reinterpret_cast<const uint4*>(lo + off);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            x[8 * c + 2 * q] = __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
            x[8 * c + 2 * q + 1] = __uint_as_float(hw[q] & 0xffff0000u) + __uint_as_float(lw[q] & 0xffff0000u);
        }
    }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
mlp_bwd_tc_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                  const __grid_constant__ den_field_grads g, const float* __restrict__ enc,
                  const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                  const int32_t* __restrict__ ray_indices, const float* __restrict__ t_starts,
                  const float* __restrict__ t_ends, const float* __restrict__ d_sigmas,
                  const float* __restrict__ d_rgbs, int64_t n, float* __restrict__ d_enc) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + BwdSmem::bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BwdSmem::tmem_ptr);
    float* s_bb1 = reinterpret_cast<float*>(smem + BwdSmem::bias);
    float* s_bb2 = s_bb1 + kWidth;
    float* s_b1 = s_bb2 + kBaseOut;
    float* s_b2 = s_b1 + kWidth;
    float* s_b3 = s_b2 + kWidth;
    float* s_w3f = reinterpret_cast<float*>(smem + BwdSmem::w3f);
    float* s_small = reinterpret_cast<float*>(smem + BwdSmem::small_acc);   // [0,16) dbb2, [16,20) db3

    const Tile<kWidth, kEncDim> Wb1(smem + BwdSmem::wb1);
    const Tile<kBaseOut, kWidth> Wb2(smem + BwdSmem::wb2);
    const Tile<kWidth, kHeadIn> W1(smem + BwdSmem::w1);
    const Tile<kWidth, kWidth> W2(smem + BwdSmem::w2);
    const Tile<kOutNb, kWidth> W3(smem + BwdSmem::w3);
    const Tile<kBwdTile, kEncDim> Tenc(smem + BwdSmem::t_enc);
    const Tile<kBwdTile, kWidth> Thb(smem + BwdSmem::t_hb);
    const Tile<kBwdTile, kHeadIn> Tin1(smem + BwdSmem::t_in1);
    const Tile<kBwdTile, kWidth> Th1(smem + BwdSmem::t_h1);
    const Tile<kBwdTile, kWidth> Th2(smem + BwdSmem::t_h2);
    const Tile<kBwdTile, kWidth> Td64(smem + BwdSmem::t_d64);
    const Tile<kBwdTile, kBaseOut> Td16(smem + BwdSmem::t_d16);
    uint8_t* ones = smem + BwdSmem::ones;

    const int enc_dim = f.grid.n_levels * 2;
    const int C = f.channels;

    // ---- setup ---------------------------------------------------------------------------
    tc::load_weight_split(Wb1.hi, Wb1.lo, p.wb1, kWidth, enc_dim, kWidth, kEncDim);
    tc::load_weight_split(Wb2.hi, Wb2.lo, p.wb2, kBaseOut, kWidth, kBaseOut, kWidth);
    tc::load_weight_split(W1.hi, W1.lo, p.w1, kWidth, kShDim + kGeo, kWidth, kHeadIn);
    tc::load_weight_split(W2.hi, W2.lo, p.w2, kWidth, kWidth, kWidth, kWidth);
    tc::load_weight_split(W3.hi, W3.lo, p.w3, C, kWidth, kOutNb, kWidth);
    load_padded(s_bb1, p.bb1, kWidth, kWidth);
    load_padded(s_bb2, p.bb2, kBaseOut, kBaseOut);
    load_padded(s_b1, p.b1, kWidth, kWidth);
    load_padded(s_b2, p.b2, kWidth, kWidth);
    load_padded(s_b3, p.b3, C, kOutNb);
    for (int i = tid; i < 4 * kWidth; i += blockDim.x)
        s_w3f[i] = (i / kWidth) < C ? __ldg(p.w3 + i) : 0.f;
    if (tid < 32) s_small[tid] = 0.f;
    for (int i = tid; i < 8 * kBwdTile; i += blockDim.x)
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
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t D = tmem_base + kColScratch;
    uint32_t phase = 0;
    bool have_acc = false;          // weight-gradient accumulators hold data from earlier tiles

#define DEN_ROUND_BEGIN()                 \
    tc::fence_smem_to_async_proxy();      \
    tc::tc_fence_before_sync();           \
    __syncthreads();                      \
    if (tid == 0) {                       \
        tc::tc_fence_after_sync();
#define DEN_ROUND_END()                   \
        tc::mma_commit(bar);              \
    }                                     \
    tc::mbar_wait(bar, phase);            \
    phase ^= 1;                           \
    tc::tc_fence_after_sync();

    const int64_t n_tiles = (n + kBwdTile - 1) / kBwdTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * kBwdTile + tid;
        const bool valid = i < n;

        // ---- forward recompute ---------------------------------------------------------
        float x32[kEncDim];
#pragma unroll
        for (int k = 0; k < kEncDim; ++k) x32[k] = 0.f;
        float dir[3] = {0.f, 0.f, 1.f};
        bool inside = false;
        float g_sigma = 0.f, g_rgb[3] = {0.f, 0.f, 0.f};
        if (valid) {
            const float4* row = reinterpret_cast<const float4*>(enc + i * enc_dim);
#pragma unroll
            for (int q = 0; q < kEncDim / 4; ++q)
                if (4 * q < enc_dim) {
                    const float4 v = __ldg(row + q);
                    x32[4 * q] = v.x; x32[4 * q + 1] = v.y; x32[4 * q + 2] = v.z; x32[4 * q + 3] = v.w;
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
            g_sigma = d_sigmas[i];
            for (int c = 0; c < C; ++c) g_rgb[c] = d_rgbs[i * C + c];
        }
        tc::store_row_split<kEncDim>(Tenc.hi, Tenc.lo, tid, x32);
        DEN_ROUND_BEGIN()
            mma_fwd<kWidth, kEncDim>(D, Tenc, Wb1);
        DEN_ROUND_END()
        float h[kWidth];
        tc::tmem_ld<kWidth>(tmem_lane + kColScratch, h);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) h[j] = hidden_act(f.hidden_act, h[j] + s_bb1[j]);
        tc::store_row_split<kWidth>(Thb.hi, Thb.lo, tid, h);
        DEN_ROUND_BEGIN()
            mma_fwd<kBaseOut, kWidth>(D, Thb, Wb2);
        DEN_ROUND_END()
        float y[kBaseOut];
        tc::tmem_ld<kBaseOut>(tmem_lane + kColScratch, y);
#pragma unroll
        for (int j = 0; j < kBaseOut; ++j) y[j] += s_bb2[j];
        const float raw = y[0];
        sh_degree4(dir, x32);
#pragma unroll
        for (int j = 0; j < kGeo; ++j) x32[kShDim + j] = y[1 + j];
        x32[kHeadIn - 1] = 0.f;
        tc::store_row_split<kHeadIn>(Tin1.hi, Tin1.lo, tid, x32);
        DEN_ROUND_BEGIN()
            mma_fwd<kWidth, kHeadIn>(D, Tin1, W1);
        DEN_ROUND_END()
        tc::tmem_ld<kWidth>(tmem_lane + kColScratch, h);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) h[j] = hidden_act(f.hidden_act, h[j] + s_b1[j]);
        tc::store_row_split<kWidth>(Th1.hi, Th1.lo, tid, h);
        DEN_ROUND_BEGIN()
            mma_fwd<kWidth, kWidth>(D, Th1, W2);
        DEN_ROUND_END()
        tc::tmem_ld<kWidth>(tmem_lane + kColScratch, h);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) h[j] = hidden_act(f.hidden_act, h[j] + s_b2[j]);
        tc::store_row_split<kWidth>(Th2.hi, Th2.lo, tid, h);           // h = h2 stays in registers
        DEN_ROUND_BEGIN()
            mma_fwd<kOutNb, kWidth>(D, Th2, W3);
        DEN_ROUND_END()

        // ---- output layer backward (SIMT: C <= 3 rows) ---------------------------------------
        float d16[kBaseOut];
        {
            float z3[kOutNb];
            tc::tmem_ld<kOutNb>(tmem_lane + kColScratch, z3);
#pragma unroll
            for (int j = 0; j < kBaseOut; ++j) d16[j] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c < C) d16[c] = g_rgb[c] * radiance_act_grad(f.radiance_act, z3[c] + s_b3[c]);
        }
        float dl[kWidth];
#pragma unroll
        for (int j = 0; j < kWidth; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc = fmaf(d16[c], s_w3f[c * kWidth + j], acc);
            dl[j] = acc * hidden_act_grad_from_out(f.hidden_act, h[j]);
        }
        tc::store_row_split<kWidth>(Td64.hi, Td64.lo, tid, dl);
        tc::store_row_split<kBaseOut>(Td16.hi, Td16.lo, tid, d16);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float s = warp_sum(d16[c]);
            if (lane == 0 && c < C) atomicAdd(&s_small[16 + c], s);
        }
        DEN_ROUND_BEGIN()
            mma_dw<kBaseOut>(tmem_base + kColDW3T, Th2, Td16, have_acc);      // dW3^T += h2^T d3
            mma_dw<kWidth>(tmem_base + kColDW2, Td64, Th1, have_acc);         // dW2  += d2^T h1
            mma_colsum(tmem_base + kColDB2, Td64, ones, have_acc);            // db2
            mma_dx<kWidth, kWidth>(D, Td64, W2);                              // dh1 = d2 W2
        DEN_ROUND_END()
        tc::tmem_ld<kWidth>(tmem_lane + kColScratch, dl);
        load_row<kWidth>(Th1.hi, Th1.lo, tid, h);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) dl[j] *= hidden_act_grad_from_out(f.hidden_act, h[j]);
        tc::store_row_split<kWidth>(Td64.hi, Td64.lo, tid, dl);
        DEN_ROUND_BEGIN()
            mma_dw<kHeadIn>(tmem_base + kColDW1, Td64, Tin1, have_acc);       // dW1 += d1^T in1
            mma_colsum(tmem_base + kColDB1, Td64, ones, have_acc);            // db1
            mma_dx<kHeadIn, kWidth>(D, Td64, W1);                             // din1 = d1 W1
        DEN_ROUND_END()
        {
            float din1[kHeadIn];
            tc::tmem_ld<kHeadIn>(tmem_lane + kColScratch, din1);
            d16[0] = inside ? g_sigma * density_act_grad(f.density_act, raw) : 0.f;
#pragma unroll
            for (int j = 0; j < kGeo; ++j) d16[1 + j] = din1[kShDim + j];
        }
        tc::store_row_split<kBaseOut>(Td16.hi, Td16.lo, tid, d16);
#pragma unroll
        for (int j = 0; j < kBaseOut; ++j) {
            const float s = warp_sum(d16[j]);
            if (lane == 0) atomicAdd(&s_small[j], s);
        }
        DEN_ROUND_BEGIN()
            mma_dw<kBaseOut>(tmem_base + kColDWb2T, Thb, Td16, have_acc);     // dWb2^T += hb^T dy
            mma_dx<kWidth, kBaseOut>(D, Td16, Wb2);                           // dhb = dy Wb2
        DEN_ROUND_END()
        tc::tmem_ld<kWidth>(tmem_lane + kColScratch, dl);
        load_row<kWidth>(Thb.hi, Thb.lo, tid, h);
#pragma unroll
        for (int j = 0; j < kWidth; ++j) dl[j] *= hidden_act_grad_from_out(f.hidden_act, h[j]);
        tc::store_row_split<kWidth>(Td64.hi, Td64.lo, tid, dl);
        DEN_ROUND_BEGIN()
            mma_dw<kEncDim>(tmem_base + kColDWb1, Td64, Tenc, have_acc);      // dWb1 += db1^T enc
            mma_colsum(tmem_base + kColDBb1, Td64, ones, have_acc);           // dbb1
            mma_dx<kEncDim, kWidth>(D, Td64, Wb1);                            // denc = db1 Wb1
        DEN_ROUND_END()
        {
            float de[kEncDim];
            tc::tmem_ld<kEncDim>(tmem_lane + kColScratch, de);
            if (valid) {
                float4* out = reinterpret_cast<float4*>(d_enc + i * enc_dim);
#pragma unroll
                for (int q = 0; q < kEncDim / 4; ++q)
                    if (4 * q < enc_dim)
                        out[q] = make_float4(de[4 * q], de[4 * q + 1], de[4 * q + 2], de[4 * q + 3]);
            }
        }
        have_acc = true;
        tc::tc_fence_before_sync();
        __syncthreads();
    }
#undef DEN_ROUND_BEGIN
#undef DEN_ROUND_END

    // ---- flush the weight-gradient accumulators (M = 64: row m in lane m%16 + 32*(m/16)) ------
    __syncthreads();
    if (have_acc) {
        tc::tc_fence_after_sync();
        const int row = warp * 16 + lane;              // valid for lane < 16
        const bool owner = lane < 16;
        float v[16];
        for (int c0 = 0; c0 < kHeadIn; c0 += 16) {     // dW1 (64, 31)
            tc::tmem_ld16(tmem_lane + kColDW1 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < kShDim + kGeo) atomicAdd(g.w1 + row * (kShDim + kGeo) + c0 + j, v[j]);
        }
        for (int c0 = 0; c0 < kWidth; c0 += 16) {      // dW2 (64, 64)
            tc::tmem_ld16(tmem_lane + kColDW2 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j) atomicAdd(g.w2 + row * kWidth + c0 + j, v[j]);
        }
        for (int c0 = 0; c0 < kEncDim; c0 += 16) {     // dWb1 (64, enc_dim)
            tc::tmem_ld16(tmem_lane + kColDWb1 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < enc_dim) atomicAdd(g.wb1 + row * enc_dim + c0 + j, v[j]);
        }
        tc::tmem_ld16(tmem_lane + kColDWb2T, v);       // dWb2^T (in 64, out 16)
        if (owner)
            for (int j = 0; j < kBaseOut; ++j) atomicAdd(g.wb2 + j * kWidth + row, v[j]);
        tc::tmem_ld16(tmem_lane + kColDW3T, v);        // dW3^T (in 64, out C)
        if (owner)
            for (int j = 0; j < C; ++j) atomicAdd(g.w3 + j * kWidth + row, v[j]);
        tc::tmem_ld16(tmem_lane + kColDB1, v);         // columns 224..239: db1 | db2
        if (owner) {
            atomicAdd(g.b1 + row, v[0]);
            atomicAdd(g.b2 + row, v[8]);
        }
        tc::tmem_ld16(tmem_lane + kColDBb1, v);
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
                           int64_t n, float* d_enc, void* stream) {
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
    DEN_CHECK_ARG((f->grid.n_levels * 2) % 4 == 0, "encoding width must be a multiple of 4");
    const int grid = grid_for((n + kBwdTile - 1) / kBwdTile, 1, 1);
    cudaFuncSetAttribute(mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)BwdSmem::total);
    mlp_bwd_tc_kernel<<<grid, kBwdThreads, BwdSmem::total, as_stream(stream)>>>(
        *f, *p, *g, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, d_sigmas, d_rgbs, n, d_enc);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}
