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
// Thread layout and pipeline: see namespace fwd below (three tiles in flight per CTA).
#include <stdlib.h>

#include "den_mlp_ops.cuh"

namespace den {

using namespace mlp;

namespace fwd {

// Pipeline (v3).  v2 ran 2 CTAs per SM, each a single tile through 5 serial rounds (epilogue ->
// hand-off -> MMA -> commit -> wait): 39 % issue-active, 39 % MUFU, every round paying ~1.2 k cycles of
// hand-shake + MMA latency.  v3 is ONE persistent CTA per SM with THREE tiles ("slots") in flight:
// 3 x 8 epilogue warps (warp quadrant q = warp % 4 serves TMEM lanes 32q.. = tile rows; half
// hf = (warp / 4) % 2 owns columns [32 hf, 32 hf + 32) of every 64-wide layer, 16 at a time) + 1 MMA
// warp serving the slots round-robin.  The C-row output layer runs on the SIMT side (dot product
// with fp32 W3 rows, halves combined through shared memory): no fifth GEMM round, no h2 operand
// tile.  Per slot: A32 (128 x 32: enc, then [SH | geo | 0]) + A64 (128 x 64: hb, then h1) = 48 KB;
// 3 slots + 36 KB of weight tiles = 181 KB of shared memory, 192 TMEM columns.
constexpr int kSlots = 3;
constexpr int kGroupThreads = 256;
constexpr int kEpiThreads = kSlots * kGroupThreads;
constexpr int kThreads = kEpiThreads;              // no dedicated MMA warp: each group issues its own GEMMs
constexpr uint32_t kTmemCols = 256;
// TMEM-A variant: per slot 64 accumulator columns + the A operands (K = 64: 32 hi + 32 lo, K = 32: 16 + 16)
constexpr uint32_t kTmemColsA = 512;
constexpr uint32_t kSlotColsA = 160;
constexpr uint32_t kColA64 = 64, kColA32 = 128;

using TA32 = OpTile<kTile, 4>;
using TA64 = OpTile<kTile, 8>;
using TWb1 = OpTile<kWidth, 4>;
using TWb2 = OpTile<kBaseOut, 8>;
using TW1 = OpTile<kWidth, 4>;
using TW2 = OpTile<kWidth, 8>;

struct Smem {
    static constexpr int wb1 = 0;
    static constexpr int wb2 = wb1 + TWb1::bytes;
    static constexpr int w1 = wb2 + TWb2::bytes;
    static constexpr int w2 = w1 + TW1::bytes;
    static constexpr int bias = w2 + TW2::bytes;          // fp32: bb1 64 | bb2 16 | b1 64 | b2 64 | b3 16
    static constexpr int w3f = bias + (3 * kWidth + kBaseOut + 16) * 4;   // fp32 W3 rows (3 x 64)
    static constexpr int zx = w3f + 3 * kWidth * 4;       // per slot (128, 2) float4: output-layer partials
    static constexpr int bars = zx + kSlots * kTile * 2 * 16;             // done[kSlots]
    static constexpr int tmem_ptr = bars + 4 * 8;
    static constexpr int slot0 = (tmem_ptr + 8 + 127) / 128 * 128;
    static constexpr int a32 = 0;
    static constexpr int a64 = a32 + TA32::bytes;
    static constexpr int slot_bytes = a64 + TA64::bytes;
    static constexpr int total = slot0 + kSlots * slot_bytes;
};
static_assert(Smem::slot_bytes % 128 == 0 && Smem::slot0 % 128 == 0, "tile alignment");
static_assert(Smem::total <= 227 * 1024, "shared-memory plan exceeds 227 KB");

}  // namespace fwd

template <bool kFull, bool kTmemA>
__global__ void __launch_bounds__(fwd::kThreads, 1)
mlp_fwd_tc_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                  const float* __restrict__ enc, const float* __restrict__ rays_o,
                  const float* __restrict__ rays_d, const int32_t* __restrict__ ray_indices,
                  const float* __restrict__ t_starts, const float* __restrict__ t_ends, int64_t n,
                  const int32_t* __restrict__ n_dev, float* __restrict__ sigmas, float* __restrict__ rgbs) {
    using namespace fwd;
    n = effective_n(n, n_dev);
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::tmem_ptr);
    const float* s_bb1 = reinterpret_cast<const float*>(smem + Smem::bias);
    const float* s_bb2 = s_bb1 + kWidth;
    const float* s_b1 = s_bb2 + kBaseOut;
    const float* s_b2 = s_b1 + kWidth;
    const float* s_b3 = s_b2 + kWidth;
    float* s_w3f = reinterpret_cast<float*>(smem + Smem::w3f);
    const int enc_dim = f.grid.n_levels * 2;
    const int C = f.channels;

    // ---- setup ---------------------------------------------------------------------------------
    tc::load_weight_split(smem + Smem::wb1, smem + Smem::wb1 + TWb1::half, p.wb1, kWidth, enc_dim, kWidth, kEncDim);
    tc::load_weight_split(smem + Smem::wb2, smem + Smem::wb2 + TWb2::half, p.wb2, kBaseOut, kWidth, kBaseOut, kWidth);
    {
        float* b = reinterpret_cast<float*>(smem + Smem::bias);
        load_padded(b, p.bb1, kWidth, kWidth);
        load_padded(b + kWidth, p.bb2, kBaseOut, kBaseOut);
        if (kFull) {
            load_padded(b + kWidth + kBaseOut, p.b1, kWidth, kWidth);
            load_padded(b + 2 * kWidth + kBaseOut, p.b2, kWidth, kWidth);
            load_padded(b + 3 * kWidth + kBaseOut, p.b3, C, 16);
        }
    }
    if (kFull) {
        tc::load_weight_split(smem + Smem::w1, smem + Smem::w1 + TW1::half, p.w1, kWidth, kShDim + kGeo, kWidth, kHeadIn);
        tc::load_weight_split(smem + Smem::w2, smem + Smem::w2 + TW2::half, p.w2, kWidth, kWidth, kWidth, kWidth);
        for (int i = tid; i < 3 * kWidth; i += kThreads) s_w3f[i] = (i / kWidth) < C ? __ldg(p.w3 + i) : 0.f;
    }
    if (tid == 0) {
        for (int b = 0; b < kSlots; ++b) tc::mbar_init(&bars[b], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemA ? kTmemColsA : kTmemCols);
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    // tiles of this CTA: blockIdx.x + k * gridDim.x, k = 0 .. my_tiles-1; slot s takes k = s, s+3, ...
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    {
        // ===================== epilogue warps: one group of 8 per slot =====================
        const int slot_id = warp >> 3;
        const int q = warp & 3, hf = (warp >> 2) & 1;
        const int row = q * 32 + lane;
        uint8_t* slot = smem + Smem::slot0 + slot_id * Smem::slot_bytes;
        uint8_t* A32 = slot + Smem::a32;
        uint8_t* A64 = slot + Smem::a64;
        uint64_t* done = &bars[slot_id];
        float* zx = reinterpret_cast<float*>(smem + Smem::zx) + slot_id * (kTile * 2 * 4);
        const uint32_t slot_cols = kTmemA ? kSlotColsA * slot_id : 64u * slot_id;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + slot_cols;
        const uint32_t Z = lane_base;
        // TMEM-A: this thread's row of the A operands (hi | lo columns), and their lane-0 addresses
        const uint32_t T64 = lane_base + kColA64, T32 = lane_base + kColA32;
        const uint32_t M64 = tmem_base + slot_cols + kColA64, M32 = tmem_base + slot_cols + kColA32;
        auto put32 = [&](int chunk0, const float (&x)[16]) {
            if constexpr (kTmemA) tstore16(T32, T32 + 16, chunk0, x);
            else store16<TA32>(A32, row, chunk0, x);
        };
        auto put64 = [&](int chunk0, const float (&x)[16]) {
            if constexpr (kTmemA) tstore16(T64, T64 + 32, chunk0, x);
            else store16<TA64>(A64, row, chunk0, x);
        };
        const int hact = f.hidden_act;
        uint32_t phase = 0;
        const uint8_t* wb1 = smem + Smem::wb1;
        const uint8_t* wb2 = smem + Smem::wb2;
        const uint8_t* w1 = smem + Smem::w1;
        const uint8_t* w2 = smem + Smem::w2;
        const uint32_t Zd = tmem_base + slot_cols;                // accumulator address (lane 0) for the MMAs
        // Hand the operand tiles to the tensor core: every thread makes its shared-memory writes
        // visible to the async proxy, the 8 warps of the group meet at a hardware named barrier, then
        // ONE elected lane of the group's first warp issues the round's GEMM and its commit — the
        // slots own disjoint accumulators, so the three issuing threads never touch the same TMEM
        // columns and no dedicated MMA warp (and no second hand-off hop) is needed.
        auto launch_round = [&](int round) {
            if constexpr (kTmemA) tmem_wait_st();
            else tc::fence_smem_to_async_proxy();
            tc::tc_fence_before_sync();
            named_sync(4 + slot_id, kGroupThreads);
            if ((warp & 7) == 0) {
                tc::tc_fence_after_sync();
                if (elect_one()) {
                    if constexpr (kTmemA) {
                        if (round == 0)
                            gemm3_ts<kEncDim / 16>(Zd, M32, M32 + 16, kmajor<TWb1>(wb1),
                                                   tc::instr_desc_bf16(128, kWidth, false, false), false);
                        else if (round == 1)
                            gemm3_ts<kWidth / 16>(Zd, M64, M64 + 32, kmajor<TWb2>(wb2),
                                                  tc::instr_desc_bf16(128, kBaseOut, false, false), false);
                        else if (round == 2)
                            gemm3_ts<kHeadIn / 16>(Zd, M32, M32 + 16, kmajor<TW1>(w1),
                                                   tc::instr_desc_bf16(128, kWidth, false, false), false);
                        else
                            gemm3_ts<kWidth / 16>(Zd, M64, M64 + 32, kmajor<TW2>(w2),
                                                  tc::instr_desc_bf16(128, kWidth, false, false), false);
                    } else
                    if (round == 0)        // z_b1 = enc Wb1^T
                        gemm3<kEncDim / 16>(Zd, kmajor<TA32>(A32), kmajor<TWb1>(wb1),
                                            tc::instr_desc_bf16(128, kWidth, false, false), false);
                    else if (round == 1)   // y = hb Wb2^T
                        gemm3<kWidth / 16>(Zd, kmajor<TA64>(A64), kmajor<TWb2>(wb2),
                                           tc::instr_desc_bf16(128, kBaseOut, false, false), false);
                    else if (round == 2)   // z1 = [SH | geo | 0] W1^T
                        gemm3<kHeadIn / 16>(Zd, kmajor<TA32>(A32), kmajor<TW1>(w1),
                                            tc::instr_desc_bf16(128, kWidth, false, false), false);
                    else                   // z2 = h1 W2^T
                        gemm3<kWidth / 16>(Zd, kmajor<TA64>(A64), kmajor<TW2>(w2),
                                           tc::instr_desc_bf16(128, kWidth, false, false), false);
                    tc::mma_commit_1t(done);
                }
                __syncwarp();
            }
        };

        // Software pipeline over this slot's tiles.  A sample's ray is a dependent chain (sample -> ray
        // index -> origin / direction) and its encoding row a 64-byte gather: the NEXT tile's ray index is
        // fetched at the start of the current tile, its origin / direction and encoding before the current
        // tile's last wait, so that no tile starts on HBM / L2 latency (ncu, profiles/r02_ncu_mlp.md: 8 % of
        // the warp samples of the un-pipelined kernel sat on the t_starts line alone).
        float xe[16], dn[3] = {0.f, 0.f, 1.f}, on[3] = {0.f, 0.f, 0.f}, tm = 0.f;
        auto load_enc = [&](int64_t i, bool valid) {
#pragma unroll
            for (int c = 0; c < 16; ++c) xe[c] = 0.f;
            if (valid) {
                const float4* src = reinterpret_cast<const float4*>(enc + i * enc_dim + 16 * hf);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4)
                    if (16 * hf + 4 * v4 < enc_dim) {
                        const float4 v = __ldg(src + v4);
                        xe[4 * v4] = v.x; xe[4 * v4 + 1] = v.y; xe[4 * v4 + 2] = v.z; xe[4 * v4 + 3] = v.w;
                    }
            }
        };
        auto load_ray = [&](int64_t r) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                dn[d] = __ldg(rays_d + 3 * r + d);
                if (hf == 0) on[d] = __ldg(rays_o + 3 * r + d);
            }
        };
        const int64_t tile_stride = (int64_t)kSlots * gridDim.x * kTile;
        int64_t i = ((int64_t)blockIdx.x + (int64_t)slot_id * gridDim.x) * kTile + row;
        bool valid = slot_id < my_tiles && i < n;
        load_enc(i, valid);
        if (valid) {
            tm = t_starts[i] + t_ends[i];
            load_ray(ray_indices[i]);
        }
        for (int64_t k = slot_id; k < my_tiles; k += kSlots) {
            // ---- operands of round 0: this thread's 16 encoding features -------------------------
            const float dir[3] = {dn[0], dn[1], dn[2]};
            bool inside = false;
            if (valid && hf == 0) {
                float pos[3], u[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) pos[d] = on[d] + (dir[d] * tm) * 0.5f;
                inside = contract_position(f, pos, u);
            }
            put32(2 * hf, xe);
            // the slot's next tile: ray index and interval now, the rest before this tile's last wait
            const int64_t i_cur = i;
            const bool valid_cur = valid;
            const int64_t ni = i + tile_stride;
            const bool valid_n = k + kSlots < my_tiles && ni < n;
            int32_t ray_n = 0;
            float tm_n = 0.f;
            if (valid_n) {
                ray_n = __ldg(ray_indices + ni);
                tm_n = __ldg(t_starts + ni) + __ldg(t_ends + ni);
            }
            auto fetch_next = [&]() {
                i = ni;
                valid = valid_n;
                tm = tm_n;
                load_enc(ni, valid_n);
                dn[0] = 0.f; dn[1] = 0.f; dn[2] = 1.f;
                if (valid_n) load_ray(ray_n);
            };
            launch_round(0);

            // ---- round 0 done: hb -> A64 ---------------------------------------------------------------
            tc::mbar_wait(done, phase); phase ^= 1; tc::tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float h[16];
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, h);
                bias_hidden_act<16>(hact, h, s_bb1 + 32 * hf + 16 * c);
                put64(4 * hf + 2 * c, h);
            }
            launch_round(1);

            // ---- round 1 done: density; [SH | geo | 0] -> A32 --------------------------------------------
            tc::mbar_wait(done, phase); phase ^= 1; tc::tc_fence_after_sync();
            if (hf == 0) {
                float y[16];
                tmem_ld_cols<16>(Z, y);
#pragma unroll
                for (int j = 0; j < kBaseOut; ++j) y[j] += s_bb2[j];
                if (valid_cur) sigmas[i_cur] = inside ? density_act(f.density_act, y[0]) : 0.f;
                if (kFull) {
                    float x[16];
#pragma unroll
                    for (int j = 0; j < kGeo; ++j) x[j] = y[1 + j];
                    x[15] = 0.f;
                    put32(2, x);
                }
            } else if (kFull) {
                float x[16];
                sh_degree4(dir, x);
                put32(0, x);
            }
            if (!kFull) {
                fetch_next();
                tc::tc_fence_before_sync();
                continue;
            }
            launch_round(2);

            // ---- round 2 done: h1 -> A64 ---------------------------------------------------------------------
            tc::mbar_wait(done, phase); phase ^= 1; tc::tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float h[16];
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, h);
                bias_hidden_act<16>(hact, h, s_b1 + 32 * hf + 16 * c);
                put64(4 * hf + 2 * c, h);
            }
            launch_round(3);
            fetch_next();

            // ---- round 3 done: h2 stays in registers; output layer = dot with the fp32 W3 rows ----------------
            tc::mbar_wait(done, phase); phase ^= 1; tc::tc_fence_after_sync();
            {
                float z3[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float h[16];
                    tmem_ld_cols<16>(Z + 32 * hf + 16 * c, h);
                    bias_hidden_act<16>(hact, h, s_b2 + 32 * hf + 16 * c);
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        if (ch < C) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                z3[ch] = fmaf(h[j], s_w3f[ch * kWidth + 32 * hf + 16 * c + j], z3[ch]);
                        }
                }
                tc::tc_fence_before_sync();
                if (hf == 1) *reinterpret_cast<float4*>(zx + row * 4) = make_float4(z3[0], z3[1], z3[2], 0.f);
                named_sync(1 + slot_id, kGroupThreads);
                if (hf == 0) {
                    const float4 other = *reinterpret_cast<const float4*>(zx + row * 4);
                    z3[0] += other.x; z3[1] += other.y; z3[2] += other.z;
                    if (valid_cur) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            if (ch < C) rgbs[i_cur * C + ch] = radiance_act(f.radiance_act, z3[ch] + s_b3[ch]);
                    }
                }
            }
        }
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kTmemA ? fwd::kTmemColsA : fwd::kTmemCols);
}

// ---- v5: FOUR tiles in flight, one warp per TMEM lane quadrant (4 warps per slot), TMEM-A only -------
// With the activations in tensor memory a slot needs no shared-memory tile at all, and the A operands of
// consecutive rounds are never live together (enc -> hb -> [SH | geo] -> h1), so they share 64 columns:
// 64 (accumulator) + 64 (A hi | lo) = 128 columns per slot, four slots fill the 512 columns.  Each thread
// owns a whole row (64 columns per round, 16 at a time), which also removes the cross-half exchange of the
// output layer; 16 warps instead of 24 leave 128 registers per thread.
namespace fwd4 {
constexpr int kSlots = 4;
constexpr int kGroupThreads = 128;
constexpr int kThreads = kSlots * kGroupThreads;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSlotCols = 128, kColA = 64;
}  // namespace fwd4

template <bool kFull>
__global__ void __launch_bounds__(fwd4::kThreads, 1)
mlp_fwd_tc4_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                   const float* __restrict__ enc, const float* __restrict__ rays_o,
                   const float* __restrict__ rays_d, const int32_t* __restrict__ ray_indices,
                   const float* __restrict__ t_starts, const float* __restrict__ t_ends, int64_t n,
                   const int32_t* __restrict__ n_dev, float* __restrict__ sigmas, float* __restrict__ rgbs) {
    using namespace fwd4;
    using fwd::Smem;
    using fwd::TWb1; using fwd::TWb2; using fwd::TW1; using fwd::TW2;
    n = effective_n(n, n_dev);
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::tmem_ptr);
    const float* s_bb1 = reinterpret_cast<const float*>(smem + Smem::bias);
    const float* s_bb2 = s_bb1 + kWidth;
    const float* s_b1 = s_bb2 + kBaseOut;
    const float* s_b2 = s_b1 + kWidth;
    const float* s_b3 = s_b2 + kWidth;
    float* s_w3f = reinterpret_cast<float*>(smem + Smem::w3f);
    const int enc_dim = f.grid.n_levels * 2;
    const int C = f.channels;

    tc::load_weight_split(smem + Smem::wb1, smem + Smem::wb1 + TWb1::half, p.wb1, kWidth, enc_dim, kWidth, kEncDim);
    tc::load_weight_split(smem + Smem::wb2, smem + Smem::wb2 + TWb2::half, p.wb2, kBaseOut, kWidth, kBaseOut, kWidth);
    {
        float* b = reinterpret_cast<float*>(smem + Smem::bias);
        load_padded(b, p.bb1, kWidth, kWidth);
        load_padded(b + kWidth, p.bb2, kBaseOut, kBaseOut);
        if (kFull) {
            load_padded(b + kWidth + kBaseOut, p.b1, kWidth, kWidth);
            load_padded(b + 2 * kWidth + kBaseOut, p.b2, kWidth, kWidth);
            load_padded(b + 3 * kWidth + kBaseOut, p.b3, C, 16);
        }
    }
    if (kFull) {
        tc::load_weight_split(smem + Smem::w1, smem + Smem::w1 + TW1::half, p.w1, kWidth, kShDim + kGeo, kWidth, kHeadIn);
        tc::load_weight_split(smem + Smem::w2, smem + Smem::w2 + TW2::half, p.w2, kWidth, kWidth, kWidth, kWidth);
        for (int i = tid; i < 3 * kWidth; i += kThreads) s_w3f[i] = (i / kWidth) < C ? __ldg(p.w3 + i) : 0.f;
    }
    if (tid == 0) {
        for (int b = 0; b < kSlots; ++b) tc::mbar_init(&bars[b], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemCols);
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    const int slot_id = warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint64_t* done = &bars[slot_id];
    const uint32_t slot_cols = kSlotCols * slot_id;
    const uint32_t Z = tmem_base + ((uint32_t)(q * 32) << 16) + slot_cols;      // this thread's accumulator row
    const uint32_t TA = Z + kColA;                                               // ... and A operand row
    const uint32_t Zd = tmem_base + slot_cols, MA = Zd + kColA;                 // lane-0 addresses for the MMAs
    const int hact = f.hidden_act;
    uint32_t phase = 0;
    const uint8_t* wb1 = smem + Smem::wb1;
    const uint8_t* wb2 = smem + Smem::wb2;
    const uint8_t* w1 = smem + Smem::w1;
    const uint8_t* w2 = smem + Smem::w2;

    // K = 32 operands: hi columns [0, 16), lo [16, 32); K = 64: hi [0, 32), lo [32, 64) of the A region
    auto launch_round = [&](int round) {
        tmem_wait_st();
        tc::tc_fence_before_sync();
        named_sync(4 + slot_id, kGroupThreads);
        if (q == 0) {
            tc::tc_fence_after_sync();
            if (elect_one()) {
                if (round == 0)
                    gemm3_ts<kEncDim / 16>(Zd, MA, MA + 16, kmajor<TWb1>(wb1),
                                           tc::instr_desc_bf16(128, kWidth, false, false), false);
                else if (round == 1)
                    gemm3_ts<kWidth / 16>(Zd, MA, MA + 32, kmajor<TWb2>(wb2),
                                          tc::instr_desc_bf16(128, kBaseOut, false, false), false);
                else if (round == 2)
                    gemm3_ts<kHeadIn / 16>(Zd, MA, MA + 16, kmajor<TW1>(w1),
                                           tc::instr_desc_bf16(128, kWidth, false, false), false);
                else
                    gemm3_ts<kWidth / 16>(Zd, MA, MA + 32, kmajor<TW2>(w2),
                                          tc::instr_desc_bf16(128, kWidth, false, false), false);
                tc::mma_commit_1t(done);
            }
            __syncwarp();
        }
    };
    // a 64-wide hidden layer: accumulator -> bias + activation -> A operand (K = 64); the TMEM load of
    // the next 16 columns is in flight while the current 16 are processed
    auto hidden_to_operand = [&](const float* bias) {
        uint32_t raw[2][16];
        tmem_ld16_nowait(Z, raw[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c < 3) tmem_ld16_nowait(Z + 16 * (c + 1), raw[(c + 1) & 1]);
            float h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = __uint_as_float(raw[c & 1][j]);
            bias_hidden_act<16>(hact, h, bias + 16 * c);
            tstore16(TA, TA + 32, 2 * c, h);
        }
    };

    for (int64_t k = slot_id; k < my_tiles; k += kSlots) {
        const int64_t tile = blockIdx.x + k * gridDim.x;
        const int64_t i = tile * kTile + row;
        const bool valid = i < n;
        float dir[3] = {0.f, 0.f, 1.f};
        bool inside = false;
        // ---- round 0 operand: the 32 encoding features of this row ------------------------------
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            float x[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = 0.f;
            if (valid) {
                const float4* src = reinterpret_cast<const float4*>(enc + i * enc_dim + 16 * hf);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4)
                    if (16 * hf + 4 * v4 < enc_dim) {
                        const float4 v = __ldg(src + v4);
                        x[4 * v4] = v.x; x[4 * v4 + 1] = v.y; x[4 * v4 + 2] = v.z; x[4 * v4 + 3] = v.w;
                    }
            }
            tstore16(TA, TA + 16, 2 * hf, x);
        }
        if (valid) {
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
        launch_round(0);

        // ---- round 0 done: hb -> A (K = 64) -------------------------------------------------------
        tc::mbar_wait(done, phase); phase ^= 1; tc::tc_fence_after_sync();
        hidden_to_operand(s_bb1);
        launch_round(1);

        // ---- round 1 done: density; [SH | geo | 0] -> A (K = 32) -------------------------------------
        tc::mbar_wait(done, phase); phase ^= 1; tc::tc_fence_after_sync();
        {
            float y[16];
            tmem_ld_cols<16>(Z, y);
#pragma unroll
            for (int j = 0; j < kBaseOut; ++j) y[j] += s_bb2[j];
            if (valid) sigmas[i] = inside ? density_act(f.density_act, y[0]) : 0.f;
            if (kFull) {
                float x[16];
                sh_degree4(dir, x);
                tstore16(TA, TA + 16, 0, x);
#pragma unroll
                for (int j = 0; j < kGeo; ++j) x[j] = y[1 + j];
                x[15] = 0.f;
                tstore16(TA, TA + 16, 2, x);
            }
        }
        if (!kFull) {
            tc::tc_fence_before_sync();
            continue;
        }
        launch_round(2);

        // ---- round 2 done: h1 -> A (K = 64) ------------------------------------------------------------
        tc::mbar_wait(done, phase); phase ^= 1; tc::tc_fence_after_sync();
        hidden_to_operand(s_b1);
        launch_round(3);

        // ---- round 3 done: h2 in registers, output layer = dot with the fp32 W3 rows ---------------------
        tc::mbar_wait(done, phase); phase ^= 1; tc::tc_fence_after_sync();
        {
            float z3[3] = {0.f, 0.f, 0.f};
            uint32_t raw[2][16];
            tmem_ld16_nowait(Z, raw[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c < 3) tmem_ld16_nowait(Z + 16 * (c + 1), raw[(c + 1) & 1]);
                float h[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = __uint_as_float(raw[c & 1][j]);
                bias_hidden_act<16>(hact, h, s_b2 + 16 * c);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    if (ch < C) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) z3[ch] = fmaf(h[j], s_w3f[ch * kWidth + 16 * c + j], z3[ch]);
                    }
            }
            tc::tc_fence_before_sync();
            if (valid) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    if (ch < C) rgbs[i * C + ch] = radiance_act(f.radiance_act, z3[ch] + s_b3[ch]);
            }
        }
    }

    tc::tc_fence_before_sync();
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
                                        const int32_t* __restrict__ n_dev, float* __restrict__ unit_pos) {
    n = effective_n(n, n_dev);
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
                                            const int32_t* __restrict__ n_dev, float* __restrict__ d_pos,
                                            float* __restrict__ d_pos_t) {
    n = effective_n(n, n_dev);
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
                             const float* d_unit, int64_t n, const int32_t* n_dev, float* d_pos,
                             float* d_pos_t, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(f != nullptr, "null descriptor");
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(rays_o && rays_d && ray_indices && t_starts && t_ends && d_unit && d_pos && d_pos_t,
                  "null pointer");
    contract_samples_bwd_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        *f, rays_o, rays_d, ray_indices, t_starts, t_ends, d_unit, n, n_dev, d_pos, d_pos_t);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_contract_samples(const den_field_desc* f, const float* rays_o, const float* rays_d,
                         const int32_t* ray_indices, const float* t_starts, const float* t_ends,
                         int64_t n, const int32_t* n_dev, float* unit_pos, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(f != nullptr, "null descriptor");
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(rays_o && rays_d && ray_indices && t_starts && t_ends && unit_pos, "null pointer");
    contract_samples_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        *f, rays_o, rays_d, ray_indices, t_starts, t_ends, n, n_dev, unit_pos);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_mlp_fwd(const den_field_desc* f, const den_field_params* p, const float* enc,
                const float* rays_o, const float* rays_d, const int32_t* ray_indices,
                const float* t_starts, const float* t_ends, int64_t n, const int32_t* n_dev,
                float* sigmas, float* rgbs, void* stream) {
    using namespace den;
    const bool full = rgbs != nullptr;
    int rc = check_field(f, p, full);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(enc && rays_o && rays_d && ray_indices && t_starts && t_ends && sigmas,
                  "null pointer");
    DEN_CHECK_ARG((f->grid.n_levels * 2) % 4 == 0, "encoding width must be a multiple of 4");
    DEN_CHECK_ARG(!full || (f->channels >= 1 && f->channels <= 3), "1 to 3 radiance channels");
    // three tiles in flight per CTA: one persistent CTA per SM, at least three tiles each when there are enough
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int grid = grid_for((n_tiles + fwd::kSlots - 1) / fwd::kSlots, 1, 1);
    // The activations reach the tensor core as the TMEM A operand (no operand tiles in shared memory).
    // DEN_MLP_FWD_TMEM_A=0 forces the round-1 kernel with shared-memory operand tiles (kept for
    // comparison), =1 / =2 one of the two TMEM-A kernels; read once per process
    // measured at 40.8 M samples (profiles/r02_time_mlp_variants.md): full evaluation 7.01 ms (SS operand
    // tiles) / 5.77 (TMEM-A, 3 slots x 8 warps) / 6.22 (TMEM-A, 4 slots x 4 warps); density only 4.13 / 3.38 /
    // 3.14 — so the default is the 3 x 8 kernel for the full evaluation and the 4 x 4 one for density only
    static const int forced = [] {               // 0: SS operands, 1: TMEM-A 3 x 8 warps, 2: TMEM-A 4 x 4 warps
        const char* e = getenv("DEN_MLP_FWD_TMEM_A");
        return e != nullptr ? (int)(e[0] - '0') : -1;
    }();
    const int variant = forced >= 0 ? forced : (full ? 1 : 2);
    const bool tmem_a = variant >= 1;
    if (variant == 2) {
        const int grid4 = grid_for((n_tiles + fwd4::kSlots - 1) / fwd4::kSlots, 1, 1);
        const size_t smem4 = (size_t)fwd::Smem::slot0;
        auto launch4 = [&](auto kernel, float* rgb_out) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
            kernel<<<grid4, fwd4::kThreads, smem4, as_stream(stream)>>>(*f, *p, enc, rays_o, rays_d, ray_indices,
                                                                       t_starts, t_ends, n, n_dev, sigmas, rgb_out);
        };
        if (full) launch4(mlp_fwd_tc4_kernel<true>, rgbs);
        else launch4(mlp_fwd_tc4_kernel<false>, nullptr);
        DEN_CHECK_LAUNCH();
        return DEN_OK;
    }
    const size_t smem = tmem_a ? (size_t)fwd::Smem::slot0 : (size_t)fwd::Smem::total;
    auto launch = [&](auto kernel, float* rgb_out) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kernel<<<grid, fwd::kThreads, smem, as_stream(stream)>>>(*f, *p, enc, rays_o, rays_d, ray_indices,
                                                                 t_starts, t_ends, n, n_dev, sigmas, rgb_out);
    };
    if (full) {
        if (tmem_a) launch(mlp_fwd_tc_kernel<true, true>, rgbs);
        else launch(mlp_fwd_tc_kernel<true, false>, rgbs);
    } else {
        if (tmem_a) launch(mlp_fwd_tc_kernel<false, true>, nullptr);
        else launch(mlp_fwd_tc_kernel<false, false>, nullptr);
    }
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
