// Backward of the tensor-core MLP: forward recompute + input gradient + weight gradients,
// fused per 128-sample tile (tcgen05.mma, TMEM accumulators; sm_100a), TWO tiles in flight.
//
// Replaces the autograd of MLP.forward / SHEncoder / activations under
// NGPradianceField.forward (external/ngp.py:269-280) — in the reference ~30 cuBLAS and
// elementwise launches that read and write every (M,64) activation twice.  Here nothing but
// enc (128 B/sample in, read twice), the two upstream gradients and dL/denc (128 B/sample out)
// touches HBM: the forward activations are recomputed into shared-memory operand tiles, which
// are then consumed three ways without being rewritten —
//   K-major  A operand        : next layer forward,  dX = dY * W
//   MN-major A/B operand      : dW += dY^T * X   (K = the 128 samples of the tile)
// Weight-gradient accumulators (64 x N, fp32) stay in TMEM across all tiles of a CTA and are
// flushed once with atomics.  Bias gradients ride along: every activation tile carries a
// column of ones (its own 8-column chunk, or the zero-padded 32nd input of the head), so
// dW += dY^T [X | 1] yields db in the extra accumulator column for free.
//
// Pipeline (v3).  v2 ran ONE tile through ten strictly serial rounds (epilogue -> bar.sync ->
// MMA issue -> commit -> mbarrier wait): ncu showed 43 % of all warp samples spinning on the
// MMA mbarrier, 20 % issue utilisation and 11 % tensor-pipe activity
// (profiles/r01_ncu_full_top_kernels.md).  v3 keeps two tiles ("slots") resident per CTA:
//   * 2 x 8 epilogue warps, one group per slot (warp quadrant q = warp % 4 serves TMEM lanes
//     32q..32q+31 = tile rows; half hf = (warp / 4) % 2 owns columns [32 hf, 32 hf + 32) of
//     every 64-wide layer, processed 16 at a time to bound registers);
//   * the four forward rounds of a tile only write the slot's own scratch columns: the slot's first
//     warp issues them itself after the slot's named barrier (no hop to another warp);
//   * 1 MMA warp issues the four backward rounds of both slots alternately — their dW GEMMs share
//     the weight-gradient accumulators, and ONE issuing thread keeps the accumulations an ordered
//     stream; hand-off slot -> MMA warp by a hardware named barrier (bar.arrive / bar.sync), back by a
//     tcgen05.commit mbarrier; per round the latency-critical dX GEMM is committed first, the dW GEMM
//     follows on its own mbarrier and is awaited only when its operand tiles are about to be
//     overwritten (the last one of a tile under the next tile's loads).  Issuing the dX GEMMs from the
//     slot as well and leaving only the dW GEMMs to the MMA warp was measured slower (the dW GEMMs then
//     start later and their waits grow): 5.00 ms against 4.85 ms;
//   * the (C <= 3)-row output layer runs on the SIMT side (no 64 x 16 GEMM, no h2 operand tile).
// (Tried and rejected, measured: all 16 warps on ONE slot's phase at a time, alternating slots, with a
// CTA-wide barrier per phase — every warp then hits the same tcgen05.ld / fence / barrier latencies at
// the same moment and nothing is left to hide them: 6.9 ms against 5.3 ms for the form below.)
// (Also tried and rejected, measured at 10.2 M samples: SAVING the forward activations in den_mlp_fwd
// (832 B/sample, tile-blocked so that every warp access is 512 contiguous bytes) and reading them back
// here instead of recomputing them — four MMA rounds per tile instead of eight, no TMEM parking, results
// bit-identical.  Forward 1.71 -> 2.05 ms, backward 4.85 -> 5.27 ms (5.68 ms with the loads hoisted
// across the MMA waits: the 96-register cap turns the extra live values into spills).  The recompute
// rounds of one slot hide under the backward rounds of the other, so removing them buys nothing; what
// bounds the kernel is the chain hand-off -> GEMM -> mbarrier -> epilogue of the four backward rounds
// with only two tiles in flight.)
// Shared memory per slot: E (128 x 40: enc | ones, later [SH | geo | 1], later dy, later enc
// again), H (128 x 72: hb | ones, later h1, later hb again) and D (128 x 64: the current
// dL/dz tile) = 88 KB; 2 slots + 36 KB of weight tiles = 223 KB.  The price of fitting two
// slots: the split hb and enc tiles are parked in TMEM (64 + 32 columns of packed bf16 words per
// slot) while h1 / [SH | geo] / dy occupy H / E, and copied back afterwards.
#include <stdlib.h>

#include "den_mlp_ops.cuh"

namespace den {

using namespace mlp;

namespace bwd {

constexpr int kSlots = 2;
constexpr int kGroupThreads = 256;                         // 8 epilogue warps per slot
constexpr int kEpiThreads = kSlots * kGroupThreads;
constexpr int kThreads = kEpiThreads + 32;                 // + the MMA warp
constexpr int kMmaWarp = kEpiThreads / 32;
constexpr uint32_t kTmemCols = 512;
// The tensor core accumulates with truncation: the error of a TMEM weight-gradient accumulator grows
// linearly with the tiles summed into it (measured against an fp64 evaluation, profiles/wgrad_check.py:
// 2.9e-4 of max|dW| after 539 tiles, 6.7e-5 after 135, 1.2e-5 after 8).  The accumulators are therefore
// flushed to global memory (fp32 atomics, round-to-nearest) every kFlushTiles tiles of a CTA.
constexpr int kFlushTiles = 128;

// TMEM column plan
constexpr uint32_t kColZ = 0;            // + 128 * slot : 64 scratch columns (forward / dX results)
constexpr uint32_t kColP = 64;           // + 128 * slot : 64 columns, the split hb tile parked as packed bf16 words
constexpr uint32_t kColDW2 = 256;        // 72: dW2 (out 64 x in 64) | db2 x 8
constexpr uint32_t kColDW1 = 336;        // 32: dW1 (64 x 31) | db1 in column 31
constexpr uint32_t kColDWb1 = 368;       // 40: dWb1 (64 x 32) | dbb1 x 8
constexpr uint32_t kColDWb2T = 408;      // 16: dWb2^T (in 64 x out 16)
constexpr uint32_t kColEncPark = 424;    // + 32 * slot : 32 columns, the split enc tile parked as packed bf16 words

using TE = OpTile<kTile, 5>;     // enc | ones     /  [SH | geo | 1]  /  dy
using TH = OpTile<kTile, 9>;     // hb | ones      /  h1
using TD = OpTile<kTile, 8>;     // dL/dz of the current 64-wide layer
using TWb1 = OpTile<kWidth, 4>;  // (64, 32)
using TWb2 = OpTile<kBaseOut, 8>;// (16, 64)
using TW1 = OpTile<kWidth, 4>;   // (64, 32)
using TW2 = OpTile<kWidth, 8>;   // (64, 64)

struct Smem {
    static constexpr int wb1 = 0;
    static constexpr int wb2 = wb1 + TWb1::bytes;
    static constexpr int w1 = wb2 + TWb2::bytes;
    static constexpr int w2 = w1 + TW1::bytes;
    static constexpr int bias = w2 + TW2::bytes;          // fp32: bb1 64 | bb2 16 | b1 64 | b2 64 | b3 16
    static constexpr int w3f = bias + (3 * kWidth + kBaseOut + 16) * 4;   // fp32 W3 rows (3 x 64)
    static constexpr int acc = w3f + 3 * kWidth * 4;      // fp32: dW3 (3 x 64) | dbb2 16 | db3 4 | pad
    static constexpr int zx = acc + (3 * kWidth + 32) * 4;               // per slot (128, 2, 4) fp32
    static constexpr int bars = zx + kSlots * kTile * 2 * 4 * 4;         // done[2], dw_done[2]
    static constexpr int tmem_ptr = bars + 4 * 8;
    static constexpr int slot0 = (tmem_ptr + 8 + 127) / 128 * 128;
    static constexpr int e = 0;                            // offsets inside a slot
    static constexpr int h = e + TE::bytes;
    static constexpr int d = h + TH::bytes;
    static constexpr int slot_bytes = d + TD::bytes;
    static constexpr int total = slot0 + kSlots * slot_bytes;
};
static_assert(Smem::slot_bytes % 128 == 0 && Smem::slot0 % 128 == 0, "tile alignment");
static_assert(Smem::total <= 227 * 1024, "shared-memory plan exceeds 227 KB");

__device__ __forceinline__ void group_sync(int slot) {       // the 8 epilogue warps of one slot
    asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(kGroupThreads) : "memory");
}

// ---- epilogue helpers ----------------------------------------------------------------------------
// hand-off epilogue group -> MMA warp through a hardware named barrier (3 + slot): the 256 producers
// arrive without blocking, the MMA warp syncs
constexpr int kHandoffThreads = kGroupThreads + 32;
__device__ __forceinline__ void publish(int slot) {
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    asm volatile("bar.arrive %0, %1;" ::"r"(3 + slot), "r"(kHandoffThreads) : "memory");
}
__device__ __forceinline__ void handoff_wait(int slot) {
    asm volatile("bar.sync %0, %1;" ::"r"(3 + slot), "r"(kHandoffThreads) : "memory");
}
__device__ __forceinline__ void commit_to(uint64_t* bar) {      // by the elected issuing thread
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void await_mma(uint64_t* done, uint32_t& phase) {
    tc::mbar_wait(done, phase);
    phase ^= 1;
    tc::tc_fence_after_sync();
}
}  // namespace bwd

using namespace bwd;

__global__ void __launch_bounds__(kThreads, 1)
mlp_bwd_tc_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                  const __grid_constant__ den_field_grads g, const float* __restrict__ enc,
                  const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                  const int32_t* __restrict__ ray_indices, const float* __restrict__ t_starts,
                  const float* __restrict__ t_ends, const float* __restrict__ d_sigmas,
                  const float* __restrict__ d_rgbs, int64_t n, const int32_t* __restrict__ n_dev,
                  const int32_t* __restrict__ enc_rows, float* __restrict__ d_enc,
                  float* __restrict__ d_dirs) {
    n = effective_n(n, n_dev);
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);       // done[0..1], dw_done[0..1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::tmem_ptr);
    const float* s_bb1 = reinterpret_cast<const float*>(smem + Smem::bias);
    const float* s_bb2 = s_bb1 + kWidth;
    const float* s_b1 = s_bb2 + kBaseOut;
    const float* s_b2 = s_b1 + kWidth;
    const float* s_b3 = s_b2 + kWidth;
    float* s_w3f = reinterpret_cast<float*>(smem + Smem::w3f);
    float* s_dw3 = reinterpret_cast<float*>(smem + Smem::acc);             // (3, 64)
    float* s_dbb2 = s_dw3 + 3 * kWidth;                                    // 16
    float* s_db3 = s_dbb2 + kBaseOut;                                      // 4

    const int enc_dim = f.grid.n_levels * 2;
    const int C = f.channels;

    // ---- setup ---------------------------------------------------------------------------------
    tc::load_weight_split(smem + Smem::wb1, smem + Smem::wb1 + TWb1::half, p.wb1, kWidth, enc_dim, kWidth, kEncDim);
    tc::load_weight_split(smem + Smem::wb2, smem + Smem::wb2 + TWb2::half, p.wb2, kBaseOut, kWidth, kBaseOut, kWidth);
    tc::load_weight_split(smem + Smem::w1, smem + Smem::w1 + TW1::half, p.w1, kWidth, kShDim + kGeo, kWidth, kHeadIn);
    tc::load_weight_split(smem + Smem::w2, smem + Smem::w2 + TW2::half, p.w2, kWidth, kWidth, kWidth, kWidth);
    {
        float* b = reinterpret_cast<float*>(smem + Smem::bias);
        load_padded(b, p.bb1, kWidth, kWidth);
        load_padded(b + kWidth, p.bb2, kBaseOut, kBaseOut);
        load_padded(b + kWidth + kBaseOut, p.b1, kWidth, kWidth);
        load_padded(b + 2 * kWidth + kBaseOut, p.b2, kWidth, kWidth);
        load_padded(b + 3 * kWidth + kBaseOut, p.b3, C, 16);
    }
    for (int i = tid; i < 3 * kWidth; i += kThreads) {
        s_w3f[i] = (i / kWidth) < C ? __ldg(p.w3 + i) : 0.f;
        s_dw3[i] = 0.f;
    }
    if (tid < 32) s_dbb2[tid] = 0.f;                       // dbb2 | db3 | pad
    // the "ones" chunks (bf16 1.0 in the hi half, 0 in the lo half): E chunk 4, H chunk 8
    for (int i = tid; i < kSlots * kTile; i += kThreads) {
        const int s = i / kTile, r = i - s * kTile;
        uint8_t* slot = smem + Smem::slot0 + s * Smem::slot_bytes;
        const uint4 one = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(slot + Smem::e + TE::off(r, 4)) = one;
        *reinterpret_cast<uint4*>(slot + Smem::e + TE::half + TE::off(r, 4)) = zero;
        *reinterpret_cast<uint4*>(slot + Smem::h + TH::off(r, 8)) = one;
        *reinterpret_cast<uint4*>(slot + Smem::h + TH::half + TH::off(r, 8)) = zero;
    }
    if (tid == 0) {
        for (int b = 0; b < 4; ++b) tc::mbar_init(&bars[b], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemCols);
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    // tiles of this CTA: blockIdx.x + k * gridDim.x, k = 0 .. my_tiles-1; slot s takes k = s, s+2, ...
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    // TMEM weight-gradient accumulators -> global (warps 0..3; M = 64: row m in lane m%16 + 32*(m/16))
    auto flush_tmem = [&]() {
        tc::tc_fence_after_sync();
        const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int row = warp * 16 + lane;              // valid for lane < 16
        const bool owner = lane < 16;
        float v[16];
        for (int c0 = 0; c0 < kHeadIn; c0 += 16) {     // dW1 (64, 31) | db1 in column 31
            tmem_ld_cols<16>(tl + kColDW1 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j) {
                    if (c0 + j < kShDim + kGeo) atomicAdd(g.w1 + row * (kShDim + kGeo) + c0 + j, v[j]);
                    else atomicAdd(g.b1 + row, v[j]);
                }
        }
        for (int c0 = 0; c0 < kWidth; c0 += 16) {      // dW2 (64, 64)
            tmem_ld_cols<16>(tl + kColDW2 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j) atomicAdd(g.w2 + row * kWidth + c0 + j, v[j]);
        }
        for (int c0 = 0; c0 < kEncDim; c0 += 16) {     // dWb1 (64, enc_dim)
            tmem_ld_cols<16>(tl + kColDWb1 + c0, v);
            if (owner)
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < enc_dim) atomicAdd(g.wb1 + row * enc_dim + c0 + j, v[j]);
        }
        {
            float b[8];
            tmem_ld_cols<8>(tl + kColDW2 + kWidth, b);           // db2
            if (owner) atomicAdd(g.b2 + row, b[0]);
            tmem_ld_cols<8>(tl + kColDWb1 + kEncDim, b);         // dbb1
            if (owner) atomicAdd(g.bb1 + row, b[0]);
        }
        tmem_ld_cols<16>(tl + kColDWb2T, v);           // dWb2^T (in 64, out 16)
        if (owner)
            for (int j = 0; j < kBaseOut; ++j) atomicAdd(g.wb2 + j * kWidth + row, v[j]);
        tc::tc_fence_before_sync();
    };

    if (warp == kMmaWarp) {
        // ===================== MMA warp =====================
        const uint8_t* wb1 = smem + Smem::wb1;
        const uint8_t* wb2 = smem + Smem::wb2;
        const uint8_t* w1 = smem + Smem::w1;
        const uint8_t* w2 = smem + Smem::w2;
        // Fixed alternation slot 0, slot 1, slot 0, ...: each hand-off is one named-barrier sync.  Per
        // round the latency-critical GEMM (forward layer / dX) is issued and committed first; the
        // weight-gradient GEMM follows with its own commit (dw_done) — nobody waits for it until the
        // operand tiles it reads are about to be overwritten.
        // per-slot counters packed in scalars (the slot loop is rolled: no dynamically indexed arrays)
        for (int64_t c0 = 0; c0 < my_tiles; c0 += kFlushTiles) {
        const int64_t nt = min((int64_t)kFlushTiles, my_tiles - c0);          // tiles of this flush period
        // the forward rounds 0..3 (private accumulators) are issued by the slots themselves; this warp
        // issues the backward rounds 4..7, whose dW GEMMs share the weight-gradient accumulators
        int64_t left0 = 4 * ((nt + 1) / 2), left1 = 4 * (nt / 2);
        int round0 = 4, round1 = 4;
        uint32_t acc_mask = 0;                           // bit r-4: the dW accumulator of round r holds earlier tiles
        while (left0 > 0 || left1 > 0) {
#pragma unroll 1
            for (int s = 0; s < kSlots; ++s) {
                if ((s == 0 ? left0 : left1) <= 0) continue;
                const int rnd = s == 0 ? round0 : round1;
                handoff_wait(s);
                tc::tc_fence_after_sync();
                uint8_t* slot = smem + Smem::slot0 + s * Smem::slot_bytes;
                const uint32_t Z = tmem_base + kColZ + 128u * s;
                const uint8_t* E = slot + Smem::e;
                const uint8_t* H = slot + Smem::h;
                const uint8_t* D = slot + Smem::d;
                uint64_t* done = &bars[s];
                uint64_t* dw_done = &bars[2 + s];
                if (elect_one()) {
                switch (rnd) {
                case 4:     // dh1 = d2 W2;  dW2 | db2 += d2^T [h1 | 1]
                    gemm3<kWidth / 16>(Z, kmajor<TD>(D), mnmajor<TW2>(w2),
                                       tc::instr_desc_bf16(128, kWidth, false, true), false);
                    commit_to(done);
                    gemm3<kTile / 16>(tmem_base + kColDW2, mnmajor<TD>(D), mnmajor<TH>(H),
                                      tc::instr_desc_bf16(64, 72, true, true), (acc_mask >> 0) & 1u);
                    commit_to(dw_done);
                    break;
                case 5:     // din1 = d1 W1;  dW1 | db1 += d1^T [SH | geo | 1]
                    gemm3<kWidth / 16>(Z, kmajor<TD>(D), mnmajor<TW1>(w1),
                                       tc::instr_desc_bf16(128, kHeadIn, false, true), false);
                    commit_to(done);
                    gemm3<kTile / 16>(tmem_base + kColDW1, mnmajor<TD>(D), mnmajor<TE>(E),
                                      tc::instr_desc_bf16(64, kHeadIn, true, true), (acc_mask >> 1) & 1u);
                    commit_to(dw_done);
                    break;
                case 6:     // dhb = dy Wb2;  dWb2^T += hb^T dy
                    gemm3<kBaseOut / 16>(Z, kmajor<TE>(E), mnmajor<TWb2>(wb2),
                                         tc::instr_desc_bf16(128, kWidth, false, true), false);
                    commit_to(done);
                    gemm3<kTile / 16>(tmem_base + kColDWb2T, mnmajor<TH>(H), mnmajor<TE>(E),
                                      tc::instr_desc_bf16(64, kBaseOut, true, true), (acc_mask >> 2) & 1u);
                    commit_to(dw_done);
                    break;
                default:    // denc = db1 Wb1;  dWb1 | dbb1 += db1^T [enc | 1]
                    gemm3<kWidth / 16>(Z, kmajor<TD>(D), mnmajor<TWb1>(wb1),
                                       tc::instr_desc_bf16(128, kEncDim, false, true), false);
                    commit_to(done);
                    gemm3<kTile / 16>(tmem_base + kColDWb1, mnmajor<TD>(D), mnmajor<TE>(E),
                                      tc::instr_desc_bf16(64, 40, true, true), (acc_mask >> 3) & 1u);
                    commit_to(dw_done);
                    break;
                }
                }
                __syncwarp();
                if (rnd >= 4) acc_mask |= 1u << (rnd - 4);
                if (s == 0) { round0 = 4 + ((rnd + 1) & 3); --left0; } else { round1 = 4 + ((rnd + 1) & 3); --left1; }
            }
        }
        // end of the flush period: the epilogue groups have awaited every GEMM; warps 0..3 drain the
        // accumulators between the two barriers, the next period starts them afresh (accumulate = 0)
        tc::tc_fence_before_sync();
        __syncthreads();
        __syncthreads();
        tc::tc_fence_after_sync();
        }
    } else {
        // ===================== epilogue warps: one group of 8 per slot =====================
        const int slot_id = warp >> 3;
        const int q = warp & 3, hf = (warp >> 2) & 1;
        const int row = q * 32 + lane;
        uint8_t* slot = smem + Smem::slot0 + slot_id * Smem::slot_bytes;
        uint8_t* E = slot + Smem::e;
        uint8_t* H = slot + Smem::h;
        uint8_t* D = slot + Smem::d;
        uint64_t* done = &bars[slot_id];
        uint64_t* dw_done = &bars[2 + slot_id];
        float* zx = reinterpret_cast<float*>(smem + Smem::zx) + slot_id * (kTile * 2 * 4);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t Z = lane_base + kColZ + 128u * slot_id, P = lane_base + kColP + 128u * slot_id;
        const int hact = f.hidden_act;
        uint32_t phase = 0, dw_phase = 0;
        bool dw_pending = false;         // the slot's previous tile left its last dW GEMM un-awaited
        // the weight-gradient GEMM of the previous round must have consumed its operand tiles before
        // they are overwritten (it was committed separately, after the latency-critical GEMM)
        auto await_dw = [&]() {
            tc::mbar_wait(dw_done, dw_phase);
            dw_phase ^= 1;
        };

        // load this thread's 16 encoding features of row i into E chunks (2 hf, 2 hf + 1)
        // forward rounds: the 8 warps of the slot meet at their named barrier and the slot's first warp
        // elects the thread that issues the GEMM (it only writes this slot's own scratch columns)
        const uint32_t Zd = tmem_base + kColZ + 128u * slot_id;
        auto launch_fwd = [&](int round) {
            tc::fence_smem_to_async_proxy();
            tc::tc_fence_before_sync();
            group_sync(slot_id);
            if ((warp & 7) == 0) {
                tc::tc_fence_after_sync();
                if (elect_one()) {
                    if (round == 0)        // z_b1 = enc Wb1^T
                        gemm3<kEncDim / 16>(Zd, kmajor<TE>(E), kmajor<TWb1>(smem + Smem::wb1),
                                            tc::instr_desc_bf16(128, kWidth, false, false), false);
                    else if (round == 1)   // y = hb Wb2^T
                        gemm3<kWidth / 16>(Zd, kmajor<TH>(H), kmajor<TWb2>(smem + Smem::wb2),
                                           tc::instr_desc_bf16(128, kBaseOut, false, false), false);
                    else if (round == 2)   // z1 = [SH | geo | 1] W1^T
                        gemm3<kHeadIn / 16>(Zd, kmajor<TE>(E), kmajor<TW1>(smem + Smem::w1),
                                            tc::instr_desc_bf16(128, kWidth, false, false), false);
                    else                   // z2 = h1 W2^T
                        gemm3<kWidth / 16>(Zd, kmajor<TH>(H), kmajor<TW2>(smem + Smem::w2),
                                           tc::instr_desc_bf16(128, kWidth, false, false), false);
                    commit_to(done);
                }
                __syncwarp();
            }
        };

        // this thread's 16 encoding features of row i: global -> registers ...
        auto load_enc = [&](int64_t i, bool valid, float (&x)[16]) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = 0.f;
            if (valid) {
                // enc_rows: sample i's encoding is row enc_rows[i] of `enc` (the survivors of the
                // visibility filter read the pre-pass encodings in place, no compacted copy)
                const int64_t src_row = enc_rows ? (int64_t)__ldg(enc_rows + i) : i;
                const float4* src = reinterpret_cast<const float4*>(enc + src_row * enc_dim + 16 * hf);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4)
                    if (16 * hf + 4 * v4 < enc_dim) {
                        const float4 v = __ldg(src + v4);
                        x[4 * v4] = v.x; x[4 * v4 + 1] = v.y; x[4 * v4 + 2] = v.z; x[4 * v4 + 3] = v.w;
                    }
            }
        };
        // ... -> split -> E chunks (2 hf, 2 hf + 1), and the packed words parked in TMEM: enc is needed
        // again for the last round (dWb1 += db1^T enc) after [SH | geo] and dy have passed through E
        const uint32_t EP = lane_base + kColEncPark + 32u * slot_id + 16u * hf;
        auto stage_enc = [&](const float (&x)[16]) {
            uint32_t words[16];
            store16_keep<TE>(E, row, 2 * hf, x, words);
            tmem_st16(EP, words);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        };
        auto restore_enc = [&]() {
            uint32_t words[16];
            tmem_ld16_words(EP, words);
            store16_words<TE>(E, row, 2 * hf, words);
        };
        // hb = act(z_b1 + bb1) for this thread's 32 columns -> H, and the packed hi/lo words parked in
        // TMEM: hb is needed again after h1 has overwritten H (dWb2^T += hb^T dy, act'(hb)), and copying
        // 32 words back costs a twentieth of re-deriving them (softplus = 2 MUFU, split = 3 instr / element)
        auto stage_hb = [&]() {
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float h[16];
                uint32_t words[16];
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, h);
                bias_hidden_act<16>(hact, h, s_bb1 + 32 * hf + 16 * c);
                store16_keep<TH>(H, row, 4 * hf + 2 * c, h, words);
                tmem_st16(P + 32 * hf + 16 * c, words);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        };
        auto restore_hb = [&]() {
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t words[16];
                tmem_ld16_words(P + 32 * hf + 16 * c, words);
                store16_words<TH>(H, row, 4 * hf + 2 * c, words);
            }
        };

        // software pipeline over this slot's tiles: the encoding row and the ray index of tile k + 2 are
        // loaded while tile k drains, so a tile never starts on an HBM miss
        float xe[16];
        int64_t i = ((int64_t)blockIdx.x + (int64_t)slot_id * gridDim.x) * kTile + row;
        bool valid = slot_id < my_tiles && i < n;
        int64_t ray = 0;
        float tmid2 = 0.f;                       // t_start + t_end
        load_enc(i, valid, xe);
        if (valid) { ray = ray_indices[i]; tmid2 = t_starts[i] + t_ends[i]; }

        for (int64_t c0 = 0; c0 < my_tiles; c0 += kFlushTiles) {
        const int64_t c1 = min(c0 + (int64_t)kFlushTiles, my_tiles);
        for (int64_t k = c0 + slot_id; k < c1; k += kSlots) {
            // ---- operands of round 0: enc -------------------------------------------------------
            float dir[3] = {0.f, 0.f, 1.f};
            bool inside = false;
            if (valid) {
                float pos[3], u[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    dir[d] = __ldg(rays_d + 3 * ray + d);
                    pos[d] = __ldg(rays_o + 3 * ray + d) + (dir[d] * tmid2) * 0.5f;
                }
                if (hf == 0) inside = contract_position(f, pos, u);
            }
            // the last weight-gradient GEMM of this slot's previous tile (dWb1: reads D and E) is awaited
            // here, not at the end of that tile
            if (dw_pending) await_dw();
            dw_pending = true;
            stage_enc(xe);
            launch_fwd(0);

            // ---- round 0 done: hb ----------------------------------------------------------------
            await_mma(done, phase);
            stage_hb();
            launch_fwd(1);

            // ---- round 1 done: y -> raw density, [SH | geo | 1] -> E -------------------------------
            float raw = 0.f;
            await_mma(done, phase);
            {
                float x[16];
                if (hf == 0) {
                    float y[16];
                    tmem_ld_cols<16>(Z, y);
#pragma unroll
                    for (int j = 0; j < kBaseOut; ++j) y[j] += s_bb2[j];
                    raw = y[0];
#pragma unroll
                    for (int j = 0; j < kGeo; ++j) x[j] = y[1 + j];
                    x[15] = 1.f;                          // the ones column of the head input
                    store16<TE>(E, row, 2, x);
                } else {
                    sh_degree4(dir, x);
                    store16<TE>(E, row, 0, x);
                }
            }
            launch_fwd(2);

            // ---- round 2 done: h1 -> H -------------------------------------------------------------
            await_mma(done, phase);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float h[16];
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, h);
                bias_hidden_act<16>(hact, h, s_b1 + 32 * hf + 16 * c);
                store16<TH>(H, row, 4 * hf + 2 * c, h);
            }
            launch_fwd(3);

            // next tile of this slot: pull its rows towards L2 while this one is in flight
            if (k + kSlots < my_tiles) {
                const int64_t ni = i + (int64_t)kSlots * gridDim.x * kTile;
                if (ni < n) {
                    if (!enc_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(enc + ni * enc_dim + 16 * hf));
                    if (hf == 0) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ray_indices + ni));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(t_starts + ni));
                    } else {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(t_ends + ni));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(d_rgbs + ni * C));
                    }
                }
            }

            // ---- round 3 done: h2 (registers), output layer forward + backward on the SIMT side ---
            float g_rgb[3] = {0.f, 0.f, 0.f};
            float g_sigma = 0.f;
            if (valid) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (c < C) g_rgb[c] = d_rgbs[i * C + c];
                if (hf == 0) g_sigma = d_sigmas[i];
            }
            await_mma(done, phase);
            {
                float h2[32];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float h[16];
                    tmem_ld_cols<16>(Z + 32 * hf + 16 * c, h);
                    bias_hidden_act<16>(hact, h, s_b2 + 32 * hf + 16 * c);
#pragma unroll
                    for (int j = 0; j < 16; ++j) h2[16 * c + j] = h[j];
                }
                float z3[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (c < C) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) z3[c] = fmaf(h2[j], s_w3f[c * kWidth + 32 * hf + j], z3[c]);
                    }
                *reinterpret_cast<float4*>(zx + (row * 2 + hf) * 4) = make_float4(z3[0], z3[1], z3[2], 0.f);
                group_sync(slot_id);
                const float4 other = *reinterpret_cast<const float4*>(zx + (row * 2 + (hf ^ 1)) * 4);
                float d3[3];
                d3[0] = z3[0] + other.x; d3[1] = z3[1] + other.y; d3[2] = z3[2] + other.z;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    d3[c] = c < C ? g_rgb[c] * radiance_act_grad(f.radiance_act, d3[c] + s_b3[c]) : 0.f;
                // d2 = (d3 W3) * act'(h2) -> D
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    float dl[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = 0.f;
#pragma unroll
                        for (int c = 0; c < 3; ++c) a = fmaf(d3[c], s_w3f[c * kWidth + 32 * hf + 16 * c2 + j], a);
                        dl[j] = a * hidden_act_grad_from_out(hact, h2[16 * c2 + j]);
                    }
                    store16<TD>(D, row, 4 * hf + 2 * c2, dl);
                }
                publish(slot_id);
                // dW3 += d3^T h2, db3 += sum d3  (off the critical path: the MMA round is running)
                for (int c = 0; c < C; ++c) {
                    float t[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] = d3[c] * h2[j];
                    const float s = warp_transpose_sum<32>(t, lane);
                    atomicAdd(&s_dw3[c * kWidth + 32 * hf + lane], s);
                    if (hf == 0) {
                        const float sb = warp_sum(d3[c]);
                        if (lane == 0) atomicAdd(&s_db3[c], sb);
                    }
                }
            }

            // ---- round 4 done: dh1 -> d1 = dh1 * act'(h1) -> D ------------------------------------
            await_mma(done, phase);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float dl[16], h[16];
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, dl);
                load16<TH>(H, row, 4 * hf + 2 * c, h);
                mul_hidden_act_grad<16>(hact, dl, h);
                if (c == 0) await_dw();                   // dW2 has read d2 (D) and h1 (H)
                store16<TD>(D, row, 4 * hf + 2 * c, dl);
            }
            publish(slot_id);

            // ---- round 5 done: din1 -> dy (E), hb again (H) -----------------------------------------
            await_mma(done, phase);
            if (hf == 0) {
                float dgeo[16], dy[16];
                tmem_ld_cols<16>(Z + kShDim, dgeo);
                dy[0] = inside ? g_sigma * density_act_grad(f.density_act, raw) : 0.f;
#pragma unroll
                for (int j = 0; j < kGeo; ++j) dy[1 + j] = dgeo[j];
                await_dw();                               // dW1 has read d1 (D) and [SH | geo | 1] (E)
                store16<TE>(E, row, 0, dy);
                restore_hb();
                publish(slot_id);
                const float s = warp_transpose_sum<16>(dy, lane & 15);
                // lanes l and l + 16 hold the two half-warp sums of column l
                const float tot = s + __shfl_xor_sync(0xffffffffu, s, 16);
                if (lane < 16) atomicAdd(&s_dbb2[lane], tot);
            } else {
                if (d_dirs != nullptr) {
                    // dL/d(view direction) through the SH encoding (only the tau path needs it)
                    float dsh[16], dd[3];
                    tmem_ld_cols<16>(Z, dsh);
                    sh_degree4_grad(dir, dsh, dd);
                    if (valid) { d_dirs[3 * i] = dd[0]; d_dirs[3 * i + 1] = dd[1]; d_dirs[3 * i + 2] = dd[2]; }
                }
                await_dw();
                restore_hb();
                publish(slot_id);
            }

            // ---- round 6 done: dhb -> db1 = dhb * act'(hb) -> D; enc again -> E -----------------------
            await_mma(done, phase);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float dl[16], h[16];
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, dl);
                load16<TH>(H, row, 4 * hf + 2 * c, h);
                mul_hidden_act_grad<16>(hact, dl, h);
                store16<TD>(D, row, 4 * hf + 2 * c, dl);
            }
            await_dw();                                   // dWb2^T has read hb (H) and dy (E)
            restore_enc();
            publish(slot_id);

            // ---- round 7 done: denc -> HBM ---------------------------------------------------------------
            const int64_t i_cur = i;
            const bool valid_cur = valid;
            i += (int64_t)kSlots * gridDim.x * kTile;     // this slot's next tile: loads in flight across the wait
            valid = k + kSlots < my_tiles && i < n;
            load_enc(i, valid, xe);
            ray = 0; tmid2 = 0.f;
            if (valid) { ray = ray_indices[i]; tmid2 = t_starts[i] + t_ends[i]; }
            await_mma(done, phase);
            {
                float de[16];
                tmem_ld_cols<16>(Z + 16 * hf, de);
                if (valid_cur) {
                    float4* out = reinterpret_cast<float4*>(d_enc + i_cur * enc_dim + 16 * hf);
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4)
                        if (16 * hf + 4 * v4 < enc_dim)
                            out[v4] = make_float4(de[4 * v4], de[4 * v4 + 1], de[4 * v4 + 2], de[4 * v4 + 3]);
                }
            }
            tc::tc_fence_before_sync();
        }
        // end of the flush period (see the MMA warp): every GEMM of both slots has been awaited
        if (dw_pending) await_dw();                       // dWb1 of the slot's last tile
        dw_pending = false;
        tc::tc_fence_before_sync();
        __syncthreads();
        if (warp < 4) flush_tmem();
        __syncthreads();
        tc::tc_fence_after_sync();
        }
    }

    // ---- the SIMT-side accumulators (dW3, dbb2, db3: fp32 shared-memory atomics) go out once ----------
    tc::tc_fence_before_sync();
    __syncthreads();
    if (my_tiles > 0 && warp < 4) {
        for (int i = tid; i < C * kWidth; i += 128) atomicAdd(g.w3 + i, s_dw3[i]);
        if (tid < kBaseOut) atomicAdd(g.bb2 + tid, s_dbb2[tid]);
        if (tid < C) atomicAdd(g.b3 + tid, s_db3[tid]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kTmemCols);
}

// pipeline v4 (den_mlp_tc_bwd4.cu): slot-issued GEMMs, private accumulators, tensor-memory A operands
int launch_mlp_bwd4(const den_field_desc* f, const den_field_params* p, const den_field_grads* g,
                    const float* enc, const float* rays_o, const float* rays_d, const int32_t* ray_indices,
                    const float* t_starts, const float* t_ends, const float* d_sigmas, const float* d_rgbs,
                    int64_t n, const int32_t* n_dev, const int32_t* enc_rows, float* d_enc, float* d_dirs,
                    cudaStream_t stream);

}  // namespace den

extern "C" int den_mlp_bwd(const den_field_desc* f, const den_field_params* p,
                           const den_field_grads* g, const float* enc, const float* rays_o,
                           const float* rays_d, const int32_t* ray_indices, const float* t_starts,
                           const float* t_ends, const float* d_sigmas, const float* d_rgbs,
                           int64_t n, const int32_t* n_dev, const int32_t* enc_rows, float* d_enc,
                           float* d_dirs, void* stream) {
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
    DEN_CHECK_ARG(f->channels >= 1 && f->channels <= 3, "1 to 3 radiance channels");
    // DEN_MLP_BWD_VARIANT: 0 = pipeline v3 (this file: 2 x 8 epilogue warps + MMA warp), 4 = pipeline v4
    // (den_mlp_tc_bwd4.cu); read once per process
    static const int forced = [] {
        const char* e = getenv("DEN_MLP_BWD_VARIANT");
        return e != nullptr ? (int)(e[0] - '0') : -1;
    }();
    const int variant = forced >= 0 ? forced : 4;       // v4 measured 17.0 ms against 19.6 ms at 40.8 M samples
    if (variant == 4)
        return launch_mlp_bwd4(f, p, g, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, d_sigmas,
                               d_rgbs, n, n_dev, enc_rows, d_enc, d_dirs, as_stream(stream));
    // two tiles in flight per CTA: at least two tiles per CTA whenever there are enough of them
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int grid = grid_for((n_tiles + 1) / 2, 1, 1);
    cudaFuncSetAttribute(mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)bwd::Smem::total);
    mlp_bwd_tc_kernel<<<grid, bwd::kThreads, bwd::Smem::total, as_stream(stream)>>>(
        *f, *p, *g, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, d_sigmas, d_rgbs, n, n_dev, enc_rows,
        d_enc, d_dirs);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}
