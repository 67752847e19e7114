// Backward of the tensor-core MLP: forward recompute + input gradient + weight gradients, fused per
// 128-sample tile (tcgen05.mma, TMEM accumulators; sm_100a), TWO tiles ("slots") in flight per CTA.
//
// Replaces the autograd of MLP.forward / SHEncoder / activations under NGPradianceField.forward
// (external/ngp.py:269-280) — in the reference ~30 cuBLAS and elementwise launches that read and write
// every (M,64) activation twice.  Here nothing but enc (128 B/sample in), the two upstream gradients and
// dL/denc (128 B/sample out) touches HBM.  Eight GEMM rounds per tile: forward recompute 0..3
// (enc -> hb -> y -> h1 -> h2), backward 4..7, each backward round = a dX GEMM (critical path) + a
// weight-gradient GEMM dW += dY^T X with K = the 128 samples of the tile.  The dW GEMMs need BOTH operands
// transposed, so dY and X live in shared memory as bf16 hi / lo operand tiles that are read K-major
// (dX = dY W) and MN-major (dW) without being rewritten; bias gradients ride along as a ones column of X.
// Weight-gradient accumulators (64 x N fp32) stay in TMEM across the tiles of a CTA and are flushed with
// atomics every kFlushTiles tiles (the tensor core accumulates with truncation: profiles/wgrad_check.py).
// The (C <= 3)-row output layer (forward, backward and dW3) runs on the SIMT side.
//
// Pipeline v4 (this file).  v3 (round 1 / early round 2, 19.6 ms at 40.8 M samples) had a 17th warp issuing
// the backward rounds of both slots because their dW GEMMs shared accumulators.  What v4 changes:
//
//   * every slot owns its OWN weight-gradient accumulators in the SAME columns: an M = 64 accumulator only
//     occupies the lower 16 lanes of each 32-lane TMEM sub-partition (row m -> lane 32 (m / 16) + m % 16),
//     so slot 1's accumulators live at lane offset 16 (the interleaved half-sub-partition allocation
//     CUTLASS calls TmemAllocMode::Interleaved).  A slot then issues all eight rounds of its tile itself,
//     straight after its own hand-off barrier: no hop to another warp, no fixed slot alternation, no
//     96-register cap (16 warps, 128 registers).  Round r is issued by warp r of the slot.
//   * ncu + an ablation (profiles/r02_mlp_bwd_v4.md) showed what bounds the kernel: not warps (2 x 16 warps
//     with 64 registers: 24.8 ms against 19.05 ms, same 43 % issue utilisation, +28 % instructions) but
//     the epilogue's own instruction stream (~5.4 k thread-instructions per sample: activation, bf16 hi / lo
//     split, TMEM <-> register traffic) plus SHARED-MEMORY BANDWIDTH — the 171 SS-form MMAs of a tile fetch
//     ~720 KB of operand tiles (the "44 cycles per small-N MMA" of v3 is exactly that fetch at 128 B/clk)
//     next to ~3.8 k LSU wavefronts of epilogue stores.  Hence:
//   * the hb / enc words that v3 parked in tensor memory (only to survive until the last rounds) are laid
//     out as tcgen05 A operands and FEED rounds 0 and 1 directly (ts-form MMA); their shared-memory tiles are
//     written only when a weight-gradient GEMM needs them.  Round 7 takes db1 from the hb columns it has just
//     consumed, round 6 takes dy from 16 spare columns, act'(hb) reads the parked words instead of the H
//     tile.  All 512 TMEM columns are in use; rounds 2..5 keep shared-memory A operands because their
//     operands (h1, d2, d1: 64 columns per slot each) have no columns left.
//   * b1 rides in column 31 of the W1 operand tile (the head input carries a ones column there already);
//     the W3 loops read only the C live rows; the SIMT-side sums (dW3, dbb2, db3) stay in registers across
//     tiles (shared-memory float atomics are CAS loops); the ray of the slot's next tile (sample -> ray
//     index -> origin / direction, a dependent chain) is fetched a tile ahead.
//   Result: 16.97 ms at 40.8 M samples (v3 19.6), parity unchanged (tests/test_gpu_mlp_tc.py).
// Tried and rejected, measured: back-off (nanosleep) in the barrier polls (the polls are 16 % of the issued
// instructions but only use idle slots: no change); per-thread register row sums for dW3 instead of a
// warp transpose per tile (spills under the 128-register cap: 17.7 ms); all 16 warps on one slot's phase
// at a time (v3: 6.9 against 5.3 ms per 10.2 M samples); SAVING the forward activations instead of
// recomputing them (v3: forward 1.71 -> 2.05 ms, backward 4.85 -> 5.27 ms per 10.2 M samples).
// Shared memory per slot: E (128 x 40: [SH | geo | 1], later dy, later enc), H (128 x 72: h1 | ones,
// later hb) and D (128 x 64: the current dL/dz tile) = 88 KB; 2 slots + 36 KB of weight tiles = 223 KB.
#include <stdlib.h>

#include "den_mlp_ops.cuh"

namespace den {

using namespace mlp;

namespace bwd {

constexpr int kSlots = 2;
constexpr int kGroupWarps = 8;               // per slot: quadrant q = warp % 4 (rows), half hf = (warp / 4) % 2 (columns)
constexpr int kGroupThreads = 32 * kGroupWarps;
constexpr int kThreads = kSlots * kGroupThreads;
constexpr uint32_t kTmemCols = 512;
constexpr int kFlushTiles = 128;             // tiles of a CTA between two flushes (see den_mlp_tc_bwd.cu)

// TMEM column plan: per slot 160 columns (scratch, hb words, enc words); the weight-gradient
// accumulators share columns, slot s at lane offset 16 s; 16 more columns per slot for the dy operand.
// The word columns are tcgen05 A operands: per K step of 16 values, 8 columns of packed bf16 pairs
// (hi) followed by 8 columns (lo).
constexpr uint32_t kSlotCols = 160;
constexpr uint32_t kColZ = 0;                // 64 scratch columns (forward / dX results)
constexpr uint32_t kColP = 64;               // 64 columns: hb (A of round 1, parked until round 6), then db1 (A of round 7)
constexpr uint32_t kColEncPark = 128;        // 32 columns: enc (A of round 0, parked until round 7's dWb1)
constexpr uint32_t kColDW2 = 320;            // 72: dW2 (out 64 x in 64) | db2 x 8
constexpr uint32_t kColDW1 = 392;            // 32: dW1 (64 x 31) | db1 in column 31
constexpr uint32_t kColDWb1 = 424;           // 40: dWb1 (64 x 32) | dbb1 x 8
constexpr uint32_t kColDWb2T = 464;          // 16: dWb2^T (in 64 x out 16)
constexpr uint32_t kColDyA = 480;            // + 16 * slot : dy (A of round 6)
static_assert(kColDyA + 16 * kSlots <= kTmemCols && kSlots * kSlotCols <= kColDW2, "TMEM plan");

using TE = OpTile<kTile, 5>;     // [SH | geo | 1] | ones  /  dy  /  enc
using TH = OpTile<kTile, 9>;     // h1 | ones      /  hb
using TD = OpTile<kTile, 8>;     // dL/dz of the current 64-wide layer
using TWb1 = OpTile<kWidth, 4>;  // (64, 32)
using TWb2 = OpTile<kBaseOut, 8>;// (16, 64)
using TW1 = OpTile<kWidth, 4>;   // (64, 32): column 31 carries b1
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

// 16 values -> the 16 operand words of one K step: w[0..7] = bf16 pairs of the hi parts, w[8..15] = lo parts
__device__ __forceinline__ void split16_words(const float (&v)[16], uint32_t (&w)[16]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint4 hi, lo;
        split8(&v[8 * c], hi, lo);
        w[4 * c] = hi.x; w[4 * c + 1] = hi.y; w[4 * c + 2] = hi.z; w[4 * c + 3] = hi.w;
        w[8 + 4 * c] = lo.x; w[8 + 4 * c + 1] = lo.y; w[8 + 4 * c + 2] = lo.z; w[8 + 4 * c + 3] = lo.w;
    }
}
// the same words -> chunks (chunk0, chunk0 + 1) of a shared-memory operand tile
template <class T>
__device__ __forceinline__ void store_kstep_words(uint8_t* tile, int r, int chunk0, const uint32_t (&w)[16]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const uint32_t o = T::off(r, chunk0 + c);
        *reinterpret_cast<uint4*>(tile + o) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        *reinterpret_cast<uint4*>(tile + T::half + o) =
            make_uint4(w[8 + 4 * c], w[8 + 4 * c + 1], w[8 + 4 * c + 2], w[8 + 4 * c + 3]);
    }
}
// ... and back to fp32 (hi + lo)
__device__ __forceinline__ void kstep_words_to_float(const uint32_t (&w)[16], float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        v[2 * q] = __uint_as_float(w[q] << 16) + __uint_as_float(w[8 + q] << 16);
        v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u) + __uint_as_float(w[8 + q] & 0xffff0000u);
    }
}
// tcgen05.wait::ld that also names the loaded registers, so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld_dep(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                   "+r"(r[15])
                 :
                 : "memory");
}
template <int N>
__device__ __forceinline__ void hidden_act_vec(int id, float (&h)[N]) {
    if (id == kActSoftplus100) {
#pragma unroll
        for (int j = 0; j < N; ++j) h[j] = softplus100(h[j]);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) h[j] = fmaxf(h[j], 0.f);
    }
}
__device__ __forceinline__ void commit_to(uint64_t* bar) {      // by the elected issuing thread
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     tc::smem_u32(bar))
                 : "memory");
}
// D = A * B, A = word columns in tensor memory (16 per K step: hi 8 | lo 8), B = hi / lo tiles in shared
// memory, 3-product split
template <int KSTEPS>
__device__ __forceinline__ void gemm3_tw(uint32_t tmem_d, uint32_t a, const OpDesc& b, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) mma1_ts(tmem_d, a + 16 * ks, b.hi.at(ks * b.kstep), idesc, ks > 0 ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) mma1_ts(tmem_d, a + 16 * ks + 8, b.hi.at(ks * b.kstep), idesc, 1u);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) mma1_ts(tmem_d, a + 16 * ks, b.lo.at(ks * b.kstep), idesc, 1u);
}

}  // namespace bwd

__global__ void __launch_bounds__(bwd::kThreads, 1)
mlp_bwd_tc_kernel(const __grid_constant__ den_field_desc f, const __grid_constant__ den_field_params p,
                   const __grid_constant__ den_field_grads g, const float* __restrict__ enc,
                   const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                   const int32_t* __restrict__ ray_indices, const float* __restrict__ t_starts,
                   const float* __restrict__ t_ends, const float* __restrict__ d_sigmas,
                   const float* __restrict__ d_rgbs, int64_t n, const int32_t* __restrict__ n_dev,
                   const int32_t* __restrict__ enc_rows, float* __restrict__ d_enc,
                   float* __restrict__ d_dirs) {
    using namespace bwd;
    using S = Smem;
    n = effective_n(n, n_dev);
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bars);          // done[0..1], dw_done[0..1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::tmem_ptr);
    const float* s_bb1 = reinterpret_cast<const float*>(smem + S::bias);
    const float* s_bb2 = s_bb1 + kWidth;
    const float* s_b2 = s_bb2 + kBaseOut + kWidth;
    const float* s_b3 = s_b2 + kWidth;
    float* s_w3f = reinterpret_cast<float*>(smem + S::w3f);
    float* s_dw3 = reinterpret_cast<float*>(smem + S::acc);                // (3, 64)
    float* s_dbb2 = s_dw3 + 3 * kWidth;                                    // 16
    float* s_db3 = s_dbb2 + kBaseOut;                                      // 4

    const int enc_dim = f.grid.n_levels * 2;
    const int C = f.channels;

    // ---- setup ---------------------------------------------------------------------------------
    tc::load_weight_split(smem + S::wb1, smem + S::wb1 + TWb1::half, p.wb1, kWidth, enc_dim, kWidth, kEncDim);
    tc::load_weight_split(smem + S::wb2, smem + S::wb2 + TWb2::half, p.wb2, kBaseOut, kWidth, kBaseOut, kWidth);
    tc::load_weight_split(smem + S::w1, smem + S::w1 + TW1::half, p.w1, kWidth, kShDim + kGeo, kWidth, kHeadIn);
    tc::load_weight_split(smem + S::w2, smem + S::w2 + TW2::half, p.w2, kWidth, kWidth, kWidth, kWidth);
    {
        float* b = reinterpret_cast<float*>(smem + S::bias);
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
        uint8_t* slot = smem + S::slot0 + s * S::slot_bytes;
        const uint4 one = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(slot + S::e + TE::off(r, 4)) = one;
        *reinterpret_cast<uint4*>(slot + S::e + TE::half + TE::off(r, 4)) = zero;
        *reinterpret_cast<uint4*>(slot + S::h + TH::off(r, 8)) = one;
        *reinterpret_cast<uint4*>(slot + S::h + TH::half + TH::off(r, 8)) = zero;
    }
    if (tid == 0) {
        for (int b = 0; b < 4; ++b) tc::mbar_init(&bars[b], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemCols);
    __syncthreads();
    // b1 rides in column 31 of the W1 operand tile: the head input [SH | geo | 1] carries its ones column
    // there (the weight-gradient accumulator of W1 collects db1 in the same column)
    if (tid < kWidth) {
        const float v = __ldg(p.b1 + tid);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        const uint32_t off = tc::chunk_offset(tid, (kHeadIn - 1) >> 3, kHeadIn) + ((kHeadIn - 1) & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(smem + S::w1 + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(smem + S::w1 + TW1::half + off) = l;
    }
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    // tiles of this CTA: blockIdx.x + k * gridDim.x, k = 0 .. my_tiles-1; slot s takes k = s, s+2, ...
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    // TMEM weight-gradient accumulators -> global (warps 0..3 = the four quadrants; lane l < 16: slot 0's
    // row 16 q + l, lane l >= 16: slot 1's row 16 q + l - 16; a slot without a tile in the period holds
    // nothing to add)
    auto flush_tmem = [&](bool has0, bool has1) {
        tc::tc_fence_after_sync();
        const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int row = warp * 16 + (lane & 15);
        const bool owner = lane < 16 ? has0 : has1;
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

    // ===================== every warp is an epilogue warp: one group of 8 per slot =====================
    const int slot_id = warp / kGroupWarps;
    const int wslot = warp - slot_id * kGroupWarps;
    const int q = warp & 3, hf = wslot >> 2;               // TMEM quadrant (rows), column half
    const int row = q * 32 + lane;
    uint8_t* slot = smem + S::slot0 + slot_id * S::slot_bytes;
    uint8_t* E = slot + S::e;
    uint8_t* H = slot + S::h;
    uint8_t* D = slot + S::d;
    uint64_t* done = &bars[slot_id];
    uint64_t* dw_done = &bars[2 + slot_id];
    float* zx = reinterpret_cast<float*>(smem + S::zx) + slot_id * (kTile * 2 * 4);
    const uint32_t slot_cols = tmem_base + kSlotCols * slot_id;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t Z = slot_cols + lane_base + kColZ, P = slot_cols + lane_base + kColP;
    const uint32_t EP = slot_cols + lane_base + kColEncPark + 16u * hf;
    const uint32_t DYA = tmem_base + lane_base + kColDyA + 16u * slot_id;
    const uint32_t Zd = slot_cols + kColZ, Pd = slot_cols + kColP;         // operand / accumulator addresses (lane 0)
    const uint32_t EPd = slot_cols + kColEncPark, DYAd = tmem_base + kColDyA + 16u * slot_id;
    const uint32_t acc_base = tmem_base + ((uint32_t)(16 * slot_id) << 16);
    const int hact = f.hidden_act;
    uint32_t phase = 0, dw_phase = 0;
    bool dw_pending = false;         // the slot's previous tile left its last dW GEMM un-awaited
    bool acc_live = false;           // the slot's weight-gradient accumulators hold earlier tiles of the period
    // SIMT-side sums kept in registers across the tiles of this thread and added to shared memory once:
    // lane l of a warp owns column 32 hf + l of dW3 (per channel), lanes < 16 of the hf = 0 warps own dbb2,
    // lane 0 of the hf = 0 warps db3  (shared-memory float atomics are CAS loops: 16 warps contending on
    // the same 32 words every tile)
    float r_dw3[3] = {0.f, 0.f, 0.f}, r_db3[3] = {0.f, 0.f, 0.f}, r_dbb2 = 0.f;
    auto await_mma = [&]() {
        tc::mbar_wait(done, phase);
        phase ^= 1;
        tc::tc_fence_after_sync();
    };
    // the weight-gradient GEMM of the previous round must have consumed its operand tiles before they
    // are overwritten (it was committed separately, after the latency-critical GEMM)
    auto await_dw = [&]() {
        tc::mbar_wait(dw_done, dw_phase);
        dw_phase ^= 1;
    };
    auto group_sync = [&]() {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + slot_id), "r"(kGroupThreads) : "memory");
    };

    // hand-off: this thread's part of the operands is written (shared memory: generic -> async proxy
    // fence; tensor memory: tcgen05.wait::st by the writer); the slot's first warp waits for the whole
    // group and one elected thread issues round `rnd`; the other warps arrive and move on.  Round r is
    // issued by warp r of the slot: issuing costs ~5 instructions per MMA (171 MMAs per tile, a third of
    // a warp's epilogue work), and a fixed issuing warp would be the straggler of every hand-off.  The
    // accumulating GEMMs of a given accumulator always come from the same thread (same round, same
    // warp), and successive rounds are ordered through the mbarrier they wait on.
    auto launch = [&](int rnd) {
        tc::fence_smem_to_async_proxy();
        tc::tc_fence_before_sync();
        if (wslot != rnd) {
            asm volatile("bar.arrive %0, %1;" ::"r"(3 + slot_id), "r"(kGroupThreads) : "memory");
            return;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(3 + slot_id), "r"(kGroupThreads) : "memory");
        tc::tc_fence_after_sync();
        const uint32_t accf = acc_live ? 1u : 0u;
        if (elect_one()) {
            const uint8_t* wb1 = smem + S::wb1;
            const uint8_t* wb2 = smem + S::wb2;
            const uint8_t* w1 = smem + S::w1;
            const uint8_t* w2 = smem + S::w2;
            switch (rnd) {
            case 0:     // z_b1 = enc Wb1^T                       (A: enc words in tensor memory)
                gemm3_tw<kEncDim / 16>(Zd, EPd, kmajor<TWb1>(wb1), tc::instr_desc_bf16(128, kWidth, false, false));
                commit_to(done);
                break;
            case 1:     // y = hb Wb2^T                            (A: hb words in tensor memory)
                gemm3_tw<kWidth / 16>(Zd, Pd, kmajor<TWb2>(wb2), tc::instr_desc_bf16(128, kBaseOut, false, false));
                commit_to(done);
                break;
            case 2:     // z1 = [SH | geo | 1] [W1 | b1]^T
                gemm3<kHeadIn / 16>(Zd, kmajor<TE>(E), kmajor<TW1>(w1),
                                    tc::instr_desc_bf16(128, kWidth, false, false), false);
                commit_to(done);
                break;
            case 3:     // z2 = h1 W2^T
                gemm3<kWidth / 16>(Zd, kmajor<TH>(H), kmajor<TW2>(w2),
                                   tc::instr_desc_bf16(128, kWidth, false, false), false);
                commit_to(done);
                break;
            case 4:     // dh1 = d2 W2;  dW2 | db2 += d2^T [h1 | 1]
                gemm3<kWidth / 16>(Zd, kmajor<TD>(D), mnmajor<TW2>(w2),
                                   tc::instr_desc_bf16(128, kWidth, false, true), false);
                commit_to(done);
                gemm3<kTile / 16>(acc_base + kColDW2, mnmajor<TD>(D), mnmajor<TH>(H),
                                  tc::instr_desc_bf16(64, 72, true, true), accf);
                commit_to(dw_done);
                break;
            case 5:     // din1 = d1 W1;  dW1 | db1 += d1^T [SH | geo | 1]
                gemm3<kWidth / 16>(Zd, kmajor<TD>(D), mnmajor<TW1>(w1),
                                   tc::instr_desc_bf16(128, kHeadIn, false, true), false);
                commit_to(done);
                gemm3<kTile / 16>(acc_base + kColDW1, mnmajor<TD>(D), mnmajor<TE>(E),
                                  tc::instr_desc_bf16(64, kHeadIn, true, true), accf);
                commit_to(dw_done);
                break;
            case 6:     // dhb = dy Wb2 (A: dy words in tensor memory);  dWb2^T += hb^T dy
                gemm3_tw<kBaseOut / 16>(Zd, DYAd, mnmajor<TWb2>(wb2), tc::instr_desc_bf16(128, kWidth, false, true));
                commit_to(done);
                gemm3<kTile / 16>(acc_base + kColDWb2T, mnmajor<TH>(H), mnmajor<TE>(E),
                                  tc::instr_desc_bf16(64, kBaseOut, true, true), accf);
                commit_to(dw_done);
                break;
            default:    // denc = db1 Wb1 (A: db1 words in tensor memory);  dWb1 | dbb1 += db1^T [enc | 1]
                gemm3_tw<kWidth / 16>(Zd, Pd, mnmajor<TWb1>(wb1), tc::instr_desc_bf16(128, kEncDim, false, true));
                commit_to(done);
                gemm3<kTile / 16>(acc_base + kColDWb1, mnmajor<TD>(D), mnmajor<TE>(E),
                                  tc::instr_desc_bf16(64, 40, true, true), accf);
                commit_to(dw_done);
                break;
            }
        }
        __syncwarp();
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
    // ... -> split -> operand words in tensor memory (K step hf of round 0's A operand).  The words stay
    // there until the last round needs enc as a SHARED-MEMORY tile (dWb1 += db1^T enc): restore_enc.
    auto stage_enc = [&](const float (&x)[16]) {
        uint32_t w[16];
        split16_words(x, w);
        tmem_st16(EP, w);
        tmem_wait_st();
    };
    auto restore_enc = [&]() {
        uint32_t w[16];
        tmem_ld16_words(EP, w);
        store_kstep_words<TE>(E, row, 2 * hf, w);
    };
    // hb = act(z_b1 + bb1) for this thread's 32 columns -> operand words in tensor memory (K steps 2 hf,
    // 2 hf + 1 of round 1's A operand); needed again as act'(hb) and, in shared memory, for dWb2^T += hb^T dy
    auto stage_hb = [&]() {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            float h[16];
            uint32_t w[16];
            tmem_ld_cols<16>(Z + 32 * hf + 16 * c, h);
            bias_hidden_act<16>(hact, h, s_bb1 + 32 * hf + 16 * c);
            split16_words(h, w);
            tmem_st16(P + 32 * hf + 16 * c, w);
        }
        tmem_wait_st();
    };
    auto restore_hb = [&]() {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t w[16];
            tmem_ld16_words(P + 32 * hf + 16 * c, w);
            store_kstep_words<TH>(H, row, 4 * hf + 2 * c, w);
        }
    };

    // software pipeline over this slot's tiles: the encoding row and the ray index of the slot's next tile
    // are loaded while the current one drains, so a tile never starts on an HBM miss
    float xe[16];
    int64_t i = ((int64_t)blockIdx.x + (int64_t)slot_id * gridDim.x) * kTile + row;
    bool valid = slot_id < my_tiles && i < n;
    float tmid2 = 0.f;                       // t_start + t_end
    // the ray of a sample is a dependent chain (sample -> ray index -> origin / direction): the index of the
    // slot's NEXT tile is fetched in the middle of the current one, origin / direction before its last
    // wait, so that no tile starts on that chain (ncu: 7 % of the warp samples sat there)
    float dn[3] = {0.f, 0.f, 1.f}, on[3] = {0.f, 0.f, 0.f};
    auto load_ray = [&](int64_t r) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            dn[d] = __ldg(rays_d + 3 * r + d);
            if (hf == 0) on[d] = __ldg(rays_o + 3 * r + d);
        }
    };
    load_enc(i, valid, xe);
    if (valid) {
        tmid2 = t_starts[i] + t_ends[i];
        load_ray(ray_indices[i]);
    }

    for (int64_t c0 = 0; c0 < my_tiles; c0 += kFlushTiles) {
        const int64_t c1 = min(c0 + (int64_t)kFlushTiles, my_tiles);
        acc_live = false;
        for (int64_t k = c0 + slot_id; k < c1; k += kSlots) {
            // ---- operands of round 0: enc -------------------------------------------------------
            const float dir[3] = {dn[0], dn[1], dn[2]};
            bool inside = false;
            if (valid && hf == 0) {
                float pos[3], u[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) pos[d] = on[d] + (dir[d] * tmid2) * 0.5f;
                inside = contract_position(f, pos, u);
            }
            stage_enc(xe);
            launch(0);

            // ---- round 0 done: hb ----------------------------------------------------------------
            await_mma();
            stage_hb();
            launch(1);

            // ---- round 1 done: y -> raw density, [SH | geo | 1] -> E -------------------------------
            // the last weight-gradient GEMM of this slot's previous tile (dWb1: reads D and E) is awaited
            // here, before E is written again, not at the end of that tile
            float raw = 0.f;
            await_mma();
            if (dw_pending) await_dw();
            dw_pending = true;
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
                    x[15] = 1.f;                          // the ones column of the head input (multiplies b1)
                    store16<TE>(E, row, 2, x);
                } else {
                    sh_degree4(dir, x);
                    store16<TE>(E, row, 0, x);
                }
            }
            launch(2);

            // ---- round 2 done: h1 -> H (b1 came through the GEMM) -------------------------------------
            await_mma();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float h[16];
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, h);
                hidden_act_vec<16>(hact, h);
                store16<TH>(H, row, 4 * hf + 2 * c, h);
            }
            launch(3);

            // next tile of this slot: its ray index and interval now (consumed before the tile's last wait),
            // its other rows pulled towards L2 while this one is in flight
            const int64_t ni = i + (int64_t)kSlots * gridDim.x * kTile;
            const bool valid_n = k + kSlots < my_tiles && ni < n;
            int32_t ray_n = 0;
            float tmid2_n = 0.f;
            if (valid_n) {
                ray_n = __ldg(ray_indices + ni);
                tmid2_n = __ldg(t_starts + ni) + __ldg(t_ends + ni);
                if (!enc_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(enc + ni * enc_dim + 16 * hf));
                if (hf == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(d_rgbs + ni * C));
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
            await_mma();
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
                group_sync();
                const float4 other = *reinterpret_cast<const float4*>(zx + (row * 2 + (hf ^ 1)) * 4);
                float d3[3];
                d3[0] = z3[0] + other.x; d3[1] = z3[1] + other.y; d3[2] = z3[2] + other.z;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    d3[c] = c < C ? g_rgb[c] * radiance_act_grad(f.radiance_act, d3[c] + s_b3[c]) : 0.f;
                // d2 = (d3 W3) * act'(h2) -> D   (only the C live rows of W3 are read)
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    float dl[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) dl[j] = d3[0] * s_w3f[32 * hf + 16 * c2 + j];
                    if (C > 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) dl[j] = fmaf(d3[1], s_w3f[kWidth + 32 * hf + 16 * c2 + j], dl[j]);
                    }
                    if (C > 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) dl[j] = fmaf(d3[2], s_w3f[2 * kWidth + 32 * hf + 16 * c2 + j], dl[j]);
                    }
                    {
                        float hh[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) hh[j] = h2[16 * c2 + j];
                        mul_hidden_act_grad<16>(hact, dl, hh);
                    }
                    store16<TD>(D, row, 4 * hf + 2 * c2, dl);
                }
                launch(4);
                // dW3 += d3^T h2, db3 += sum d3  (off the critical path: the MMA round is running)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (c >= C) break;
                    float t[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] = d3[c] * h2[j];
                    r_dw3[c] += warp_transpose_sum<32>(t, lane);
                    if (hf == 0) r_db3[c] += warp_sum(d3[c]);
                }
            }

            // ---- round 4 done: dh1 -> d1 = dh1 * act'(h1) -> D ------------------------------------
            await_mma();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float dl[16], h[16];
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, dl);
                load16<TH>(H, row, 4 * hf + 2 * c, h);
                mul_hidden_act_grad<16>(hact, dl, h);
                if (c == 0) await_dw();                   // dW2 has read d2 (D) and h1 (H)
                store16<TD>(D, row, 4 * hf + 2 * c, dl);
            }
            launch(5);

            // ---- round 5 done: din1 -> dy (tensor memory + E), hb again (H) ---------------------------
            await_mma();
            if (hf == 0) {
                float dgeo[16], dy[16];
                uint32_t w[16];
                tmem_ld_cols<16>(Z + kShDim, dgeo);
                dy[0] = inside ? g_sigma * density_act_grad(f.density_act, raw) : 0.f;
#pragma unroll
                for (int j = 0; j < kGeo; ++j) dy[1 + j] = dgeo[j];
                split16_words(dy, w);
                tmem_st16(DYA, w);
                await_dw();                               // dW1 has read d1 (D) and [SH | geo | 1] (E)
                store_kstep_words<TE>(E, row, 0, w);
                restore_hb();
                tmem_wait_st();
                launch(6);
                const float s = warp_transpose_sum<16>(dy, lane & 15);
                // lanes l and l + 16 hold the two half-warp sums of column l
                r_dbb2 += s + __shfl_xor_sync(0xffffffffu, s, 16);
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
                launch(6);
            }

            // ---- round 6 done: dhb -> db1 = dhb * act'(hb) -> D and, as round 7's A operand, over the hb
            // words it was derived from; enc again -> E -----------------------------------------------------
            await_mma();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float dl[16], h[16];
                uint32_t w[16];
                tmem_ld16_nowait(P + 32 * hf + 16 * c, w);
                tmem_ld_cols<16>(Z + 32 * hf + 16 * c, dl);       // (its wait covers both loads ...
                tmem_wait_ld_dep(w);                              //  ... this one ties the words to it)
                kstep_words_to_float(w, h);
                mul_hidden_act_grad<16>(hact, dl, h);
                split16_words(dl, w);
                tmem_st16(P + 32 * hf + 16 * c, w);
                store_kstep_words<TD>(D, row, 4 * hf + 2 * c, w);
            }
            await_dw();                                   // dWb2^T has read hb (H) and dy (E)
            restore_enc();
            tmem_wait_st();
            launch(7);
            acc_live = true;

            // ---- round 7 done: denc -> HBM ---------------------------------------------------------------
            const int64_t i_cur = i;
            const bool valid_cur = valid;
            i = ni;                                       // this slot's next tile: loads in flight across the wait
            valid = valid_n;
            load_enc(i, valid, xe);
            tmid2 = tmid2_n;
            dn[0] = 0.f; dn[1] = 0.f; dn[2] = 1.f;
            if (valid) load_ray(ray_n);
            await_mma();
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
        // end of the flush period: every GEMM of the slot has been awaited; the quadrant warps 0..3 drain
        // both slots' accumulators between the two barriers
        if (dw_pending) await_dw();                       // dWb1 of the slot's last tile
        dw_pending = false;
        tc::tc_fence_before_sync();
        __syncthreads();
        if (warp < 4) flush_tmem(c0 < c1, c0 + 1 < c1);
        __syncthreads();
        tc::tc_fence_after_sync();
    }

    // ---- the SIMT-side accumulators (dW3, dbb2, db3) go out once: registers -> shared -> global ------
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c >= C) break;
        atomicAdd(&s_dw3[c * kWidth + 32 * hf + lane], r_dw3[c]);
        if (hf == 0 && lane == 0) atomicAdd(&s_db3[c], r_db3[c]);
    }
    if (hf == 0 && lane < 16) atomicAdd(&s_dbb2[lane], r_dbb2);
    tc::tc_fence_before_sync();
    __syncthreads();
    if (my_tiles > 0 && warp < 4) {
        for (int i2 = tid; i2 < C * kWidth; i2 += 128) atomicAdd(g.w3 + i2, s_dw3[i2]);
        if (tid < kBaseOut) atomicAdd(g.bb2 + tid, s_dbb2[tid]);
        if (tid < C) atomicAdd(g.b3 + tid, s_db3[tid]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kTmemCols);
}

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
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int grid = grid_for((n_tiles + 1) / 2, 1, 1);       // one persistent CTA per SM, two tiles in flight each
    cudaFuncSetAttribute(mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd::Smem::total);
    mlp_bwd_tc_kernel<<<grid, bwd::kThreads, bwd::Smem::total, as_stream(stream)>>>(
        *f, *p, *g, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, d_sigmas, d_rgbs, n, n_dev, enc_rows, d_enc,
        d_dirs);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}
