// Shared machinery of the tensor-core MLP kernels (forward and backward).
//
// Thread organisation of a CTA that owns a tile of 128 samples:
//   * 4*S "epilogue" warps.  Warp w serves TMEM lane quadrant q = w % 4 (rows 32q .. 32q+31 of
//     the tile — a hardware rule: a warp can only read the TMEM lanes of its quadrant) and
//     column group cg = w / 4: thread (w, lane) owns row 32q + lane and the 64/S accumulator
//     columns [cg*64/S, (cg+1)*64/S) of every 64-wide layer.  S = 2 (forward) or 4 (backward)
//     keeps 8-16 warps resident per CTA so that the elementwise epilogue (bias, activation,
//     bf16 hi/lo split, operand-tile store) hides its own instruction latencies — v1 had one
//     thread per row and was issue-latency bound (profiles/r01_ncu_mlp_tc_v1_summary.md).
//   * 1 MMA warp whose lane 0 issues every tcgen05.mma of a round from pre-folded descriptors
//     and commits them to the mbarrier the epilogue warps wait on.
// One bar.sync per round hands the operand tiles from the epilogue warps to the MMA warp.
#pragma once
#include "den_common.cuh"
#include "den_field.cuh"
#include "den_tc.cuh"

namespace den {
namespace mlp {

constexpr int kTile = 128;
constexpr int kOutN = 16;      // N of the radiance / dy GEMMs (C, 16 padded to the UMMA minimum)

// operand tile: bf16 hi half followed by bf16 lo half, canonical K-major no-swizzle layout
template <int ROWS, int COLS>
struct Tile {
    uint8_t* hi;
    uint8_t* lo;
    static constexpr uint32_t row_group = (COLS / 8) * 128;   // bytes between groups of 8 rows
    static constexpr int bytes = 2 * ROWS * COLS * 2;
    __device__ explicit Tile(uint8_t* base) : hi(base), lo(base + ROWS * COLS * 2) {}
};

// pre-folded shared-memory matrix descriptor: at(off) costs one 32-bit add
struct Desc {
    uint32_t lo, hi;
    __device__ __forceinline__ uint64_t at(uint32_t byte_off) const {
        return ((uint64_t)hi << 32) | (uint64_t)(lo + (byte_off >> 4));
    }
};
__device__ __forceinline__ Desc make_desc(const uint8_t* p, uint32_t lbo, uint32_t sbo) {
    Desc d;
    d.lo = ((tc::smem_u32(p) & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
    d.hi = (sbo >> 4) | (1u << 14);
    return d;
}

// ---- GEMM issue (one lane of the MMA warp) ------------------------------------------------
// D[128 x N] = A[128 x K] * W^T, W tile (N, K) K-major                          — forward
template <int N, int K>
__device__ __forceinline__ void mma_fwd(uint32_t tmem_d, const Tile<kTile, K>& a, const Tile<N, K>& w) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(128, N, false, false);
    const Desc ah = make_desc(a.hi, 128, Tile<kTile, K>::row_group);
    const Desc al = make_desc(a.lo, 128, Tile<kTile, K>::row_group);
    const Desc wh = make_desc(w.hi, 128, Tile<N, K>::row_group);
    const Desc wl = make_desc(w.lo, 128, Tile<N, K>::row_group);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16(tmem_d, ah.at(ks * 256), wh.at(ks * 256), idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16(tmem_d, al.at(ks * 256), wh.at(ks * 256), idesc, true);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16(tmem_d, ah.at(ks * 256), wl.at(ks * 256), idesc, true);
}

// the same GEMM issued by the single elected thread (tc::elect_one)
template <int N, int K>
__device__ __forceinline__ void mma_fwd_1t(uint32_t tmem_d, const Tile<kTile, K>& a, const Tile<N, K>& w) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(128, N, false, false);
    const Desc ah = make_desc(a.hi, 128, Tile<kTile, K>::row_group);
    const Desc al = make_desc(a.lo, 128, Tile<kTile, K>::row_group);
    const Desc wh = make_desc(w.hi, 128, Tile<N, K>::row_group);
    const Desc wl = make_desc(w.lo, 128, Tile<N, K>::row_group);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16_1t(tmem_d, ah.at(ks * 256), wh.at(ks * 256), idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16_1t(tmem_d, al.at(ks * 256), wh.at(ks * 256), idesc, 1u);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16_1t(tmem_d, ah.at(ks * 256), wl.at(ks * 256), idesc, 1u);
}

// D[128 x N] = dY[128 x K] * W, W tile (K rows = out, N feats = in) read MN-major   — dX
template <int N, int K>
__device__ __forceinline__ void mma_dx(uint32_t tmem_d, const Tile<kTile, K>& dy, const Tile<K, N>& w) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(128, N, false, true);
    constexpr uint32_t wrg = Tile<K, N>::row_group;
    const Desc ah = make_desc(dy.hi, 128, Tile<kTile, K>::row_group);
    const Desc al = make_desc(dy.lo, 128, Tile<kTile, K>::row_group);
    const Desc wh = make_desc(w.hi, wrg, 128);
    const Desc wl = make_desc(w.lo, wrg, 128);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16(tmem_d, ah.at(ks * 256), wh.at(ks * 2 * wrg), idesc, ks > 0);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16(tmem_d, al.at(ks * 256), wh.at(ks * 2 * wrg), idesc, true);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) tc::mma_bf16(tmem_d, ah.at(ks * 256), wl.at(ks * 2 * wrg), idesc, true);
}

// D[64 x N] (+)= A^T * B over the tile's 128 rows; A (128, 64), B (128, N) read MN-major — dW
template <int N>
__device__ __forceinline__ void mma_dw(uint32_t tmem_d, const Tile<kTile, 64>& a, const Tile<kTile, N>& b,
                                       bool accumulate) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(64, N, true, true);
    constexpr uint32_t arg = Tile<kTile, 64>::row_group;
    constexpr uint32_t brg = Tile<kTile, N>::row_group;
    const Desc ah = make_desc(a.hi, arg, 128), al = make_desc(a.lo, arg, 128);
    const Desc bh = make_desc(b.hi, brg, 128), bl = make_desc(b.lo, brg, 128);
#pragma unroll
    for (int ks = 0; ks < kTile / 16; ++ks)
        tc::mma_bf16(tmem_d, ah.at(ks * 2 * arg), bh.at(ks * 2 * brg), idesc, accumulate || ks > 0);
#pragma unroll
    for (int ks = 0; ks < kTile / 16; ++ks) tc::mma_bf16(tmem_d, al.at(ks * 2 * arg), bh.at(ks * 2 * brg), idesc, true);
#pragma unroll
    for (int ks = 0; ks < kTile / 16; ++ks) tc::mma_bf16(tmem_d, ah.at(ks * 2 * arg), bl.at(ks * 2 * brg), idesc, true);
}

// D[64 x 8] (+)= A^T * ones : column sums of a (128, 64) tile (bias gradient)
__device__ __forceinline__ void mma_colsum(uint32_t tmem_d, const Tile<kTile, 64>& a, const uint8_t* ones,
                                           bool accumulate) {
    constexpr uint32_t idesc = tc::instr_desc_bf16(64, 8, true, false);
    constexpr uint32_t arg = Tile<kTile, 64>::row_group;
    const Desc ah = make_desc(a.hi, arg, 128), al = make_desc(a.lo, arg, 128);
    const Desc ob = make_desc(ones, 128, (kTile / 8) * 128);
#pragma unroll
    for (int ks = 0; ks < kTile / 16; ++ks)
        tc::mma_bf16(tmem_d, ah.at(ks * 2 * arg), ob.at(ks * 256), idesc, accumulate || ks > 0);
#pragma unroll
    for (int ks = 0; ks < kTile / 16; ++ks) tc::mma_bf16(tmem_d, al.at(ks * 2 * arg), ob.at(ks * 256), idesc, true);
}

// ---- epilogue helpers -------------------------------------------------------------------------
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hp = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        h[i] = *reinterpret_cast<const uint32_t*>(&hp);
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        const __nv_bfloat162 lp = __floats2bfloat162_rn(x[2 * i] - h0, x[2 * i + 1] - h1);
        l[i] = *reinterpret_cast<const uint32_t*>(&lp);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// store columns [col0, col0 + COLS) of row `r` into a (128, K) tile
template <int COLS, int K>
__device__ __forceinline__ void store_cols(const Tile<kTile, K>& t, int r, int col0, const float (&v)[COLS]) {
#pragma unroll
    for (int c = 0; c < COLS / 8; ++c) {
        uint4 hi, lo;
        split8(&v[8 * c], hi, lo);
        const uint32_t off = tc::chunk_offset(r, (col0 >> 3) + c, K);
        *reinterpret_cast<uint4*>(t.hi + off) = hi;
        *reinterpret_cast<uint4*>(t.lo + off) = lo;
    }
}

template <int COLS, int K>
__device__ __forceinline__ void load_cols(const Tile<kTile, K>& t, int r, int col0, float (&v)[COLS]) {
#pragma unroll
    for (int c = 0; c < COLS / 8; ++c) {
        const uint32_t off = tc::chunk_offset(r, (col0 >> 3) + c, K);
        const uint4 h = *reinterpret_cast<const uint4*>(t.hi + off);
        const uint4 l = *reinterpret_cast<const uint4*>(t.lo + off);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[8 * c + 2 * q] = __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
            v[8 * c + 2 * q + 1] = __uint_as_float(hw[q] & 0xffff0000u) + __uint_as_float(lw[q] & 0xffff0000u);
        }
    }
}

// TMEM -> registers, COLS in {8, 16, 32}: all loads issued, then one wait
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float (&v)[COLS]) {
    static_assert(COLS == 8 || COLS == 16 || COLS == 32, "COLS must be 8, 16 or 32");
    uint32_t r[COLS];
    if constexpr (COLS == 8) {
        tmem_ld8_nowait(taddr, r);
    } else {
#pragma unroll
        for (int c = 0; c < COLS / 16; ++c) tmem_ld16_nowait(taddr + 16 * c, r + 16 * c);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < COLS; ++i) v[i] = __uint_as_float(r[i]);
}

// hand the operand tiles written by this thread to the MMA warp
__device__ __forceinline__ void publish() {
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
}
// the same hand-off through named barrier 1 without blocking the producers: the epilogue threads
// arrive and go on to their mbarrier wait, the MMA warp syncs (n_threads = epilogue + 32)
__device__ __forceinline__ void publish_arrive(int n_threads) {
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    asm volatile("bar.arrive 1, %0;" ::"r"(n_threads) : "memory");
}
__device__ __forceinline__ void handoff_sync(int n_threads) {
    asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
    tc::tc_fence_after_sync();
}
__device__ __forceinline__ void await(uint64_t* bar, uint32_t& phase) {
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::tc_fence_after_sync();
}

// weight tiles shared by both kernels
struct Weights {
    static constexpr int wb1 = 0;                                  // (64, 32)
    static constexpr int wb2 = wb1 + Tile<kWidth, kEncDim>::bytes; // (16, 64)
    static constexpr int w1 = wb2 + Tile<kBaseOut, kWidth>::bytes; // (64, 32)
    static constexpr int w2 = w1 + Tile<kWidth, kHeadIn>::bytes;   // (64, 64)
    static constexpr int w3 = w2 + Tile<kWidth, kWidth>::bytes;    // (16, 64)
    static constexpr int bias = w3 + Tile<kOutN, kWidth>::bytes;   // fp32: bb1 64 | bb2 16 | b1 64 | b2 64 | b3 16
    static constexpr int end = bias + (3 * kWidth + kBaseOut + kOutN) * 4;
};

__device__ __forceinline__ void load_all_weights(uint8_t* smem, const den_field_desc& f,
                                                 const den_field_params& p, bool full) {
    const int enc_dim = f.grid.n_levels * 2;
    float* b = reinterpret_cast<float*>(smem + Weights::bias);
    const Tile<kWidth, kEncDim> Wb1(smem + Weights::wb1);
    const Tile<kBaseOut, kWidth> Wb2(smem + Weights::wb2);
    tc::load_weight_split(Wb1.hi, Wb1.lo, p.wb1, kWidth, enc_dim, kWidth, kEncDim);
    tc::load_weight_split(Wb2.hi, Wb2.lo, p.wb2, kBaseOut, kWidth, kBaseOut, kWidth);
    load_padded(b, p.bb1, kWidth, kWidth);
    load_padded(b + kWidth, p.bb2, kBaseOut, kBaseOut);
    if (full) {
        const Tile<kWidth, kHeadIn> W1(smem + Weights::w1);
        const Tile<kWidth, kWidth> W2(smem + Weights::w2);
        const Tile<kOutN, kWidth> W3(smem + Weights::w3);
        tc::load_weight_split(W1.hi, W1.lo, p.w1, kWidth, kShDim + kGeo, kWidth, kHeadIn);
        tc::load_weight_split(W2.hi, W2.lo, p.w2, kWidth, kWidth, kWidth, kWidth);
        tc::load_weight_split(W3.hi, W3.lo, p.w3, f.channels, kWidth, kOutN, kWidth);
        load_padded(b + kWidth + kBaseOut, p.b1, kWidth, kWidth);
        load_padded(b + 2 * kWidth + kBaseOut, p.b2, kWidth, kWidth);
        load_padded(b + 3 * kWidth + kBaseOut, p.b3, f.channels, kOutN);
    }
}

}  // namespace mlp
}  // namespace den
