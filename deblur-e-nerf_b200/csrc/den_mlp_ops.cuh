// Operand tiles, descriptors and single-thread GEMM issue shared by the multi-slot tensor-core MLP
// kernels (den_mlp_tc.cu forward: three tiles in flight, den_mlp_tc_bwd.cu backward: two).  Every GEMM is
// issued by one elected thread of the slot that owns the tile; there is no MMA warp.
#pragma once
#include "den_mlp_tc.cuh"

namespace den {
namespace mlp {

// operand tile with CH 16-byte chunks (8 bf16 columns each) per row: hi half then lo half
template <int ROWS, int CH>
struct OpTile {
    static constexpr uint32_t rg = CH * 128;               // bytes between groups of 8 rows
    static constexpr uint32_t half = ROWS * CH * 16;       // bytes of the hi (or lo) half
    static constexpr uint32_t bytes = 2 * half;
    static __device__ __forceinline__ uint32_t off(int r, int chunk) {
        return (uint32_t)((r >> 3) * (int)rg + chunk * 128 + (r & 7) * 16);
    }
};
// ---- descriptors / GEMM issue (all lanes of the MMA warp, one elected lane issues) ------------
struct OpDesc {                     // hi and lo descriptors of one operand view
    Desc hi, lo;
    uint32_t kstep;                 // byte advance per K step of 16 elements
};
template <class T>
__device__ __forceinline__ OpDesc kmajor(const uint8_t* base) {         // K along the tile's columns
    return {make_desc(base, 128, T::rg), make_desc(base + T::half, 128, T::rg), 256u};
}
template <class T>
__device__ __forceinline__ OpDesc mnmajor(const uint8_t* base) {        // K along the tile's rows
    return {make_desc(base, T::rg, 128), make_desc(base + T::half, T::rg, 128), 2u * T::rg};
}
// D (+)= A * B with the 3-product bf16 split, KSTEPS steps of 16 along K
__device__ __forceinline__ bool elect_one() {
    uint32_t leader;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(leader));
    return leader != 0;
}
// one tcgen05.mma from the calling (single, elected) thread
__device__ __forceinline__ void mma1(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// D (+)= A * B with the 3-product bf16 split, KSTEPS steps of 16 along K.  Called by the elected
// lane only (one branch per GEMM instead of one elect.sync + vote per MMA).
template <int KSTEPS>
__device__ __forceinline__ void gemm3(uint32_t tmem_d, const OpDesc& a, const OpDesc& b, uint32_t idesc,
                                      bool accumulate) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
        mma1(tmem_d, a.hi.at(ks * a.kstep), b.hi.at(ks * b.kstep), idesc, (accumulate || ks > 0) ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
        mma1(tmem_d, a.lo.at(ks * a.kstep), b.hi.at(ks * b.kstep), idesc, 1u);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
        mma1(tmem_d, a.hi.at(ks * a.kstep), b.lo.at(ks * b.kstep), idesc, 1u);
}

template <class T>
__device__ __forceinline__ void store16(uint8_t* tile, int r, int chunk0, const float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint4 hi, lo;
        split8(&v[8 * c], hi, lo);
        const uint32_t o = T::off(r, chunk0 + c);
        *reinterpret_cast<uint4*>(tile + o) = hi;
        *reinterpret_cast<uint4*>(tile + T::half + o) = lo;
    }
}
// registers <-> TMEM, 16 words per thread (warp-collective; thread i addresses lane base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_words(uint32_t taddr, uint32_t (&r)[16]) {
    tmem_ld16_nowait(taddr, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- activations as the A operand IN TENSOR MEMORY (ts-form tcgen05.mma) ------------------------------
// An M = 128 A operand of K bf16 columns occupies K / 2 TMEM columns: lane = row, column c = the pair
// (a[2c], a[2c+1]), low half first (pinned by tests/test_gpu_mlp_tc.py, probe mode 3).  The epilogue
// thread of a row writes its packed hi / lo words with tcgen05.st; no shared-memory store, no
// generic->async proxy fence, and the MMA fetches only B from shared memory.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// 16 values of this thread's row, K columns [8 chunk0, 8 chunk0 + 16) -> hi / lo operand columns
__device__ __forceinline__ void tstore16(uint32_t t_hi, uint32_t t_lo, int chunk0, const float (&v)[16]) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint4 h, l;
        split8(&v[8 * c], h, l);
        hi[4 * c] = h.x; hi[4 * c + 1] = h.y; hi[4 * c + 2] = h.z; hi[4 * c + 3] = h.w;
        lo[4 * c] = l.x; lo[4 * c + 1] = l.y; lo[4 * c + 2] = l.z; lo[4 * c + 3] = l.w;
    }
    tmem_st8(t_hi + 4 * chunk0, hi);
    tmem_st8(t_lo + 4 * chunk0, lo);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma1_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                        uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// D (+)= A * B, A = (hi, lo) columns in TMEM, B = hi / lo tiles in shared memory, 3-product split
template <int KSTEPS>
__device__ __forceinline__ void gemm3_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, const OpDesc& b,
                                         uint32_t idesc, bool accumulate) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
        mma1_ts(tmem_d, a_hi + 8 * ks, b.hi.at(ks * b.kstep), idesc, (accumulate || ks > 0) ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) mma1_ts(tmem_d, a_lo + 8 * ks, b.hi.at(ks * b.kstep), idesc, 1u);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) mma1_ts(tmem_d, a_hi + 8 * ks, b.lo.at(ks * b.kstep), idesc, 1u);
}

template <class T>
__device__ __forceinline__ void load16(const uint8_t* tile, int r, int chunk0, float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const uint32_t o = T::off(r, chunk0 + c);
        const uint4 h = *reinterpret_cast<const uint4*>(tile + o);
        const uint4 l = *reinterpret_cast<const uint4*>(tile + T::half + o);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[8 * c + 2 * q] = __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
            v[8 * c + 2 * q + 1] = __uint_as_float(hw[q] & 0xffff0000u) + __uint_as_float(lw[q] & 0xffff0000u);
        }
    }
}
// butterfly "reduce-scatter": on return lane l holds, in v[0], the sum over the warp of v[l]
template <int N>
__device__ __forceinline__ float warp_transpose_sum(float (&v)[N], int lane) {
    static_assert(N == 32 || N == 16, "N must be 16 or 32");
#pragma unroll
    for (int off = N / 2; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const float send = up ? v[k] : v[k + off];
            const float keep = up ? v[k + off] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}


__device__ __forceinline__ void named_sync(int barrier_id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(barrier_id), "r"(n_threads) : "memory");
}

}  // namespace mlp
}  // namespace den
