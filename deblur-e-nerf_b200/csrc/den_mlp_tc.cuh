// Shared machinery of the tensor-core MLP kernels (forward and backward): the canonical UMMA
// "no swizzle" descriptor, the bf16 hi/lo split of eight values and the TMEM -> register loads.
// The multi-slot pipelines themselves (operand tiles, GEMM issue, hand-offs) live in den_mlp_ops.cuh
// and the two kernel files.
#pragma once
#include "den_common.cuh"
#include "den_field.cuh"
#include "den_tc.cuh"

namespace den {
namespace mlp {

constexpr int kTile = 128;
constexpr int kOutN = 16;      // N of the radiance / dy GEMMs (C, 16 padded to the UMMA minimum)

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

}  // namespace mlp
}  // namespace den
