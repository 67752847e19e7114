// Hash-grid level arithmetic shared by the stand-alone encoding kernels and the fused
// field kernels.  Semantics: tcnn grid.h (grid_index, pos_fract, coherent_prime_hash),
// restated in oracle/tcnn_ref.py.
#pragma once
#include "den_common.cuh"

namespace den {

constexpr uint32_t kPrime1 = 2654435761u;
constexpr uint32_t kPrime2 = 805459861u;

struct LevelInfo {
    float scale;
    uint32_t size;      // entries in this level
    uint32_t st0, st1, st2;  // dense strides (0 when the upstream loop has stopped)
    uint32_t mask;      // size - 1 when size is a power of two, else 0
    bool hashed;
};

__device__ __forceinline__ LevelInfo make_level(const den_hashgrid_desc& g, int level) {
    LevelInfo li;
    li.scale = g.scale[level];
    li.size = g.size[level];
    const uint32_t res = g.resolution[level];
    uint32_t stride = 1;
    li.st0 = li.st1 = li.st2 = 0;
    if (stride <= li.size) { li.st0 = stride; stride *= res; }
    if (stride <= li.size) { li.st1 = stride; stride *= res; } else { li.st1 = 0; }
    if (li.st1 != 0 && stride <= li.size) { li.st2 = stride; stride *= res; }
    li.hashed = li.size < stride;
    li.mask = (li.size & (li.size - 1)) == 0 ? li.size - 1 : 0;
    return li;
}

__device__ __forceinline__ uint32_t entry_index(const LevelInfo& li, uint32_t cx, uint32_t cy, uint32_t cz) {
    uint32_t idx = li.hashed ? (cx ^ (cy * kPrime1) ^ (cz * kPrime2))
                             : (cx * li.st0 + cy * li.st1 + cz * li.st2);
    if (li.mask) return idx & li.mask;
    // dense level: the linear index of an in-range cell is already < size; the (slow) modulo only runs
    // for out-of-range positions, where tcnn's wrap-around must be reproduced bit for bit
    return idx < li.size ? idx : idx % li.size;
}

struct CellFrac {
    uint32_t c[3];
    float f[3];
};

__device__ __forceinline__ CellFrac locate(float scale, float x, float y, float z) {
    CellFrac cf;
    const float p[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float pos = fmaf(scale, p[d], 0.5f);
        float fl = floorf(pos);
        cf.c[d] = (uint32_t)(int)fl;
        cf.f[d] = pos - fl;
    }
    return cf;
}


__device__ __forceinline__ float corner_weight(const CellFrac& cf, int c) {
    return ((c & 1) ? cf.f[0] : 1.f - cf.f[0]) * (((c >> 1) & 1) ? cf.f[1] : 1.f - cf.f[1]) *
           (((c >> 2) & 1) ? cf.f[2] : 1.f - cf.f[2]);
}

__device__ __forceinline__ uint32_t corner_index(const LevelInfo& li, const CellFrac& cf, int c) {
    return entry_index(li, cf.c[0] + (c & 1), cf.c[1] + ((c >> 1) & 1), cf.c[2] + ((c >> 2) & 1));
}

__device__ __forceinline__ void red_add_v2(float2* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ void red_add_v4(float2* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
// The two x-neighbours (cx, cx + 1) of a cell edge: when their entries are the aligned pair
// {2k, 2k + 1} (dense levels with an even / odd start, hashed power-of-two levels whenever cx is even:
// (cx ^ h) and ((cx | 1) ^ h) differ in bit 0 only) one 16-byte vector reduction replaces two 8-byte
// ones — the L2 atomic unit is the bound of the scatter, not the bytes.
__device__ __forceinline__ void red_add_pair(float2* base, uint32_t idx0, uint32_t idx1, float a0, float b0,
                                             float a1, float b1) {
    if ((idx0 ^ idx1) == 1u) {
        if (idx0 & 1u) red_add_v4(base + idx1, a1, b1, a0, b0);
        else red_add_v4(base + idx0, a0, b0, a1, b1);
    } else {
        red_add_v2(base + idx0, a0, b0);
        red_add_v2(base + idx1, a1, b1);
    }
}

}  // namespace den
