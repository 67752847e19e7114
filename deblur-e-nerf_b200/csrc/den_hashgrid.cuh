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

// All eight corner entries of a cell at once: the dense / hashed decision is taken once per level
// (warp-uniform), the neighbours come from additions (dense: + stride, hashed: + prime before the
// XOR) instead of eight independent index evaluations.  Bit-identical to corner_index(): unsigned
// arithmetic is modular, (c + 1) * k == c * k + k.  Corner c = dx + 2 dy + 4 dz.
__device__ __forceinline__ void corner_indices(const LevelInfo& li, const CellFrac& cf, uint32_t (&idx)[8]) {
    if (li.hashed) {
        const uint32_t x[2] = {cf.c[0], cf.c[0] + 1u};
        const uint32_t y0 = cf.c[1] * kPrime1, z0 = cf.c[2] * kPrime2;
        const uint32_t y[2] = {y0, y0 + kPrime1};
        const uint32_t z[2] = {z0, z0 + kPrime2};
#pragma unroll
        for (int c = 0; c < 8; ++c) idx[c] = x[c & 1] ^ y[(c >> 1) & 1] ^ z[c >> 2];
        if (li.mask) {
#pragma unroll
            for (int c = 0; c < 8; ++c) idx[c] &= li.mask;
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) idx[c] %= li.size;
        }
    } else {
        const uint32_t b = cf.c[0] * li.st0 + cf.c[1] * li.st1 + cf.c[2] * li.st2;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t i = b + ((c & 1) ? li.st0 : 0u) + ((c & 2) ? li.st1 : 0u) + ((c & 4) ? li.st2 : 0u);
            if (li.mask) i &= li.mask;
            else if (i >= li.size) i %= li.size;       // only out-of-range positions wrap (tcnn semantics)
            idx[c] = i;
        }
    }
}
// trilinear weights of the eight corners from three pairs
__device__ __forceinline__ void corner_weights(const CellFrac& cf, float (&w)[8]) {
    const float wx[2] = {1.f - cf.f[0], cf.f[0]};
    const float wy[2] = {1.f - cf.f[1], cf.f[1]};
    const float wz[2] = {1.f - cf.f[2], cf.f[2]};
#pragma unroll
    for (int c = 0; c < 8; ++c) w[c] = wx[c & 1] * wy[(c >> 1) & 1] * wz[c >> 2];
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
