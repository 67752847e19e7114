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
    return li.mask ? (idx & li.mask) : (idx % li.size);
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

}  // namespace den
