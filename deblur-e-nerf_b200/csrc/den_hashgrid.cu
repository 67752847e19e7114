// Multiresolution hash-grid encoding (forward gather, backward scatter) for sm_100a.
//
// Replaces tinycudann.Encoding(HashGrid, Linear, fp32) — reference call sites
// external/ngp.py:166-170 (construction) and :240 (evaluation); semantics per tcnn
// include/tiny-cuda-nn/encodings/grid.h (kernel_grid, kernel_grid_backward,
// kernel_grid_backward_input), restated in oracle/tcnn_ref.py.
//
// Layout / mapping
//   * table: per-level tables concatenated, entry = float2 (F = 2), 8-byte gathers.
//   * work item = (tile of 32 consecutive samples, level).  lane <-> sample,
//     warp <-> level: consecutive samples lie along a ray, so on coarse levels a
//     warp's 8 gathers hit a handful of cells (L1/L2 broadcast), and the 48 MiB
//     table stays L2-resident (126 MB L2).
//   * forward: the 32 x (L*2) tile is staged in shared memory and written back as
//     one contiguous, float4-coalesced block (tcnn writes 8-byte pieces at a
//     128-byte stride instead).
//   * backward: dL/denc is staged through shared memory the same way; coarse
//     levels (level < n_agg_levels) merge runs of consecutive samples that share a
//     cell with a segmented warp scan before issuing `red.global.add.v2.f32`
//     (one vector atomic per run and corner instead of one per sample and corner).
#include "den_common.cuh"
#include "den_hashgrid.cuh"

namespace den {

constexpr int kTile = 32;            // samples per tile (one per lane)
constexpr int kHashThreads = 256;    // 8 warps

// ------------------------------------------------------------------ forward --
__global__ void __launch_bounds__(kHashThreads, 6)
hashgrid_fwd_kernel(const __grid_constant__ den_hashgrid_desc g, const float* __restrict__ x,
                    const float2* __restrict__ table, float* __restrict__ out, int64_t n,
                    const int32_t* __restrict__ n_dev) {
    n = effective_n(n, n_dev);
    extern __shared__ float s_tile[];            // kTile x (LF + 1)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int LF = g.n_levels * 2;
    const int ld = LF + 1;
    const int64_t n_tiles = (n + kTile - 1) / kTile;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m = tile * kTile + lane;
        const bool valid = m < n;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (valid) {
            px = __ldg(x + 3 * m + 0);
            py = __ldg(x + 3 * m + 1);
            pz = __ldg(x + 3 * m + 2);
        }
        for (int level = warp; level < g.n_levels; level += nwarp) {
            const LevelInfo li = make_level(g, level);
            const float2* __restrict__ base = table + g.offset[level];
            const CellFrac cf = locate(li.scale, px, py, pz);
            uint32_t idx[8];
            float w[8];
            float2 v[8];
            corner_indices(li, cf, idx);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = valid ? __ldg(base + idx[c]) : make_float2(0.f, 0.f);
            corner_weights(cf, w);
            float ax = 0.f, ay = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                ax = fmaf(w[c], v[c].x, ax);
                ay = fmaf(w[c], v[c].y, ay);
            }
            s_tile[lane * ld + 2 * level + 0] = ax;
            s_tile[lane * ld + 2 * level + 1] = ay;
        }
        __syncthreads();
        // coalesced write-back of the tile: kTile*LF contiguous floats
        const int64_t tile_base = tile * kTile * (int64_t)LF;
        const int64_t tile_elems = min((int64_t)kTile, n - tile * kTile) * LF;
        for (int i = threadIdx.x * 4; i < tile_elems; i += blockDim.x * 4) {
            const int row = i / LF, col = i - row * LF;   // LF % 4 == 0 -> same row
            float4 o;
            o.x = s_tile[row * ld + col + 0];
            o.y = s_tile[row * ld + col + 1];
            o.z = s_tile[row * ld + col + 2];
            o.w = s_tile[row * ld + col + 3];
            *reinterpret_cast<float4*>(out + tile_base + i) = o;
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------- backward --

template <bool kInputGrad>
__global__ void __launch_bounds__(kHashThreads)
hashgrid_bwd_kernel(const __grid_constant__ den_hashgrid_desc g, const float* __restrict__ x,
                    const float* __restrict__ dout, const float2* __restrict__ table,
                    float2* __restrict__ dtable, float* __restrict__ dx, int64_t n,
                    const int32_t* __restrict__ n_dev) {
    n = effective_n(n, n_dev);
    extern __shared__ float s_tile[];            // kTile x (LF + 1)  (+ kTile x 3 for dx)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int LF = g.n_levels * 2;
    const int ld = LF + 1;
    float* s_dx = s_tile + kTile * ld;           // kTile x 3, only if kInputGrad
    const int64_t n_tiles = (n + kTile - 1) / kTile;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // coalesced load of the dL/denc tile
        const int64_t tile_base = tile * kTile * (int64_t)LF;
        const int64_t tile_elems = min((int64_t)kTile, n - tile * kTile) * LF;
        for (int i = threadIdx.x * 4; i < tile_elems; i += blockDim.x * 4) {
            const int row = i / LF, col = i - row * LF;
            float4 v = __ldg(reinterpret_cast<const float4*>(dout + tile_base + i));
            s_tile[row * ld + col + 0] = v.x;
            s_tile[row * ld + col + 1] = v.y;
            s_tile[row * ld + col + 2] = v.z;
            s_tile[row * ld + col + 3] = v.w;
        }
        if (kInputGrad && threadIdx.x < kTile * 3) s_dx[threadIdx.x] = 0.f;
        __syncthreads();

        const int64_t m = tile * kTile + lane;
        const bool valid = m < n;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (valid) {
            px = __ldg(x + 3 * m + 0);
            py = __ldg(x + 3 * m + 1);
            pz = __ldg(x + 3 * m + 2);
        }
        // coarse (run-merging, expensive) and fine (cheap) levels are paired on the same warp:
        // warp w takes levels w, L-1-w, w + 2*nwarp, L-1-w - 2*nwarp, ...
        for (int it = warp; it < g.n_levels; it += nwarp) {
            const int pair = it / nwarp;                  // 0, 1, 2, ...
            const int level = (pair & 1) ? g.n_levels - 1 - (it - pair * nwarp) - (pair >> 1) * nwarp
                                         : (it - pair * nwarp) + (pair >> 1) * nwarp;
            const LevelInfo li = make_level(g, level);
            float2* __restrict__ gbase = dtable + g.offset[level];
            const CellFrac cf = locate(li.scale, px, py, pz);
            const float gx = valid ? s_tile[lane * ld + 2 * level + 0] : 0.f;
            const float gy = valid ? s_tile[lane * ld + 2 * level + 1] : 0.f;
            uint32_t idx[8];
            corner_indices(li, cf, idx);

            if (level < g.n_agg_levels) {
                // runs of consecutive lanes in the same cell -> one atomic per run & corner
                const uint32_t key = valid ? (cf.c[0] * 73856093u) ^ (cf.c[1] * 19349663u) ^ (cf.c[2] * 83492791u)
                                           : 0xffffffffu - lane;
                const uint32_t kprev = __shfl_up_sync(0xffffffffu, key, 1);
                const uint32_t c0p = __shfl_up_sync(0xffffffffu, cf.c[0], 1);
                const uint32_t c1p = __shfl_up_sync(0xffffffffu, cf.c[1], 1);
                const uint32_t c2p = __shfl_up_sync(0xffffffffu, cf.c[2], 1);
                const bool head = lane == 0 || !valid || key != kprev || cf.c[0] != c0p ||
                                  cf.c[1] != c1p || cf.c[2] != c2p;
                const uint32_t heads = __ballot_sync(0xffffffffu, head);
                const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
                const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
                // scan steps beyond the longest run of this warp cannot contribute (warp-uniform bound)
                const int max_run = __reduce_max_sync(0xffffffffu, lane - start + 1);
#pragma unroll
                for (int c = 0; c < 8; c += 2) {           // x-neighbour pairs (c, c + 1)
                    const float wyz = (((c >> 1) & 1) ? cf.f[1] : 1.f - cf.f[1]) *
                                      (((c >> 2) & 1) ? cf.f[2] : 1.f - cf.f[2]);
                    const float w0 = (1.f - cf.f[0]) * wyz, w1 = cf.f[0] * wyz;
                    float a0 = w0 * gx, b0 = w0 * gy, a1 = w1 * gx, b1 = w1 * gy;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        if (d >= max_run) break;
                        const float t0 = __shfl_up_sync(0xffffffffu, a0, d);
                        const float t1 = __shfl_up_sync(0xffffffffu, b0, d);
                        const float t2 = __shfl_up_sync(0xffffffffu, a1, d);
                        const float t3 = __shfl_up_sync(0xffffffffu, b1, d);
                        if (lane - d >= start) { a0 += t0; b0 += t1; a1 += t2; b1 += t3; }
                    }
                    if (tail && valid) red_add_pair(gbase, idx[c], idx[c + 1], a0, b0, a1, b1);
                }
            } else if (valid) {
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    const float wyz = (((c >> 1) & 1) ? cf.f[1] : 1.f - cf.f[1]) *
                                      (((c >> 2) & 1) ? cf.f[2] : 1.f - cf.f[2]);
                    const float w0 = (1.f - cf.f[0]) * wyz, w1 = cf.f[0] * wyz;
                    red_add_pair(gbase, idx[c], idx[c + 1], w0 * gx, w0 * gy, w1 * gx, w1 * gy);
                }
            }

            if (kInputGrad && valid) {
                // dL/dx_d = scale * sum_corners (+-1 on axis d) * prod_{other} w * <table, g>
                const float2* __restrict__ base = table + g.offset[level];
                float dpos[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float2 t = __ldg(base + idx[c]);
                    float dotv = t.x * gx + t.y * gy;
                    float w0 = (c & 1) ? cf.f[0] : 1.f - cf.f[0];
                    float w1 = ((c >> 1) & 1) ? cf.f[1] : 1.f - cf.f[1];
                    float w2 = ((c >> 2) & 1) ? cf.f[2] : 1.f - cf.f[2];
                    dpos[0] += ((c & 1) ? 1.f : -1.f) * w1 * w2 * dotv;
                    dpos[1] += (((c >> 1) & 1) ? 1.f : -1.f) * w0 * w2 * dotv;
                    dpos[2] += (((c >> 2) & 1) ? 1.f : -1.f) * w0 * w1 * dotv;
                }
                atomicAdd(&s_dx[lane * 3 + 0], li.scale * dpos[0]);
                atomicAdd(&s_dx[lane * 3 + 1], li.scale * dpos[1]);
                atomicAdd(&s_dx[lane * 3 + 2], li.scale * dpos[2]);
            }
        }
        __syncthreads();
        if (kInputGrad) {
            const int64_t rows = min((int64_t)kTile, n - tile * kTile);
            if (threadIdx.x < rows * 3) dx[tile * kTile * 3 + threadIdx.x] = s_dx[threadIdx.x];
            __syncthreads();
        }
    }
}

static int check_desc(const den_hashgrid_desc* d) {
    if (!d) { set_error("hashgrid: null descriptor"); return DEN_ERR_INVALID_ARGUMENT; }
    if (d->n_features != 2) { set_error("hashgrid: n_features must be 2, got %d", d->n_features); return DEN_ERR_UNSUPPORTED; }
    if (d->n_levels < 1 || d->n_levels > DEN_MAX_LEVELS) { set_error("hashgrid: n_levels %d out of range", d->n_levels); return DEN_ERR_INVALID_ARGUMENT; }
    if ((d->n_levels * 2) % 4 != 0) { set_error("hashgrid: n_levels must be even"); return DEN_ERR_UNSUPPORTED; }
    for (int l = 0; l < d->n_levels; ++l)
        if (d->size[l] == 0) { set_error("hashgrid: level %d has zero entries", l); return DEN_ERR_INVALID_ARGUMENT; }
    return DEN_OK;
}

}  // namespace den

extern "C" {

int den_hashgrid_fwd(const den_hashgrid_desc* desc, const float* x, const float* table, float* out,
                     int64_t n, const int32_t* n_dev, void* stream) {
    using namespace den;
    int rc = check_desc(desc);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(x && table && out, "null pointer");
    const int LF = desc->n_levels * 2;
    const size_t smem = (size_t)kTile * (LF + 1) * sizeof(float);
    const int grid = grid_for((n + kTile - 1) / kTile, 1, 16);
    hashgrid_fwd_kernel<<<grid, kHashThreads, smem, as_stream(stream)>>>(
        *desc, x, reinterpret_cast<const float2*>(table), out, n, n_dev);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_hashgrid_bwd(const den_hashgrid_desc* desc, const float* x, const float* dout,
                     const float* table, float* dtable, float* dx, int64_t n, const int32_t* n_dev,
                     void* stream) {
    using namespace den;
    int rc = check_desc(desc);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(x && dout && dtable, "null pointer");
    DEN_CHECK_ARG(dx == nullptr || table != nullptr, "input gradient needs the table");
    const int LF = desc->n_levels * 2;
    const size_t smem = (size_t)kTile * (LF + 1 + 3) * sizeof(float);
    const int grid = grid_for((n + kTile - 1) / kTile, 1, 16);
    if (dx) {
        hashgrid_bwd_kernel<true><<<grid, kHashThreads, smem, as_stream(stream)>>>(
            *desc, x, dout, reinterpret_cast<const float2*>(table),
            reinterpret_cast<float2*>(dtable), dx, n, n_dev);
    } else {
        hashgrid_bwd_kernel<false><<<grid, kHashThreads, smem, as_stream(stream)>>>(
            *desc, x, dout, nullptr, reinterpret_cast<float2*>(dtable), nullptr, n, n_dev);
    }
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
