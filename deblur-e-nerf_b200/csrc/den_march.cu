// Occupancy-grid ray marching, visibility filtering and sample compaction (sm_100a).
//
// Replaces nerfacc.ray_marching / ray_aabb_intersect / render_visibility — reference
// call sites external/utils.py:106-119 and models/nerf.py:248-251; semantics per
// nerfacc 0.3.1 cuda/csrc/{ray_marching,intersection,render_transmittance}.cu and
// include/helpers_contraction.h, restated (with every fp32 rounding explicit) in
// oracle/nerfacc_ref.py.  Sample indices must match the oracle BIT-EXACTLY, so all
// marching arithmetic uses the non-contracting intrinsics __fadd_rn/__fmul_rn/
// __fdiv_rn (nvcc would otherwise fuse mul+add into FMA).
//
// Data flow (no host synchronisation inside the library):
//   march_count (1 thread/ray) -> exclusive scan -> march_write into a caller-owned
//   arena -> [density pre-pass elsewhere] -> visibility (sequential T per ray) ->
//   scan -> compaction (1 warp/ray, ballot-packed, coalesced).
#include "den_common.cuh"

namespace den {

// ---------------------------------------------------------------- helpers ----
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

struct MarchCtx {
    float roi_min[3], extent[3], roi_max[3];
    float resf[3];
    int res[3];
    int contraction;
    float dt_min, cone;
};

__device__ __forceinline__ float calc_dt(float t, const MarchCtx& c) {
    // clamp(t * cone_angle, dt_min, dt_max = 1e10)
    return fmaxf(c.dt_min, fminf(fmul(t, c.cone), 1e10f));
}

// helpers_contraction.h apply_contraction + ray_marching.cu grid_occupied_at
__device__ __forceinline__ bool occupied_at(const float xyz[3], const MarchCtx& c,
                                            const uint8_t* __restrict__ binary) {
    if (c.contraction == DEN_CONTRACT_AABB) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (!(xyz[d] >= c.roi_min[d] && xyz[d] <= c.roi_max[d])) return false;
    }
    float u[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = fdiv(fsub(xyz[d], c.roi_min[d]), c.extent[d]);
    if (c.contraction == DEN_CONTRACT_SPHERE) {
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = fsub(fmul(u[d], 2.0f), 1.0f);
        const float nsq = fadd(fadd(fmul(u[0], u[0]), fmul(u[1], u[1])), fmul(u[2], u[2]));
        const float norm = __fsqrt_rn(nsq);
        if (norm > 1.0f) {
            const float k = fsub(2.0f, fdiv(1.0f, norm));
#pragma unroll
            for (int d = 0; d < 3; ++d) u[d] = fmul(k, fdiv(u[d], norm));
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = fadd(fmul(u[d], 0.25f), 0.5f);
    } else if (c.contraction == DEN_CONTRACT_TANH) {
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = fadd(fmul(tanhf(fsub(u[d], 0.5f)), 0.5f), 0.5f);
    }
    int ijk[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float s = fmul(u[d], c.resf[d]);
        int i = isfinite(s) ? (int)s : 0;         // truncation toward zero
        ijk[d] = min(max(i, 0), c.res[d] - 1);
    }
    const int64_t idx = ((int64_t)ijk[0] * c.res[1] + ijk[1]) * c.res[2] + ijk[2];
    return binary[idx] != 0;
}

// kMode 0: count only.  1: write at offsets[r] (second pass of the two-pass scheme).  2: single pass —
// write into the ray's own segment [offsets[r], offsets[r+1]) of an upper-bound arena AND count.
template <int kMode>
__global__ void __launch_bounds__(128)
march_kernel(const __grid_constant__ den_march_params p, const float* __restrict__ rays_o,
             const float* __restrict__ rays_d, const float* __restrict__ t_min,
             const float* __restrict__ t_max, const uint8_t* __restrict__ binary,
             const int32_t* __restrict__ offsets, int32_t* __restrict__ num_steps,
             int32_t* __restrict__ ray_indices, float* __restrict__ t_starts,
             float* __restrict__ t_ends, int64_t n_rays, int64_t capacity) {
    MarchCtx c;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        c.roi_min[d] = p.roi[d];
        c.roi_max[d] = p.roi[d + 3];
        c.extent[d] = fsub(p.roi[d + 3], p.roi[d]);
        c.res[d] = p.res[d];
        c.resf[d] = (float)p.res[d];
    }
    c.contraction = p.contraction;
    c.dt_min = p.step_size;
    c.cone = p.cone_angle;

    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rays;
         r += (int64_t)gridDim.x * blockDim.x) {
        const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
        const float dir[3] = {rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2]};
        const float far = t_max[r];
        int64_t base = 0, seg = 0;
        if (kMode >= 1) base = offsets[r];
        if (kMode == 2) seg = offsets[r + 1] - base;

        int j = 0;
        float t0 = t_min[r];
        float t1 = fadd(t0, calc_dt(t0, c));
        float tm = fmul(fadd(t0, t1), 0.5f);
        if (c.contraction == DEN_CONTRACT_AABB) {
            // AABB: the unit-cube coordinate u of the sample is needed both by the occupancy lookup and by
            // the voxel skip — computed once per step; 1/dir, sign(dir) are per-ray constants; the division
            // by a power-of-two grid resolution is an exact multiplication.  Every value is bit-identical
            // to the per-use evaluation (same IEEE operations on the same operands).
            float inv[3], half_sgn[3], rres[3];
            bool pow2[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                inv[d] = fdiv(1.0f, dir[d]);
                half_sgn[d] = fmul(0.5f, copysignf(1.0f, dir[d]));
                pow2[d] = (c.res[d] & (c.res[d] - 1)) == 0;
                rres[d] = fdiv(1.0f, c.resf[d]);
            }
            while (tm < far) {
                float xyz[3], u[3];
                bool in_box = true;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    xyz[d] = fadd(o[d], fmul(tm, dir[d]));
                    in_box = in_box && (xyz[d] >= c.roi_min[d] && xyz[d] <= c.roi_max[d]);
                    u[d] = fdiv(fsub(xyz[d], c.roi_min[d]), c.extent[d]);
                }
                bool occ = false;
                if (in_box) {
                    int ijk[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float sc = fmul(u[d], c.resf[d]);
                        const int i = isfinite(sc) ? (int)sc : 0;
                        ijk[d] = min(max(i, 0), c.res[d] - 1);
                    }
                    occ = binary[((int64_t)ijk[0] * c.res[1] + ijk[1]) * c.res[2] + ijk[2]] != 0;
                }
                if (occ) {
                    if (kMode == 1) {
                        const int64_t k = base + j;
                        if (k < capacity) {
                            t_starts[k] = t0;
                            t_ends[k] = t1;
                            ray_indices[k] = (int32_t)r;
                        }
                    } else if (kMode == 2) {
                        if (j < seg) {
                            t_starts[base + j] = t0;
                            t_ends[base + j] = t1;
                        }
                    }
                    ++j;
                    t0 = t1;
                    t1 = fadd(t0, calc_dt(t0, c));
                    tm = fmul(fadd(t0, t1), 0.5f);
                } else {
                    // DDA-like skip to the next voxel face, then catch up in whole steps
                    float tnext = 3.4e38f;
                    bool any = false;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float ug = fmul(u[d], c.resf[d]);
                        const float face = floorf(fadd(fadd(ug, 0.5f), half_sgn[d]));
                        const float num = fmul(fsub(face, ug), inv[d]);
                        const float tx = fmul(pow2[d] ? fmul(num, rres[d]) : fdiv(num, c.resf[d]), c.extent[d]);
                        // fminf semantics: ignore NaN
                        if (!isnan(tx)) { tnext = any ? fminf(tnext, tx) : tx; any = true; }
                    }
                    if (!any) tnext = __int_as_float(0x7fc00000);   // all-NaN -> NaN, fmaxf(NaN,0)=0
                    const float dist = fmaxf(tnext, 0.0f);
                    const float target = fadd(tm, dist);
                    float t = tm;
                    do { t = fadd(t, c.dt_min); } while (t < target);
                    tm = t;
                    const float dt = calc_dt(tm, c);
                    t0 = fsub(tm, fmul(dt, 0.5f));
                    t1 = fadd(tm, fmul(dt, 0.5f));
                }
            }
        } else
        while (tm < far) {
            float xyz[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) xyz[d] = fadd(o[d], fmul(tm, dir[d]));
            if (occupied_at(xyz, c, binary)) {
                if (kMode == 1) {
                    const int64_t k = base + j;
                    if (k < capacity) {
                        t_starts[k] = t0;
                        t_ends[k] = t1;
                        ray_indices[k] = (int32_t)r;
                    }
                } else if (kMode == 2) {
                    if (j < seg) {
                        t_starts[base + j] = t0;
                        t_ends[base + j] = t1;
                    }
                }
                ++j;
                t0 = t1;
                t1 = fadd(t0, calc_dt(t0, c));
                tm = fmul(fadd(t0, t1), 0.5f);
            } else {
                t0 = t1;
                t1 = fadd(t0, calc_dt(t0, c));
                tm = fmul(fadd(t0, t1), 0.5f);
            }
        }
        if (kMode != 1) num_steps[r] = j;
    }
}

// Upper bound of the samples ray r can emit: every emitted sample advances t0 by dt >= dt_min and has
// its midpoint below t_max, so count <= (t_max - t_min) / dt_min + 1 (+1 rounding slack).
__global__ void march_bound_kernel(const float* __restrict__ t_min, const float* __restrict__ t_max,
                                   float dt_min, int32_t cap, int32_t* __restrict__ bound, int64_t n_rays) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rays;
         r += (int64_t)gridDim.x * blockDim.x) {
        const float span = t_max[r] - t_min[r];
        int32_t b = 0;
        if (span > 0.f) {
            const float steps = span / dt_min + 3.0f;
            b = steps < (float)cap ? (int32_t)steps : cap;
        }
        bound[r] = b;
    }
}

// One warp per ray: copy the first count[r] samples of the ray's arena segment to their packed place.
__global__ void __launch_bounds__(256)
march_pack_kernel(const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ offsets,
                  const float* __restrict__ arena_t0, const float* __restrict__ arena_t1,
                  int64_t n_rays, int32_t* __restrict__ ray_indices, float* __restrict__ t_starts,
                  float* __restrict__ t_ends) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rays; r += nwarps) {
        const int64_t src = seg_offsets[r];
        const int64_t dst = offsets[r];
        const int count = offsets[r + 1] - offsets[r];
        for (int j = lane; j < count; j += 32) {
            t_starts[dst + j] = arena_t0[src + j];
            t_ends[dst + j] = arena_t1[src + j];
            ray_indices[dst + j] = (int32_t)r;
        }
    }
}

__global__ void aabb_kernel(const float* __restrict__ o, const float* __restrict__ d, float ax0,
                            float ay0, float az0, float ax1, float ay1, float az1,
                            float* __restrict__ tmin_out, float* __restrict__ tmax_out, int64_t n) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x) {
        const float ox = o[3 * r], oy = o[3 * r + 1], oz = o[3 * r + 2];
        const float dx = d[3 * r], dy = d[3 * r + 1], dz = d[3 * r + 2];
        float tmin = fdiv(fsub(ax0, ox), dx), tmax = fdiv(fsub(ax1, ox), dx);
        if (tmin > tmax) { float t = tmin; tmin = tmax; tmax = t; }
        float tymin = fdiv(fsub(ay0, oy), dy), tymax = fdiv(fsub(ay1, oy), dy);
        if (tymin > tymax) { float t = tymin; tymin = tymax; tymax = t; }
        bool miss = (tmin > tymax) || (tymin > tmax);
        if (tymin > tmin) tmin = tymin;
        if (tymax < tmax) tmax = tymax;
        float tzmin = fdiv(fsub(az0, oz), dz), tzmax = fdiv(fsub(az1, oz), dz);
        if (tzmin > tzmax) { float t = tzmin; tzmin = tzmax; tzmax = t; }
        miss = miss || (tmin > tzmax) || (tzmin > tmax);
        if (tzmin > tmin) tmin = tzmin;
        if (tzmax < tmax) tmax = tzmax;
        tmin_out[r] = miss ? 1e10f : tmin;
        tmax_out[r] = miss ? 1e10f : tmax;
    }
}

__global__ void clamp_jitter_kernel(float* __restrict__ tmin, float* __restrict__ tmax,
                                    const float* __restrict__ jitter, int has_near, float near_plane,
                                    int has_far, float far_plane, float step, int64_t n) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x) {
        float a = tmin[r], b = tmax[r];
        if (has_near) a = fmaxf(a, near_plane);
        if (has_far) b = fminf(b, far_plane);
        if (jitter) a = fadd(a, fmul(jitter[r], step));
        tmin[r] = a;
        tmax[r] = b;
    }
}

// ------------------------------------------------------------------- scan ----
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        s_warp[lane] = winc - w;       // exclusive warp offsets
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    block_total = s_warp[32];
    const int res = s_warp[warp] + inc - v;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads)
scan_local_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                  int32_t* __restrict__ block_sums, int64_t n) {
    __shared__ int s_warp[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        sum += v[i];
    }
    int total;
    int excl = block_exclusive_scan(sum, s_warp, total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
scan_sums_kernel(int32_t* __restrict__ block_sums, int64_t n_blocks, int32_t* __restrict__ total_out) {
    __shared__ int s_warp[33];
    int carry = 0;
    for (int64_t base = 0; base < n_blocks; base += kScanThreads) {
        const int64_t i = base + threadIdx.x;
        int v = i < n_blocks ? block_sums[i] : 0;
        int total;
        int excl = block_exclusive_scan(v, s_warp, total);
        if (i < n_blocks) block_sums[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ block_sums, int64_t n) {
    const int add = block_sums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) out[base + i] += add;
}

// ------------------------------------------------------------- visibility ----
__global__ void alpha_kernel(const float* __restrict__ sigmas, const float* __restrict__ t0,
                             const float* __restrict__ t1, float* __restrict__ alphas, int64_t n,
                             const int32_t* __restrict__ n_dev) {
    n = effective_n(n, n_dev);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        alphas[i] = 1.0f - expf(-sigmas[i] * (t1[i] - t0[i]));
}

// one thread per ray: T is the sequential fp32 product (bit-exact vs the oracle)
__global__ void __launch_bounds__(128)
visibility_kernel(const float* __restrict__ alphas, const int32_t* __restrict__ offsets,
                  int64_t n_rays, float eps, float alpha_thre, uint8_t* __restrict__ mask,
                  int32_t* __restrict__ vis_count) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rays;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int beg = offsets[r], end = offsets[r + 1];
        float T = 1.0f;
        int cnt = 0;
        for (int i = beg; i < end; ++i) {
            const float a = alphas[i];
            bool vis = T >= eps;
            if (alpha_thre > 0.0f) vis = vis && (a >= alpha_thre);
            mask[i] = vis ? 1 : 0;
            cnt += vis ? 1 : 0;
            T = fmul(T, fsub(1.0f, a));
        }
        vis_count[r] = cnt;
    }
}

// one warp per ray, ballot-packed
__global__ void __launch_bounds__(256)
compact_kernel(const uint8_t* __restrict__ mask, const int32_t* __restrict__ off_in,
               const int32_t* __restrict__ off_out, const int32_t* __restrict__ ray_in,
               const float* __restrict__ t0_in, const float* __restrict__ t1_in,
               int32_t* __restrict__ ray_out, float* __restrict__ t0_out,
               float* __restrict__ t1_out, int64_t n_rays, const float* __restrict__ sig_in,
               const float* __restrict__ rgb_in, int channels, float* __restrict__ sig_out,
               float* __restrict__ rgb_out, int32_t* __restrict__ src_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rays; r += nwarps) {
        const int beg = off_in[r], end = off_in[r + 1];
        int dst = off_out[r];
        if (off_out[r + 1] == dst) continue;
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const bool in = i < end;
            // the row is loaded together with its mask byte (one memory latency per chunk instead of two:
            // mask -> ballot -> loads); the rows of culled samples are read in vain, the usual case keeps most
            const uint8_t m = in ? mask[i] : (uint8_t)0;
            int32_t rv = 0;
            float a = 0.f, b = 0.f, sg = 0.f, col[4] = {0.f, 0.f, 0.f, 0.f};
            if (in) {
                rv = ray_in[i];
                a = t0_in[i];
                b = t1_in[i];
                if (sig_out) sg = sig_in[i];
                if (rgb_out) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < channels) col[c] = rgb_in[i * (int64_t)channels + c];
                }
            }
            const bool keep = m != 0;
            const unsigned ball = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int k = dst + __popc(ball & ((1u << lane) - 1u));
                ray_out[k] = rv;
                t0_out[k] = a;
                t1_out[k] = b;
                if (sig_out) sig_out[k] = sg;
                if (rgb_out) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < channels) rgb_out[k * (int64_t)channels + c] = col[c];
                }
                if (src_rows) src_rows[k] = i;
            }
            dst += __popc(ball);
        }
    }
}

// offsets (exclusive prefix of per-ray counts) limited to the capacity of the sample buffers
__global__ void clamp_offsets_kernel(int32_t* __restrict__ offsets, int64_t n, int32_t capacity,
                                     int32_t* __restrict__ overflow) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = offsets[i];
        if (v > capacity) {
            offsets[i] = capacity;
            if (i == n - 1 && overflow) *overflow = 1;
        }
    }
}

}  // namespace den

extern "C" {

int den_ray_aabb_intersect(const float* o, const float* d, const float* aabb, float* tmin,
                           float* tmax, int64_t n, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0, "negative ray count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(o && d && aabb && tmin && tmax, "null pointer");
    aabb_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        o, d, aabb[0], aabb[1], aabb[2], aabb[3], aabb[4], aabb[5], tmin, tmax, n);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_clamp_jitter(float* tmin, float* tmax, const float* jitter, int has_near, float near_plane,
                     int has_far, float far_plane, float step, int64_t n, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0, "negative ray count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(tmin && tmax, "null pointer");
    clamp_jitter_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        tmin, tmax, jitter, has_near, near_plane, has_far, far_plane, step, n);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

static int check_march(const den_march_params* p) {
    using namespace den;
    DEN_CHECK_ARG(p != nullptr, "null params");
    DEN_CHECK_ARG(p->res[0] > 0 && p->res[1] > 0 && p->res[2] > 0, "grid resolution must be positive");
    DEN_CHECK_ARG(p->contraction >= 0 && p->contraction <= 2, "unknown contraction type");
    DEN_CHECK_ARG(p->step_size > 0.0f, "render_step_size must be positive");
    DEN_CHECK_ARG(p->cone_angle >= 0.0f, "cone_angle must be non-negative");
    return DEN_OK;
}

int den_march_count(const den_march_params* p, const float* o, const float* d, const float* tmin,
                    const float* tmax, const uint8_t* binary, int32_t* num_steps, int64_t n,
                    void* stream) {
    using namespace den;
    int rc = check_march(p);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative ray count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(o && d && tmin && tmax && binary && num_steps, "null pointer");
    march_kernel<0><<<grid_for(n, 128, 16), 128, 0, as_stream(stream)>>>(
        *p, o, d, tmin, tmax, binary, nullptr, num_steps, nullptr, nullptr, nullptr, n, 0);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_march_write(const den_march_params* p, const float* o, const float* d, const float* tmin,
                    const float* tmax, const uint8_t* binary, const int32_t* offsets,
                    int32_t* ray_indices, float* t_starts, float* t_ends, int64_t n,
                    int64_t capacity, void* stream) {
    using namespace den;
    int rc = check_march(p);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0 && capacity >= 0, "negative size");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(o && d && tmin && tmax && binary && offsets, "null pointer");
    DEN_CHECK_ARG(capacity == 0 || (ray_indices && t_starts && t_ends), "null output");
    march_kernel<1><<<grid_for(n, 128, 16), 128, 0, as_stream(stream)>>>(
        *p, o, d, tmin, tmax, binary, offsets, nullptr, ray_indices, t_starts, t_ends, n, capacity);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_march_bound(const den_march_params* p, const float* tmin, const float* tmax,
                    int32_t max_per_ray, int32_t* bound, int64_t n, void* stream) {
    using namespace den;
    int rc = check_march(p);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0 && max_per_ray > 0, "bad size");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(tmin && tmax && bound, "null pointer");
    march_bound_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(tmin, tmax, p->step_size,
                                                                          max_per_ray, bound, n);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_march_single(const den_march_params* p, const float* o, const float* d, const float* tmin,
                     const float* tmax, const uint8_t* binary, const int32_t* seg_offsets,
                     int32_t* num_steps, float* arena_t0, float* arena_t1, int64_t n, void* stream) {
    using namespace den;
    int rc = check_march(p);
    if (rc) return rc;
    DEN_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(o && d && tmin && tmax && binary && seg_offsets && num_steps && arena_t0 && arena_t1,
                  "null pointer");
    march_kernel<2><<<grid_for(n, 128, 16), 128, 0, as_stream(stream)>>>(
        *p, o, d, tmin, tmax, binary, seg_offsets, num_steps, nullptr, arena_t0, arena_t1, n, 0);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_march_pack(const int32_t* seg_offsets, const int32_t* offsets, const float* arena_t0,
                   const float* arena_t1, int64_t n, int32_t* ray_indices, float* t_starts,
                   float* t_ends, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(seg_offsets && offsets && arena_t0 && arena_t1, "null pointer");
    march_pack_kernel<<<grid_for(n, 8, 8), 256, 0, as_stream(stream)>>>(
        seg_offsets, offsets, arena_t0, arena_t1, n, ray_indices, t_starts, t_ends);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

size_t den_scan_workspace_bytes(int64_t n) {
    const int64_t blocks = (n + den::kScanTile - 1) / den::kScanTile;
    return (size_t)(blocks > 0 ? blocks : 1) * sizeof(int32_t);
}

int den_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* workspace,
                           size_t workspace_bytes, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0, "negative length");
    DEN_CHECK_ARG(out != nullptr, "null output");
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(out, 0, sizeof(int32_t), as_stream(stream));
        if (e != cudaSuccess) return cuda_fail(e, "den_exclusive_scan_i32");
        return DEN_OK;
    }
    DEN_CHECK_ARG(in && workspace, "null pointer");
    if (workspace_bytes < den_scan_workspace_bytes(n)) {
        set_error("den_exclusive_scan_i32: workspace too small (%zu < %zu)", workspace_bytes,
                  den_scan_workspace_bytes(n));
        return DEN_ERR_WORKSPACE;
    }
    const int64_t blocks = (n + kScanTile - 1) / kScanTile;
    int32_t* sums = static_cast<int32_t*>(workspace);
    scan_local_kernel<<<(unsigned)blocks, kScanThreads, 0, as_stream(stream)>>>(in, out, sums, n);
    scan_sums_kernel<<<1, kScanThreads, 0, as_stream(stream)>>>(sums, blocks, out + n);
    scan_add_kernel<<<(unsigned)blocks, kScanThreads, 0, as_stream(stream)>>>(out, sums, n);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_alpha_from_sigma(const float* sigmas, const float* t0, const float* t1, float* alphas,
                         int64_t n, const int32_t* n_dev, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0, "negative sample count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(sigmas && t0 && t1 && alphas, "null pointer");
    alpha_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(sigmas, t0, t1, alphas, n, n_dev);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_visibility(const float* alphas, const int32_t* offsets, int64_t n_rays, float eps,
                   float alpha_thre, uint8_t* mask, int32_t* vis_count, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "negative ray count");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(offsets && vis_count, "null pointer");
    visibility_kernel<<<grid_for(n_rays, 128, 16), 128, 0, as_stream(stream)>>>(
        alphas, offsets, n_rays, eps, alpha_thre, mask, vis_count);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_compact_samples(const uint8_t* mask, const int32_t* off_in, const int32_t* off_out,
                        const int32_t* ray_in, const float* t0_in, const float* t1_in,
                        int32_t* ray_out, float* t0_out, float* t1_out, int64_t n_rays,
                        void* stream) {
    return den_compact_samples_ex(mask, off_in, off_out, ray_in, t0_in, t1_in, ray_out, t0_out, t1_out,
                                  n_rays, nullptr, nullptr, 0, nullptr, nullptr, nullptr, stream);
}

int den_compact_samples_ex(const uint8_t* mask, const int32_t* off_in, const int32_t* off_out,
                           const int32_t* ray_in, const float* t0_in, const float* t1_in,
                           int32_t* ray_out, float* t0_out, float* t1_out, int64_t n_rays,
                           const float* sig_in, const float* rgb_in, int32_t channels, float* sig_out,
                           float* rgb_out, int32_t* src_rows, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "negative ray count");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(off_in && off_out, "null pointer");
    DEN_CHECK_ARG((sig_out == nullptr || sig_in != nullptr) && (rgb_out == nullptr || rgb_in != nullptr),
                  "an output row array needs its input");
    DEN_CHECK_ARG(rgb_out == nullptr || (channels >= 1 && channels <= 4), "1 to 4 channels");
    compact_kernel<<<grid_for(n_rays, 8, 8), 256, 0, as_stream(stream)>>>(
        mask, off_in, off_out, ray_in, t0_in, t1_in, ray_out, t0_out, t1_out, n_rays, sig_in, rgb_in,
        channels, sig_out, rgb_out, src_rows);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_clamp_offsets(int32_t* offsets, int64_t n_plus_1, int32_t capacity, int32_t* overflow,
                      void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_plus_1 >= 1 && capacity >= 0, "bad sizes");
    DEN_CHECK_ARG(offsets != nullptr, "null pointer");
    clamp_offsets_kernel<<<grid_for(n_plus_1, 256, 8), 256, 0, as_stream(stream)>>>(offsets, n_plus_1,
                                                                                   capacity, overflow);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
