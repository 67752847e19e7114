// Occupancy-grid update — replaces nerfacc.OccupancyGrid._update
// (grid.py, called through every_n_step at models/nerf.py:200-204) and the cone-aware
// occ_eval_fn of NeRF.update_occ_grid (models/nerf.py:171-198).
//
// Upstream is ~25 eager torch launches over res^3 cells (gather of grid_coords, add, divide, norm,
// boolean compaction, contract_inv kernel, indexed read-modify-write, mean, compare).  Here:
//   den_occgrid_cell_points   cell index -> jittered unit point -> inverse contraction -> world point
//   den_occgrid_occ           density -> density * (cone-aware) step size
//   den_occgrid_ema_update    occs[idx] = max(occs[idx] * decay, occ), mean, binary = occs > min(mean, thre)
// The random draws (cell indices, jitter, camera ids) stay torch calls in upstream order and are
// passed in.  Arithmetic is written with non-contracting intrinsics in the order of the upstream
// elementwise ops so the points are bit-identical to the torch evaluation.
#include <limits.h>

#include "den_common.cuh"

namespace den {

constexpr int kOccBlocks = 1024;      // fixed grid of the reduction: the fp64 sum order is run-to-run fixed
constexpr int kOccThreads = 256;

__global__ void occgrid_cell_points_kernel(const __grid_constant__ den_occgrid_desc g,
                                           const int64_t* __restrict__ indices,
                                           const float* __restrict__ jitter, int64_t n,
                                           float* __restrict__ world, uint8_t* __restrict__ keep) {
    const int64_t ryz = (int64_t)g.res[1] * g.res[2];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = indices ? indices[i] : i;
        const int64_t cx = cell / ryz, rem = cell - cx * ryz;
        const int64_t cy = rem / g.res[2], cz = rem - cy * g.res[2];
        const int64_t c[3] = {cx, cy, cz};
        float x[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)     // (grid_coords + rand) / resolution
            x[d] = __fdiv_rn(__fadd_rn((float)c[d], jitter[3 * i + d]), (float)g.res[d]);
        bool inside = true;
        float u[3];
        if (g.contraction == DEN_CONTRACT_SPHERE) {
            // upstream drops points with ||x - 0.5|| >= 0.5 before the inverse contraction
            const float a = __fadd_rn(x[0], -0.5f), b = __fadd_rn(x[1], -0.5f), c2 = __fadd_rn(x[2], -0.5f);
            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c2, c2)));
            inside = nrm < 0.5f;
            // helpers_contraction.h unit_sphere_to_inf: v = 4 (x - 0.5); n > 1: v /= max(2n - n^2, eps)
            float v[3] = {__fmul_rn(a, 4.f), __fmul_rn(b, 4.f), __fmul_rn(c2, 4.f)};
            const float nsq = __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2]));
            const float nv = sqrtf(nsq);
            if (nv > 1.f) {
                const float den = fmaxf(__fadd_rn(__fmul_rn(2.f, nv), -nsq), 1e-10f);
#pragma unroll
                for (int d = 0; d < 3; ++d) v[d] = __fdiv_rn(v[d], den);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) u[d] = __fadd_rn(__fmul_rn(v[d], 0.5f), 0.5f);
        } else if (g.contraction == DEN_CONTRACT_TANH) {
#pragma unroll
            for (int d = 0; d < 3; ++d) u[d] = __fadd_rn(atanhf(__fmul_rn(__fadd_rn(x[d], -0.5f), 2.f)), 0.5f);
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) u[d] = x[d];
        }
#pragma unroll
        for (int d = 0; d < 3; ++d)
            world[3 * i + d] = __fadd_rn(__fmul_rn(u[d], __fadd_rn(g.roi[d + 3], -g.roi[d])), g.roi[d]);
        if (keep) keep[i] = inside ? 1 : 0;
    }
}

__global__ void occgrid_occ_kernel(const float* __restrict__ sigma, const float* __restrict__ world,
                                   const int64_t* __restrict__ camera_ids,
                                   const float* __restrict__ camera_pos, float cone_angle, float step,
                                   int has_planes, float near_plane, float far_plane, int64_t n,
                                   float* __restrict__ occ) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float s = step;
        if (cone_angle > 0.f) {
            const int64_t cam = camera_ids[i];
            const float a = __fadd_rn(camera_pos[3 * cam], -world[3 * i]);
            const float b = __fadd_rn(camera_pos[3 * cam + 1], -world[3 * i + 1]);
            const float c = __fadd_rn(camera_pos[3 * cam + 2], -world[3 * i + 2]);
            const float t = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
            s = fmaxf(__fmul_rn(t, cone_angle), step);
            if (has_planes && !(t > near_plane && t < far_plane)) s = 0.f;
        }
        occ[i] = __fmul_rn(sigma[i], s);
    }
}

// order-preserving float <-> int key (atomicMax on the key == max on the float)
__device__ __forceinline__ int float_key(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void occgrid_scatter_max_kernel(const int64_t* __restrict__ indices,
                                           const uint8_t* __restrict__ keep,
                                           const float* __restrict__ occ, int64_t n, int64_t n_cells,
                                           int32_t* __restrict__ scratch) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (keep && !keep[i]) continue;
        const int64_t cell = indices ? indices[i] : i;
        if (cell < 0 || cell >= n_cells) continue;
        const float v = occ[i];
        if (v != v) continue;            // a NaN density never raises a cell
        atomicMax(scratch + cell, float_key(v));
    }
}

// occs[c] = max(occs[c] * decay, candidate) for the touched cells, scratch reset, fp64 block sums
__global__ void __launch_bounds__(kOccThreads)
occgrid_ema_kernel(float* __restrict__ occs, int32_t* __restrict__ scratch, float decay, int64_t n_cells,
                   double* __restrict__ partials) {
    double acc = 0.0;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_cells;
         c += (int64_t)gridDim.x * blockDim.x) {
        float v = occs[c];
        const int key = scratch[c];
        if (key != INT_MIN) {
            v = fmaxf(__fmul_rn(v, decay), key_float(key));
            occs[c] = v;
            scratch[c] = INT_MIN;
        }
        acc += (double)v;
    }
    __shared__ double sh[kOccThreads];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kOccThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}

// every block re-derives the mean from the block sums (fixed order), then thresholds its cells
__global__ void __launch_bounds__(kOccThreads)
occgrid_threshold_kernel(const float* __restrict__ occs, const double* __restrict__ partials,
                         int n_partials, float occ_thre, int64_t n_cells, uint8_t* __restrict__ binary,
                         float* __restrict__ mean_out) {
    __shared__ double sh[kOccThreads];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += kOccThreads) acc += partials[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kOccThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    const float mean = (float)(sh[0] / (double)n_cells);
    const float thr = fminf(mean, occ_thre);        // torch.clamp(occs.mean(), max=occ_thre)
    if (blockIdx.x == 0 && threadIdx.x == 0 && mean_out) *mean_out = mean;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_cells;
         c += (int64_t)gridDim.x * blockDim.x)
        binary[c] = occs[c] > thr ? 1 : 0;
}

__global__ void occgrid_fill_kernel(int32_t* __restrict__ scratch, int64_t n) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n;
         c += (int64_t)gridDim.x * blockDim.x)
        scratch[c] = INT_MIN;
}

}  // namespace den

extern "C" {

size_t den_occgrid_workspace_bytes(int64_t n_cells) {
    return (size_t)(n_cells > 0 ? n_cells : 0) * sizeof(int32_t) + den::kOccBlocks * sizeof(double);
}

int den_occgrid_workspace_init(void* workspace, int64_t n_cells, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(workspace != nullptr && n_cells > 0, "null workspace or no cells");
    int32_t* scratch = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) +
                                                  kOccBlocks * sizeof(double));
    occgrid_fill_kernel<<<grid_for(n_cells, 256, 8), 256, 0, as_stream(stream)>>>(scratch, n_cells);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_occgrid_cell_points(const den_occgrid_desc* g, const int64_t* indices, const float* jitter,
                            int64_t n, float* world, uint8_t* keep, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(g != nullptr, "null descriptor");
    DEN_CHECK_ARG(n >= 0, "negative point count");
    DEN_CHECK_ARG(g->res[0] > 0 && g->res[1] > 0 && g->res[2] > 0, "grid resolution must be positive");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(jitter && world, "null pointer");
    occgrid_cell_points_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(*g, indices, jitter, n,
                                                                                 world, keep);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_occgrid_occ(const float* sigma, const float* world, const int64_t* camera_ids,
                    const float* camera_pos, float cone_angle, float step_size, int has_planes,
                    float near_plane, float far_plane, int64_t n, float* occ, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0, "negative point count");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(sigma && occ, "null pointer");
    DEN_CHECK_ARG(!(cone_angle > 0.f) || (world && camera_ids && camera_pos),
                  "a cone angle needs the points, the camera ids and the camera positions");
    occgrid_occ_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        sigma, world, camera_ids, camera_pos, cone_angle, step_size, has_planes, near_plane, far_plane, n, occ);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

int den_occgrid_ema_update(const int64_t* indices, const uint8_t* keep, const float* occ, int64_t n,
                           float ema_decay, float occ_thre, float* occs, int64_t n_cells,
                           uint8_t* binary, float* mean_out, void* workspace, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0 && n_cells > 0, "bad sizes");
    DEN_CHECK_ARG(occs && binary && workspace, "null pointer");
    DEN_CHECK_ARG(n == 0 || occ, "null occupancy values");
    DEN_CHECK_ARG(indices != nullptr || n <= n_cells, "more points than cells without an index list");
    double* partials = reinterpret_cast<double*>(workspace);
    int32_t* scratch = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) +
                                                  kOccBlocks * sizeof(double));
    cudaStream_t s = as_stream(stream);
    if (n > 0) {
        occgrid_scatter_max_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(indices, keep, occ, n, n_cells, scratch);
        DEN_CHECK_LAUNCH();
    }
    const int blocks = (int)((n_cells + kOccThreads - 1) / kOccThreads < kOccBlocks
                                 ? (n_cells + kOccThreads - 1) / kOccThreads
                                 : kOccBlocks);
    occgrid_ema_kernel<<<blocks, kOccThreads, 0, s>>>(occs, scratch, ema_decay, n_cells, partials);
    DEN_CHECK_LAUNCH();
    occgrid_threshold_kernel<<<grid_for(n_cells, kOccThreads, 8), kOccThreads, 0, s>>>(
        occs, partials, blocks, occ_thre, n_cells, binary, mean_out);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
