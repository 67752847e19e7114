// Raw event stream -> queued events on the device (the data format in front of the hot path).
//
// Replaces the per-event Python loops of Event.queue_raw_events (data/datasets.py:186-276) and
// Event.extract_max_refractory_period (:131-183): both walk the time-ordered raw events once and keep, for
// every pixel, a two-entry sliding window of the last events seen there.  What they compute per event i is a
// function of the PREVIOUS raw event at the same pixel, prev(i):
//   queue      valid_i    = prev(i) exists and timestamp[prev(i)] != timestamp[i]            (:246-253)
//              start_ts_i = timestamp[prev(i)], end_ts_i = timestamp[i]                       (:255-257)
//              num_pos_i  = polarity_i, num_neg_i = 1 - polarity_i   (window of 2: the earlier event only
//                           dates the interval, :260-267) — formed by the caller, no kernel needed
//   refractory max_refractory_period = min over i with prev(i) and timestamp[i] != timestamp[prev(i)] of
//              timestamp[i] - timestamp[prev(i)]: an event equal in time to the window's last entry is not
//              appended (:163-168), so the window's last entry always equals the previous event's time.
// prev(i) comes from a STABLE sort of the event indices by pixel id (y * width + x): within a pixel the
// indices stay in stream order, so prev(i) is the left neighbour in the sorted order when it shares the key.
//
//   den_radix_sort_pairs_u32   stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass:
//                              per-tile digit histograms -> one exclusive scan over [digit][tile] ->
//                              stable scatter (rows of 256 consecutive elements in order; inside a row
//                              __match_any_sync ranks the lanes of a warp, per-warp digit counts rank the warps)
//   den_queue_raw_events       pixel keys -> sort -> neighbour pass (keep flag, start_ts, min interval) ->
//                              exclusive scan of the flags (where each kept event goes)
//   den_compact_queued_events  the kept events in stream order, in the reference's layout
// Integer work, bit-exact against the reference's loops (oracle/events_ref.py, tests/golden/raw_events.npz).
#include <limits.h>

#include "den_common.cuh"

namespace den {

constexpr int kSortThreads = 256;                       // == radix: thread t owns digit t in the tables
constexpr int kSortRows = 8;                            // rows of 256 consecutive elements per tile
constexpr int kSortTile = kSortThreads * kSortRows;
constexpr int kRadix = 256;
constexpr int kSortWarps = kSortThreads / 32;

__device__ __forceinline__ uint32_t digit_of(uint32_t key, int shift) { return (key >> shift) & 0xffu; }

// hist[digit][tile] = number of keys of the tile with that digit
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int num_tiles,
                  int32_t* __restrict__ hist) {
    __shared__ int s_hist[kRadix];
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        s_hist[threadIdx.x] = 0;
        __syncthreads();
        const int64_t base = (int64_t)tile * kSortTile;
#pragma unroll
        for (int r = 0; r < kSortRows; ++r) {
            const int64_t i = base + r * kSortThreads + threadIdx.x;
            if (i < n) atomicAdd(&s_hist[digit_of(__ldg(keys + i), shift)], 1);
        }
        __syncthreads();
        hist[(int64_t)threadIdx.x * num_tiles + tile] = s_hist[threadIdx.x];
        __syncthreads();
    }
}

// offsets = exclusive scan of hist in [digit][tile] order: where the tile's first key of a digit goes.
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                     int num_tiles, const int32_t* __restrict__ offsets) {
    __shared__ int s_base[kRadix];                      // destination of the tile's next key of each digit
    __shared__ int s_warp[kSortWarps][kRadix];          // digit counts of the current row, per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        s_base[threadIdx.x] = offsets[(int64_t)threadIdx.x * num_tiles + tile];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) s_warp[w][threadIdx.x] = 0;
        __syncthreads();
        const int64_t base = (int64_t)tile * kSortTile;
        for (int r = 0; r < kSortRows; ++r) {
            const int64_t i = base + r * kSortThreads + threadIdx.x;
            const bool valid = i < n;
            uint32_t key = 0, val = 0;
            if (valid) {
                key = __ldg(keys_in + i);
                val = __ldg(vals_in + i);
            }
            const uint32_t d = valid ? digit_of(key, shift) : (uint32_t)kRadix;   // idle lanes group apart
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank_in_warp == 0) s_warp[warp][d] = __popc(peers);
            __syncthreads();
            if (valid) {
                int dst = s_base[d] + rank_in_warp;
                for (int w = 0; w < warp; ++w) dst += s_warp[w][d];
                keys_out[dst] = key;
                vals_out[dst] = val;
            }
            __syncthreads();
            int add = 0;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                add += s_warp[w][threadIdx.x];
                s_warp[w][threadIdx.x] = 0;
            }
            s_base[threadIdx.x] += add;
            __syncthreads();
        }
    }
}

// key = y * width + x, value = stream index; positions outside the sensor raise the flag
__global__ void event_keys_kernel(const int32_t* __restrict__ position_xy, int64_t n, int width, int height,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                  int32_t* __restrict__ out_of_range) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int x = position_xy[2 * i + 0], y = position_xy[2 * i + 1];
        if (x < 0 || x >= width || y < 0 || y >= height) {
            *out_of_range = 1;
            x = min(max(x, 0), width - 1);
            y = min(max(y, 0), height - 1);
        }
        keys[i] = (uint32_t)y * (uint32_t)width + (uint32_t)x;
        vals[i] = (uint32_t)i;
    }
}

// sorted (key, index) pairs -> per event (stream order): keep flag, start_ts; global min of the non-zero
// intervals.  Each thread fetches the timestamp of ITS event (one random 8-byte read); the left neighbour's
// comes over a shuffle (lane 0 fetches it).  One atomicMin per CTA.
__global__ void __launch_bounds__(256)
queue_events_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ order,
                    const int64_t* __restrict__ timestamp, int64_t n, int64_t* __restrict__ start_ts,
                    int32_t* __restrict__ keep, long long* __restrict__ min_interval) {
    __shared__ long long s_min[8];
    const int lane = threadIdx.x & 31;
    long long local_min = LLONG_MAX;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (n + stride - 1) / stride;               // every lane runs every round (shuffles)
    for (int64_t r = 0; r < rounds; ++r) {
        const int64_t s = r * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        const bool in = s < n;
        uint32_t i = 0, key = 0;
        long long ts = 0;
        if (in) {
            i = order[s];
            key = keys[s];
            ts = timestamp[i];
        }
        uint32_t key_left = __shfl_up_sync(0xffffffffu, key, 1);
        long long ts_left = __shfl_up_sync(0xffffffffu, ts, 1);
        if (lane == 0 && in && s > 0) {
            key_left = keys[s - 1];
            ts_left = timestamp[order[s - 1]];
        }
        if (in) {
            const bool ok = s > 0 && key_left == key && ts_left != ts;
            if (ok) local_min = min(local_min, ts - ts_left);
            keep[i] = ok ? 1 : 0;
            start_ts[i] = ok ? ts_left : 0;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) local_min = min(local_min, __shfl_xor_sync(0xffffffffu, local_min, d));
    if (lane == 0) s_min[threadIdx.x >> 5] = local_min;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = s_min[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = min(m, s_min[w]);
        if (m != LLONG_MAX) atomicMin(min_interval, m);
    }
}

// kept events, in stream order, in the reference's layout (position int64 (M, 2), the rest int64 (M))
__global__ void __launch_bounds__(256)
compact_queued_kernel(const int32_t* __restrict__ position_xy, const int64_t* __restrict__ timestamp,
                      const uint8_t* __restrict__ polarity, const int64_t* __restrict__ start_ts,
                      const int32_t* __restrict__ kept_offsets, int64_t n,
                      int64_t* __restrict__ out_position, int64_t* __restrict__ out_start_ts,
                      int64_t* __restrict__ out_end_ts, int64_t* __restrict__ out_num_pos,
                      int64_t* __restrict__ out_num_neg) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = kept_offsets[i];
        if (kept_offsets[i + 1] == m) continue;                    // not kept: the prefix sum did not move
        const int2 xy = __ldg(reinterpret_cast<const int2*>(position_xy) + i);
        reinterpret_cast<longlong2*>(out_position)[m] = make_longlong2(xy.x, xy.y);
        out_start_ts[m] = start_ts[i];
        out_end_ts[m] = timestamp[i];
        const int64_t pos = polarity[i] ? 1 : 0;
        out_num_pos[m] = pos;
        out_num_neg[m] = 1 - pos;
    }
}

static int sort_tiles(int64_t n) { return (int)((n + kSortTile - 1) / kSortTile); }

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// result always lands in (keys_out, vals_out); (keys_tmp, vals_tmp) is the other side of the ping-pong
static int radix_sort_pairs(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                            uint32_t* keys_tmp, uint32_t* vals_tmp, int64_t n, int key_bits, void* workspace,
                            size_t workspace_bytes, cudaStream_t stream) {
    const int passes = key_bits <= 0 ? 1 : (key_bits + 7) / 8;
    const int tiles = sort_tiles(n);
    const int64_t cells = (int64_t)kRadix * tiles;
    const size_t table = align256((size_t)(cells + 1) * sizeof(int32_t));
    const size_t scan_ws = den_scan_workspace_bytes(cells);
    if (workspace_bytes < 2 * table + scan_ws) {
        set_error("radix sort: workspace too small (%zu < %zu)", workspace_bytes, 2 * table + scan_ws);
        return DEN_ERR_INVALID_ARGUMENT;
    }
    int32_t* hist = reinterpret_cast<int32_t*>(workspace);
    int32_t* offsets = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) + table);
    void* scan_space = reinterpret_cast<uint8_t*>(workspace) + 2 * table;
    const int grid = grid_for(tiles, 1, 8);
    const uint32_t* src_k = keys_in;
    const uint32_t* src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        const bool to_out = ((passes - 1 - p) & 1) == 0;          // the last pass writes the output pair
        uint32_t* dst_k = to_out ? keys_out : keys_tmp;
        uint32_t* dst_v = to_out ? vals_out : vals_tmp;
        radix_hist_kernel<<<grid, kSortThreads, 0, stream>>>(src_k, n, 8 * p, tiles, hist);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "radix_hist_kernel");
        int rc = den_exclusive_scan_i32(hist, offsets, cells, scan_space, scan_ws, stream);
        if (rc) return rc;
        radix_scatter_kernel<<<grid, kSortThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n, 8 * p, tiles, offsets);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "radix_scatter_kernel");
        src_k = dst_k;
        src_v = dst_v;
    }
    return DEN_OK;
}

static size_t sort_workspace_bytes(int64_t n) {
    const int64_t cells = (int64_t)kRadix * sort_tiles(n < 1 ? 1 : n);
    return 2 * align256((size_t)(cells + 1) * sizeof(int32_t)) + den_scan_workspace_bytes(cells);
}

static int bits_for(int64_t count) {        // bits needed to hold 0 .. count - 1
    int b = 0;
    while (((int64_t)1 << b) < count) ++b;
    return b;
}

}  // namespace den

extern "C" {

size_t den_radix_sort_workspace_bytes(int64_t n) { return den::sort_workspace_bytes(n); }

int den_radix_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                             uint32_t* vals_out, uint32_t* keys_tmp, uint32_t* vals_tmp, int64_t n,
                             int32_t key_bits, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 31), "element count out of range");
    DEN_CHECK_ARG(key_bits >= 0 && key_bits <= 32, "key_bits out of range");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(keys_in && vals_in && keys_out && vals_out && keys_tmp && vals_tmp && workspace, "null pointer");
    return radix_sort_pairs(keys_in, vals_in, keys_out, vals_out, keys_tmp, vals_tmp, n, key_bits, workspace,
                            workspace_bytes, as_stream(stream));
}

size_t den_queue_events_workspace_bytes(int64_t n) {
    const size_t arr = den::align256((size_t)(n < 1 ? 1 : n) * sizeof(uint32_t));
    return 7 * arr + den::sort_workspace_bytes(n) + den::align256(den_scan_workspace_bytes(n < 1 ? 1 : n));
}

int den_queue_raw_events(const int32_t* position_xy, const int64_t* timestamp, int64_t n, int32_t width,
                         int32_t height, void* workspace, size_t workspace_bytes, int64_t* start_ts,
                         int32_t* kept_offsets, int64_t* min_interval, int32_t* out_of_range, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 31), "event count out of range");
    DEN_CHECK_ARG(width >= 1 && height >= 1 && (int64_t)width * height <= ((int64_t)1 << 32), "bad sensor size");
    DEN_CHECK_ARG(kept_offsets != nullptr, "null pointer");
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(kept_offsets, 0, sizeof(int32_t), as_stream(stream));
        if (e != cudaSuccess) return cuda_fail(e, "den_queue_raw_events");
        return DEN_OK;
    }
    DEN_CHECK_ARG(position_xy && timestamp && workspace && start_ts && min_interval && out_of_range, "null pointer");
    DEN_CHECK_ARG(workspace_bytes >= den_queue_events_workspace_bytes(n), "workspace too small");
    const size_t arr = align256((size_t)n * sizeof(uint32_t));
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    uint32_t* k0 = reinterpret_cast<uint32_t*>(w + 0 * arr);
    uint32_t* v0 = reinterpret_cast<uint32_t*>(w + 1 * arr);
    uint32_t* k1 = reinterpret_cast<uint32_t*>(w + 2 * arr);
    uint32_t* v1 = reinterpret_cast<uint32_t*>(w + 3 * arr);
    uint32_t* k2 = reinterpret_cast<uint32_t*>(w + 4 * arr);
    uint32_t* v2 = reinterpret_cast<uint32_t*>(w + 5 * arr);
    int32_t* valid32 = reinterpret_cast<int32_t*>(w + 6 * arr);
    const size_t sort_ws = sort_workspace_bytes(n);
    cudaStream_t s = as_stream(stream);
    event_keys_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(position_xy, n, width, height, k0, v0, out_of_range);
    DEN_CHECK_LAUNCH();
    int rc = radix_sort_pairs(k0, v0, k1, v1, k2, v2, n, bits_for((int64_t)width * height), w + 7 * arr, sort_ws, s);
    if (rc) return rc;
    queue_events_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(k1, v1, timestamp, n, start_ts, valid32,
                                                           reinterpret_cast<long long*>(min_interval));
    DEN_CHECK_LAUNCH();
    // kept_offsets[i] = number of kept events before i, kept_offsets[n] = M
    return den_exclusive_scan_i32(valid32, kept_offsets, n, w + 7 * arr + sort_ws,
                                  workspace_bytes - 7 * arr - sort_ws, s);
}

int den_compact_queued_events(const int32_t* position_xy, const int64_t* timestamp, const uint8_t* polarity,
                              const int64_t* start_ts, const int32_t* kept_offsets, int64_t n,
                              int64_t* out_position, int64_t* out_start_ts, int64_t* out_end_ts,
                              int64_t* out_num_pos, int64_t* out_num_neg, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 31), "event count out of range");
    if (n == 0) return DEN_OK;
    DEN_CHECK_ARG(position_xy && timestamp && polarity && start_ts && kept_offsets, "null pointer");
    DEN_CHECK_ARG(out_position && out_start_ts && out_end_ts && out_num_pos && out_num_neg, "null output");
    compact_queued_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
        position_xy, timestamp, polarity, start_ts, kept_offsets, n, out_position, out_start_ts,
        out_end_ts, out_num_pos, out_num_neg);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

}  // extern "C"
