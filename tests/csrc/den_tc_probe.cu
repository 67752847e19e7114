// Probe kernel for the three tcgen05 GEMM flavours the tensor-core MLP uses; exercised by
// tests/test_gpu_mlp_tc.py so that descriptor conventions (K-major vs MN-major reuse of one
// shared-memory tile, M = 64 accumulator lane mapping) are pinned by a test of their own.
//   mode 0:  D[128 x N] = X[128 x K] * W[N x K]^T        X, W tiles K-major        (forward)
//   mode 1:  D[128 x N] = X[128 x K] * W[K x N]          W tile (K rows, N feats) read MN-major
//                                                        (dX = dY * W)
//   mode 2:  D[ 64 x N] = X[128 x 64]^T * Y[128 x N]     both tiles read MN-major, K = 128 rows
//                                                        (dW = dY^T * X), M = 64 lane mapping
//   mode 3:  D[128 x N] = X[128 x K] * W[N x K]^T        X is the A operand IN TENSOR MEMORY: lane = row,
//                                                        column c holds the bf16 pair (x[2c], x[2c+1])
//                                                        (written with tcgen05.st), W tile K-major in smem
// Operands are plain bf16 (single tile each); inputs are fp32 arrays rounded to bf16.
#include "den_common.cuh"
#include "den_tc.cuh"

namespace den {

__device__ __forceinline__ void store_tile_bf16(uint8_t* tile, const float* __restrict__ src,
                                                int rows, int cols) {
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
        const int r = i / cols, k = i - r * cols;
        const uint32_t off = tc::chunk_offset(r, k >> 3, cols) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(tile + off) = __float2bfloat16_rn(src[i]);
    }
}

__global__ void __launch_bounds__(128)
tc_probe_kernel(int mode, const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ d,
                int n, int k) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* xt = smem;                       // up to 128 x 64 bf16 = 16 KB
    uint8_t* wt = smem + 16384;               // up to 128 x 64 bf16 = 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32776);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (mode == 0 || mode == 3) {
        store_tile_bf16(xt, x, 128, k);       // X (128, K)  (mode 3: unused, X goes to TMEM below)
        store_tile_bf16(wt, w, n, k);         // W (N, K)
    } else if (mode == 1) {
        store_tile_bf16(xt, x, 128, k);       // X (128, K)
        store_tile_bf16(wt, w, k, n);         // W (K, N): K rows, N features
    } else {
        store_tile_bf16(xt, x, 128, 64);      // X (128, 64)
        store_tile_bf16(wt, w, 128, n);       // Y (128, N)
    }
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(slot, 128);
    tc::fence_smem_to_async_proxy();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = *slot;
    constexpr uint32_t kColA = 64;            // A operand columns (K / 2 of them) behind the accumulator
    if (mode == 3) {
        // thread tid = row tid: its K values as K / 2 packed bf16 words, 8 columns per tcgen05.st
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16) + kColA;
        for (int c0 = 0; c0 < k / 2; c0 += 8) {
            uint32_t wds[8];
            for (int j = 0; j < 8; ++j) {
                const __nv_bfloat162 pr = __floats2bfloat162_rn(x[tid * k + 2 * (c0 + j)], x[tid * k + 2 * (c0 + j) + 1]);
                wds[j] = *reinterpret_cast<const uint32_t*>(&pr);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tl + c0),
                         "r"(wds[0]), "r"(wds[1]), "r"(wds[2]), "r"(wds[3]), "r"(wds[4]), "r"(wds[5]), "r"(wds[6]),
                         "r"(wds[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc::tc_fence_before_sync();
        __syncthreads();
        tc::tc_fence_after_sync();
    }

    if (warp == 0) {
        if (mode == 0) {
            const uint32_t idesc = tc::instr_desc_bf16(128, n, false, false);
            const uint32_t sbo = (k / 8) * 128;
            for (int ks = 0; ks < k / 16; ++ks)
                tc::mma_bf16(tmem, tc::smem_desc(tc::smem_u32(xt) + ks * 256, 128, sbo),
                             tc::smem_desc(tc::smem_u32(wt) + ks * 256, 128, sbo), idesc, ks > 0);
        } else if (mode == 3) {
            const uint32_t idesc = tc::instr_desc_bf16(128, n, false, false);
            const uint32_t sbo = (k / 8) * 128;
            for (int ks = 0; ks < k / 16; ++ks) {
                const uint32_t acc = ks > 0 ? 1u : 0u;
                const uint64_t bdesc = tc::smem_desc(tc::smem_u32(wt) + ks * 256, 128, sbo);
                asm volatile(
                    "{\n\t.reg .pred p, leader;\n\t"
                    "elect.sync _|leader, 0xffffffff;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "@leader tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                    :
                    : "r"(tmem), "r"(tmem + kColA + ks * 8), "l"(bdesc), "r"(idesc), "r"(acc)
                    : "memory");
            }
        } else if (mode == 1) {
            // A: X K-major (K cols).  B: W tile has K rows x N feats -> MN-major, MN = feature
            const uint32_t idesc = tc::instr_desc_bf16(128, n, false, true);
            const uint32_t a_sbo = (k / 8) * 128;
            const uint32_t w_row_group = (n / 8) * 128;    // stride between groups of 8 tile rows
            for (int ks = 0; ks < k / 16; ++ks)
                tc::mma_bf16(tmem, tc::smem_desc(tc::smem_u32(xt) + ks * 256, 128, a_sbo),
                             tc::smem_desc(tc::smem_u32(wt) + ks * 2 * w_row_group,
                                           /*LBO: K groups*/ w_row_group, /*SBO: MN groups*/ 128),
                             idesc, ks > 0);
        } else {
            // A: X tile (128 rows, 64 feats) MN-major -> M = 64.  B: Y tile (128 rows, N feats)
            const uint32_t idesc = tc::instr_desc_bf16(64, n, true, true);
            const uint32_t x_row_group = (64 / 8) * 128;
            const uint32_t y_row_group = (n / 8) * 128;
            for (int ks = 0; ks < 128 / 16; ++ks)
                tc::mma_bf16(tmem,
                             tc::smem_desc(tc::smem_u32(xt) + ks * 2 * x_row_group, x_row_group, 128),
                             tc::smem_desc(tc::smem_u32(wt) + ks * 2 * y_row_group, y_row_group, 128),
                             idesc, ks > 0);
        }
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, 0);
    tc::tc_fence_after_sync();
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < n; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tl + c0, v);
        if (mode == 2) {
            // M = 64: row m lives in lane (m % 16) + 32 * (m / 16)
            if (lane < 16) {
                const int row = warp * 16 + lane;
                for (int j = 0; j < 16; ++j) d[row * n + c0 + j] = v[j];
            }
        } else {
            for (int j = 0; j < 16; ++j) d[tid * n + c0 + j] = v[j];
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace den

extern "C" int den_tc_probe_gemm(int mode, const float* x, const float* w, float* d, int n, int k,
                                 void* stream) {
    using namespace den;
    DEN_CHECK_ARG(mode >= 0 && mode <= 3, "mode must be 0 .. 3");
    DEN_CHECK_ARG(n % 16 == 0 && n >= 16 && n <= 64, "N must be 16..64, multiple of 16");
    DEN_CHECK_ARG(mode == 2 || (k % 16 == 0 && k >= 16 && k <= 64), "K must be 16..64");
    DEN_CHECK_ARG(x && w && d, "null pointer");
    tc_probe_kernel<<<1, 128, 32768 + 64, as_stream(stream)>>>(mode, x, w, d, n, k);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}
