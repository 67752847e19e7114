// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) used by the tensor-core MLP
// kernels.  Operand tiles live in shared memory in the canonical UMMA "no swizzle" layout:
// core matrix = 8 rows x 16 bytes, contiguous (128 B); for a K-major operand of K elements
// (bf16: 8 per 16 B chunk) element (r, k) sits at
//     (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2        [bytes]
// with LBO = 128 (adjacent K chunks are adjacent core matrices) and SBO = (K / 8) * 128.
// The SAME bytes read with the MN-major flag describe the transposed operand (MN = k,
// K = r) with the roles of LBO and SBO exchanged, which is how the backward GEMMs
// (dX = dY W, dW = dY^T X) reuse the tiles the forward wrote.
#pragma once
#include <cuda_bf16.h>

#include "den_common.cuh"

namespace den {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Bounded wait: a barrier that never completes traps instead of hanging the GPU.  The poll loop
// is kept ROLLED (nvcc unrolled it into ~50 try_wait copies per call site: 190 KB of SASS for the
// backward kernel, instruction-cache misses on every epilogue phase) and the time-out report is
// out of line.
constexpr uint32_t kMbarSuspendNs = 20000;      // upper bound of one parked wait (the MMA rounds take ~1 us)
static __device__ __noinline__ void mbar_timeout() {
    printf("den_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
    __trap();
}
// The wait blocks INSIDE the instruction: `suspendTimeHint` lets the hardware park the thread until the
// phase completes (or the hint expires) instead of returning after its short default period.  ncu on the
// round-1 kernels showed the poll loop itself (try_wait, predicate, branch, spin counter, convergence
// barrier) issuing ~25 % of all warp instructions of den_mlp_bwd while the other tiles' warps competed
// for the same issue slots; with the hint a wait is a handful of instructions.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t spin = 0;
#pragma unroll 1
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(kMbarSuspendNs)
            : "memory");
        if (done) return;
        if (++spin == (1u << 22)) mbar_timeout();
    }
}

// ---- proxies / fences -----------------------------------------------------------
__device__ __forceinline__ void fence_smem_to_async_proxy() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM allocation (one warp) ---------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t n_cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(n_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t n_cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols)
                 : "memory");
}

// ---- descriptors -------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor for kind::f16 with bf16 operands and fp32 accumulation
__host__ __device__ constexpr uint32_t instr_desc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) |
           ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem].  Called by ALL lanes of the (converged) MMA warp: the
// descriptor arithmetic stays warp-uniform (uniform datapath, no R2UR per operand) and one
// elected lane issues the instruction.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p, leader;\n\t"
        "elect.sync _|leader, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@leader tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// make the mbarrier track completion of all MMAs issued so far (same elected lane)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred leader;\n\t"
        "elect.sync _|leader, 0xffffffff;\n\t"
        "@leader tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        :
        : "r"(smem_u32(bar))
        : "memory");
}

// Single-thread issue path: the MMA warp elects ONE lane per GEMM round (elect_one), which then
// issues every tcgen05.mma of the round and the commit from straight-line code.  Descriptor
// arithmetic lands on the uniform datapath (5 SASS instructions per MMA; the per-MMA elect.sync
// form above costs 14).
__device__ __forceinline__ bool elect_one() {
    uint32_t leader;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(leader));
    return leader != 0;
}
__device__ __forceinline__ void mma_commit_1t(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// ---- TMEM -> registers (warp-collective; thread i of the warp gets lane base+i) -------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]) {
    static_assert(N % 16 == 0, "N must be a multiple of 16");
#pragma unroll
    for (int c = 0; c < N / 16; ++c) {
        float t[16];
        tmem_ld16(taddr + 16 * c, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[16 * c + i] = t[i];
    }
}

// ---- operand tiles --------------------------------------------------------------------
// byte offset of the 16-byte chunk holding elements (r, 8*chunk .. 8*chunk+7) of a K-major
// bf16 operand with `k_elems` columns
__device__ __forceinline__ uint32_t chunk_offset(int r, int chunk, int k_elems) {
    return (uint32_t)((r >> 3) * (k_elems >> 3) * 128 + chunk * 128 + (r & 7) * 16);
}

// split 8 fp32 values into bf16 hi / lo chunks (x ~= hi + lo to ~2^-17 relative)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(x[2 * i + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
        h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// cooperative load of an nn.Linear weight (n_out, n_in) fp32 row-major from global into hi / lo
// bf16 operand tiles of logical shape (n_pad, k_pad), zero padded
__device__ __forceinline__ void load_weight_split(uint8_t* tile_hi, uint8_t* tile_lo,
                                                  const float* __restrict__ w, int n_out, int n_in,
                                                  int n_pad, int k_pad) {
    for (int i = threadIdx.x; i < n_pad * k_pad; i += blockDim.x) {
        const int n = i / k_pad, k = i - n * k_pad;
        const float v = (n < n_out && k < n_in) ? __ldg(w + n * n_in + k) : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        const uint32_t off = chunk_offset(n, k >> 3, k_pad) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(tile_hi + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(tile_lo + off) = l;
    }
}

}  // namespace tc
}  // namespace den
