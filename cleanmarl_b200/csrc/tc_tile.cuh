// Shared pieces of the tcgen05 / TMEM kernels (tc_chain.cu: MLP chains, tc_gru.cu: recurrent actor): tile constants,
// operand-image offsets, TMEM stores, the compute <-> issuer hand-off and the warp reduce-scatter.
#pragma once

#include "common.cuh"
#include "tc_ptx.cuh"

namespace tctile {

constexpr int M = 128;                        // samples per tile = UMMA M = TMEM lanes
constexpr int LBO_K = 128;                    // feature-major (weights): next chunk of 4 K elements
constexpr int LBO_S = 144;                    // sample-major: next chunk of 4 samples (128 B + 16 B pad: conflict-free STS.32)
constexpr int SBO_S = (M / 4) * LBO_S;        // 4608: next group of 8 feature rows
constexpr int KSTEP_S = 2 * LBO_S;            // one MMA consumes 8 samples
// ------------------------------------------------------------------------------------------------
// TMEM helpers on 16-column chunks
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
                 : "memory");
}

// byte offset of element (feature row r, sample s) in a sample-major image
__device__ __forceinline__ int smaj(int r, int s) { return (r >> 3) * SBO_S + (r & 7) * 16 + (s >> 2) * LBO_S + (s & 3) * 4; }
// byte offset of element (row n, k) in a K-major weight image with KTOT columns
__device__ __forceinline__ int kmaj(int n, int k, int ktot) { return (n >> 3) * (ktot / 4) * LBO_K + (n & 7) * 16 + (k >> 2) * LBO_K + (k & 3) * 4; }

// u / d and u % d for u < 2^24 from a precomputed 1.0f / d (the estimate is within one of the quotient; the compiler's
// own 32-bit division is ~30 instructions, and every tile decodes its index twice)
__device__ __forceinline__ void fast_divmod(int u, int d, float inv, int& q, int& r) {
    q = __float2int_rz(__int2float_rz(u) * inv);
    r = u - q * d;
    if (r < 0) { --q; r += d; }
    else if (r >= d) { ++q; r -= d; }
}

__device__ __forceinline__ void mbar_wait_trap(uint64_t* bar, uint32_t parity) {
    // bounded: a descriptor / protocol bug must surface as a launch failure, never as a hung GPU
    if (!tc::mbar_wait_bounded(bar, parity, 1u << 26)) __trap();
}

constexpr int NCOMP = 256;                    // compute threads
constexpr int NTHREADS = NCOMP + 32;          // + the issue warp
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// operands written by this thread (TMEM stores and/or generic-proxy shared stores) -> visible to the issuer's MMAs
__device__ __forceinline__ void publish(uint64_t* bar) {
    tc::tmem_wait_st();
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    tc::mbar_arrive(bar);
}
__device__ __forceinline__ void acquire(uint64_t* bar, uint32_t parity) {
    mbar_wait_trap(bar, parity);
    tc::tcgen05_fence_after();
}

// Sum over the 32 lanes of NV per-lane values: butterfly reduce-scatter while more than one value is left
// (N/2 shuffles per step), plain butterfly afterwards.  v[0] ends up as the total of original index `idx`
// (every lane a different one when NV == 32; for NV == 16 lanes 2i and 2i+1 hold the same index).
// Template recursion keeps every array index a compile-time constant (registers, no local memory).
template <int N, int W, int NV>
__device__ __forceinline__ void rs_step(float (&v)[NV], int lane, int& idx) {
    const bool upper = (lane & W) != 0;
    if constexpr (N > 1) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float send = upper ? v[i] : v[i + N / 2];
            const float keep = upper ? v[i + N / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, W);
        }
        if (upper) idx += N / 2;
        if constexpr (W > 1) rs_step<N / 2, W / 2, NV>(v, lane, idx);
    } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], W);
        if constexpr (W > 1) rs_step<1, W / 2, NV>(v, lane, idx);
    }
}
template <int NV>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[NV], int lane, int& idx_out) {
    static_assert(NV == 16 || NV == 32, "16 or 32 values per lane");
    int idx = 0;
    rs_step<NV, 16, NV>(v, lane, idx);
    idx_out = idx;
}

}  // namespace tctile
