// Diagnostic: issues tcgen05.mma.kind::tf32 on a caller-supplied shared-memory image with
// caller-supplied matrix / instruction descriptors and dumps the TMEM accumulator.  Built as its own
// library (libcmarl_umma_probe.so); profiles/tools/umma_explore.py uses it to pin the descriptor and
// canonical-layout conventions that tc_chain.cu relies on (K-major / MN-major, swizzle modes, M = 64
// vs 128 accumulator layouts) against a host GEMM.  Not part of the training path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tc_ptx.cuh"

struct ProbeArgs {
    const uint8_t* image;      // global: bytes copied to dynamic shared memory offset 0
    uint32_t image_bytes;
    uint32_t a_off, b_off;     // byte offsets of the operands inside the image (16-B aligned)
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;   // bytes
    uint32_t a_layout, b_layout;           // UMMA layout_type (0 none, 1 128B_base32B, 2 128B, 4 64B, 6 32B)
    uint32_t idesc;            // instruction descriptor
    uint32_t ksteps;           // MMAs issued; accumulate = (k > 0)
    uint32_t a_kstep, b_kstep; // bytes added to the start address per MMA
    uint32_t n_cols;           // accumulator columns to dump (<= 256)
    uint32_t passes;           // > 1: a second/third pass with other operand offsets (3xTF32 check)
    uint32_t a_off2, b_off2, a_off3, b_off3;
    float* out;                // [128][n_cols]
    int* status;               // 0 ok, 1 timed out waiting for the MMA
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 16; i < p.image_bytes; i += 128 * 16)
        *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(p.image + i);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    tc::fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = tmem_base_s;
    // sentinel fill so untouched lanes / columns are recognisable
    for (uint32_t c = 0; c < p.n_cols; ++c)
        tc::tmem_st1(tmem + ((uint32_t)(warp * 32) << 16) + c, __float_as_uint(-12345.0f));
    // a_layout == 99: A comes from TMEM.  The image holds A row-major [128][ksteps*8] at a_off (pass 2/3: a_off2/3);
    // thread m stores row m to columns [128 + pass*64, ...): one 32-bit element per column.
    const bool a_tmem = p.a_layout == 99;
    if (a_tmem) {
        const uint32_t kt = p.ksteps * 8;
        for (uint32_t pass = 0; pass < (p.passes ? p.passes : 1); ++pass) {
            const uint32_t ao = pass == 0 ? p.a_off : (pass == 1 ? p.a_off2 : p.a_off3);
            const uint32_t* row = reinterpret_cast<const uint32_t*>(smem + ao) + (size_t)tid * kt;
            for (uint32_t k = 0; k < kt; ++k)
                tc::tmem_st1(tmem + ((uint32_t)(warp * 32) << 16) + 128 + pass * 64 + k, row[k]);
        }
    }
    tc::tmem_wait_st();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    if (tid == 0) {
        const uint32_t base = tc::smem_u32(smem);
        uint32_t acc = 0;
        for (uint32_t pass = 0; pass < (p.passes ? p.passes : 1); ++pass) {
            const uint32_t ao = pass == 0 ? p.a_off : (pass == 1 ? p.a_off2 : p.a_off3);
            const uint32_t bo = pass == 0 ? p.b_off : (pass == 1 ? p.b_off2 : p.b_off3);
            for (uint32_t k = 0; k < p.ksteps; ++k) {
                const uint64_t db = tc::make_smem_desc(base + bo + k * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout);
                if (a_tmem) {
                    tc::mma_tf32_ts(tmem, tmem + 128 + pass * 64 + k * 8, db, p.idesc, acc);
                } else {
                    const uint64_t da = tc::make_smem_desc(base + ao + k * p.a_kstep, p.a_lbo, p.a_sbo, p.a_layout);
                    tc::mma_tf32(tmem, da, db, p.idesc, acc);
                }
                acc = 1;
            }
        }
        tc::mma_commit(&bar);
    }
    int ok = tc::mbar_wait_bounded(&bar, 0, 1u << 24);
    tc::tcgen05_fence_after();
    if (ok) {
        for (uint32_t c = 0; c < p.n_cols; ++c) {
            const uint32_t v = tc::tmem_ld1(tmem + ((uint32_t)(warp * 32) << 16) + c);
            p.out[(size_t)tid * p.n_cols + c] = __uint_as_float(v);
        }
    }
    if (tid == 0) *p.status = ok ? 0 : 1;
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

extern "C" int cmarl_umma_probe(const ProbeArgs* args, uint32_t smem_bytes, void* stream) {
    cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return (int)e;
    umma_probe_kernel<<<1, 128, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(*args);
    return (int)cudaGetLastError();
}

// ---- micro-benchmark: cycles for `count` back-to-back MMAs rotating over `nacc` accumulators -----------------
struct BenchArgs {
    uint32_t m, n, a_tmem, count, nacc, lbo, sbo, kstep;
    long long* out;     // [2]: cycles from first issue to completion seen, cycles spent issuing
};

__global__ void __launch_bounds__(128, 1) umma_bench_kernel(BenchArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid * 4; i < 96 * 1024; i += 128 * 4) *reinterpret_cast<float*>(smem + i) = 1.0f;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async_smem();
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = tmem_base_s;
    for (uint32_t c = 0; c < 512; ++c) tc::tmem_st1(tmem + ((uint32_t)(warp * 32) << 16) + c, 0u);
    tc::tmem_wait_st();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    if (warp == 0) {     // warp-uniform issue loop, one elected lane executes the MMA itself
        const uint32_t base = tc::smem_u32(smem);
        const uint32_t idesc = tc::make_idesc_tf32(p.m, p.n, 0, 0);
        const uint64_t db0 = tc::make_smem_desc(base + 48 * 1024, p.lbo, p.sbo, 0);
        const uint64_t da0 = tc::make_smem_desc(base, p.lbo, p.sbo, 0);
        const uint32_t kq = p.kstep >> 4;
        const long long t0 = clock64();
        uint32_t acc_i = 0;
        for (uint32_t i = 0; i < p.count; i += 8) {
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t d = tmem + acc_i * p.n;
                if (p.a_tmem) { if (tc::elect_one()) tc::mma_tf32_ts(d, tmem + 384 + k * 8, db0 + k * kq, idesc, 1); }
                else { if (tc::elect_one()) tc::mma_tf32(d, da0 + k * kq, db0 + k * kq, idesc, 1); }
                acc_i = acc_i + 1 == p.nacc ? 0 : acc_i + 1;
            }
        }
        const long long t1 = clock64();
        if (tc::elect_one()) tc::mma_commit(&bar);
        tc::mbar_wait_bounded(&bar, 0, 1u << 24);
        const long long t2 = clock64();
        if (tid == 0) { p.out[0] = t2 - t0; p.out[1] = t1 - t0; }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

extern "C" int cmarl_umma_bench(const BenchArgs* args, void* stream) {
    cudaError_t e = cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return (int)e;
    umma_bench_kernel<<<1, 128, 96 * 1024, reinterpret_cast<cudaStream_t>(stream)>>>(*args);
    return (int)cudaGetLastError();
}
