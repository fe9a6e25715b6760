// Recurrent-actor path on the 5th-gen tensor cores (tcgen05 / TMEM): one truncated-BPTT chunk of
// cleanmarl/mappo_lstm_multienvs.py:551-620 ("LSTM" below) as TWO kernels that meet in global memory (h_seq + gate stash,
// the formats of gru.cu, so the backward half can be exchanged for the fp32 FFMA kernel of gru.cu -- CMARL_TBPTT=tcfwd):
//
//   tc_gru_fwd_kernel   t = t0 .. t1-1: x1 = relu(W1 x + b1), GRUCell, h_{t+1} -> h_seq, (x1, r, z, n, Whn h + bhn) -> stash;
//                       the head of every step (loss terms, statistics, dlogits -> workspace, dW2 / db2)
//   tc_gru_bwd_kernel   t = t1-1 .. t0: gate gradients from the dlogits, dx1 / dh through the recurrence, weight gradients
//                       accumulated in TMEM -> one partial row per CTA (+ the forward CTAs' dW2 / db2 / statistics rows)
//
// Both: one CTA = 8 compute warps + 1 MMA-issue warp (backward: + 3 idle warps that complete the issue warp's warpgroup, see
// there), one tile = 128 consecutive envs of one agent = 128 TMEM lanes;
// compute thread (q = warp & 3, hf = warp >> 2, lane) owns sample s = 32 q + lane and hidden units 16 hf .. 16 hf + 15.
// Precision: kind::tf32 with the 3-term split of tc_chain.cu (x = hi + lo; lo*hi, hi*lo, hi*hi; fp32 accumulation).
//
// Forward (2 CTAs / SM; 256 TMEM columns each): the activations x1 | h are the A operand in TMEM, the gate pre-activations
// r | z | Whn h | Win x1 one 128-column accumulator; the recurrent half of the gate GEMM (h Whh^T) is issued as soon as h_t is
// published, fc1 (X W1^T, X as a K-major shared-memory image) right behind it, the input half (x1 Wih^T) when x1 is.
// Backward (1 CTA / SM; 496 TMEM columns): the four gate gradients da = [da_r | da_z | da_n | da_hn] are the A operand of
// ONE GEMM against [Wih | Whh] (dx1 | dh); the weight gradients contract over the samples and take sample-major
// shared-memory images like tc_chain.cu's: A = two gates' hi images stacked on their lo images (M = 128), B = [x1 | h | 1],
// two rounds (r, z) and (n, hn) per step plus one for dW1 = dx1^T [x | 1]; their accumulators stay in TMEM for the steps of a
// tile and are then added into the CTA's partial row.
#include <stdlib.h>

#include "chain.cuh"
#include "heads.cuh"
#include "tc_ptx.cuh"
#include "tc_tile.cuh"

namespace tcgru {

using namespace chain;
using namespace tctile;

constexpr int H = 32, G3 = 96, NA = 5;
constexpr int K1P = 24;                       // padded input rows
constexpr int NCX = K1P / 8;                  // 8-row chunks of X; a thread owns chunks c with (c & 1) == hf
constexpr int NXO = 2;

using Args = GruChunkArgs;

// sigmoid / tanh on the special-function unit (ex2.approx + rcp.approx: ~2 ulp each; |error| <= 2e-7 on outputs in (0, 1) /
// (-1, 1)): the libm forms (expf, tanhf, IEEE division: ~100 instructions per hidden unit) made the forward epilogue
// issue-bound -- 8.8 k cycles per step with two co-resident CTAs.
__device__ __forceinline__ float sigmoid_sfu(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_sfu(float x) {
    const float e = __expf(-2.0f * fabsf(x));
    return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}

// D[128 x N] (+)= A(TMEM, K columns at a_hi / a_lo) * B(K-major image [N][K])^T, 3xTF32; `keep`: accumulate from the first MMA on
template <int N, int K>
__device__ __forceinline__ void issue_ts(bool leader, uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                         uint32_t keep) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
    constexpr uint32_t sbo = (K / 4) * LBO_K;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        uint32_t a = pass == 0 ? a_lo : a_hi;
        uint64_t db = tc::make_smem_desc(pass == 1 ? b_lo : b_hi, LBO_K, sbo, 0);
#pragma unroll 2
        for (int ks = 0; ks < K / 8; ++ks) {
            if (leader) tc::mma_tf32_ts(d, a, db, idesc, keep | (uint32_t)(pass | ks));
            a += 8;
            db += (uint64_t)((2 * LBO_K) >> 4);
        }
    }
}
// same with A a K-major shared-memory image [128][K] (fc1: the input rows)
template <int N, int K>
__device__ __forceinline__ void issue_ss_k(bool leader, uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
    constexpr uint32_t sbo = (K / 4) * LBO_K;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        uint64_t da = tc::make_smem_desc(pass == 0 ? a_lo : a_hi, LBO_K, sbo, 0);
        uint64_t db = tc::make_smem_desc(pass == 1 ? b_lo : b_hi, LBO_K, sbo, 0);
#pragma unroll 1
        for (int ks = 0; ks < K / 8; ++ks) {
            if (leader) tc::mma_tf32(d, da, db, idesc, (uint32_t)(pass | ks));
            da += (uint64_t)((2 * LBO_K) >> 4);
            db += (uint64_t)((2 * LBO_K) >> 4);
        }
    }
}
// D[MROWS x N] (+)= [A_hi ; A_lo](sample-major image, MROWS rows) * B(sample-major image [N][128])^T over the 128 samples
// (tc_chain.cu's weight-gradient form: pass 0 by B_lo, pass 1 by B_hi; the hi / lo row blocks are added at the flush)
template <int MROWS, int N>
__device__ __forceinline__ void issue_ss(bool leader, uint32_t d, uint32_t a, uint32_t b_hi, uint32_t b_lo, uint32_t keep) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(MROWS, N, 0, 0);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        uint64_t da = tc::make_smem_desc(a, LBO_S, SBO_S, 0);
        uint64_t db = tc::make_smem_desc(pass == 0 ? b_lo : b_hi, LBO_S, SBO_S, 0);
#pragma unroll 2
        for (int ks = 0; ks < M / 8; ++ks) {
            if (leader) tc::mma_tf32(d, da, db, idesc, keep | (uint32_t)(pass | ks));
            da += (uint64_t)(KSTEP_S >> 4);
            db += (uint64_t)(KSTEP_S >> 4);
        }
    }
}

// Row r of a thread's column: one CTA-uniform 64-bit base + a 32-bit per-thread offset (an element index < 2^32), so that
// a batch of row loads keeps ONE address register alive instead of a 64-bit pair per row.
__device__ __forceinline__ float ld_row(const float* base_u, uint32_t off, bool ok) { return ok ? __ldcg(base_u + off) : 0.0f; }

__device__ __forceinline__ void st_split(uint8_t* hi_img, uint8_t* lo_img, int off, float v) {
    float h, l;
    tc::split_tf32(v, h, l);
    *reinterpret_cast<float*>(hi_img + off) = h;
    *reinterpret_cast<float*>(lo_img + off) = l;
}

// debug timeline of CTA 0 (compute thread 0), second step of its first tile
__device__ long long g_tcgru_tl[32];
__device__ int g_tcgru_tl_on = 0;
#define BTL(slot) do { if (tl_on) g_tcgru_tl[slot] = clock64(); } while (0)

// ================================================================================================
// Forward
// ================================================================================================
namespace fwd {
// shared memory (bytes)
constexpr int oBar = 0;                         // 5 mbarriers + tmem base
constexpr int szW1 = H * K1P * 4;               // 3 072
constexpr int oW1 = 128;                        // W1 hi | lo   [32][24]   K-major image (B of fc1)
constexpr int szWg = G3 * H * 4;                // 12 288
constexpr int oWhh = oW1 + 2 * szW1;            // Whh hi | lo  [96][32]   rows r, z, n
constexpr int oWih = oWhh + 2 * szWg;           // Wih hi | lo  [96][32]
constexpr int oB1 = oWih + 2 * szWg;            // f32 [4][32]  b1 (+ folded id column) per agent
constexpr int oBg = oB1 + 4 * H * 4;            // f32 [32][4]  bir + bhr, biz + bhz, bin, bhn
constexpr int oW2T = oBg + H * 4 * 4;           // f32 [32][8]  output layer, unit-major
constexpr int oB2 = oW2T + H * 8 * 4;           // f32 [8]
constexpr int oDW2 = oB2 + 32;                  // f32 [4 quadrants][8][32] + [4][8]: dW2 / db2 partial sums
constexpr int oRed = oDW2 + (4 * 8 * H + 32) * 4;
constexpr int oZx = oRed + 256;                 // f32 [5][128]: partial logits half 1 -> half 0, then dlogits half 0 -> half 1
constexpr int oX = ((oZx + NA * M * 4 + 127) / 128) * 128;     // X hi | lo    [128][24]  K-major image (A of fc1)
constexpr int szX = M * K1P * 4;                // 12 288
constexpr int SMEM = oX + 2 * szX;
static_assert(oX % 128 == 0 && SMEM <= 113 * 1024, "two CTAs per SM");
// TMEM columns
constexpr int cAh = 0, cAl = 2 * H;             // [x1 | h] hi, lo
constexpr int cG = 4 * H;                       // r | z | Whn h | Win x1
constexpr int cD1 = cG + 3 * H;                 // fc1 output shares the Win x1 columns (consumed before they are written)
constexpr int TMEM_COLS = 256;
enum { R_H = 0, R_X, R_X1, D_F1, D_G, N_BARS };
}  // namespace fwd

__global__ void __launch_bounds__(NTHREADS, 2) tc_gru_fwd_kernel(Args a) {
    using namespace fwd;
    extern __shared__ __align__(1024) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + oBar + N_BARS * 8);
    const GruLayout& L = a.L;
    const int tiles_b = (a.B + M - 1) / M;
    const int units = a.N * tiles_b;
    const int nsteps = a.t1 - a.t0;

    if (tid == 0) {
        for (int i = 0; i < N_BARS; ++i) tc::mbar_init(&bars[i], i < D_F1 ? NCOMP : 1);
        tc::fence_mbar_init();
    }
    if (warp == 8) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    for (int i = tid; i < 4 * 8 * H + 32; i += NTHREADS) reinterpret_cast<float*>(sm + oDW2)[i] = 0.0f;
    pdl_wait_then_trigger();
    // ---- weights -> K-major hi | lo images ---------------------------------------------------------
    {
        const float* P = a.params;
        const int O = L.in;
        CMARL_STRIDED(i, H * 8, NTHREADS) {
            const int j = i / 8, c = i - j * 8;
            reinterpret_cast<float*>(sm + oW2T)[i] = (c < NA) ? __ldcg(P + L.w2 + c * H + j) : 0.0f;
        }
        if (tid < 8) reinterpret_cast<float*>(sm + oB2)[tid] = (tid < NA) ? __ldcg(P + L.b2 + tid) : 0.0f;
        CMARL_STRIDED(i, H * K1P, NTHREADS) {
            const int j = i / K1P, k = i - j * K1P;
            const float w = k < a.in_rows ? __ldcg(P + L.w1 + j * O + k) : 0.0f;
            st_split(sm + oW1, sm + oW1 + szW1, kmaj(j, k, K1P), w);
        }
        CMARL_STRIDED(i, G3 * H, NTHREADS) {
            const int n = i / H, k = i - n * H;
            st_split(sm + oWhh, sm + oWhh + szWg, kmaj(n, k, H), __ldcg(P + L.whh + i));
            st_split(sm + oWih, sm + oWih + szWg, kmaj(n, k, H), __ldcg(P + L.wih + i));
        }
        CMARL_STRIDED(i, 4 * H, NTHREADS) {
            const int g = i / H, j = i - g * H;
            float v = __ldcg(P + L.b1 + j);
            if (a.fold_ids && g < a.N) v += __ldcg(P + L.w1 + j * O + a.in_rows + g);
            reinterpret_cast<float*>(sm + oB1)[i] = v;
        }
        CMARL_STRIDED(j, H, NTHREADS) {
            float4 b;
            b.x = __ldcg(P + L.bih + j) + __ldcg(P + L.bhh + j);
            b.y = __ldcg(P + L.bih + H + j) + __ldcg(P + L.bhh + H + j);
            b.z = __ldcg(P + L.bih + 2 * H + j);
            b.w = __ldcg(P + L.bhh + 2 * H + j);
            reinterpret_cast<float4*>(sm + oBg)[j] = b;
        }
    }
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sbase = tc::smem_u32(sm);

    if (warp == 8) {
        // ================================ MMA issue warp ==================================================
        const bool leader = tc::elect_one();
        uint32_t par = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            for (int i = 0; i < nsteps; ++i, par ^= 1) {
                acquire(&bars[R_H], par);          // h_t in the A columns; every read of the previous step's accumulator is done
                acquire(&bars[R_X], par);          // (staged a step ahead: normally complete long before)
                issue_ss_k<H, K1P>(leader, tmem + cD1, sbase + oX, sbase + oX + szX, sbase + oW1, sbase + oW1 + szW1);
                if (leader) tc::mma_commit(&bars[D_F1]);
                // the recurrent half runs under the fc1 epilogue: r | z | Whn h  = h [Whr; Whz; Whn]^T
                issue_ts<G3, H>(leader, tmem + cG, tmem + cAh + H, tmem + cAl + H, sbase + oWhh, sbase + oWhh + szWg, 0u);
                acquire(&bars[R_X1], par);         // x1 in the A columns, fc1's output consumed
                issue_ts<2 * H, H>(leader, tmem + cG, tmem + cAh, tmem + cAl, sbase + oWih, sbase + oWih + szWg, 1u);
                issue_ts<H, H>(leader, tmem + cG + 3 * H, tmem + cAh, tmem + cAl, sbase + oWih + kmaj(2 * H, 0, H),
                               sbase + oWih + szWg + kmaj(2 * H, 0, H), 0u);
                if (leader) tc::mma_commit(&bars[D_G]);
            }
        }
        __syncwarp();
    } else {
        // ================================ compute warps ====================================================
        const int q = warp & 3, hf = warp >> 2;
        const int s = q * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        const float* fb1 = reinterpret_cast<const float*>(sm + oB1);
        const float4* fbg = reinterpret_cast<const float4*>(sm + oBg);
        const int xoff = (s >> 3) * (K1P / 4) * LBO_K + (s & 7) * 16;     // this sample's row in the K-major X image
        const float* fw2 = reinterpret_cast<const float*>(sm + oW2T);
        const float* fb2 = reinterpret_cast<const float*>(sm + oB2);
        float* zx = reinterpret_cast<float*>(sm + oZx);
        float* dw2acc = reinterpret_cast<float*>(sm + oDW2) + q * 8 * H;
        float* db2acc = reinterpret_cast<float*>(sm + oDW2) + 4 * 8 * H + q * 8;
        const int ct = warp * 32 + lane;
        float st[PolicyHeadT<true>::NSTAT];
#pragma unroll
        for (int k = 0; k < PolicyHeadT<true>::NSTAT; ++k) st[k] = 0.0f;

        float xr[NXO * 8];
        auto load_x = [&](int t, int g, int b) {
            const float* xp = a.x + (size_t)t * a.stride_t + (size_t)g * a.stride_g + b + (size_t)(8 * hf) * a.B;
            const uint32_t step = (uint32_t)a.B;
            const int rmax = (b < a.B) ? a.in_rows - 8 * hf : 0;
#pragma unroll
            for (int i = 0; i < NXO; ++i)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int rr = 16 * i + e;
                    xr[i * 8 + e] = (rr < rmax) ? __ldcg(xp + (uint32_t)rr * step) : 0.0f;
                }
        };
        // the input rows of a later step -> L2, one row per lane (a warp's 32 samples of a row are one 128-B line)
        auto prefetch_x = [&](int t, int g, int b) {
            const int bw = b - lane;
            if (lane < a.in_rows && bw < a.B)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x + (size_t)t * a.stride_t + (size_t)g * a.stride_g + (size_t)lane * a.B + bw));
        };
        auto stage_x = [&]() {
#pragma unroll
            for (int i = 0; i < NXO; ++i) {
                const int c = 2 * i + hf;
                if (c < NCX) {
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        float4 hi, lo;
                        tc::split_tf32(xr[i * 8 + 4 * m + 0], hi.x, lo.x);
                        tc::split_tf32(xr[i * 8 + 4 * m + 1], hi.y, lo.y);
                        tc::split_tf32(xr[i * 8 + 4 * m + 2], hi.z, lo.z);
                        tc::split_tf32(xr[i * 8 + 4 * m + 3], hi.w, lo.w);
                        const int o = xoff + (2 * c + m) * LBO_K;
                        *reinterpret_cast<float4*>(sm + oX + o) = hi;
                        *reinterpret_cast<float4*>(sm + oX + szX + o) = lo;
                    }
                }
            }
            publish(&bars[R_X]);
        };

        uint32_t par = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int g = u / tiles_b, b = (u - g * tiles_b) * M + s;
            const bool inb = b < a.B;
            load_x(a.t0, g, b);
            float hp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                hp[i] = 0.0f;       // zeros at the start of an epoch (LSTM:558), else h_seq[t0]
                if (a.t0 > 0 && inb) hp[i] = __ldcg(a.h_seq + (((size_t)a.t0 * a.N + g) * H + 16 * hf) * a.B + b + (uint32_t)i * (uint32_t)a.B);
            }
            {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float h, l;
                    tc::split_tf32(hp[i], h, l);
                    hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(l);
                }
                tmem_st16(tl + cAh + H + 16 * hf, hi);
                tmem_st16(tl + cAl + H + 16 * hf, lo);
                publish(&bars[R_H]);
            }
            stage_x();
            if (nsteps > 1) prefetch_x(a.t0 + 1, g, b);

            for (int i = 0; i < nsteps; ++i, par ^= 1) {
                const int t = a.t0 + i;
                const bool tl_on = g_tcgru_tl_on && blockIdx.x == 0 && u == (int)blockIdx.x && i == 1 && tid == 0;
                BTL(16);
                // one 64-bit base per step and 32-bit row offsets (the obvious 64-bit products are ~10 instructions per access)
                const size_t rs = (size_t)a.B;
                // running pointers into this thread's rows of the stash (x1, r, z, n, Whn h) and of h_seq[t + 1]
                float* q0 = a.stash + ((size_t)t * a.N + g) * (5 * H) * a.B + (size_t)(16 * hf) * rs + b;
                float* q1 = q0 + (size_t)H * rs;
                float* q2 = q1 + (size_t)H * rs;
                float* q3 = q2 + (size_t)H * rs;
                float* q4 = q3 + (size_t)H * rs;
                float* q5 = a.h_seq + (((size_t)(t + 1) * a.N + g) * H) * a.B + (size_t)(16 * hf) * rs + b;
                // ---- x1 = relu(fc1 + b1[g]) -> stash, split -> A columns -----------------------------------
                float bv1[16];
#pragma unroll
                for (int i4 = 0; i4 < 16; i4 += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(fb1 + g * H + 16 * hf + i4);
                    bv1[i4] = bb.x; bv1[i4 + 1] = bb.y; bv1[i4 + 2] = bb.z; bv1[i4 + 3] = bb.w;
                }
                acquire(&bars[D_F1], par);
                BTL(17);
                {
                    uint32_t v[16], hi[16], lo[16];
                    tc::tmem_ld16(tl + cD1 + 16 * hf, v);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float x1 = fmaxf(__uint_as_float(v[k]) + bv1[k], 0.0f);
                        if (inb) *q0 = x1;
                        q0 += rs;
                        float h, l;
                        tc::split_tf32(x1, h, l);
                        hi[k] = __float_as_uint(h); lo[k] = __float_as_uint(l);
                    }
                    tmem_st16(tl + cAh + 16 * hf, hi);
                    tmem_st16(tl + cAl + 16 * hf, lo);
                }
                publish(&bars[R_X1]);
                BTL(18);
                // fc1 of this step has completed: the X image is free for the next step's rows
                if (i + 1 < nsteps) {
                    // (the rows are L2 hits, prefetched a step ago, requested only now: held in registers across a step -- or
                    // just across the fc1 epilogue -- they were parked on the stack as they arrived; the wait falls under
                    // the gate MMAs that the hand-off above released)
                    load_x(t + 1, g, b);
                    stage_x();
                    if (i + 2 < nsteps) prefetch_x(t + 2, g, b);
                }
                // ---- gates -> h_{t+1} -----------------------------------------------------------------------
                BTL(19);
                acquire(&bars[D_G], par);
                BTL(20);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    uint32_t vr[8], vz[8], vh[8], vi[8];
                    const uint32_t c0 = tl + cG + 16 * hf + 8 * k;
                    tc::tmem_ld8(c0, vr);
                    tc::tmem_ld8(c0 + H, vz);
                    tc::tmem_ld8(c0 + 2 * H, vh);
                    tc::tmem_ld8(c0 + 3 * H, vi);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int j = 16 * hf + 8 * k + e;
                        const float4 bb = fbg[j];
                        const float ar = __uint_as_float(vr[e]) + bb.x, az = __uint_as_float(vz[e]) + bb.y;
                        const float ai = __uint_as_float(vi[e]) + bb.z, ah = __uint_as_float(vh[e]) + bb.w;
                        const float r = sigmoid_sfu(ar), z = sigmoid_sfu(az);
                        const float n = tanh_sfu(fmaf(r, ah, ai));
                        const float hn = fmaf(hp[8 * k + e] - n, z, n);
                        hp[8 * k + e] = hn;
                        if (inb) { *q1 = r; *q2 = z; *q3 = n; *q4 = ah; *q5 = hn; }
                        q1 += rs; q2 += rs; q3 += rs; q4 += rs; q5 += rs;
                    }
                }
                if (i + 1 < nsteps) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        float h, l;
                        tc::split_tf32(hp[k], h, l);
                        hi[k] = __float_as_uint(h); lo[k] = __float_as_uint(l);
                    }
                    tmem_st16(tl + cAh + H + 16 * hf, hi);
                    tmem_st16(tl + cAl + H + 16 * hf, lo);
                    publish(&bars[R_H]);
                }
                BTL(21);
                // ---- the head of this step (LSTM:574-593, 628-638), behind the hand-off of h_{t+1} so that the next step's
                //      MMAs run under it: logits = W2 relu(h') + b2 -> loss terms, statistics, dlogits (-> global, for the
                //      backward kernel), dW2 / db2 ------------------------------------------------------------------------
                {
                    const PolicyHeadT<true>::In hin = PolicyHeadT<true>::load(a.head, t, g, b, a.N, a.B, inb && hf == 0);
                    float rh[16], z[NA], dz[NA];
#pragma unroll
                    for (int c = 0; c < NA; ++c) z[c] = 0.0f;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int j = 16 * hf + k;
                        rh[k] = fmaxf(hp[k], 0.0f);
                        const float4 w = *reinterpret_cast<const float4*>(fw2 + j * 8);
                        const float w4 = fw2[j * 8 + 4];
                        z[0] = fmaf(w.x, rh[k], z[0]); z[1] = fmaf(w.y, rh[k], z[1]); z[2] = fmaf(w.z, rh[k], z[2]);
                        z[3] = fmaf(w.w, rh[k], z[3]); z[4] = fmaf(w4, rh[k], z[4]);
                    }
                    if (hf == 1) {
#pragma unroll
                        for (int c = 0; c < NA; ++c) zx[c * M + s] = z[c];
                    }
                    compute_bar();
                    if (hf == 0) {
#pragma unroll
                        for (int c = 0; c < NA; ++c) z[c] = (z[c] + zx[c * M + s]) + fb2[c];
                        PolicyHeadT<true>::compute(a.head, hin, z, true, dz, st);
                        float* dzp = a.dlogits + ((size_t)t * a.N + g) * 8 * a.B + b;
#pragma unroll
                        for (int c = 0; c < NA; ++c) {
                            zx[c * M + s] = dz[c];
                            if (inb) dzp[(size_t)c * a.B] = dz[c];
                        }
                    }
                    compute_bar();
                    if (hf == 1) {
#pragma unroll
                        for (int c = 0; c < NA; ++c) dz[c] = zx[c * M + s];
                    }
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < NA; ++c) {
                        float p[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) p[k] = dz[c] * rh[k];
                        int idx;
                        warp_reduce_scatter<16>(p, lane, idx);
                        if ((lane & 1) == 0) dw2acc[c * H + 16 * hf + idx] += p[0];
                        if (hf == 0) {
                            float d = dz[c];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                            if (lane == 0) db2acc[c] += d;
                        }
                    }
                }
            }
        }
        // ---- this CTA's row of fwd_partials: dW2 (160) | db2 (5) | statistics (8) ------------------------------------
        {
            compute_bar();
            float* out = a.fwd_partials + (size_t)blockIdx.x * (NA * H + NA + CMARL_N_STATS);
            const float* w2all = reinterpret_cast<const float*>(sm + oDW2);
            for (int i = ct; i < NA * H; i += NCOMP)
                out[i] = ((w2all[i] + w2all[8 * H + i]) + w2all[2 * 8 * H + i]) + w2all[3 * 8 * H + i];
            if (ct < NA) {
                const float* b2all = w2all + 4 * 8 * H;
                out[NA * H + ct] = ((b2all[ct] + b2all[8 + ct]) + b2all[16 + ct]) + b2all[24 + ct];
            }
            float* red = reinterpret_cast<float*>(sm + oRed);
#pragma unroll
            for (int k = 0; k < PolicyHeadT<true>::NSTAT; ++k) {
                const float v = warp_sum_f(st[k]);              // half-1 warps carry zeros
                compute_bar();
                if (lane == 0) red[warp] = v;
                compute_bar();
                if (ct == 0) out[NA * H + NA + k] = ((red[0] + red[1]) + red[2]) + red[3];
            }
            if (ct == 0)
                for (int k = PolicyHeadT<true>::NSTAT; k < CMARL_N_STATS; ++k) out[NA * H + NA + k] = 0.0f;
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem, TMEM_COLS);
}

// ================================================================================================
// Backward
// ================================================================================================
namespace bwd {
constexpr int NB = 72;                          // B operand of the gate weight gradients: x1 (32) | h (32) | ones group (8)
constexpr int A_BYTES = 16 * SBO_S;             // 128 feature rows
constexpr int B_BYTES = (NB / 8) * SBO_S;       // per hi / lo image
// shared memory (bytes)
constexpr int oBar = 0;                         // 6 mbarriers + tmem base
constexpr int szWb = 2 * H * 4 * H * 4;         // 32 768
constexpr int oWb = 128;                        // [Wih | Whh] hi | lo: K-major image [64 outputs: dx1 | dh][128: da_r, da_z, da_n, da_hn]
constexpr int oW2T = oWb + 2 * szWb;            // f32 [32][8]
constexpr int oB2 = oW2T + H * 8 * 4;           // f32 [8]
constexpr int oDW2 = oB2 + 32;                  // f32 [4 quadrants][8][32] + [4][8]: dW2 / db2 partial sums
constexpr int oRed = oDW2 + (4 * 8 * H + 32) * 4;
constexpr int oZx = oRed + 256;                 // f32 [5][128]: partial logits half 1 -> half 0, then dlogits half 0 -> half 1
constexpr int oAs = ((oZx + NA * M * 4 + 127) / 128) * 128;     // gate-gradient pair, sample-major: hi rows 0..63 | lo rows 64..127
constexpr int oBs = oAs + A_BYTES;              // x1 | h | ones, sample-major, hi | lo   (round 3: x | ones in rows 0..31)
constexpr int SMEM = oBs + 2 * B_BYTES;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
// TMEM columns
constexpr int cDAh = 0, cDAl = 4 * H;           // da hi | lo: the A operand of the dx1 | dh GEMM
constexpr int cDB = 8 * H;                      // dx1 (before relu') | dh through Whh
constexpr int cWrz = cDB + 2 * H;               // rows (r | z) hi, (r | z) lo  x  [x1 | h | 1]
constexpr int cWn = cWrz + NB;                  // rows (n | hn) hi, (n | hn) lo x  [x1 | h | 1]
constexpr int cW1 = cWn + NB;                   // rows dx1 hi, dx1 lo (M = 64)  x  [x | 1]
constexpr int cEnd = cW1 + H;
static_assert(cEnd <= 512, "TMEM budget");
constexpr int TMEM_COLS = 512;
enum { R_1 = 0, R_2, R_3, D_1, D_2, D_3, N_BARS };
}  // namespace bwd

// Weight-gradient accumulators of one tile -> the CTA's partial row.  Scratch = the gradient vector up to W2 in parameter
// order (the A image, free between two tiles): zero, the lo row blocks, the hi row blocks added to them (fixed order), one
// coalesced pass.  Every TMEM chunk of a thread is requested before the first one is used (chunk-by-chunk round trips made
// the flush 24 k cycles).  NOT inlined: inside the step loop its register demand made the compiler park the prefetched gate
// operands on the stack as they arrived; as a call, live registers are saved around it only when a tile ends.
__device__ __noinline__ void bwd_flush(uint8_t* sm, const Args& a, uint32_t tl, int q, int hf, int lane, int ct, int g,
                                       float* part_out, bool flushed, bool ktl) {
    using namespace bwd;
    const GruLayout& L = a.L;
        if (ktl && !flushed) g_tcgru_tl[12] = clock64();
        float* S = reinterpret_cast<float*>(sm + oAs);
        const int nflush = L.w2;
        // scratch index of a Wih / Whh element: rows of 33 floats (a warp's lanes are 32 different rows of one column: 32
        // floats apart they all hit one bank -- the flush was 31 k cycles); bih / bhh follow, shifted by the padding
        constexpr int GP = H + 1, GSHIFT = 2 * G3 * (GP - H);
        const int oG = L.wih;                            // Wih rows 0..95, Whh rows 96..191
        // (no zeroing pass: every scratch element the last pass reads is written by the lo phase, except the folded id columns
        // of the OTHER agents, which that pass takes as zero)
        const int gsel = q & 1;                      // rows of this quadrant: gate r / n (0) or z / hn (1)
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {
            if ((q >= 2) == (ph == 0)) {             // warp-uniform: lo row blocks first
                uint32_t vz[5][8], vn[3][8], v1[2][8];
                // this thread's chunks: (r | z) rows: 8-column chunks 2 kk + hf < 9; (n | hn) rows: the two chunks of the
                // wanted 32 columns (x1 columns for n rows, h columns for hn rows) + the ones chunk (half 0); dx1 rows: 2 kk + hf
#pragma unroll
                for (int kk = 0; kk < 5; ++kk)
                    if (2 * kk + hf < NB / 8) tc::tmem_ld8(tl + cWrz + 8 * (2 * kk + hf), vz[kk]);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) tc::tmem_ld8(tl + cWn + 8 * (4 * gsel + 2 * kk + hf), vn[kk]);
                if (hf == 0) tc::tmem_ld8(tl + cWn + 64, vn[2]);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) tc::tmem_ld8(tl + cW1 + 8 * (2 * kk + hf), v1[kk]);
                tc::tmem_wait_ld();
                {
                    const int R = gsel * H + lane;   // gate row in Wih / Whh / bih / bhh
#pragma unroll
                    for (int kk = 0; kk < 5; ++kk) {
                        const int c8 = 2 * kk + hf;
                        if (c8 < 8) {
                            float* p = S + oG + ((c8 < 4 ? 0 : G3) + R) * GP + 8 * (c8 & 3);
                            float o8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) o8[e] = ph == 0 ? 0.0f : p[e];     // every load before the first store
#pragma unroll
                            for (int e = 0; e < 8; ++e) p[e] = o8[e] + __uint_as_float(vz[kk][e]);
                        } else if (c8 == 8) {
                            const float x = __uint_as_float(vz[kk][0]);
                            S[L.bih + GSHIFT + R] = ph == 0 ? x : S[L.bih + GSHIFT + R] + x;
                            S[L.bhh + GSHIFT + R] = ph == 0 ? x : S[L.bhh + GSHIFT + R] + x;
                        }
                    }
                }
                {
                    const int R = 2 * H + lane;      // n rows -> Wih_n, bih_n; hn rows -> Whh_n, bhh_n
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        float* p = S + oG + ((gsel == 0 ? 0 : G3) + R) * GP + 8 * (2 * kk + hf);
                        float o8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) o8[e] = ph == 0 ? 0.0f : p[e];
#pragma unroll
                        for (int e = 0; e < 8; ++e) p[e] = o8[e] + __uint_as_float(vn[kk][e]);
                    }
                    if (hf == 0) {
                        const int o = (gsel == 0 ? L.bih : L.bhh) + GSHIFT + R;
                        const float x = __uint_as_float(vn[2][0]);
                        S[o] = ph == 0 ? x : S[o] + x;
                    }
                }
                if (lane < 16) {                     // dx1 rows (M = 64 accumulator: row r sits in lane (r / 16) * 32 + r % 16)
                    const int j = gsel * 16 + lane;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int c8 = 2 * kk + hf;
                        if (c8 < 3) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int k = 8 * c8 + e;
                                if (k < a.in_rows) {
                                    float* p = S + L.w1 + j * L.in + k;
                                    *p = ph == 0 ? __uint_as_float(v1[kk][e]) : *p + __uint_as_float(v1[kk][e]);
                                }
                            }
                        } else {
                            const float x = __uint_as_float(v1[kk][0]);
                            S[L.b1 + j] = ph == 0 ? x : S[L.b1 + j] + x;
                            if (a.fold_ids) {
                                float* p = S + L.w1 + j * L.in + a.in_rows + g;
                                *p = ph == 0 ? x : *p + x;
                            }
                        }
                    }
                }
            }
            compute_bar();
        }
        {
            constexpr int NF = 28;                   // >= ceil(7 040 / 256)
            float v[NF];
#pragma unroll
            for (int r = 0; r < NF; ++r) {
                const int i = ct + NCOMP * r;
                const int g0 = i - oG;
            const int si = i < oG ? i : (i < L.bih ? oG + (g0 >> 5) * GP + (g0 & 31) : i + GSHIFT);
            bool other_id = false;                   // W1 column of another agent's folded one-hot id: no gradient from this tile
            if (i < L.b1) { const int kcol = i % L.in; other_id = kcol >= a.in_rows && kcol != a.in_rows + g; }
            v[r] = (i < nflush && !other_id) ? S[si] : 0.0f;
            }
            if (flushed) {
                float o[NF];
#pragma unroll
                for (int r = 0; r < NF; ++r) {
                    const int i = ct + NCOMP * r;
                    o[r] = i < nflush ? part_out[i] : 0.0f;
                }
#pragma unroll
                for (int r = 0; r < NF; ++r) v[r] += o[r];
            }
#pragma unroll
            for (int r = 0; r < NF; ++r) {
                const int i = ct + NCOMP * r;
                if (i < nflush) part_out[i] = v[r];
            }
        }
        if (ktl && !flushed) g_tcgru_tl[13] = clock64();
        compute_bar();          // scratch reads done before the next step's gate gradients go to the A image
}

// THREE warpgroups: two of compute warps and one that holds the MMA-issue warp (its other three warps idle), so that the
// groups can trade registers (setmaxnreg).  A 9-warp CTA is capped at 168 registers per thread (three warps share one
// scheduler partition's 16 K registers), and at that cap every attempt to hold a step's 96 operands in registers ended with
// the compiler parking them on the stack as they arrived (LDG; STL pairs: one L2 round trip per value, 10-16 k cycles per
// step).  The issue group gives registers back (40), the compute groups take 232 -- ptxas budgets the two roles separately
// only if each role's whole branch is dominated by its own setmaxnreg (an if / else that merges before the roles split again
// left the compute code at 168 with 1.5 KB of spills).  Folding the issuing into compute warp 0 instead (8 warps, 254
// registers) was slower: tcgen05.mma issue blocks while the pipe's queue is full, warp 0 arrived ~3 k cycles late at every
// hand-off.
constexpr int BWD_THREADS = NCOMP + 128;
__global__ void __launch_bounds__(BWD_THREADS, 1) tc_gru_bwd_kernel(Args a) {
    using namespace bwd;
    extern __shared__ __align__(1024) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + oBar + N_BARS * 8);
    const GruLayout& L = a.L;
    const int tiles_b = (a.B + M - 1) / M;
    const int units = a.N * tiles_b;
    const int nsteps = a.t1 - a.t0;
    const bool ktl = g_tcgru_tl_on && blockIdx.x == 0 && tid == 0;
    if (ktl) g_tcgru_tl[10] = clock64();

    // ---- set-up that touches no global memory --------------------------------------------------------
    for (int i = tid * 16; i < A_BYTES + 2 * B_BYTES; i += BWD_THREADS * 16) *reinterpret_cast<uint4*>(sm + oAs + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < N_BARS; ++i) tc::mbar_init(&bars[i], i < D_1 ? NCOMP : 1);
        tc::fence_mbar_init();
    }
    if (warp == 8) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    pdl_wait_then_trigger();
    {
        const float* P = a.params;
        // B of the dx1 | dh GEMM: output n < 32: dx1 unit n = sum_k da_i[k] Wih[k][n] (k < 96: r, z, n gate rows);
        //                         output 32 + n: dh unit n = sum_k da_h[k] Whh[k][n] (A columns 0..63 and 96..127: r, z, hn)
        CMARL_STRIDED(i, 2 * H * 4 * H, BWD_THREADS) {
            const int k = i / (2 * H), n = i - k * (2 * H);      // n fastest: consecutive threads read consecutive floats
            float w = 0.0f;
            if (n < H) { if (k < G3) w = __ldcg(P + L.wih + k * H + n); }
            else if (k < 2 * H) w = __ldcg(P + L.whh + k * H + (n - H));
            else if (k >= G3) w = __ldcg(P + L.whh + (k - H) * H + (n - H));
            st_split(sm + oWb, sm + oWb + szWb, kmaj(n, k, 4 * H), w);
        }
        CMARL_STRIDED(i, H * 8, BWD_THREADS) {
            const int j = i / 8, c = i - j * 8;
            reinterpret_cast<float*>(sm + oW2T)[i] = (c < NA) ? __ldcg(P + L.w2 + c * H + j) : 0.0f;
        }
        if (tid < 8) reinterpret_cast<float*>(sm + oB2)[tid] = (tid < NA) ? __ldcg(P + L.b2 + tid) : 0.0f;
    }
    __syncthreads();
    // the ones row of the B image (row 64: hi = 1, lo = 0): its products with the gate gradients are the bias gradients
    if (tid < M) *reinterpret_cast<float*>(sm + oBs + smaj(2 * H, tid)) = 1.0f;
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sbase = tc::smem_u32(sm);
    if (ktl) g_tcgru_tl[11] = clock64();

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 8) {
        // ================================ MMA issue warp ==================================================
        const bool leader = tc::elect_one();
        const uint32_t As = sbase + oAs, Bs_h = sbase + oBs, Bs_l = Bs_h + B_BYTES;
        uint32_t par = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            for (int i = 0; i < nsteps; ++i, par ^= 1) {
                const uint32_t keep = i != 0 ? 1u : 0u;      // the accumulators leave TMEM at the end of every tile
                acquire(&bars[R_1], par);
                issue_ss<128, NB>(leader, tmem + cWrz, As, Bs_h, Bs_l, keep);
                if (leader) tc::mma_commit(&bars[D_1]);
                issue_ts<2 * H, 4 * H>(leader, tmem + cDB, tmem + cDAh, tmem + cDAl, sbase + oWb, sbase + oWb + szWb, 0u);
                acquire(&bars[R_2], par);
                issue_ss<128, NB>(leader, tmem + cWn, As, Bs_h, Bs_l, keep);
                if (leader) tc::mma_commit(&bars[D_2]);          // (covers the dx1 | dh GEMM as well)
                acquire(&bars[R_3], par);
                issue_ss<64, H>(leader, tmem + cW1, As, Bs_h, Bs_l, keep);
                if (leader) tc::mma_commit(&bars[D_3]);
            }
        }
        __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ================================ compute warps ====================================================
        const int q = warp & 3, hf = warp >> 2;
        const int s = q * 32 + lane;
        const int ct = warp * 32 + lane;                 // 0..255
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        const float* fw2 = reinterpret_cast<const float*>(sm + oW2T);
        uint8_t* As = sm + oAs;
        uint8_t* Bs_h = sm + oBs; uint8_t* Bs_l = Bs_h + B_BYTES;
        const int so = smaj(0, s);
        float* part_out = a.partials + (size_t)blockIdx.x * (L.count + CMARL_N_STATS);

        // Loads: running pointers (one 64-bit add per row; 64-bit products per row were ~10 instructions and a register pair
        // each); whatever is wanted a step later is first pulled into L2, one row per lane.
        float xr[NXO * 8];
        auto load_x = [&](int t, int g, int b) {
            const float* xp = a.x + (size_t)t * a.stride_t + (size_t)g * a.stride_g + b + (size_t)(8 * hf) * a.B;
            const uint32_t step = (uint32_t)a.B;
            const int rmax = (b < a.B) ? a.in_rows - 8 * hf : 0;
#pragma unroll
            for (int i = 0; i < NXO; ++i)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int rr = 16 * i + e;
                    xr[i * 8 + e] = (rr < rmax) ? __ldcg(xp + (uint32_t)rr * step) : 0.0f;
                }
        };



        // A step's operands: x1, r, z, n, Whn h + bhn (stash), h_t (h_seq), this thread's 16 units, and its 5 dlogits.  With the
        // compute groups at 232 registers they are requested a step AHEAD -- right after the hand-off of the (n, hn) round, in
        // the shadow of that round and the dx1 | dh GEMM -- and sit in registers until the gate-gradient stage of their step.
        float ox1[16], orr[16], oz[16], on[16], og[16], ohp[16], odz[NA];
        auto load_ops = [&](int t, int g, int b, int half) {
            const bool inb = b < a.B, hb = inb && t > 0;
            const size_t rs = (size_t)a.B;
            const float* p0 = a.stash + ((size_t)t * a.N + g) * (5 * H) * a.B + (size_t)(16 * hf) * rs + b;
            if (half != 1) {
                const float* p1 = p0 + (size_t)H * rs;
                const float* p2 = p1 + (size_t)H * rs;
                const float* dzp = a.dlogits + ((size_t)t * a.N + g) * 8 * a.B + b;
#pragma unroll
                for (int c = 0; c < NA; ++c, dzp += rs) odz[c] = inb ? __ldcg(dzp) : 0.0f;
#pragma unroll
                for (int e = 0; e < 16; ++e, p0 += rs, p1 += rs, p2 += rs) {
                    ox1[e] = inb ? __ldcg(p0) : 0.0f;
                    orr[e] = inb ? __ldcg(p1) : 0.0f;
                    oz[e] = inb ? __ldcg(p2) : 0.0f;
                }
            }
            if (half != 0) {
                const float* p3 = a.stash + ((size_t)t * a.N + g) * (5 * H) * a.B + (size_t)(3 * H + 16 * hf) * rs + b;
                const float* p4 = p3 + (size_t)H * rs;
                const float* p5 = a.h_seq + (((size_t)t * a.N + g) * H) * a.B + (size_t)(16 * hf) * rs + b;
#pragma unroll
                for (int e = 0; e < 16; ++e, p3 += rs, p4 += rs, p5 += rs) {
                    on[e] = inb ? __ldcg(p3) : 0.0f;
                    og[e] = inb ? __ldcg(p4) : 0.0f;
                    ohp[e] = hb ? __ldcg(p5) : 0.0f;
                }
            }
        };

        // ONE flat loop over the steps of all tiles of this CTA (every stage exists once in the code).  The head of every step
        // (logits, loss terms, dlogits, dW2) was evaluated by the forward kernel: this one reads 5 dlogits per sample.
        const int ntiles = (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int total = ntiles * nsteps;
        uint32_t par = 0;
        bool flushed = false;
        float carry[16];                                 // dL/dh_{t+1} carried down the chunk (this thread's units)
        int ci = 0, cu = blockIdx.x;                    // cur: step index within its tile, tile
#pragma unroll 1
        for (int k = 0; k < total; ++k) {
            const bool active = true;
            const int g = cu / tiles_b, b = (cu - g * tiles_b) * M + s, t = a.t1 - 1 - ci;
            const bool inb = b < a.B;
            const bool tl_on = g_tcgru_tl_on && blockIdx.x == 0 && k == 1 && tid == 0;
            BTL(0);
            uint32_t x1mask = 0u;
            // L2 prefetch of the step BELOW this one (its lines were written by the forward kernel up to 10 steps x 0.8 KB x
            // 24 k samples ago: more than L2 holds), one row per LANE (a warp's 32 samples of a row are one 128-B line): this
            // warp's 80 stash rows + 16 h rows, the 5 dlogits rows and the input rows of its tile -- 4 instructions per warp
            if (ci + 1 < nsteps) {
                const int bw = b - lane;
                if (bw < a.B) {
                    const size_t rs = (size_t)a.B;
                    const float* slab = a.stash + ((size_t)(t - 1) * a.N + g) * (5 * H) * rs + bw;
                    const float* hrow = a.h_seq + (((size_t)(t - 1) * a.N + g) * H) * rs + bw;
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        const int r = lane + 32 * m, arr = r >> 4, k = r & 15;       // 96 rows: 5 stash arrays + h, 16 units each
                        const float* pf = arr < 5 ? slab + (size_t)(arr * H + 16 * hf + k) * rs : hrow + (size_t)(16 * hf + k) * rs;
                        if (arr < 5 || t > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                    }
                    const float* pf2 = lane < 8 ? a.dlogits + (((size_t)(t - 1) * a.N + g) * 8 + lane) * rs + bw
                                                : a.x + (size_t)(t - 1) * a.stride_t + (size_t)g * a.stride_g + (size_t)(lane - 8) * rs + bw;
                    if (lane < NA || (lane >= 8 && lane - 8 < a.in_rows)) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf2));
                }
            }
            if (active) {
                // ---- x1, h_t -> B image rows 0..31 / 32..63; gate gradients da -> TMEM A columns, their (r, z) pair -> A image
                //      (four units at a time; operands prefetched a step ago, a tile's first step requests them here) ----------
                {
                    if (ci == 0) load_ops(t, g, b, 2);
                    if (ci > 0) acquire(&bars[D_3], par ^ 1);
                    else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) carry[i] = 0.0f;
                    }
                    BTL(1);
                    float dz[NA];
#pragma unroll
                    for (int c = 0; c < NA; ++c) dz[c] = odz[c];
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        float sx1[4], shp[4], sr[4], sz[4], sn[4], sg[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            sx1[e] = ox1[4 * k4 + e]; shp[e] = ohp[4 * k4 + e]; sr[e] = orr[4 * k4 + e];
                            sz[e] = oz[4 * k4 + e]; sn[e] = on[4 * k4 + e]; sg[e] = og[4 * k4 + e];
                        }
                        const int j0 = 16 * hf + 4 * k4;
                        float dan[4], daz[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int kk = 4 * k4 + e, j = j0 + e;
                            st_split(Bs_h, Bs_l, smaj(H + j, 0) + so, shp[e]);
                            st_split(Bs_h, Bs_l, smaj(j, 0) + so, sx1[e]);
                            x1mask |= (sx1[e] > 0.0f ? 1u : 0u) << kk;
                            const float4 w = *reinterpret_cast<const float4*>(fw2 + j * 8);
                            const float w4 = fw2[j * 8 + 4];
                            float up = w.x * dz[0];
                            up = fmaf(w.y, dz[1], up); up = fmaf(w.z, dz[2], up); up = fmaf(w.w, dz[3], up); up = fmaf(w4, dz[4], up);
                            const float hc = fmaf(shp[e] - sn[e], sz[e], sn[e]);      // h_{t+1}, exactly as the forward pass formed it
                            const float dh = carry[kk] + (hc > 0.0f ? up : 0.0f);
                            const float dn = dh * (1.0f - sz[e]);
                            const float dzg = dh * (shp[e] - sn[e]);
                            carry[kk] = dh * sz[e];
                            dan[e] = dn * (1.0f - sn[e] * sn[e]);
                            daz[e] = dzg * (sz[e] * (1.0f - sz[e]));
                        }
                        const uint32_t c0 = tl + 16 * hf + 4 * k4;
                        uint32_t vh[4], vl[4];
                        auto split4 = [&](auto&& f) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float h, l;
                                tc::split_tf32(f(e), h, l);
                                vh[e] = __float_as_uint(h); vl[e] = __float_as_uint(l);
                            }
                        };
                        split4([&](int e) { return dan[e] * sg[e] * (sr[e] * (1.0f - sr[e])); });      // da_r
                        tmem_st4(c0 + cDAh, vh); tmem_st4(c0 + cDAl, vl);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            *reinterpret_cast<uint32_t*>(As + smaj(j0 + e, 0) + so) = vh[e];
                            *reinterpret_cast<uint32_t*>(As + smaj(2 * H + j0 + e, 0) + so) = vl[e];
                        }
                        split4([&](int e) { return daz[e]; });                                          // da_z
                        tmem_st4(c0 + cDAh + H, vh); tmem_st4(c0 + cDAl + H, vl);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            *reinterpret_cast<uint32_t*>(As + smaj(H + j0 + e, 0) + so) = vh[e];
                            *reinterpret_cast<uint32_t*>(As + smaj(3 * H + j0 + e, 0) + so) = vl[e];
                        }
                        split4([&](int e) { return dan[e]; });                                          // da_n
                        tmem_st4(c0 + cDAh + 2 * H, vh); tmem_st4(c0 + cDAl + 2 * H, vl);
                        split4([&](int e) { return dan[e] * sr[e]; });                                  // da_hn
                        tmem_st4(c0 + cDAh + 3 * H, vh); tmem_st4(c0 + cDAl + 3 * H, vl);
                    }
                }
                BTL(2);
                publish(&bars[R_1]);            // issuer: (r, z) weight-gradient round, then the dx1 | dh GEMM
                BTL(3);
            }
            BTL(4);
            if (active) {
                // ---- (n, hn) pair: back from the TMEM A columns into the A image once the first round is done ----
                acquire(&bars[D_1], par);
                BTL(5);
                {
                    uint32_t v0[16], v1[16], v2[16], v3[16];
                    tc::tmem_ld16(tl + cDAh + 2 * H + 16 * hf, v0);
                    tc::tmem_ld16(tl + cDAh + 3 * H + 16 * hf, v1);
                    tc::tmem_ld16(tl + cDAl + 2 * H + 16 * hf, v2);
                    tc::tmem_ld16(tl + cDAl + 3 * H + 16 * hf, v3);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = 16 * hf + i;
                        *reinterpret_cast<uint32_t*>(As + smaj(j, 0) + so) = v0[i];
                        *reinterpret_cast<uint32_t*>(As + smaj(H + j, 0) + so) = v1[i];
                        *reinterpret_cast<uint32_t*>(As + smaj(2 * H + j, 0) + so) = v2[i];
                        *reinterpret_cast<uint32_t*>(As + smaj(3 * H + j, 0) + so) = v3[i];
                    }
                }
                publish(&bars[R_2]);
                BTL(6);
                if (ci + 1 < nsteps) load_ops(t - 1, g, b, 2);  // in flight under the (n, hn) round and the dx1 | dh GEMM (split over both
                                                                // rounds' shadows it delayed the restaging: 13.5 k vs 13.0 k cycles per step)
                load_x(t, g, b);                // this step's input rows (dW1 round): requested here, not at the first hand-off
                                                // (held across the restaging they were spilled as they arrived)
            }
            if (active) {
                // ---- dx1 = (da_i Wih) . relu'(x1), dh carry += da_h Whh; dW1 round: A image <- dx1, B image <- x | 1 ----
                acquire(&bars[D_2], par);       // second round and the dx1 | dh GEMM complete
                BTL(7);

                {
                    uint32_t vx[16], vh[16];
                    tc::tmem_ld16(tl + cDB + 16 * hf, vx);
                    tc::tmem_ld16(tl + cDB + H + 16 * hf, vh);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = 16 * hf + i;
                        carry[i] += __uint_as_float(vh[i]);
                        const float dx1 = ((x1mask >> i) & 1u) ? __uint_as_float(vx[i]) : 0.0f;
                        float h, l;
                        tc::split_tf32(dx1, h, l);
                        *reinterpret_cast<float*>(As + smaj(j, 0) + so) = h;
                        *reinterpret_cast<float*>(As + smaj(H + j, 0) + so) = l;
                    }
                }
#pragma unroll
                for (int ii = 0; ii < NXO; ++ii) {
                    const int c = 2 * ii + hf;
                    if (c < NCX) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) st_split(Bs_h, Bs_l, smaj(8 * c + e, 0) + so, xr[ii * 8 + e]);
                    }
                }
                if (hf == 1) {                  // ones row of this round: row 24 (hi = 1, lo = 0)
                    *reinterpret_cast<float*>(Bs_h + smaj(K1P, 0) + so) = 1.0f;
                    *reinterpret_cast<float*>(Bs_l + smaj(K1P, 0) + so) = 0.0f;
                }
                publish(&bars[R_3]);
                BTL(8);
                // ---- end of a tile: its weight-gradient sums leave TMEM before the next tile's operands are requested ------
                if (ci + 1 == nsteps) {
                    acquire(&bars[D_3], par);
                    bwd_flush(sm, a, tl, q, hf, lane, ct, g, part_out, flushed, ktl);
                    flushed = true;
                }
                par ^= 1;
            }
            BTL(9);
            if (++ci == nsteps) { ci = 0; cu += gridDim.x; }
        }
        if (ktl) g_tcgru_tl[14] = clock64();
        // ---- the rest of the partial row: W2 / b2 / statistics = this CTA's share of the forward kernel's rows (fixed order) --
        {
            constexpr int NFW = NA * H + NA + CMARL_N_STATS;        // 173: contiguous behind bhh in the parameter order
            for (int i = ct; i < NFW; i += NCOMP) {
                float v = 0.0f;
                for (int r = blockIdx.x; r < a.grid_fwd; r += gridDim.x) v += __ldcg(a.fwd_partials + (size_t)r * NFW + i);
                part_out[L.w2 + i] = v;
            }
        }
    }
    if (ktl) g_tcgru_tl[15] = clock64();
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tcgru

int cmarl_tc_gru_setup() {
    int e = cmarl_check_cuda(cudaFuncSetAttribute(tcgru::tc_gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcgru::fwd::SMEM),
                             "cudaFuncSetAttribute(tc_gru_fwd_kernel)");
    if (!e) e = cmarl_check_cuda(cudaFuncSetAttribute(tcgru::tc_gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcgru::bwd::SMEM),
                                 "cudaFuncSetAttribute(tc_gru_bwd_kernel)");
    return e;
}

// which = 1: forward (hidden states + stash), 2: backward (partial rows), 3: both.  Returns the grid of the backward launch
// (the number of partial rows) in *grid_out.
int cmarl_tc_gru_launch(cmarl_ctx* ctx, const chain::GruChunkArgs& a, int which, int* grid_out, cudaStream_t st) {
    const int units = a.N * ceil_div(a.B, tctile::M);
    if (units >= (1 << 24)) { cmarl_set_error("tc_gru: too many tiles"); return -1; }
    if (which & 1) {
        const int grid = a.grid_fwd;
        CMARL_CUDA(cmarl_launch(ctx, tcgru::tc_gru_fwd_kernel, dim3(grid), dim3(tctile::NTHREADS), tcgru::fwd::SMEM, st, a));
    }
    if (which & 2) {
        const int grid = units < ctx->sm_count ? units : ctx->sm_count;
        CMARL_CUDA(cmarl_launch(ctx, tcgru::tc_gru_bwd_kernel, dim3(grid), dim3(tcgru::BWD_THREADS), tcgru::bwd::SMEM, st, a));
        if (grid_out) *grid_out = grid;
    }
    return 0;
}

extern "C" int cmarl_debug_tcgru_timeline(int enable, long long* out_host32) {
    cudaError_t e = cudaMemcpyToSymbol(tcgru::g_tcgru_tl_on, &enable, sizeof(int));
    if (e == cudaSuccess && out_host32) e = cudaMemcpyFromSymbol(out_host32, tcgru::g_tcgru_tl, sizeof(long long) * 32);
    return (int)e;
}
