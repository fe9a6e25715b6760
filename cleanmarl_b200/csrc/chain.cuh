// Block-cooperative fused MLP chain for sm_100a: in -> H -> H -> out, forward and backward, with
// every activation resident in shared memory (feature-major [row][sample], LD = M + 4).
//
// One CTA processes a tile of M samples (M consecutive envs b at one (t, agent-group g)):
//   TMA bulk copies (cp.async.bulk + mbarrier, double buffered) bring the input rows in,
//   S1/S2  H1 = relu(W1 x + b1), H2 = relu(W2 H1 + b2)      register-tiled block GEMMs (8 x TN)
//   S3     z = W3 H2 + b3, per-sample head (policy loss / value loss) -> dz
//   S4     dW3 += dz H2^T, db3               S5  dH2 = (W3^T dz) . relu'(H2)   (in place)
//   S6     dW2 += dH2 H1^T, db2              S7  dH1 = (W2^T dH2) . relu'(H1)  (in place)
//   S8     dW1 += dH1 x^T,  db1
// Weight gradients accumulate in shared memory for the CTA's whole lifetime (persistent CTAs), are
// written once as a per-CTA partial and summed in a fixed order by reduce_partials_kernel, so the
// result is deterministic run to run.
//
// Shared-memory traffic is the limiter of an fp32 FFMA kernel on smem operands (128 B/clk/SM to
// the register file vs 128 FMA lanes/clk): an 8x8 register tile needs (8+8)*4 B per 64 FMA, i.e.
// exactly 1 B per lane-FMA -- the tile shapes below are chosen to sit at or near that balance.
#pragma once

#include "common.cuh"

namespace chain {

constexpr int OUTP = 8;   // padded rows of the output layer (5 logits / 1 value)

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1D bulk async copy (TMA engine, SASS UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float warp_sum_f(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// ------------------------------------------------------------------------------------------------
// Compile-time configuration
// ------------------------------------------------------------------------------------------------
template <int H_, int KIN_>
struct Cfg {
    static constexpr int H = H_;
    static constexpr int KIN = KIN_;                                  // padded input rows (24 or 56)
    static constexpr bool BIG = (H_ == 32 && KIN_ == 24);             // small net: afford a 256-sample tile
    static constexpr int M = BIG ? 256 : 128;                         // samples per tile
    static constexpr int NT = BIG ? 256 : 128;                        // threads per CTA
    static constexpr int TN = (H_ == 64) ? 8 : 4;                     // outputs per thread in S1/S2/S5/S7
    static constexpr int LD = M + 4;                                  // row stride (floats), 16-B aligned rows
    static constexpr int SG = M / 8;                                  // sample groups (8 samples each)
    static constexpr int OG = H / TN;                                 // output groups
    static_assert(SG * OG == NT, "thread tiling must cover the tile");
    static constexpr int PT = (H_ == 64) ? 8 : 4;                     // dW patch is PT x PT
    static constexpr int SX = NT / 128;                               // cross-warp sample splits in the dW stages
    // shared-memory map (float offsets)
    static constexpr int oBar = 0;                     // 2 mbarriers (16 B)
    static constexpr int oW1T = 4;                     // [KIN][H]   in-major  (S1)
    static constexpr int oB1 = oW1T + KIN * H;         // [4][H]     b1 (+ folded id column) per agent group
    static constexpr int oW2T = oB1 + 4 * H;           // [H][H]     in-major  (S2)
    static constexpr int oW2 = oW2T + H * H;           // [H][H]     out-major (S7)
    static constexpr int oB2 = oW2 + H * H;            // [H]
    static constexpr int oW3T = oB2 + H;               // [H][OUTP]
    static constexpr int oB3 = oW3T + H * OUTP;        // [OUTP]
    static constexpr int oX = oB3 + OUTP;              // [2][KIN][LD]
    static constexpr int oH1 = oX + 2 * KIN * LD;      // [H][LD]
    static constexpr int oH2 = oH1 + H * LD;           // [H][LD]
    static constexpr int oZ = oH2 + H * LD;            // [OUTP][LD]
    static constexpr int oFwdEnd = oZ + OUTP * LD;
    static constexpr int PMAX = H * (KIN + 4) + H + H * H + H + OUTP * H + OUTP;   // >= real param count
    static constexpr int oDW = oFwdEnd;                // [PMAX]  gradient accumulators (train)
    static constexpr int oScr = oDW + PMAX;            // scratch for the fixed-order cross-warp combines
    static constexpr int SCR_PATCH = SX > 1 ? H * (H > KIN ? H : KIN) : 0;
    static constexpr int SCR = SCR_PATCH + 8 * OUTP * H;
    static constexpr int oRed = oScr + SCR;            // [32] block-reduce scratch
    static constexpr int oTrainEnd = oRed + 64;
    static_assert(PMAX % 4 == 0 && SCR % 4 == 0, "keep 16-B alignment");
    static constexpr size_t smem_fwd = (size_t)oFwdEnd * 4;
    static constexpr size_t smem_train = (size_t)oTrainEnd * 4;
};

struct NetDesc {
    const float* params;   // flat params of this net
    int in_rows;           // rows loaded per tile (18, 21 or 54)
    int in_dim;            // fan-in of W1 in the parameter layout (in_rows, or in_rows + n_groups when ids fold)
    int fold_ids;          // 1: the one-hot id columns [in_rows, in_rows+G) of W1 act as a per-group bias
    int out_dim;           // 5 or 1
};

struct TileSrc {
    const float* x;        // row r of tile (t, g, b0) lives at x + t*stride_t + g*stride_g + r*B + b0
    size_t stride_t, stride_g;
    int T, G, B;           // B = row stride (envs of the whole buffer)
    int nb;                // envs of this launch: [0, nb) relative to the (pre-offset) base pointers -- the whole buffer
                           // (nb == B) or one minibatch (contiguous env block, cmarl_ppo_epoch_grads_ex)
    int indep;             // tc_chain_kernel only: 1 = reads nothing the launch in front of it writes (see tc_chain.cu)
    int flush;             // tc_chain_kernel only: tiles whose weight-gradient products accumulate in TMEM between two flushes
};

struct PolicyHeadArgs {
    const int32_t* actions;    // [T][N][B]
    const float* logp_old;     // [T][N][B]
    const float* adv;          // [T][V][B]
    const uint8_t* mask;       // [T][B] or null
    const uint8_t* avail;      // [T][N][A][B] or null
    int V, A;
    float clip, ent_coef, inv_groups;
};

// One truncated-BPTT chunk of the recurrent actor (gru.cu: fp32 FFMA kernel, tc_gru.cu: tcgen05 kernels)
struct GruChunkArgs {
    const float* params;      // recurrent actor, torch order
    GruLayout L;
    const float* x;           // state [T][S][B] (rows 18 g + k) or obs [T][N][O][B]
    size_t stride_t, stride_g;
    int in_rows;              // 18 (ids folded into the bias) or O
    int fold_ids;
    int T, N, B;
    int t0, t1;
    float* h_seq;             // [T+1][N][H][B]
    float* stash;             // [T][N][5H][B] or null: x1, r, z, n, ghn of every step (pass 1 writes, pass 2 reads instead
                              // of recomputing the gates: 640 B per sample-step through L2 / HBM for 27 % fewer instructions)
    float* partials;          // [grid][P + 8]
    PolicyHeadArgs head;
    int passes;               // FFMA kernel: bit 0 = pass 1 (hidden states, stash), bit 1 = pass 2 (head + backward; alone it needs the stash)
    int flush;                // (unused: the tcgen05 backward kernel flushes its TMEM accumulators at the end of every tile)
    // tcgen05 pair (tc_gru.cu): the forward kernel evaluates the head as well -- dlogits of every step go to `dlogits`
    // [T][N][8][B], its dW2 / db2 / statistics to one row of 176 floats per forward CTA in `fwd_partials`
    float* dlogits;
    float* fwd_partials;
    int grid_fwd;
};

struct ValueHeadArgs {
    const float* returns;      // [T][V][B]  (train)
    const uint8_t* mask;
    float* values_out;         // [T][V][B]  (forward)
    float inv_heads;
    const float* values_old;   // [T][V][B]  values at rollout time (value clipping only)
    float vclip;               // <= 0: the reference's plain MSE (MME:554-558); > 0: PPO2-style clipped value loss
};

// ------------------------------------------------------------------------------------------------
// Weights -> shared memory (transposed where the consumer wants k-major)
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ void load_weights(float* sm, const NetDesc& nd, int n_groups) {
    constexpr int H = C::H;
    const float* P = nd.params;
    const int in_dim = nd.in_dim;
    const float* W1 = P;
    const float* b1 = W1 + H * in_dim;
    const float* W2 = b1 + H;
    const float* b2 = W2 + H * H;
    const float* W3 = b2 + H;
    const float* b3 = W3 + nd.out_dim * H;
    for (int i = threadIdx.x; i < C::KIN * H; i += C::NT) {
        const int k = i / H, j = i - k * H;
        sm[C::oW1T + i] = (k < nd.in_rows) ? W1[j * in_dim + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < 4 * H; i += C::NT) {
        const int g = i / H, j = i - g * H;
        float v = b1[j];
        if (nd.fold_ids && g < n_groups) v += W1[j * in_dim + nd.in_rows + g];
        sm[C::oB1 + i] = v;
    }
    for (int i = threadIdx.x; i < H * H; i += C::NT) {
        const int a = i / H, b = i - a * H;
        sm[C::oW2 + i] = W2[i];                 // [j][k]
        sm[C::oW2T + b * H + a] = W2[i];        // [k][j]
    }
    for (int i = threadIdx.x; i < H; i += C::NT) sm[C::oB2 + i] = b2[i];
    for (int i = threadIdx.x; i < H * OUTP; i += C::NT) {
        const int j = i / OUTP, a = i - j * OUTP;
        sm[C::oW3T + i] = (a < nd.out_dim) ? W3[a * H + j] : 0.0f;
    }
    if (threadIdx.x < OUTP) sm[C::oB3 + threadIdx.x] = (threadIdx.x < nd.out_dim) ? b3[threadIdx.x] : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// Tile input: rows of M contiguous floats.  Full, 16-B aligned tiles go through the TMA engine;
// the ragged last tile (or an unaligned B) falls back to guarded loads, zero-filling the tail.
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ bool tile_is_bulk(const TileSrc& src, int b0) {
    return (b0 + C::M <= src.nb) && ((src.B & 3) == 0) && ((reinterpret_cast<uintptr_t>(src.x) & 15) == 0);
}

template <class C>
__device__ void issue_tile(float* xbuf, uint64_t* bar, const TileSrc& src, int in_rows, int t, int g, int b0) {
    // called by all threads; only full tiles use the barrier
    const float* base = src.x + (size_t)t * src.stride_t + (size_t)g * src.stride_g + b0;
    if (tile_is_bulk<C>(src, b0)) {
        if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)(in_rows * C::M * 4));
        if (threadIdx.x < in_rows) {
            fence_proxy_async();   // earlier generic-proxy accesses of this buffer precede the async write
            bulk_g2s(xbuf + threadIdx.x * C::LD, base + (size_t)threadIdx.x * src.B, C::M * 4, bar);
        }
    } else {
        const int valid = src.nb - b0;
        for (int i = threadIdx.x; i < in_rows * C::M; i += C::NT) {
            const int r = i / C::M, s = i - r * C::M;
            xbuf[r * C::LD + s] = (s < valid) ? __ldcg(base + (size_t)r * src.B + s) : 0.0f;
        }
        if (threadIdx.x == 0) mbar_arrive(bar);   // keeps the phase bookkeeping uniform
    }
}

// ------------------------------------------------------------------------------------------------
// Register-tiled block GEMM over smem operands:  Out[n][s] = epi( sum_k Wk[k][n] * A[k][s] )
//   thread tile = 8 samples (two float4 chunks, M/2 apart: conflict-free LDS.128) x TN outputs
//   FWD:  epi = relu(acc + bias[n])           BWD: epi = Out_old[n][s] > 0 ? acc : 0   (in place)
// ------------------------------------------------------------------------------------------------
template <class C, int NOUT, bool BWD>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ A, int K, const float* __restrict__ Wk,
                                          const float* __restrict__ bias, float* __restrict__ Out) {
    constexpr int TN = C::TN, LD = C::LD, SG = C::SG;
    static_assert(NOUT == C::H, "hidden layers only");
    const int sg = threadIdx.x % SG, og = threadIdx.x / SG;
    const int s0 = 4 * sg, s1 = C::M / 2 + 4 * sg, n0 = og * TN;
    float acc[TN][8];
#pragma unroll
    for (int i = 0; i < TN; ++i) {
        const float b = BWD ? 0.0f : bias[n0 + i];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = b;
    }
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(A + k * LD + s0);
        const float4 a1 = *reinterpret_cast<const float4*>(A + k * LD + s1);
        float w[TN];
#pragma unroll
        for (int i = 0; i < TN; i += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(Wk + k * NOUT + n0 + i);
            w[i] = wv.x; w[i + 1] = wv.y; w[i + 2] = wv.z; w[i + 3] = wv.w;
        }
#pragma unroll
        for (int i = 0; i < TN; ++i) {
            acc[i][0] = fmaf(w[i], a0.x, acc[i][0]); acc[i][1] = fmaf(w[i], a0.y, acc[i][1]);
            acc[i][2] = fmaf(w[i], a0.z, acc[i][2]); acc[i][3] = fmaf(w[i], a0.w, acc[i][3]);
            acc[i][4] = fmaf(w[i], a1.x, acc[i][4]); acc[i][5] = fmaf(w[i], a1.y, acc[i][5]);
            acc[i][6] = fmaf(w[i], a1.z, acc[i][6]); acc[i][7] = fmaf(w[i], a1.w, acc[i][7]);
        }
    }
#pragma unroll
    for (int i = 0; i < TN; ++i) {
        float4* p0 = reinterpret_cast<float4*>(Out + (n0 + i) * LD + s0);
        float4* p1 = reinterpret_cast<float4*>(Out + (n0 + i) * LD + s1);
        float4 r0, r1;
        if (BWD) {
            const float4 o0 = *p0, o1 = *p1;
            r0.x = o0.x > 0.0f ? acc[i][0] : 0.0f; r0.y = o0.y > 0.0f ? acc[i][1] : 0.0f;
            r0.z = o0.z > 0.0f ? acc[i][2] : 0.0f; r0.w = o0.w > 0.0f ? acc[i][3] : 0.0f;
            r1.x = o1.x > 0.0f ? acc[i][4] : 0.0f; r1.y = o1.y > 0.0f ? acc[i][5] : 0.0f;
            r1.z = o1.z > 0.0f ? acc[i][6] : 0.0f; r1.w = o1.w > 0.0f ? acc[i][7] : 0.0f;
        } else {
            r0.x = fmaxf(acc[i][0], 0.0f); r0.y = fmaxf(acc[i][1], 0.0f);
            r0.z = fmaxf(acc[i][2], 0.0f); r0.w = fmaxf(acc[i][3], 0.0f);
            r1.x = fmaxf(acc[i][4], 0.0f); r1.y = fmaxf(acc[i][5], 0.0f);
            r1.z = fmaxf(acc[i][6], 0.0f); r1.w = fmaxf(acc[i][7], 0.0f);
        }
        *p0 = r0;
        *p1 = r1;
    }
}

// ------------------------------------------------------------------------------------------------
// Weight-gradient stage:  dW[j][k] += sum_s dY[j][s] * X[k][s]   (J = H rows of dY, KP padded rows of X)
//   thread = (patch, split): PT x PT register patch with interleaved rows (j = jg + JG*a,
//   k = kg + KG*b) so the 8 lanes of an LDS.128 phase hit 8 consecutive rows (distinct bank groups);
//   samples are split 2 ways inside the warp (lane bit 4, combined with one shuffle per value)
//   and SX ways across warps (combined through `scr` in a fixed order).
//   Also accumulates the bias gradient db[j] = sum_s dY[j][s] (+ the folded id column).
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void dw_stage(const float* __restrict__ dY, const float* __restrict__ X, int KP,
                                         int k_real, int k_ld, float* __restrict__ dW, float* __restrict__ db,
                                         float* __restrict__ dWid, float* __restrict__ scr) {
    constexpr int PT = C::PT, LD = C::LD, H = C::H, SX = C::SX, M = C::M;
    constexpr int JG = H / PT;
    const int KG = KP / PT;
    const int P = JG * KG;                       // patches
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WPS = C::NT / 32 / SX;         // warps per cross-warp split (= 4)
    const int sx = warp / WPS, wl = warp - sx * WPS;
    const int sw = lane >> 4, pl = lane & 15;
    constexpr int NSPLIT = 2 * SX;
    const int q = sx * 2 + sw;
    constexpr int NQ = M / 4;                    // float4 quads per row
    const int rounds = (P + WPS * 16 - 1) / (WPS * 16);      // uniform trip count: the loop holds barriers
    for (int rnd = 0; rnd < rounds; ++rnd) {
        const int p = (rnd * WPS + wl) * 16 + pl;
        const bool active = p < P;
        const int jg = active ? p / KG : 0, kg = active ? p - jg * KG : 0;
        float acc[PT][PT];
#pragma unroll
        for (int a = 0; a < PT; ++a)
#pragma unroll
            for (int b = 0; b < PT; ++b) acc[a][b] = 0.0f;
        if (active) {
#pragma unroll 2
            for (int c = q; c < NQ; c += NSPLIT) {
                float4 d[PT], x[PT];
#pragma unroll
                for (int a = 0; a < PT; ++a) d[a] = *reinterpret_cast<const float4*>(dY + (jg + JG * a) * LD + 4 * c);
#pragma unroll
                for (int b = 0; b < PT; ++b) x[b] = *reinterpret_cast<const float4*>(X + (kg + KG * b) * LD + 4 * c);
#pragma unroll
                for (int a = 0; a < PT; ++a)
#pragma unroll
                    for (int b = 0; b < PT; ++b) {
                        acc[a][b] = fmaf(d[a].x, x[b].x, acc[a][b]);
                        acc[a][b] = fmaf(d[a].y, x[b].y, acc[a][b]);
                        acc[a][b] = fmaf(d[a].z, x[b].z, acc[a][b]);
                        acc[a][b] = fmaf(d[a].w, x[b].w, acc[a][b]);
                    }
            }
        }
        // combine the two in-warp halves (all lanes participate in the shuffle)
#pragma unroll
        for (int a = 0; a < PT; ++a)
#pragma unroll
            for (int b = 0; b < PT; ++b) acc[a][b] += __shfl_xor_sync(0xffffffffu, acc[a][b], 16);
        if (SX > 1) {
            // cross-warp halves: sx == 1 parks its sums in scr, sx == 0 adds them (fixed order)
            if (sx == 1 && sw == 0 && active) {
#pragma unroll
                for (int a = 0; a < PT; ++a)
#pragma unroll
                    for (int b = 0; b < PT; ++b) scr[(jg + JG * a) * KP + kg + KG * b] = acc[a][b];
            }
            __syncthreads();
        }
        if (sx == 0 && sw == 0 && active) {
#pragma unroll
            for (int a = 0; a < PT; ++a)
#pragma unroll
                for (int b = 0; b < PT; ++b) {
                    const int j = jg + JG * a, k = kg + KG * b;
                    float v = acc[a][b];
                    if (SX > 1) v += scr[j * KP + k];
                    if (k < k_real) dW[j * k_ld + k] += v;
                }
        }
        if (SX > 1) __syncthreads();
    }
    // bias gradient: one thread per row, fixed order over the samples
    if (threadIdx.x < H) {
        const int j = threadIdx.x;
        float s = 0.0f;
        for (int c = 0; c < NQ; ++c) {
            const float4 d = *reinterpret_cast<const float4*>(dY + j * LD + 4 * c);
            s += (d.x + d.y) + (d.z + d.w);
        }
        db[j] += s;
        if (dWid) dWid[j * k_ld] += s;           // folded one-hot id column of this agent group
    }
}

}  // namespace chain
