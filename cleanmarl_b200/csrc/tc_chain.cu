// Tensor-core (tcgen05 / TMEM) version of the fused MLP chain: K4 critic forward and K7 PPO epoch
// (actor and critic forward + head + backward) with every GEMM on the 5th-gen tensor cores.
//
// Precision: kind::tf32 with a 3-term split ("3xTF32"): x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi)
// (residual <= 2^-23 |x|), D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi with fp32 accumulation in TMEM, the two
// small products issued FIRST (the accumulator add truncates, so the big term goes last: measured
// 2.4e-7 vs fp64 on K = 64, plain fp32 GEMM 1.8e-7; profiles/umma_layouts_r1.md).
//
// One CTA = 128 threads = one tile of 128 samples (consecutive envs b at one (t, agent group g)); thread s
// owns sample s = TMEM lane s.  Activations that are the M x K operand of a GEMM (X, H1, dH2) live in
// TMEM (tcgen05.mma with A from TMEM); weights are K-major core-matrix images in shared memory.
//
//   X (global, coalesced LDG) -> split -> TMEM X            F1: D1 = X W1^T
//   E1: H1 = relu(D1 + b1) -> split -> TMEM A               F2: D2 = H1 W2^T
//   E2: H2 = relu(D2 + b2), z = W3 H2 + b3, head -> dz, dH2 = (W3^T dz) . relu'(H2) -> TMEM A
//                                                           B1: D3 = dH2 W2           (D3 reuses D2's columns)
//   E3: dH1 = D3 . relu'(H1)
//   weight gradients: dW2 = dH2^T H1, dW1 = dH1^T X  contract over the SAMPLES, so both operands are
//   needed sample-major ([feature][sample]) in shared memory: A = the hi image of dH stacked on its lo image
//   (2H feature rows: ONE M = 2H MMA forms the A_hi and the A_lo products together), B = H1 / X in rounds of
//   32 feature rows (+ a "ones" row group: its products are the bias gradient and, row 1 + g, the gradient of the
//   folded one-hot id column of agent group g), two passes per round (B_lo, then B_hi).  The accumulators stay in
//   TMEM across `src.flush` tiles of a CTA and are then added into the CTA's partial row.
#include "chain.cuh"
#include "heads.cuh"
#include "tc_ptx.cuh"
#include "tc_tile.cuh"

namespace tcchain {

using namespace chain;
using namespace tctile;

constexpr int B_GROUPS = 5;                   // H1 / X operand: 32 feature rows + the ones group
constexpr int B_S_BYTES = B_GROUPS * SBO_S;

template <int H_, int K1P_, bool TRAIN_, int OUT_>
struct TCfg {
    static constexpr int ZX_BYTES = OUT_ * 128 * 4;
    // dH operand of the weight-gradient GEMMs: the hi image (H feature rows) directly followed by the lo image: together
    // the 2H rows of ONE A operand (M = 2H = 64 or 128)
    static constexpr int A_S_BYTES = (H_ / 8) * SBO_S;        // per hi / lo image
    static constexpr int H = H_, K1P = K1P_;
    static constexpr bool TRAIN = TRAIN_;
    static constexpr int NR2 = H / 32;                    // rounds of the dW2 GEMM (32 H1 features each)
    static constexpr int NR1 = (K1P + 31) / 32;           // rounds of the dW1 GEMM
    // TMEM columns
    static constexpr int cXh = 0, cXl = K1P, cAh = 2 * K1P, cAl = cAh + H, cD1 = cAl + H, cD2 = cD1 + H;
    static constexpr int cW2 = cD2 + H;                   // dW2 | db2 accumulators: rows 0..H-1 = A_hi products, H..2H-1 = A_lo products
    static constexpr int nW2 = 32 * NR2 + 8;
    static constexpr int cW1 = cW2 + nW2;
    static constexpr int nW1 = 32 * NR1 + 8;
    static constexpr int cEnd = TRAIN ? cW1 + nW1 : cD2 + H;
    static constexpr int TMEM_COLS = cEnd <= 32 ? 32 : cEnd <= 64 ? 64 : cEnd <= 128 ? 128 : cEnd <= 256 ? 256 : 512;
    static_assert(cEnd <= 512, "TMEM budget");
    // shared memory (bytes)
    static constexpr int oBar = 0;                        // 13 mbarriers + tmem base
    static constexpr int oW1 = 128;                        // W1 hi | lo   [H][K1P]   K-major core-matrix image
    static constexpr int szW1 = H * K1P * 4;
    static constexpr int oW2 = oW1 + 2 * szW1;            // W2 hi | lo   [H][H]
    static constexpr int szW2 = H * H * 4;
    static constexpr int oW2T = oW2 + 2 * szW2;           // W2^T hi | lo (train)
    static constexpr int oB1 = oW2T + (TRAIN ? 2 * szW2 : 0);   // f32 [4][H]
    static constexpr int oB2 = oB1 + 4 * H * 4;           // f32 [H]
    static constexpr int oW3T = oB2 + H * 4;              // f32 [H][8]
    static constexpr int oB3 = oW3T + H * 8 * 4;          // f32 [8]
    static constexpr int oDW3 = oB3 + 32;                 // f32 [4 warps][8][H] + [4][8]: dW3 / db3 partial sums per warp
    static constexpr int oRed = oDW3 + (TRAIN ? (4 * 8 * H + 32) * 4 : 0);      // f32 [64]
    static constexpr int oZx = oRed + 64;                                       // f32 [OUT][128]: partial outputs half 1 -> half 0, then dz half 0 -> half 1
    static constexpr int oAs = ((oZx + ZX_BYTES + 127) / 128) * 128;            // dH sample-major hi | lo
    static constexpr int oBs = oAs + (TRAIN ? 2 * A_S_BYTES : 0);               // H1 / X sample-major hi | lo
    static constexpr int smem_bytes = oBs + (TRAIN ? 2 * B_S_BYTES : 0);
    static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
    // flush scratch (the A images between two tiles): [H][nW2 + 1] | [H][nW1 + 1] floats
    static constexpr int pS2 = nW2 + 1, pS1 = nW1 + 1;
    static_assert(!TRAIN || (H * (pS2 + pS1)) * 4 <= 2 * A_S_BYTES, "flush scratch fits the A images");

    // two co-resident CTAs double the warps that hide latency -- if BOTH fit: a CTA that needs more than half of the SM's 512
    // TMEM columns makes its neighbour wait in tcgen05.alloc until it has exited (the 64-wide critic forward, 368 -> 512
    // columns, ran as two consecutive waves of 148 CTAs, each paying the 12 k-cycle set-up for three tiles: 31 us)
    static constexpr int CTAS_PER_SM = (smem_bytes <= 113 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
};

// ------------------------------------------------------------------------------------------------
// MMA issue: called by the whole (converged) issue warp so that descriptors stay in uniform registers; one
// elected lane executes each tcgen05.mma.  3xTF32 pass order: (lo,hi), (hi,lo), (hi,hi).
// ------------------------------------------------------------------------------------------------
// The loops stay ROLLED (two MMAs per iteration): fully unrolled, the 261 MMAs of a critic tile are 36 KB of straight-line
// code that the issue warp streams through the instruction cache once per tile, next to the compute warps' own ~90 KB.
// D[128 x N] = A(TMEM, K columns at a_hi / a_lo) * B(smem K-major image [N][K])^T
template <int N, int K>
__device__ __forceinline__ void issue_ts(bool leader, uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
    constexpr uint32_t sbo = (K / 4) * LBO_K;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        uint32_t a = pass == 0 ? a_lo : a_hi;
        uint64_t db = tc::make_smem_desc(pass == 1 ? b_lo : b_hi, LBO_K, sbo, 0);
#pragma unroll 2
        for (int ks = 0; ks < K / 8; ++ks) {
            if (leader) tc::mma_tf32_ts(d, a, db, idesc, (uint32_t)(pass | ks));
            a += 8;
            db += (uint64_t)((2 * LBO_K) >> 4);
        }
    }
}
// D[MROWS x N] (+)= [A_hi ; A_lo](smem sample-major, MROWS = 2H rows) * B(smem sample-major [N][128])^T over the 128 samples
// of the tile.  Pass 0 multiplies by B_lo (rows 0..H-1 collect A_hi B_lo, rows H.. the negligible A_lo B_lo), pass 1 by
// B_hi (A_hi B_hi | A_lo B_hi): small products first within a tile, as in issue_ts; the row blocks are added when the
// accumulators are flushed.  `keep`: the accumulators already hold earlier tiles of the flush group.
// (Round 1 issued three M = 64 passes (A_lo B_hi, A_hi B_lo, A_hi B_hi): 48 MMAs per round instead of 32.  The SS MMAs
// are bound by their shared-memory operand reads, not by the tensor pipe -- measured 10.5 k cycles per critic tile for
// the four rounds in that form, 7.1 k in this one.)
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

// ------------------------------------------------------------------------------------------------
// Weights -> shared memory images (hi | lo)
// ------------------------------------------------------------------------------------------------
// Every global load of a thread is issued before the first value is used: as a plain strided loop the compiler keeps one
// load in flight per iteration, i.e. ~27 serial L2 round trips (~12 us) in front of the critic chain's first tile.
// A warp fills one core matrix (8 rows x 4 k = 128 contiguous bytes) per round: lane = (row l & 7, k l >> 3), so the hi / lo
// stores are conflict-free (consecutive k per lane, the first mapping, put 8 lanes on every bank: the 31 k scalar stores of
// the critic's images were most of its 10.5 k-cycle set-up) and a load instruction still covers whole 16-byte pieces of 8 rows.
template <class C>
__device__ void load_weights_tc(uint8_t* sm, const NetDesc& nd, int n_groups) {
    constexpr int H = C::H, K1P = C::K1P, NT = 288, NW = NT / 32;
    constexpr int M1 = (H / 8) * (K1P / 4), M2 = (H / 8) * (H / 4);          // core matrices of the W1 / W2 images
    constexpr int N1 = (M1 + NW - 1) / NW, N2 = (M2 + NW - 1) / NW, N3 = (H * 8 + NT - 1) / NT;
    static_assert(4 * H <= NT, "one b1 element per thread");
    const float* P = nd.params;
    const int in_dim = nd.in_dim;
    const float* W1 = P;
    const float* b1 = W1 + H * in_dim;
    const float* W2 = b1 + H;
    const float* b2 = W2 + H * H;
    const float* W3 = b2 + H;
    const float* b3 = W3 + nd.out_dim * H;
    const int tid = threadIdx.x;
    if (tid >= NT) return;                       // the idle warps of the three-warpgroup launch
    const int wi = tid >> 5, jl = tid & 7, kk = (tid & 31) >> 3;
    const int eo = jl * 16 + kk * 4;             // byte offset of this lane's element inside a core matrix
    float w1[N1], w2[N2], w3[N3], vb1 = 0.0f, vid = 0.0f, vb2 = 0.0f, vb3 = 0.0f;
#pragma unroll
    for (int r = 0; r < N1; ++r) {
        const int m = wi + r * NW, jg = m / (K1P / 4), kc = m - jg * (K1P / 4);
        const int j = 8 * jg + jl, k = 4 * kc + kk;
        w1[r] = (m < M1 && k < nd.in_rows) ? __ldcg(W1 + j * in_dim + k) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < N2; ++r) {
        const int m = wi + r * NW, jg = m / (H / 4), kc = m - jg * (H / 4);
        w2[r] = m < M2 ? __ldcg(W2 + (8 * jg + jl) * H + 4 * kc + kk) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < N3; ++r) {
        const int i = tid + r * NT, j = i / 8, a = i - j * 8;
        w3[r] = (i < H * 8 && a < nd.out_dim) ? __ldcg(W3 + a * H + j) : 0.0f;
    }
    if (tid < 4 * H) {
        const int g = tid / H, j = tid - g * H;
        vb1 = __ldcg(b1 + j);
        if (nd.fold_ids && g < n_groups) vid = __ldcg(W1 + j * in_dim + nd.in_rows + g);
    }
    if (tid < H) vb2 = __ldcg(b2 + tid);
    if (tid < 8 && tid < nd.out_dim) vb3 = __ldcg(b3 + tid);

#pragma unroll
    for (int r = 0; r < N1; ++r) {
        const int m = wi + r * NW;
        if (m < M1) {
            float hi, lo;
            tc::split_tf32(w1[r], hi, lo);
            const int o = m * 128 + eo;                      // == kmaj(j, k, K1P)
            *reinterpret_cast<float*>(sm + C::oW1 + o) = hi;
            *reinterpret_cast<float*>(sm + C::oW1 + C::szW1 + o) = lo;
        }
    }
#pragma unroll
    for (int r = 0; r < N2; ++r) {
        const int m = wi + r * NW;
        if (m < M2) {
            float hi, lo;
            tc::split_tf32(w2[r], hi, lo);
            const int o = m * 128 + eo;                      // B of F2: N = j, K = k: kmaj(j, k, H)
            *reinterpret_cast<float*>(sm + C::oW2 + o) = hi;
            *reinterpret_cast<float*>(sm + C::oW2 + C::szW2 + o) = lo;
            if (C::TRAIN) {
                const int jg = m / (H / 4), kc = m - jg * (H / 4);
                const int ot = kmaj(4 * kc + kk, 8 * jg + jl, H);   // B of B1: N = k, K = j
                *reinterpret_cast<float*>(sm + C::oW2T + ot) = hi;
                *reinterpret_cast<float*>(sm + C::oW2T + C::szW2 + ot) = lo;
            }
        }
    }
    float* fw3 = reinterpret_cast<float*>(sm + C::oW3T);
#pragma unroll
    for (int r = 0; r < N3; ++r) {
        const int i = tid + r * NT;
        if (i < H * 8) fw3[i] = w3[r];
    }
    if (tid < 4 * H) reinterpret_cast<float*>(sm + C::oB1)[tid] = vb1 + vid;
    if (tid < H) reinterpret_cast<float*>(sm + C::oB2)[tid] = vb2;
    if (tid < 8) reinterpret_cast<float*>(sm + C::oB3)[tid] = vb3;
}

// ------------------------------------------------------------------------------------------------
// The kernel: 8 compute warps + 1 MMA-issue warp.
//   compute thread (cw, lane): TMEM quadrant q = cw & 3, sample s = 32 q + lane, column half hf = cw >> 2:
//   it owns the 16-column chunks c of the H-wide matrices with (c & 1) == hf and the 8-column chunks of X
//   with (c & 1) == hf, so both halves carry the same load in every round of the weight-gradient GEMMs.
//   Hand-offs: compute -> issuer through "ready" mbarriers (256 arrivals), issuer -> compute through
//   tcgen05.commit on "done" mbarriers; every barrier completes exactly one phase per tile.
// ------------------------------------------------------------------------------------------------
// debug timeline: clock64 stamps of CTA 0 (compute thread 0: slots 0..31, issue thread: slots 32..63) for its 2nd tile
__device__ long long g_tc_timeline[64];
__device__ int g_tc_timeline_on = 0;
#define TL_STAMP(slot) do { if (tl_on) g_tc_timeline[slot] = clock64(); } while (0)

enum { R_X = 0, R_H1, R_DH2, R_W2B, R_W1A, R_W1B, D_F1, D_F2, D_B1, D_W2A, D_W2B, D_W1A, D_W1B, N_BARS };

template <class C, class Head>
__global__ void __launch_bounds__(C::CTAS_PER_SM == 1 ? NTHREADS + 96 : NTHREADS, C::CTAS_PER_SM)
tc_chain_kernel(NetDesc nd, TileSrc src, typename Head::Args ha, float* __restrict__ partials, int p_net) {
    extern __shared__ __align__(1024) uint8_t sm[];
    constexpr int H = C::H, K1P = C::K1P, OUT = Head::OUT;
    constexpr bool TRAIN = C::TRAIN;
    constexpr int NOWN = H / 32;                 // 16-column chunks of an H-wide matrix owned by a thread
    constexpr int NCX = K1P / 8;                 // 8-column chunks of X
    constexpr int NXO = (NCX + 1) / 2;           // ... owned by a thread (at most)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ktl = g_tc_timeline_on == (OUT > 1 ? 2 : 1) && blockIdx.x == 0 && tid == 0;     // kernel-level stamps, slots 20..
    if (ktl) g_tc_timeline[20] = clock64();
    // `src.indep` (the critic chain of an epoch, launched as a programmatic dependent of the actor chain): this grid
    // reads nothing the grid in front of it writes, so whatever is launched behind it may be scheduled right away ...
    if (src.indep) pdl_trigger();
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::oBar + N_BARS * 8);
    const int tiles_b = (src.nb + M - 1) / M;
    const int units = src.T * src.G * tiles_b;
    const float inv_tb = 1.0f / (float)tiles_b, inv_g = 1.0f / (float)src.G;

    // ---- one-time setup (all 288 threads) -------------------------------------------------------------
    // first what touches no global memory: under launch chaining this part runs while the kernel in front still does
    if (TRAIN) {
        for (int i = tid * 16; i < 2 * C::A_S_BYTES + 2 * B_S_BYTES; i += NTHREADS * 16)
            *reinterpret_cast<uint4*>(sm + C::oAs + i) = make_uint4(0, 0, 0, 0);
        float* z0 = reinterpret_cast<float*>(sm + C::oDW3);
        for (int i = tid; i < 4 * 8 * H + 32; i += NTHREADS) z0[i] = 0.0f;
    }
    if (tid == 0) {
        for (int i = 0; i < N_BARS; ++i) tc::mbar_init(&bars[i], i < D_F1 ? NCOMP : 1);
        tc::fence_mbar_init();
    }
    if (warp == 8) tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
    // the grid in front of this one (and, transitively, everything before it) has completed past this point
    if (!src.indep) pdl_wait_then_trigger();
    // compute thread (warp < 8): TMEM quadrant q, column half hf, sample s of the tile
    const int q = warp & 3, hf = warp >> 2;
    const int s = q * 32 + lane;                                        // sample within the tile = TMEM lane
    // X chunks owned by this thread: chunk c = 2 i + hf, i < NXO (chunks >= NCX do not exist)
    float xr[NXO * 8];
    auto load_x = [&](int u) {
        int bt, r, t, g;
        fast_divmod(u, tiles_b, inv_tb, r, bt);
        fast_divmod(r, src.G, inv_g, t, g);
        const int b = bt * M + s;
        // 32-bit row offsets from one 64-bit base and one bound per thread: ~5 instructions per load (the 64-bit
        // products and double predicates of the obvious form were 14, a fifth of the critic kernel's instructions)
        const float* xp = src.x + (size_t)t * src.stride_t + (size_t)g * src.stride_g + b + (size_t)(8 * hf) * src.B;
        const uint32_t step = (uint32_t)src.B;
        const int rmax = (b < src.nb) ? nd.in_rows - 8 * hf : 0;     // this thread reads rows 16 i + e < rmax of its half
#pragma unroll
        for (int i = 0; i < NXO; ++i) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int rr = 16 * i + e;
                xr[i * 8 + e] = (rr < rmax) ? __ldcg(xp + (uint32_t)rr * step) : 0.0f;
            }
        }
    };
    // the first tile's rows are requested BEFORE the weights: their DRAM latency (5-6 k cycles between set-up and the first
    // tile's start, profiles/tc_timeline_r2.txt) then runs under the set-up instead of behind it
    if (warp < 8 && (int)blockIdx.x < units) load_x(blockIdx.x);
    load_weights_tc<C>(sm, nd, src.G);
    __syncthreads();
    if (TRAIN && tid < M) {
        // the ones row (row 0 of group 4 of the B image, hi = 1, lo = 0): its product with dH is the bias gradient
        *reinterpret_cast<float*>(sm + C::oBs + smaj(32, tid)) = 1.0f;
    }
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sbase = tc::smem_u32(sm);
    if (ktl) g_tc_timeline[21] = clock64();

    // One CTA per SM (the 64-wide training chains): the CTA is launched as THREE warpgroups -- the issue warp's group is
    // completed by three idle warps -- so that the groups can trade registers: a 9-warp CTA is capped at 168 per thread,
    // the issue group gives back what it does not need, the compute groups take 232 (see tc_gru.cu).
    // (ptxas budgets the two roles separately only if each role's whole branch is dominated by its own setmaxnreg)
    if (warp >= 8) {
        if (C::CTAS_PER_SM == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 8) {
        // ================================ MMA issue warp ==================================================
        {
            const uint32_t As = sbase + C::oAs, Bs_h = sbase + C::oBs, Bs_l = Bs_h + B_S_BYTES;
            constexpr int MR = 2 * H;                    // rows of the stacked dH operand
            int gpos = 0;                                // position of the tile in its flush group
            uint32_t par = 0;
            int it = 0;
            const bool leader = tc::elect_one();
            for (int u = blockIdx.x; u < units; u += gridDim.x, par ^= 1, ++it, gpos = (gpos + 1 == src.flush) ? 0 : gpos + 1) {
                const bool tl_on = g_tc_timeline_on == (OUT > 1 ? 2 : 1) && blockIdx.x == 0 && it == 1 && lane == 0;
                const uint32_t keep = gpos != 0 ? 1u : 0u;
                TL_STAMP(32);
                acquire(&bars[R_X], par);
                TL_STAMP(33);
                issue_ts<H, K1P>(leader, tmem + C::cD1, tmem + C::cXh, tmem + C::cXl, sbase + C::oW1, sbase + C::oW1 + C::szW1);
                if (leader) tc::mma_commit(&bars[D_F1]);
                TL_STAMP(34);
                acquire(&bars[R_H1], par);
                TL_STAMP(35);
                issue_ts<H, H>(leader, tmem + C::cD2, tmem + C::cAh, tmem + C::cAl, sbase + C::oW2, sbase + C::oW2 + C::szW2);
                if (leader) tc::mma_commit(&bars[D_F2]);
                TL_STAMP(36);
                if (TRAIN) {
                    acquire(&bars[R_DH2], par);
                    TL_STAMP(37);
                    issue_ts<H, H>(leader, tmem + C::cD2, tmem + C::cAh, tmem + C::cAl, sbase + C::oW2T, sbase + C::oW2T + C::szW2);
                    if (leader) tc::mma_commit(&bars[D_B1]);
                    TL_STAMP(38);
                    if (C::NR2 == 1) issue_ss<MR, 40>(leader, tmem + C::cW2, As, Bs_h, Bs_l, keep);
                    else             issue_ss<MR, 32>(leader, tmem + C::cW2, As, Bs_h, Bs_l, keep);
                    if (leader) tc::mma_commit(&bars[D_W2A]);
                    TL_STAMP(39);
                    if (C::NR2 == 2) {
                        acquire(&bars[R_W2B], par);
                        TL_STAMP(40);
                        issue_ss<MR, 40>(leader, tmem + C::cW2 + 32, As, Bs_h, Bs_l, keep);
                        if (leader) tc::mma_commit(&bars[D_W2B]);
                        TL_STAMP(41);
                    }
                    acquire(&bars[R_W1A], par);
                    TL_STAMP(42);
                    if (C::NR1 == 1) issue_ss<MR, 40>(leader, tmem + C::cW1, As, Bs_h, Bs_l, keep);
                    else             issue_ss<MR, 32>(leader, tmem + C::cW1, As, Bs_h, Bs_l, keep);
                    if (leader) tc::mma_commit(&bars[D_W1A]);
                    TL_STAMP(43);
                    if (C::NR1 == 2) {
                        acquire(&bars[R_W1B], par);
                        TL_STAMP(44);
                        issue_ss<MR, 40>(leader, tmem + C::cW1 + 32, As, Bs_h, Bs_l, keep);
                        if (leader) tc::mma_commit(&bars[D_W1B]);
                        TL_STAMP(45);
                    }
                }
            }
        }
        __syncwarp();
        }
    } else {
        if (C::CTAS_PER_SM == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ================================ compute warps ====================================================
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        const float* fb1 = reinterpret_cast<const float*>(sm + C::oB1);
        const float* fb2 = reinterpret_cast<const float*>(sm + C::oB2);
        const float* fw3 = reinterpret_cast<const float*>(sm + C::oW3T);
        const float* fb3 = reinterpret_cast<const float*>(sm + C::oB3);
        float* zx = reinterpret_cast<float*>(sm + C::oZx);              // [2][OUT][128] partial outputs of the two halves
        float* dw3acc = reinterpret_cast<float*>(sm + C::oDW3) + q * 8 * H;                     // [8][H] of this quadrant
        float* db3acc = reinterpret_cast<float*>(sm + C::oDW3) + 4 * 8 * H + q * 8;
        uint8_t* As_h = sm + C::oAs; uint8_t* As_l = As_h + C::A_S_BYTES;
        uint8_t* Bs_h = sm + C::oBs; uint8_t* Bs_l = Bs_h + B_S_BYTES;
        const int so = smaj(0, s);                                      // this sample's offset inside a feature row

        // this thread's row of the weight-gradient accumulators (M = 2H): feature `feat` of the A_hi (quadrants 0, 1) or A_lo
        // (quadrants 2, 3) products; an M = 64 accumulator keeps row r in lane (r / 16) * 32 + r % 16
        const int feat = (q & 1) * (H / 2) + lane;
        const bool lo_rows = q >= 2, row_ok = lane < H / 2;
        float* part_out = TRAIN ? partials + (size_t)blockIdx.x * (p_net + CMARL_N_STATS) : nullptr;
        bool flushed = false;                                        // the partial row holds earlier flushes

        float st[Head::NSTAT];
#pragma unroll
        for (int k = 0; k < Head::NSTAT; ++k) st[k] = 0.0f;


        // X stage: the prefetched rows (xr) -> split -> TMEM X columns, hand-off to the issuer (F1), then prefetch the
        // rows of the tile after.  It runs one tile AHEAD of the rest of the chain: as soon as the current tile no longer
        // reads the X columns (forward: after F1; train: after the last dW1 round has been staged) the next tile's X is
        // published, so its F1 runs under this tile's remaining epilogues instead of after them.
        auto stage_x = [&](int u_next_next) {
#pragma unroll
            for (int i = 0; i < NXO; ++i) {
                const int c = 2 * i + hf;
                if (c < NCX) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float h, l;
                        tc::split_tf32(xr[i * 8 + e], h, l);
                        hi[e] = __float_as_uint(h); lo[e] = __float_as_uint(l);
                    }
                    tmem_st8(tl + C::cXh + 8 * c, hi);
                    tmem_st8(tl + C::cXl + 8 * c, lo);
                }
            }
            publish(&bars[R_X]);
            if (u_next_next < units) load_x(u_next_next);               // latency hidden behind a whole tile
        };
        if ((int)blockIdx.x < units) stage_x(blockIdx.x + gridDim.x);

        uint32_t par = 0;
        int it = 0, gpos = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, par ^= 1, ++it, gpos = (gpos + 1 == src.flush) ? 0 : gpos + 1) {
            int bt, r, t, g;
            fast_divmod(u, tiles_b, inv_tb, r, bt);
            fast_divmod(r, src.G, inv_g, t, g);
            const int b = bt * M + s;
            const bool inb = b < src.nb;
            const bool tl_on = g_tc_timeline_on == (OUT > 1 ? 2 : 1) && blockIdx.x == 0 && it == 1 && tid == 0;
            if (ktl && it < 8) g_tc_timeline[48 + it] = clock64();       // start of tile `it`
            const bool has_next = u + (int)gridDim.x < units;
            TL_STAMP(0);
            TL_STAMP(1);
            // the head's per-sample inputs (half 0 evaluates the head): loaded now, used after F2
            const typename Head::In hin = Head::load(ha, t, g, b, src.G, src.B, inb && hf == 0);

            // ---- E1: H1 = relu(D1 + b1[g]) -> split -> TMEM A ---------------------------------------------
            // (every asm statement is a compiler barrier for memory operations: shared-memory constants are read BEFORE the
            // wait / TMEM load they would otherwise queue behind, so their latency hides under it)
            float bv1[NOWN][16];
            auto load_b1 = [&]() {
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) {
#pragma unroll
                    for (int i4 = 0; i4 < 16; i4 += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(fb1 + g * H + 16 * (2 * ci + hf) + i4);
                        bv1[ci][i4] = bb.x; bv1[ci][i4 + 1] = bb.y; bv1[ci][i4 + 2] = bb.z; bv1[ci][i4 + 3] = bb.w;
                    }
                }
            };
            load_b1();
            acquire(&bars[D_F1], par);
            TL_STAMP(2);
            uint32_t h1h[NOWN][16], h1l[NOWN][16];                       // kept for the sample-major copies
            {
                uint32_t v[NOWN][16];
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) tc::tmem_ld16(tl + C::cD1 + 16 * (2 * ci + hf), v[ci]);
                tc::tmem_wait_ld();
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float h1 = fmaxf(__uint_as_float(v[ci][i]) + bv1[ci][i], 0.0f);
                        float h, l;
                        tc::split_tf32(h1, h, l);
                        h1h[ci][i] = __float_as_uint(h); h1l[ci][i] = __float_as_uint(l);
                    }
                }
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) {
                    const int c0 = 16 * (2 * ci + hf);
                    tmem_st16(tl + C::cAh + c0, h1h[ci]);
                    tmem_st16(tl + C::cAl + c0, h1l[ci]);
                }
            }
            publish(&bars[R_H1]);
            TL_STAMP(3);
            if (!TRAIN && has_next) stage_x(u + 2 * gridDim.x);          // F1 of this tile has completed: X columns are free
            if (TRAIN) {
                // round-0 features of H1 (chunk hf), sample-major, while F2 runs (B image is free: the previous
                // tile's last weight-gradient round has completed)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    *reinterpret_cast<uint32_t*>(Bs_h + smaj(16 * hf + i, 0) + so) = h1h[0][i];
                    *reinterpret_cast<uint32_t*>(Bs_l + smaj(16 * hf + i, 0) + so) = h1l[0][i];
                }
                // ones group, rows 1 + g': 1 for the samples of agent group g' (= this tile's g): the product with dH1 is the
                // gradient of the id column folded into b1[g'].  (The previous tile's last round has completed: B is free.)
                if (nd.fold_ids && hf == 1) {
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg)
                        *reinterpret_cast<float*>(Bs_h + smaj(33 + gg, 0) + so) = (gg == g) ? 1.0f : 0.0f;
                }
            }

            // ---- E2: H2, output layer, head ---------------------------------------------------------------
            TL_STAMP(4);
            float bv2[NOWN][16];
#pragma unroll
            for (int ci = 0; ci < NOWN; ++ci) {
#pragma unroll
                for (int i4 = 0; i4 < 16; i4 += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(fb2 + 16 * (2 * ci + hf) + i4);
                    bv2[ci][i4] = bb.x; bv2[ci][i4 + 1] = bb.y; bv2[ci][i4 + 2] = bb.z; bv2[ci][i4 + 3] = bb.w;
                }
            }
            acquire(&bars[D_F2], par);
            TL_STAMP(5);
            float h2[NOWN * 16];
            float z[OUT], dz[OUT];
#pragma unroll
            for (int a = 0; a < OUT; ++a) z[a] = 0.0f;
            {
                uint32_t v[NOWN][16];
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) tc::tmem_ld16(tl + C::cD2 + 16 * (2 * ci + hf), v[ci]);
                tc::tmem_wait_ld();
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = 16 * (2 * ci + hf) + i;
                        const float hv = fmaxf(__uint_as_float(v[ci][i]) + bv2[ci][i], 0.0f);
                        h2[ci * 16 + i] = hv;
                        if (OUT > 1) {
                            const float4 w = *reinterpret_cast<const float4*>(fw3 + j * 8);
                            const float w4 = fw3[j * 8 + 4];
                            z[0] = fmaf(w.x, hv, z[0]);
                            z[OUT > 1 ? 1 : 0] = fmaf(w.y, hv, z[OUT > 1 ? 1 : 0]);
                            z[OUT > 2 ? 2 : 0] = fmaf(w.z, hv, z[OUT > 2 ? 2 : 0]);
                            z[OUT > 3 ? 3 : 0] = fmaf(w.w, hv, z[OUT > 3 ? 3 : 0]);
                            z[OUT > 4 ? 4 : 0] = fmaf(w4, hv, z[OUT > 4 ? 4 : 0]);
                        } else {
                            z[0] = fmaf(fw3[j * 8], hv, z[0]);
                        }
                    }
                }
            }
            // half 1 hands its partial outputs to half 0, which evaluates the head and hands dz back
            if (hf == 1) {
#pragma unroll
                for (int a = 0; a < OUT; ++a) zx[a * M + s] = z[a];
            }
            compute_bar();
            if (hf == 0) {
#pragma unroll
                for (int a = 0; a < OUT; ++a) z[a] = (z[a] + zx[a * M + s]) + fb3[a];
                Head::compute(ha, hin, z, TRAIN, dz, st);
                if (TRAIN) {
#pragma unroll
                    for (int a = 0; a < OUT; ++a) zx[a * M + s] = dz[a];
                }
            }
            if (TRAIN) {
                compute_bar();
                if (hf == 1) {
#pragma unroll
                    for (int a = 0; a < OUT; ++a) dz[a] = zx[a * M + s];
                }
            }
            __syncwarp();
            TL_STAMP(6);

            if (TRAIN) {
                // dH2 = (W3^T dz) . relu'(H2) -> split -> TMEM A (operand of B1) + sample-major A image (operand of dW2).
                // A chunk of 16 is computed before its first store: a shared-memory store in between would pin every later
                // W3 load behind it (the compiler cannot prove they do not alias) and serialise 16 LDS round trips per chunk.
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) {
                    const int c0 = 16 * (2 * ci + hf);
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = c0 + i;
                        float acc;
                        if (OUT > 1) {
                            const float4 w = *reinterpret_cast<const float4*>(fw3 + j * 8);
                            const float w4 = fw3[j * 8 + 4];
                            acc = w.x * dz[0];
                            acc = fmaf(w.y, dz[OUT > 1 ? 1 : 0], acc);
                            acc = fmaf(w.z, dz[OUT > 2 ? 2 : 0], acc);
                            acc = fmaf(w.w, dz[OUT > 3 ? 3 : 0], acc);
                            acc = fmaf(w4, dz[OUT > 4 ? 4 : 0], acc);
                        } else {
                            acc = fw3[j * 8] * dz[0];
                        }
                        const float d = h2[ci * 16 + i] > 0.0f ? acc : 0.0f;
                        float h, l;
                        tc::split_tf32(d, h, l);
                        hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(l);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        *reinterpret_cast<uint32_t*>(As_h + smaj(c0 + i, 0) + so) = hi[i];
                        *reinterpret_cast<uint32_t*>(As_l + smaj(c0 + i, 0) + so) = lo[i];
                    }
                    tmem_st16(tl + C::cAh + c0, hi);
                    tmem_st16(tl + C::cAl + c0, lo);
                }
                publish(&bars[R_DH2]);          // issuer: B1, then dW2 round 0
                TL_STAMP(8);
                // (runs under B1 / dW2 round 0 on the tensor pipe)
                // dW3[a][j] += sum_s dz[s][a] h2[s][j] over this thread's columns; db3[a] += sum_s dz[s][a]
#pragma unroll
                for (int a = 0; a < OUT; ++a) {
                    float p[NOWN * 16];
#pragma unroll
                    for (int i = 0; i < NOWN * 16; ++i) p[i] = dz[a] * h2[i];
                    int idx;
                    warp_reduce_scatter<NOWN * 16>(p, lane, idx);
                    const int j = 16 * (2 * (idx >> 4) + hf) + (idx & 15);
                    if (NOWN * 16 >= 32 || (lane & 1) == 0) dw3acc[a * H + j] += p[0];
                    if (hf == 0) {
                        float d = dz[a];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                        if (lane == 0) db3acc[a] += d;
                    }
                }
                TL_STAMP(7);

                // ---- E3: dH1 = D3 . relu'(H1) (H1 > 0 <=> D1 + b1 > 0) --------------------------------------
                load_b1();
                acquire(&bars[D_B1], par);
                TL_STAMP(9);
                float dh1[NOWN * 16];
                uint32_t r1h[16], r1l[16];                               // round-1 features of H1 (NR2 == 2)
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) {
                    const int c0 = 16 * (2 * ci + hf);
                    uint32_t v[16], d1[16];
                    tc::tmem_ld16(tl + C::cD2 + c0, v);
                    tc::tmem_ld16(tl + C::cD1 + c0, d1);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float pre = __uint_as_float(d1[i]) + bv1[ci][i];
                        dh1[ci * 16 + i] = pre > 0.0f ? __uint_as_float(v[i]) : 0.0f;
                        if (C::NR2 == 2 && ci == 1) {
                            float h, l;
                            tc::split_tf32(fmaxf(pre, 0.0f), h, l);
                            r1h[i] = __float_as_uint(h); r1l[i] = __float_as_uint(l);
                        }
                    }
                }
                TL_STAMP(10);
                acquire(&bars[D_W2A], par);     // dW2 round 0 complete: the B image is free
                TL_STAMP(11);
                if (C::NR2 == 2) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        *reinterpret_cast<uint32_t*>(Bs_h + smaj(16 * hf + i, 0) + so) = r1h[i];
                        *reinterpret_cast<uint32_t*>(Bs_l + smaj(16 * hf + i, 0) + so) = r1l[i];
                    }
                    publish(&bars[R_W2B]);
                    TL_STAMP(12);
                    acquire(&bars[D_W2B], par);
                    TL_STAMP(13);
                }
                // ---- dW1 = dH1^T X: A image <- dH1, B image <- X in rounds of 32 features -----------------------
#pragma unroll
                for (int ci = 0; ci < NOWN; ++ci) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = 16 * (2 * ci + hf) + i;
                        float h, l;
                        tc::split_tf32(dh1[ci * 16 + i], h, l);
                        *reinterpret_cast<float*>(As_h + smaj(j, 0) + so) = h;
                        *reinterpret_cast<float*>(As_l + smaj(j, 0) + so) = l;
                    }
                }
#pragma unroll
                for (int rd = 0; rd < C::NR1; ++rd) {
                    // this thread's X chunks of the round: 4 rd + hf and 4 rd + 2 + hf -> B rows 8 hf.. and 16 + 8 hf..
#pragma unroll
                    for (int h2i = 0; h2i < 2; ++h2i) {
                        const int c = 4 * rd + 2 * h2i + hf;             // X chunk (may not exist: zeros)
                        uint32_t xh[8], xl[8];
                        if (c < NCX) {
                            tc::tmem_ld8(tl + C::cXh + 8 * c, xh);
                            tc::tmem_ld8(tl + C::cXl + 8 * c, xl);
                            tc::tmem_wait_ld();
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) { xh[i] = 0u; xl[i] = 0u; }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = 16 * h2i + 8 * hf + i;
                            *reinterpret_cast<uint32_t*>(Bs_h + smaj(row, 0) + so) = xh[i];
                            *reinterpret_cast<uint32_t*>(Bs_l + smaj(row, 0) + so) = xl[i];
                        }
                    }
                    publish(&bars[rd == 0 ? R_W1A : R_W1B]);
                    TL_STAMP(14 + 2 * rd);
                    // last read of this tile's X columns is behind us: stage the next tile's X under this round's MMAs
                    if (rd == C::NR1 - 1 && has_next) stage_x(u + 2 * gridDim.x);
                    acquire(&bars[rd == 0 ? D_W1A : D_W1B], par);
                    TL_STAMP(15 + 2 * rd);
                }
                // ---- every src.flush tiles (and after the CTA's last one): accumulators -> the CTA's partial row ------------
                if (gpos + 1 == src.flush || !has_next) {
                    // the A images are free between two tiles: scratch [H][pS2] | [H][pS1]; the A_lo row block first, the A_hi
                    // block added to it (fixed order), then one coalesced pass over the parameter vector
                    if (ktl && !flushed) g_tc_timeline[24] = clock64();
                    float* S2 = reinterpret_cast<float*>(sm + C::oAs);
                    float* S1 = S2 + H * C::pS2;
                    float* Sm = hf == 0 ? S2 : S1;
                    const int cb = hf == 0 ? C::cW2 : C::cW1, ps = hf == 0 ? C::pS2 : C::pS1, cn = hf == 0 ? C::nW2 : C::nW1;
                    constexpr int FCH = (C::nW2 > C::nW1 ? C::nW2 : C::nW1) / 8;   // 8-column chunks per block (5 or 9)
#pragma unroll 1
                    for (int ph = 0; ph < 2; ++ph) {
                        if (lo_rows == (ph == 0)) {                     // warp-uniform
                            uint32_t v[FCH][8];
#pragma unroll
                            for (int k = 0; k < FCH; ++k)
                                if (8 * k < cn) tc::tmem_ld8(tl + cb + 8 * k, v[k]);          // warp-uniform guard
                            tc::tmem_wait_ld();
                            if (row_ok) {
                                float* p = Sm + feat * ps;
#pragma unroll
                                for (int k = 0; k < FCH; ++k)
                                    if (8 * k < cn) {
#pragma unroll
                                        for (int i = 0; i < 8; ++i)
                                            p[8 * k + i] = ph == 0 ? __uint_as_float(v[k][i]) : p[8 * k + i] + __uint_as_float(v[k][i]);
                                    }
                            }
                        }
                        compute_bar();
                        if (ktl && !flushed) g_tc_timeline[25 + ph] = clock64();
                    }
                    {
                        // one coalesced pass over W1 | b1 | W2 | b2 of the partial row: straight-line, predicated code (every
                        // scratch read and every load of the row's earlier value in flight before the first store)
                        const int in_dim = nd.in_dim, in_rows = nd.in_rows;
                        const int nW1e = H * in_dim;
                        const float inv_in = 1.0f / (float)in_dim;
                        const int ct = warp * 32 + lane;
                        constexpr int N1 = (H * (K1P + 4) + 255) / 256, N2 = H * H / 256;
                        float* oW1 = part_out;
                        float* oB = part_out + nW1e;                     // b1 [H], then (after W2) b2 [H]
                        float* oW2 = oB + H;
                        float v1[N1], v2[N2], vb = 0.0f;
#pragma unroll
                        for (int r = 0; r < N1; ++r) {
                            const int i = ct + 256 * r;
                            int jj, k;
                            fast_divmod(i < nW1e ? i : 0, in_dim, inv_in, jj, k);
                            const int col = k < in_rows ? k : 32 * C::NR1 + 1 + (k - in_rows);     // k >= in_rows: folded id column
                            v1[r] = S1[jj * C::pS1 + col];
                        }
#pragma unroll
                        for (int r = 0; r < N2; ++r) {
                            const int i = ct + 256 * r;
                            v2[r] = S2[(i / H) * C::pS2 + (i % H)];
                        }
                        if (ct < H) vb = S1[ct * C::pS1 + 32 * C::NR1];
                        else if (ct < 2 * H) vb = S2[(ct - H) * C::pS2 + 32 * C::NR2];
                        float* ob = ct < H ? oB + ct : oW2 + H * H + (ct - H);
                        if (flushed) {
                            float o1[N1], o2[N2], ob0 = 0.0f;
#pragma unroll
                            for (int r = 0; r < N1; ++r) o1[r] = (ct + 256 * r < nW1e) ? oW1[ct + 256 * r] : 0.0f;
#pragma unroll
                            for (int r = 0; r < N2; ++r) o2[r] = oW2[ct + 256 * r];
                            if (ct < 2 * H) ob0 = *ob;
#pragma unroll
                            for (int r = 0; r < N1; ++r) v1[r] += o1[r];
#pragma unroll
                            for (int r = 0; r < N2; ++r) v2[r] += o2[r];
                            vb += ob0;
                        }
#pragma unroll
                        for (int r = 0; r < N1; ++r)
                            if (ct + 256 * r < nW1e) oW1[ct + 256 * r] = v1[r];
#pragma unroll
                        for (int r = 0; r < N2; ++r) oW2[ct + 256 * r] = v2[r];
                        if (ct < 2 * H) *ob = vb;
                    }
                    if (ktl && !flushed) g_tc_timeline[27] = clock64();
                    flushed = true;
                    TL_STAMP(18);
                    compute_bar();      // scratch reads done before the next tile's dH2 goes to the A images
                }
            }
        }

        if (ktl) g_tc_timeline[22] = clock64();
        // ---- the rest of this CTA's partial row: output-layer gradients (shared-memory sums), then the statistics; the
        //      hidden-layer gradients went there with the flushes (the last tile always flushes) ------------------------
        if (TRAIN) {
            compute_bar();
            float* out = part_out;
            float* gW3 = out + (H * nd.in_dim + H + H * H + H);
            float* gb3 = gW3 + nd.out_dim * H;
            const int ct = warp * 32 + lane;                        // 0..255
            const float* w3all = reinterpret_cast<const float*>(sm + C::oDW3);
            for (int i = ct; i < nd.out_dim * H; i += NCOMP)
                gW3[i] = ((w3all[i] + w3all[8 * H + i]) + w3all[2 * 8 * H + i]) + w3all[3 * 8 * H + i];
            if (ct < nd.out_dim) {
                const float* b3all = w3all + 4 * 8 * H;
                gb3[ct] = ((b3all[ct] + b3all[8 + ct]) + b3all[16 + ct]) + b3all[24 + ct];
            }
            float* red = reinterpret_cast<float*>(sm + C::oRed);
#pragma unroll
            for (int k = 0; k < Head::NSTAT; ++k) {
                const float v = warp_sum_f(st[k]);                  // half-1 warps carry zeros
                compute_bar();
                if (lane == 0) red[warp] = v;
                compute_bar();
                if (ct == 0) out[p_net + k] = ((red[0] + red[1]) + red[2]) + red[3];
            }
            if (ct == 0)
                for (int k = Head::NSTAT; k < CMARL_N_STATS; ++k) out[p_net + k] = 0.0f;
        }
    }
    if (ktl) g_tc_timeline[23] = clock64();
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem, C::TMEM_COLS);
    // ... but whoever follows must see BOTH this grid and the one in front of it complete: it completes only after that one
    if (src.indep) pdl_wait();
}

// The tensor-core chains evaluate the policy head with exp / log / the softmax divisions on the special-function unit
// (heads.cuh, PolicyHeadT<true>: <= 2 ulp each; the head runs on half of a CTA's compute warps while the others wait).
template <class Hd> struct TcHead { using type = Hd; };
template <> struct TcHead<PolicyHead> { using type = PolicyHeadT<true>; };

template <class C, class Head0>
static int tc_set_attr() {
    using Head = typename TcHead<Head0>::type;
    return cmarl_check_cuda(cudaFuncSetAttribute(tc_chain_kernel<C, Head>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 C::smem_bytes),
                            "cudaFuncSetAttribute(tc_chain_kernel)");
}

// `src.indep` launches (the critic chain of an epoch) are programmatic dependents of the kernel in front of them whatever
// the context's launch-chaining switch says; everything else follows the switch (cmarl_launch).
template <class C, class Head0>
static int tc_launch(const cmarl_ctx* ctx, const NetDesc& nd, const TileSrc& src, const typename Head0::Args& ha, float* partials,
                     int p_net, int grid, cudaStream_t st) {
    using Head = typename TcHead<Head0>::type;
    return cmarl_check_cuda(cmarl_launch_pdl(src.indep || cmarl_chained(ctx), tc_chain_kernel<C, Head>, dim3(grid), dim3(C::CTAS_PER_SM == 1 ? NTHREADS + 96 : NTHREADS),
                                             C::smem_bytes, st, nd, src, ha, partials, p_net),
                            "tc_chain_kernel launch");
}

}  // namespace tcchain

using namespace tcchain;

int cmarl_tc_setup() {
    int e = 0;
#define SET(Hh, Kk)                                                             \
    if (!e) e = tc_set_attr<TCfg<Hh, Kk, true, 5>, PolicyHead>();               \
    if (!e) e = tc_set_attr<TCfg<Hh, Kk, true, 1>, ValueHead>();                \
    if (!e) e = tc_set_attr<TCfg<Hh, Kk, false, 1>, ValueHead>();
    SET(32, 24) SET(32, 56) SET(64, 24) SET(64, 56)
#undef SET
    return e;
}

int cmarl_tc_tile() { return M; }

// co-resident CTAs per SM of the kernel that dispatch<Head, TRAIN> would launch (sizes the persistent grid)
int cmarl_tc_ctas_per_sm(int H, int in_rows, bool train, int out) {
    const int kin = in_rows <= 24 ? 24 : 56;
#define CPS(Hh, Kk)                                                                                         \
    if (H == Hh && kin == Kk)                                                                               \
        return train ? (out > 1 ? TCfg<Hh, Kk, true, 5>::CTAS_PER_SM : TCfg<Hh, Kk, true, 1>::CTAS_PER_SM)  \
                     : TCfg<Hh, Kk, false, 1>::CTAS_PER_SM;
    CPS(32, 24) CPS(32, 56) CPS(64, 24) CPS(64, 56)
#undef CPS
    return 1;
}

// debug aid (not in the public header): enable / read the clock64 timeline of CTA 0's second tile
extern "C" int cmarl_debug_tc_timeline(int enable, long long* out_host64) {
    cudaError_t e = cudaMemcpyToSymbol(g_tc_timeline_on, &enable, sizeof(int));
    if (e == cudaSuccess && out_host64) e = cudaMemcpyFromSymbol(out_host64, g_tc_timeline, sizeof(long long) * 64);
    return (int)e;
}

template <class Head, bool TRAIN>
int cmarl_tc_dispatch(const cmarl_ctx* ctx, int H, const NetDesc& nd, const TileSrc& src, const typename Head::Args& ha,
                      float* partials, int p_net, int grid, cudaStream_t st) {
    const int kin = nd.in_rows <= 24 ? 24 : 56;
    if (H == 32 && kin == 24) return tc_launch<TCfg<32, 24, TRAIN, Head::OUT>, Head>(ctx, nd, src, ha, partials, p_net, grid, st);
    if (H == 32 && kin == 56) return tc_launch<TCfg<32, 56, TRAIN, Head::OUT>, Head>(ctx, nd, src, ha, partials, p_net, grid, st);
    if (H == 64 && kin == 24) return tc_launch<TCfg<64, 24, TRAIN, Head::OUT>, Head>(ctx, nd, src, ha, partials, p_net, grid, st);
    if (H == 64 && kin == 56) return tc_launch<TCfg<64, 56, TRAIN, Head::OUT>, Head>(ctx, nd, src, ha, partials, p_net, grid, st);
    cmarl_set_error("tc dispatch: unsupported hidden=%d in_rows=%d", H, nd.in_rows);
    return -1;
}

template int cmarl_tc_dispatch<PolicyHead, true>(const cmarl_ctx*, int, const NetDesc&, const TileSrc&, const PolicyHeadArgs&, float*, int, int, cudaStream_t);
template int cmarl_tc_dispatch<ValueHead, true>(const cmarl_ctx*, int, const NetDesc&, const TileSrc&, const ValueHeadArgs&, float*, int, int, cudaStream_t);
template int cmarl_tc_dispatch<ValueHead, false>(const cmarl_ctx*, int, const NetDesc&, const TileSrc&, const ValueHeadArgs&, float*, int, int, cudaStream_t);
