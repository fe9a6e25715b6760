// Tensor-core (tcgen05 / TMEM) version of the fused MLP chain: K4 critic forward and K7 PPO epoch
// (actor and critic forward + head + backward) with every GEMM on the 5th-gen tensor cores.
//
// Precision: kind::tf32 with a 3-term split ("3xTF32"): x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi)
// (residual <= 2^-24 |x|), D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi with fp32 accumulation in TMEM, the two
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
//   needed sample-major ([feature][sample]) in shared memory: A = dH (64 feature rows), B = H1 / X in rounds
//   of 32 feature rows (+ a constant "ones" row group whose product is the bias gradient), M = 64 MMAs
//   into per-tile accumulators; the per-tile sums are added to running sums kept in the TMEM lanes the
//   M = 64 accumulator layout leaves unused (lanes 16..31 of every quadrant).
#include "chain.cuh"
#include "heads.cuh"
#include "tc_ptx.cuh"

namespace tcchain {

using namespace chain;

constexpr int M = 128;                        // samples per tile = UMMA M = TMEM lanes
constexpr int NT = 128;                       // threads per CTA
constexpr int LBO_K = 128;                    // feature-major (weights): next chunk of 4 K elements
constexpr int LBO_S = 144;                    // sample-major: next chunk of 4 samples (128 B + 16 B pad: conflict-free STS.32)
constexpr int SBO_S = (M / 4) * LBO_S;        // 4608: next group of 8 feature rows
constexpr int KSTEP_S = 2 * LBO_S;            // one MMA consumes 8 samples
constexpr int A_GROUPS = 8;                   // dH operand: 64 feature rows
constexpr int B_GROUPS = 5;                   // H1 / X operand: 32 feature rows + the ones group
constexpr int A_S_BYTES = A_GROUPS * SBO_S;   // per hi / lo image
constexpr int B_S_BYTES = B_GROUPS * SBO_S;

template <int H_, int K1P_, bool TRAIN_>
struct TCfg {
    static constexpr int H = H_, K1P = K1P_;
    static constexpr bool TRAIN = TRAIN_;
    static constexpr int NR2 = H / 32;                    // rounds of the dW2 GEMM (32 H1 features each)
    static constexpr int NR1 = (K1P + 31) / 32;           // rounds of the dW1 GEMM
    // TMEM columns
    static constexpr int cXh = 0, cXl = K1P, cAh = 2 * K1P, cAl = cAh + H, cD1 = cAl + H, cD2 = cD1 + H;
    static constexpr int cW2 = cD2 + H;                   // per-tile dW2 | db2 (lanes 0-15 of each quadrant), running sums in lanes 16-31
    static constexpr int nW2 = 32 * NR2 + 8;
    static constexpr int cW1 = cW2 + nW2;
    static constexpr int nW1 = 32 * NR1 + 8;
    static constexpr int cEnd = TRAIN ? cW1 + nW1 : cD2 + H;
    static constexpr int TMEM_COLS = cEnd <= 32 ? 32 : cEnd <= 64 ? 64 : cEnd <= 128 ? 128 : cEnd <= 256 ? 256 : 512;
    static_assert(cEnd <= 512, "TMEM budget");
    // shared memory (bytes)
    static constexpr int oBar = 0;                        // mbarrier (8 B) + tmem base (4 B)
    static constexpr int oW1 = 64;                        // W1 hi | lo   [H][K1P]   K-major core-matrix image
    static constexpr int szW1 = H * K1P * 4;
    static constexpr int oW2 = oW1 + 2 * szW1;            // W2 hi | lo   [H][H]
    static constexpr int szW2 = H * H * 4;
    static constexpr int oW2T = oW2 + 2 * szW2;           // W2^T hi | lo (train)
    static constexpr int oB1 = oW2T + (TRAIN ? 2 * szW2 : 0);   // f32 [4][H]
    static constexpr int oB2 = oB1 + 4 * H * 4;           // f32 [H]
    static constexpr int oW3T = oB2 + H * 4;              // f32 [H][8]
    static constexpr int oB3 = oW3T + H * 8 * 4;          // f32 [8]
    static constexpr int oDW3 = oB3 + 32;                 // f32 [4 warps][8][H] + [4][8]: dW3 / db3 partial sums per warp
    static constexpr int oDId = oDW3 + (TRAIN ? (4 * 8 * H + 32) * 4 : 0);       // f32 [4][H] folded id-column gradients per agent group
    static constexpr int oRed = oDId + (TRAIN ? 4 * H * 4 : 0);                 // f32 [64]
    static constexpr int oAs = ((oRed + 256 + 127) / 128) * 128;                // dH sample-major hi | lo
    static constexpr int oBs = oAs + (TRAIN ? 2 * A_S_BYTES : 0);               // H1 / X sample-major hi | lo
    static constexpr int smem_bytes = oBs + (TRAIN ? 2 * B_S_BYTES : 0);
    static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
};

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

// byte offset of element (feature row r, sample s) in a sample-major image
__device__ __forceinline__ int smaj(int r, int s) { return (r >> 3) * SBO_S + (r & 7) * 16 + (s >> 2) * LBO_S + (s & 3) * 4; }
// byte offset of element (row n, k) in a K-major weight image with KTOT columns
__device__ __forceinline__ int kmaj(int n, int k, int ktot) { return (n >> 3) * (ktot / 4) * LBO_K + (n & 7) * 16 + (k >> 2) * LBO_K + (k & 3) * 4; }

__device__ __forceinline__ void mbar_wait_trap(uint64_t* bar, uint32_t parity) {
    // bounded: a descriptor / protocol bug must surface as a launch failure, never as a hung GPU
    if (!tc::mbar_wait_bounded(bar, parity, 1u << 26)) __trap();
}

// ------------------------------------------------------------------------------------------------
// MMA issue (ONE thread).  3xTF32 pass order: (lo,hi), (hi,lo), (hi,hi).
// ------------------------------------------------------------------------------------------------
// D[128 x N] = A(TMEM, K columns at a_hi / a_lo) * B(smem K-major image [N][K])^T
template <int N, int K>
__device__ __forceinline__ void issue_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
    constexpr uint32_t sbo = (K / 4) * LBO_K;
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = pass == 0 ? a_lo : a_hi;
        const uint64_t db0 = tc::make_smem_desc(pass == 1 ? b_lo : b_hi, LBO_K, sbo, 0);
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks) {
            tc::mma_tf32_ts(d, a + ks * 8, db0 + (uint64_t)((ks * 2 * LBO_K) >> 4), idesc, acc);
            acc = 1;
        }
    }
}
// D[64 x N] = A(smem sample-major [64][128]) * B(smem sample-major [N][128])^T, contraction over the 128 samples
template <int N>
__device__ __forceinline__ void issue_ss(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(64, N, 0, 0);
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint64_t da0 = tc::make_smem_desc(pass == 0 ? a_lo : a_hi, LBO_S, SBO_S, 0);
        const uint64_t db0 = tc::make_smem_desc(pass == 1 ? b_lo : b_hi, LBO_S, SBO_S, 0);
#pragma unroll
        for (int ks = 0; ks < M / 8; ++ks) {
            const uint64_t off = (uint64_t)((ks * KSTEP_S) >> 4);
            tc::mma_tf32(d, da0 + off, db0 + off, idesc, acc);
            acc = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Weights -> shared memory images (hi | lo)
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ void load_weights_tc(uint8_t* sm, const NetDesc& nd, int n_groups) {
    constexpr int H = C::H, K1P = C::K1P;
    const float* P = nd.params;
    const int in_dim = nd.in_dim;
    const float* W1 = P;
    const float* b1 = W1 + H * in_dim;
    const float* W2 = b1 + H;
    const float* b2 = W2 + H * H;
    const float* W3 = b2 + H;
    const float* b3 = W3 + nd.out_dim * H;
    for (int i = threadIdx.x; i < H * K1P; i += NT) {
        const int j = i / K1P, k = i - j * K1P;
        const float w = (k < nd.in_rows) ? W1[j * in_dim + k] : 0.0f;
        float hi, lo;
        tc::split_tf32(w, hi, lo);
        const int o = kmaj(j, k, K1P);
        *reinterpret_cast<float*>(sm + C::oW1 + o) = hi;
        *reinterpret_cast<float*>(sm + C::oW1 + C::szW1 + o) = lo;
    }
    for (int i = threadIdx.x; i < H * H; i += NT) {
        const int j = i / H, k = i - j * H;
        float hi, lo;
        tc::split_tf32(W2[i], hi, lo);
        const int o = kmaj(j, k, H);                     // B of F2: N = j, K = k
        *reinterpret_cast<float*>(sm + C::oW2 + o) = hi;
        *reinterpret_cast<float*>(sm + C::oW2 + C::szW2 + o) = lo;
        if (C::TRAIN) {
            const int ot = kmaj(k, j, H);                // B of B1: N = k, K = j
            *reinterpret_cast<float*>(sm + C::oW2T + ot) = hi;
            *reinterpret_cast<float*>(sm + C::oW2T + C::szW2 + ot) = lo;
        }
    }
    float* fb1 = reinterpret_cast<float*>(sm + C::oB1);
    for (int i = threadIdx.x; i < 4 * H; i += NT) {
        const int g = i / H, j = i - g * H;
        float v = b1[j];
        if (nd.fold_ids && g < n_groups) v += W1[j * in_dim + nd.in_rows + g];
        fb1[i] = v;
    }
    float* fb2 = reinterpret_cast<float*>(sm + C::oB2);
    for (int i = threadIdx.x; i < H; i += NT) fb2[i] = b2[i];
    float* fw3 = reinterpret_cast<float*>(sm + C::oW3T);
    for (int i = threadIdx.x; i < H * 8; i += NT) {
        const int j = i / 8, a = i - j * 8;
        fw3[i] = (a < nd.out_dim) ? W3[a * H + j] : 0.0f;
    }
    float* fb3 = reinterpret_cast<float*>(sm + C::oB3);
    if (threadIdx.x < 8) fb3[threadIdx.x] = (threadIdx.x < nd.out_dim) ? b3[threadIdx.x] : 0.0f;
}

// sum over the 32 lanes of NV per-lane values: afterwards lane l holds the totals of values
// {l, l + 32, ...} in v[0], v[1], ... (butterfly reduce-scatter: NV/2 + NV/4 + ... shuffles)
template <int NV>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[NV], int lane) {
    static_assert(NV % 32 == 0, "NV must be a multiple of 32");
#pragma unroll
    for (int w = 16, n = NV; w >= 1; w >>= 1, n >>= 1) {
        const bool upper = (lane & w) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            // lanes with bit w clear keep the even half [i], the others the odd half [i + n/2]
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
template <class C, class Head>
__global__ void __launch_bounds__(NT, 1)
tc_chain_kernel(NetDesc nd, TileSrc src, typename Head::Args ha, float* __restrict__ partials, int p_net) {
    extern __shared__ __align__(1024) uint8_t sm[];
    constexpr int H = C::H, K1P = C::K1P, OUT = Head::OUT;
    constexpr bool TRAIN = C::TRAIN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + C::oBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::oBar + 16);
    const int tiles_b = (src.B + M - 1) / M;
    const int units = src.T * src.G * tiles_b;

    load_weights_tc<C>(sm, nd, src.G);
    if (TRAIN) {
        for (int i = tid * 16; i < 2 * A_S_BYTES + 2 * B_S_BYTES; i += NT * 16)
            *reinterpret_cast<uint4*>(sm + C::oAs + i) = make_uint4(0, 0, 0, 0);
        float* z0 = reinterpret_cast<float*>(sm + C::oDW3);
        for (int i = tid; i < 4 * 8 * H + 32 + 4 * H; i += NT) z0[i] = 0.0f;
    }
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
    __syncthreads();
    if (TRAIN) {
        // the ones row (row 0 of group 4 of the B image, hi = 1, lo = 0): its product with dH is the bias gradient
        float* ones = reinterpret_cast<float*>(sm + C::oBs + smaj(32, tid));
        *ones = 1.0f;
    }
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);       // this thread's lane
    const uint32_t sbase = tc::smem_u32(sm);
    uint32_t phase = 0;

    if (TRAIN) {   // running gradient sums (lanes 16-31 of each quadrant) start at zero
        uint32_t zero[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) zero[i] = 0u;
        for (int c = C::cW2; c < C::cEnd; c += 8) tmem_st8(tl + c, zero);
        tc::tmem_wait_st();
    }

    const float* fb1 = reinterpret_cast<const float*>(sm + C::oB1);
    const float* fb2 = reinterpret_cast<const float*>(sm + C::oB2);
    const float* fw3 = reinterpret_cast<const float*>(sm + C::oW3T);
    const float* fb3 = reinterpret_cast<const float*>(sm + C::oB3);
    float* dw3acc = reinterpret_cast<float*>(sm + C::oDW3) + warp * 8 * H;                 // [8][H] of this warp
    float* db3acc = reinterpret_cast<float*>(sm + C::oDW3) + 4 * 8 * H + warp * 8;
    float* didacc = reinterpret_cast<float*>(sm + C::oDId);

    float st[Head::NSTAT];
#pragma unroll
    for (int k = 0; k < Head::NSTAT; ++k) st[k] = 0.0f;

    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int bt = u % tiles_b, r = u / tiles_b;
        const int t = r / src.G, g = r % src.G, b0 = bt * M;
        const int b = b0 + tid;
        const bool inb = b < src.B;

        // ---- X: coalesced loads (row k of the tile = 128 consecutive floats), split, TMEM ------------
        {
            const float* xp = src.x + (size_t)t * src.stride_t + (size_t)g * src.stride_g + b;
            float x[K1P];
#pragma unroll
            for (int k = 0; k < K1P; ++k) x[k] = (inb && k < nd.in_rows) ? __ldg(xp + (size_t)k * src.B) : 0.0f;
#pragma unroll
            for (int c0 = 0; c0 < K1P; c0 += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float h, l;
                    tc::split_tf32(x[c0 + i], h, l);
                    hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(l);
                }
                tmem_st8(tl + C::cXh + c0, hi);
                tmem_st8(tl + C::cXl + c0, lo);
            }
            tc::tmem_wait_st();
        }
        tc::tcgen05_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tcgen05_fence_after();
            issue_ts<H, K1P>(tmem + C::cD1, tmem + C::cXh, tmem + C::cXl, sbase + C::oW1, sbase + C::oW1 + C::szW1);
            tc::mma_commit(bar);
        }
        mbar_wait_trap(bar, phase); phase ^= 1;
        tc::tcgen05_fence_after();

        // ---- E1: H1 = relu(D1 + b1[g]) -> split -> TMEM A; (train) round-0 features also sample-major into B ----
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 16) {
            uint32_t v[16], hi[16], lo[16];
            tc::tmem_ld16(tl + C::cD1 + c0, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float h1 = fmaxf(__uint_as_float(v[i]) + fb1[g * H + c0 + i], 0.0f);
                float h, l;
                tc::split_tf32(h1, h, l);
                hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(l);
                if (TRAIN && c0 < 32) {
                    *reinterpret_cast<float*>(sm + C::oBs + smaj(c0 + i, tid)) = h;
                    *reinterpret_cast<float*>(sm + C::oBs + B_S_BYTES + smaj(c0 + i, tid)) = l;
                }
            }
            tmem_st16(tl + C::cAh + c0, hi);
            tmem_st16(tl + C::cAl + c0, lo);
        }
        tc::tmem_wait_st();
        tc::tcgen05_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tcgen05_fence_after();
            issue_ts<H, H>(tmem + C::cD2, tmem + C::cAh, tmem + C::cAl, sbase + C::oW2, sbase + C::oW2 + C::szW2);
            tc::mma_commit(bar);
        }
        mbar_wait_trap(bar, phase); phase ^= 1;
        tc::tcgen05_fence_after();

        // ---- E2: H2, output layer, head ------------------------------------------------------------------
        float h2[H];
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 16) {
            uint32_t v[16];
            tc::tmem_ld16(tl + C::cD2 + c0, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) h2[c0 + i] = fmaxf(__uint_as_float(v[i]) + fb2[c0 + i], 0.0f);
        }
        float z[OUT], dz[OUT];
#pragma unroll
        for (int a = 0; a < OUT; ++a) z[a] = fb3[a];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            if (OUT > 1) {
                const float4 w = *reinterpret_cast<const float4*>(fw3 + j * 8);
                const float w4 = fw3[j * 8 + 4];
                z[0] = fmaf(w.x, h2[j], z[0]);
                z[OUT > 1 ? 1 : 0] = fmaf(w.y, h2[j], z[OUT > 1 ? 1 : 0]);
                z[OUT > 2 ? 2 : 0] = fmaf(w.z, h2[j], z[OUT > 2 ? 2 : 0]);
                z[OUT > 3 ? 3 : 0] = fmaf(w.w, h2[j], z[OUT > 3 ? 3 : 0]);
                z[OUT > 4 ? 4 : 0] = fmaf(w4, h2[j], z[OUT > 4 ? 4 : 0]);
            } else {
                z[0] = fmaf(fw3[j * 8], h2[j], z[0]);
            }
        }
        Head::apply(ha, z, t, g, b, src.G, src.B, inb, TRAIN, dz, st);
        __syncwarp();

        if (TRAIN) {
            // dW3[a][j] += sum_s dz[s][a] h2[s][j], db3[a] += sum_s dz[s][a]: warp reduce-scatter, per-warp sums
#pragma unroll
            for (int a = 0; a < OUT; ++a) {
                float p[H];
#pragma unroll
                for (int j = 0; j < H; ++j) p[j] = dz[a] * h2[j];
                warp_reduce_scatter<H>(p, lane);
#pragma unroll
                for (int i = 0; i < H / 32; ++i) {
                    // after the butterfly lane l holds feature index bit-reversed order: recover it
                    // (lane bit 4 selected the low index bit of the first split, ...)
                    int j = 0;
                    {
                        // value index path: at step with width w (16,8,4,2,1) and remaining n (H, H/2, ...),
                        // upper lanes kept the odd half [i + n/2]; so the original index is
                        // i + sum over steps of (bit ? n_step/2 : 0)
                        int n = H;
#pragma unroll
                        for (int w = 16; w >= 1; w >>= 1) {
                            if (lane & w) j += n / 2;
                            n >>= 1;
                        }
                        j += i;
                    }
                    dw3acc[a * H + j] += p[i];
                }
                float d = dz[a];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (lane == 0) db3acc[a] += d;
            }
            // dH2 = (W3^T dz) . relu'(H2) -> split -> TMEM A (operand of B1) and sample-major A image (operand of dW2)
#pragma unroll
            for (int c0 = 0; c0 < H; c0 += 16) {
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
                    const float d = h2[j] > 0.0f ? acc : 0.0f;
                    float h, l;
                    tc::split_tf32(d, h, l);
                    hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(l);
                    *reinterpret_cast<float*>(sm + C::oAs + smaj(j, tid)) = h;
                    *reinterpret_cast<float*>(sm + C::oAs + A_S_BYTES + smaj(j, tid)) = l;
                }
                tmem_st16(tl + C::cAh + c0, hi);
                tmem_st16(tl + C::cAl + c0, lo);
            }
            tc::tmem_wait_st();
            tc::fence_proxy_async_smem();
            tc::tcgen05_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc::tcgen05_fence_after();
                // B1: D3 (D2's columns) = dH2 W2
                issue_ts<H, H>(tmem + C::cD2, tmem + C::cAh, tmem + C::cAl, sbase + C::oW2T, sbase + C::oW2T + C::szW2);
                // dW2 round 0: features 0..31 of H1 (+ the ones group when it is the only round)
                if (C::NR2 == 1)
                    issue_ss<40>(tmem + C::cW2, sbase + C::oAs, sbase + C::oAs + A_S_BYTES, sbase + C::oBs, sbase + C::oBs + B_S_BYTES);
                else
                    issue_ss<32>(tmem + C::cW2, sbase + C::oAs, sbase + C::oAs + A_S_BYTES, sbase + C::oBs, sbase + C::oBs + B_S_BYTES);
                tc::mma_commit(bar);
            }
            mbar_wait_trap(bar, phase); phase ^= 1;
            tc::tcgen05_fence_after();

            // ---- E3: dH1 = D3 . relu'(H1) (H1 > 0  <=>  D1 + b1 > 0) -------------------------------------
            float dh1[H];
#pragma unroll
            for (int c0 = 0; c0 < H; c0 += 16) {
                uint32_t v[16], d1[16];
                tc::tmem_ld16(tl + C::cD2 + c0, v);
                tc::tmem_ld16(tl + C::cD1 + c0, d1);
                tc::tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float pre = __uint_as_float(d1[i]) + fb1[g * H + c0 + i];
                    dh1[c0 + i] = pre > 0.0f ? __uint_as_float(v[i]) : 0.0f;
                    if (C::NR2 == 2 && c0 >= 32) {      // round-1 features of H1, sample-major
                        float h, l;
                        tc::split_tf32(fmaxf(pre, 0.0f), h, l);
                        *reinterpret_cast<float*>(sm + C::oBs + smaj(c0 - 32 + i, tid)) = h;
                        *reinterpret_cast<float*>(sm + C::oBs + B_S_BYTES + smaj(c0 - 32 + i, tid)) = l;
                    }
                }
            }
            if (C::NR2 == 2) {
                tc::fence_proxy_async_smem();
                tc::tcgen05_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc::tcgen05_fence_after();
                    issue_ss<40>(tmem + C::cW2 + 32, sbase + C::oAs, sbase + C::oAs + A_S_BYTES, sbase + C::oBs, sbase + C::oBs + B_S_BYTES);
                    tc::mma_commit(bar);
                }
                mbar_wait_trap(bar, phase); phase ^= 1;
                tc::tcgen05_fence_after();
            }
            // ---- dW1 = dH1^T X: A image <- dH1 (dW2 is complete), B image <- X in rounds of 32 features ----
#pragma unroll
            for (int j = 0; j < H; ++j) {
                float h, l;
                tc::split_tf32(dh1[j], h, l);
                *reinterpret_cast<float*>(sm + C::oAs + smaj(j, tid)) = h;
                *reinterpret_cast<float*>(sm + C::oAs + A_S_BYTES + smaj(j, tid)) = l;
            }
#pragma unroll
            for (int rd = 0; rd < C::NR1; ++rd) {
#pragma unroll
                for (int c0 = 0; c0 < 32; c0 += 8) {
                    const int k0 = rd * 32 + c0;
                    uint32_t xh[8], xl[8];
                    if (k0 < K1P) {
                        tc::tmem_ld8(tl + C::cXh + k0, xh);
                        tc::tmem_ld8(tl + C::cXl + k0, xl);
                        tc::tmem_wait_ld();
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { xh[i] = 0u; xl[i] = 0u; }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        *reinterpret_cast<uint32_t*>(sm + C::oBs + smaj(c0 + i, tid)) = xh[i];
                        *reinterpret_cast<uint32_t*>(sm + C::oBs + B_S_BYTES + smaj(c0 + i, tid)) = xl[i];
                    }
                }
                tc::fence_proxy_async_smem();
                tc::tcgen05_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc::tcgen05_fence_after();
                    if (rd == C::NR1 - 1)
                        issue_ss<40>(tmem + C::cW1 + 32 * rd, sbase + C::oAs, sbase + C::oAs + A_S_BYTES, sbase + C::oBs, sbase + C::oBs + B_S_BYTES);
                    else
                        issue_ss<32>(tmem + C::cW1 + 32 * rd, sbase + C::oAs, sbase + C::oAs + A_S_BYTES, sbase + C::oBs, sbase + C::oBs + B_S_BYTES);
                    tc::mma_commit(bar);
                }
                mbar_wait_trap(bar, phase); phase ^= 1;
                tc::tcgen05_fence_after();
            }
            // ---- per-tile sums (lanes 0-15 of the quadrant) -> running sums (lanes 16-31) ------------------
#pragma unroll 1
            for (int c = C::cW2; c < C::cEnd; c += 8) {
                uint32_t v[8];
                tc::tmem_ld8(tl + c, v);
                tc::tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float part = __shfl_sync(0xffffffffu, __uint_as_float(v[i]), lane & 15);
                    if (lane >= 16) v[i] = __float_as_uint(__uint_as_float(v[i]) + part);
                }
                // folded one-hot id column of this agent group: gradient = bias-gradient column of dW1
                if (nd.fold_ids && c == C::cW1 + 32 * C::NR1 && lane < 16) {
                    const int row = warp * 16 + lane;
                    if (row < H) didacc[g * H + row] += __uint_as_float(v[0]);
                }
                tmem_st8(tl + c, v);
            }
            tc::tmem_wait_st();
        }
        // every thread is done with this tile's TMEM / shared operands before the next tile overwrites them
        tc::tcgen05_fence_before();
        __syncthreads();
        tc::tcgen05_fence_after();
    }

    // ---- write this CTA's partial: gradients in torch parameter order, then the statistics ------------
    if (TRAIN) {
        float* out = partials + (size_t)blockIdx.x * (p_net + CMARL_N_STATS);
        const int in_dim = nd.in_dim;
        float* gW1 = out;
        float* gb1 = gW1 + H * in_dim;
        float* gW2 = gb1 + H;
        float* gb2 = gW2 + H * H;
        float* gW3 = gb2 + H;
        float* gb3 = gW3 + nd.out_dim * H;
        {
            // running sums live in lanes 16-31 of each quadrant; every lane executes the (warp-aligned) loads
            const int row = warp * 16 + (lane - 16);           // gradient row j held by this lane
            const bool valid = lane >= 16 && row < H;
            for (int c = 0; c < C::nW2; ++c) {
                const float v = __uint_as_float(tc::tmem_ld1(tl + C::cW2 + c));
                if (valid) {
                    if (c < H) gW2[row * H + c] = v;
                    else if (c == 32 * C::NR2) gb2[row] = v;
                }
            }
            for (int c = 0; c < C::nW1; ++c) {
                const float v = __uint_as_float(tc::tmem_ld1(tl + C::cW1 + c));
                if (valid) {
                    if (c < nd.in_rows) gW1[row * in_dim + c] = v;
                    else if (c == 32 * C::NR1) gb1[row] = v;
                }
            }
        }
        __syncthreads();
        if (nd.fold_ids)
            for (int i = tid; i < src.G * H; i += NT) {
                const int g = i / H, j = i - g * H;
                gW1[j * in_dim + nd.in_rows + g] = didacc[i];
            }
        const float* w3all = reinterpret_cast<const float*>(sm + C::oDW3);
        for (int i = tid; i < nd.out_dim * H; i += NT)
            gW3[i] = ((w3all[i] + w3all[8 * H + i]) + w3all[2 * 8 * H + i]) + w3all[3 * 8 * H + i];
        if (tid < nd.out_dim) {
            const float* b3all = w3all + 4 * 8 * H;
            gb3[tid] = ((b3all[tid] + b3all[8 + tid]) + b3all[16 + tid]) + b3all[24 + tid];
        }
        float* red = reinterpret_cast<float*>(sm + C::oRed);
#pragma unroll
        for (int k = 0; k < Head::NSTAT; ++k) {
            const float v = warp_sum_f(st[k]);
            __syncthreads();
            if (lane == 0) red[warp] = v;
            __syncthreads();
            if (tid == 0) out[p_net + k] = ((red[0] + red[1]) + red[2]) + red[3];
        }
        if (tid == 0)
            for (int k = Head::NSTAT; k < CMARL_N_STATS; ++k) out[p_net + k] = 0.0f;
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, C::TMEM_COLS);
}

template <class C, class Head>
static int tc_set_attr() {
    return cmarl_check_cuda(cudaFuncSetAttribute(tc_chain_kernel<C, Head>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 C::smem_bytes),
                            "cudaFuncSetAttribute(tc_chain_kernel)");
}

template <class C, class Head>
static int tc_launch(const NetDesc& nd, const TileSrc& src, const typename Head::Args& ha, float* partials, int p_net,
                     int grid, cudaStream_t st) {
    tc_chain_kernel<C, Head><<<grid, NT, C::smem_bytes, st>>>(nd, src, ha, partials, p_net);
    return cmarl_check_cuda(cudaGetLastError(), "tc_chain_kernel launch");
}

}  // namespace tcchain

using namespace tcchain;

int cmarl_tc_setup() {
    int e = 0;
#define SET(Hh, Kk)                                                             \
    if (!e) e = tc_set_attr<TCfg<Hh, Kk, true>, PolicyHead>();                  \
    if (!e) e = tc_set_attr<TCfg<Hh, Kk, true>, ValueHead>();                   \
    if (!e) e = tc_set_attr<TCfg<Hh, Kk, false>, ValueHead>();
    SET(32, 24) SET(32, 56) SET(64, 24) SET(64, 56)
#undef SET
    return e;
}

int cmarl_tc_tile() { return M; }

template <class Head, bool TRAIN>
int cmarl_tc_dispatch(int H, const NetDesc& nd, const TileSrc& src, const typename Head::Args& ha, float* partials,
                      int p_net, int grid, cudaStream_t st) {
    const int kin = nd.in_rows <= 24 ? 24 : 56;
    if (H == 32 && kin == 24) return tc_launch<TCfg<32, 24, TRAIN>, Head>(nd, src, ha, partials, p_net, grid, st);
    if (H == 32 && kin == 56) return tc_launch<TCfg<32, 56, TRAIN>, Head>(nd, src, ha, partials, p_net, grid, st);
    if (H == 64 && kin == 24) return tc_launch<TCfg<64, 24, TRAIN>, Head>(nd, src, ha, partials, p_net, grid, st);
    if (H == 64 && kin == 56) return tc_launch<TCfg<64, 56, TRAIN>, Head>(nd, src, ha, partials, p_net, grid, st);
    cmarl_set_error("tc dispatch: unsupported hidden=%d in_rows=%d", H, nd.in_rows);
    return -1;
}

template int cmarl_tc_dispatch<PolicyHead, true>(int, const NetDesc&, const TileSrc&, const PolicyHeadArgs&, float*, int, int, cudaStream_t);
template int cmarl_tc_dispatch<ValueHead, true>(int, const NetDesc&, const TileSrc&, const ValueHeadArgs&, float*, int, int, cudaStream_t);
template int cmarl_tc_dispatch<ValueHead, false>(int, const NetDesc&, const TileSrc&, const ValueHeadArgs&, float*, int, int, cudaStream_t);
