// Recurrent-actor path (BASELINE config 4, cleanmarl/mappo_lstm_multienvs.py, "LSTM" below):
//   K7a  tbptt_chunk_kernel   actor forward + clipped-PPO head + backward through time for ONE truncated-BPTT
//                             chunk (LSTM:563-607), weight gradients accumulated in shared memory
//   K2r  actor_act_gru_kernel Actor.act(x, h, avail) for one time step (LSTM:170-184)
//
// Actor: x1 = relu(W1 x + b1); GRUCell (torch gate order r, z, n):
//   r = sigmoid(Wir x1 + bir + Whr h + bhr)      z = sigmoid(Wiz x1 + biz + Whz h + bhz)
//   n = tanh(Win x1 + bin + r * (Whn h + bhn))   h' = (h - n) * z + n           (ATen gru_cell, CPU path)
//   logits = W2 relu(h') + b2
//
// tbptt_chunk_kernel: one CTA owns a tile of M = 64 consecutive envs of one agent and walks the chunk twice:
//   pass 1 (t = t0 .. t1-1)  h_{t+1} from (x_t, h_t); the hidden state goes to h_seq[t+1] (global, L2-resident) and,
//                            with a gate stash, x1 / r / z / n / ghn of the step to the stash (640 B per sample-step)
//   pass 2 (t = t1-1 .. t0)  gets the gates of step t back -- from the stash, or (no stash) by recomputing them from
//                            (x_t, h_seq[t]), 0.7x the cost of a forward -- then the head (loss terms, statistics,
//                            dlogits) and the backward step:
//     dh   = dh_carry + (W2^T dz) . relu'(h')            dW2 += dz relu(h')^T
//     dn   = dh (1-z)   dz_g = dh (h - n)   dh_carry = dh z
//     da_n = dn (1-n^2) da_hn = da_n r      da_r = da_n ghn r(1-r)     da_z = dz_g z(1-z)
//     dWih += [da_r, da_z, da_n] x1^T       dWhh += [da_r, da_z, da_hn] h^T      (+ bias sums)
//     dx1  = Wih^T [da_r, da_z, da_n] . relu'(x1)        dh_carry += Whh^T [da_r, da_z, da_hn]
//     dW1 += dx1 x^T (+ the folded one-hot id column of this agent)
// All activations are feature-major in shared memory ([row][sample], LD = 68): thread tiles of 4 samples x 2 units
// read them as LDS.128 and the weights as warp-broadcast LDS.64/128.  Input rows arrive by cp.async.bulk (TMA
// engine) + mbarrier, double buffered across steps; h_seq[t] of the next backward step is prefetched into registers.
#include <stdlib.h>
#include <string.h>

#include "chain.cuh"
#include "heads.cuh"

int cmarl_tc_gru_setup();
int cmarl_tc_gru_launch(cmarl_ctx* ctx, const chain::GruChunkArgs& a, int which, int* grid_out, cudaStream_t st);
int cmarl_reduce_one_net(cmarl_ctx* ctx, const float* pa, int grid_a, int Pa, const float* pc, int grid_c, int Pc,
                         float count_div, float* out, cudaStream_t st);

namespace gru {

using namespace chain;

constexpr int H = 32;            // hidden units (fc1 out = GRU in = GRU hidden)
constexpr int G3 = 3 * H;
constexpr int M = 64;            // samples (envs) per tile
constexpr int NTMAX = 512;       // threads per CTA: 512 / NU (NU = hidden units per thread in the 4-sample x NU-unit tiles)
constexpr int LD = M + 4;        // row stride in floats (16-B aligned rows, conflict-free LDS.128)
constexpr int NA = 5;            // actions
constexpr int KIN = 24;          // padded input rows (18 raw / 21 with explicit ids)
constexpr int NQ = M / 4;        // float4 quads per row

// shared-memory map (float offsets)
constexpr int oBar = 0;                          // 2 mbarriers
constexpr int oW1T = 4;                          // [KIN][H]     fc1, in-major
constexpr int oB1 = oW1T + KIN * H;              // [4][H]       b1 (+ folded id column) per agent
constexpr int oWgT = oB1 + 4 * H;                // [2H][H][4]   k-major gate weights: k < H from x1 (Wir, Wiz, Win, 0), k >= H from h (Whr, Whz, Whn, 0)
constexpr int oBg = oWgT + 2 * H * H * 4;        // [H][4]       bir+bhr, biz+bhz, bin, bhn
constexpr int oWih = oBg + H * 4;                // [3H][H/2][4] backward weights, packed per (gate row, output pair k0 = 2 og):
constexpr int oWhh = oWih + G3 * H;              //              (Wih[row][k0], Wih[row][k0+1], Whh[row][k0], Whh[row][k0+1]): one LDS.128
constexpr int oW2T = oWhh + G3 * H;              // [H][8]
constexpr int oB2 = oW2T + H * 8;                // [8]
constexpr int oX = oB2 + 8;                      // [2][KIN][LD] input rows, double buffered
constexpr int oHp = oX + 2 * KIN * LD;           // [H][LD]      h before the step
constexpr int oX1 = oHp + H * LD;                // [H][LD]      x1, then dx1
constexpr int oG = oX1 + H * LD;                 // [4H][LD]     r, z, n, ghn -> da_r, da_z, da_n, da_hn
constexpr int oHc = oG + 4 * H * LD;             // [H][LD]      h after the step
constexpr int oDH = oHc + H * LD;                // [H][LD]      dh carried to the previous step
constexpr int oZ = oDH + H * LD;                 // [8][LD]      dlogits
constexpr int PMAXG = 7232;                      // >= 7 205 parameters, multiple of 4
constexpr int oDW = oZ + 8 * LD;                 // [PMAXG]      gradient accumulators, torch parameter order
constexpr int oScr = oDW + PMAXG;                // [4][H][KIN]  dW1 sample-split partials
constexpr int GLD = H + 1;                       // row stride of the dWih / dWhh accumulators: lanes of a warp add to
                                                 // rows jg, jg+1, .. at the same column -> distinct banks (32 would be 8-way)
constexpr int oDG = oScr + 4 * H * KIN;          // [2][3H][GLD] dWih, dWhh accumulators
constexpr int oRed = oDG + 2 * G3 * GLD + 4;     // [64]
constexpr int oEnd = oRed + 64;
constexpr size_t SMEM_BYTES = (size_t)oEnd * 4;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(oX % 4 == 0 && oHp % 4 == 0 && oG % 4 == 0 && oDW % 4 == 0 && oWgT % 4 == 0 && oWih % 4 == 0, "16-B alignment");

using ChunkArgs = chain::GruChunkArgs;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int NT>
__device__ void load_weights(float* sm, const ChunkArgs& a) {
    const GruLayout& L = a.L;
    const float* __restrict__ P = a.params;
    const int O = L.in;
    // every loop is straight-line code (CMARL_STRIDED): the ~30 global loads of a thread are in flight together
    CMARL_STRIDED(i, KIN * H, NT) {
        const int k = i / H, j = i - k * H;
        sm[oW1T + i] = (k < a.in_rows) ? __ldcg(P + L.w1 + j * O + k) : 0.0f;
    }
    CMARL_STRIDED(i, 4 * H, NT) {
        const int g = i / H, j = i - g * H;
        float v = __ldcg(P + L.b1 + j);
        if (a.fold_ids && g < a.N) v += __ldcg(P + L.w1 + j * O + a.in_rows + g);
        sm[oB1 + i] = v;
    }
    CMARL_STRIDED(i, 2 * H * H, NT) {
        const int k = i / H, j = i - k * H;           // k: input index (x1 then h), j: unit
        const float* W = (k < H) ? P + L.wih : P + L.whh;
        const int kk = (k < H) ? k : k - H;
        float4 w;
        w.x = __ldcg(W + (0 * H + j) * H + kk);
        w.y = __ldcg(W + (1 * H + j) * H + kk);
        w.z = __ldcg(W + (2 * H + j) * H + kk);
        w.w = 0.0f;
        *reinterpret_cast<float4*>(sm + oWgT + (size_t)i * 4) = w;
    }
    CMARL_STRIDED(j, H, NT) {
        float4 b;
        b.x = __ldcg(P + L.bih + j) + __ldcg(P + L.bhh + j);
        b.y = __ldcg(P + L.bih + H + j) + __ldcg(P + L.bhh + H + j);
        b.z = __ldcg(P + L.bih + 2 * H + j);
        b.w = __ldcg(P + L.bhh + 2 * H + j);
        *reinterpret_cast<float4*>(sm + oBg + j * 4) = b;
    }
    CMARL_STRIDED(i, G3 * H, NT) {
        const int row = i / H, k = i - row * H;
        sm[oWih + (row * (H / 2) + (k >> 1)) * 4 + (k & 1)] = __ldcg(P + L.wih + i);
        sm[oWih + (row * (H / 2) + (k >> 1)) * 4 + 2 + (k & 1)] = __ldcg(P + L.whh + i);
    }
    CMARL_STRIDED(i, H * 8, NT) {
        const int j = i / 8, c = i - j * 8;
        sm[oW2T + i] = (c < NA) ? __ldcg(P + L.w2 + c * H + j) : 0.0f;
    }
    if (threadIdx.x < 8) sm[oB2 + threadIdx.x] = (threadIdx.x < NA) ? __ldcg(P + L.b2 + threadIdx.x) : 0.0f;
}

// input rows of (t, g, b0) -> xbuf; full aligned tiles by TMA bulk copies, ragged ones by guarded loads
__device__ __forceinline__ bool tile_bulk(const ChunkArgs& a, int b0) {
    return (b0 + M <= a.B) && ((a.B & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
}
__device__ void issue_x(float* xbuf, uint64_t* bar, const ChunkArgs& a, int t, int g, int b0) {
    const int NT = blockDim.x;
    const float* base = a.x + (size_t)t * a.stride_t + (size_t)g * a.stride_g + b0;
    if (tile_bulk(a, b0)) {
        if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)(a.in_rows * M * 4));
        if (threadIdx.x < a.in_rows) {
            fence_proxy_async();
            bulk_g2s(xbuf + threadIdx.x * LD, base + (size_t)threadIdx.x * a.B, M * 4, bar);
        }
    } else {
        const int valid = a.B - b0;
        for (int i = threadIdx.x; i < a.in_rows * M; i += NT) {
            const int r = i / M, s = i - r * M;
            xbuf[r * LD + s] = (s < valid) ? __ldcg(base + (size_t)r * a.B + s) : 0.0f;
        }
        if (threadIdx.x == 0) mbar_arrive(bar);
    }
}

// x1 = relu(W1 x + b1[g]) : thread (sg, og) = 4 samples x NU units
template <int NU>
__device__ __forceinline__ void fc1(const float* __restrict__ sm, const float* __restrict__ X, int in_rows, int g,
                                    float* __restrict__ X1) {
    const int sg = threadIdx.x & 15, og = threadIdx.x >> 4;
    const int s0 = 4 * sg, j0 = NU * og;
    float acc[NU][4];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const float b = sm[oB1 + g * H + j0 + i];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = b;
    }
#pragma unroll 6
    for (int k = 0; k < in_rows; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(X + k * LD + s0);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const float w = sm[oW1T + k * H + j0 + i];
            acc[i][0] = fmaf(w, x.x, acc[i][0]); acc[i][1] = fmaf(w, x.y, acc[i][1]);
            acc[i][2] = fmaf(w, x.z, acc[i][2]); acc[i][3] = fmaf(w, x.w, acc[i][3]);
        }
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        float4 o;
        o.x = fmaxf(acc[i][0], 0.0f); o.y = fmaxf(acc[i][1], 0.0f); o.z = fmaxf(acc[i][2], 0.0f); o.w = fmaxf(acc[i][3], 0.0f);
        *reinterpret_cast<float4*>(X1 + (j0 + i) * LD + s0) = o;
    }
}

// GRUCell: (X1, Hp) -> Hc ; STASH also keeps r, z, n, ghn for the backward step
template <int NU, bool STASH>
__device__ __forceinline__ void gru_cell(const float* __restrict__ sm, const float* __restrict__ X1,
                                         const float* __restrict__ Hp, float* __restrict__ Hc, float* __restrict__ G) {
    const int sg = threadIdx.x & 15, og = threadIdx.x >> 4;
    const int s0 = 4 * sg, j0 = NU * og;
    float ar[NU][4], az[NU][4], ai[NU][4], ah[NU][4];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const float4 b = *reinterpret_cast<const float4*>(sm + oBg + (j0 + i) * 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) { ar[i][c] = b.x; az[i][c] = b.y; ai[i][c] = b.z; ah[i][c] = b.w; }
    }
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(X1 + k * LD + s0);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const float4 w = *reinterpret_cast<const float4*>(sm + oWgT + ((size_t)k * H + j0 + i) * 4);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ar[i][c] = fmaf(w.x, xs[c], ar[i][c]);
                az[i][c] = fmaf(w.y, xs[c], az[i][c]);
                ai[i][c] = fmaf(w.z, xs[c], ai[i][c]);
            }
        }
    }
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(Hp + k * LD + s0);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const float4 w = *reinterpret_cast<const float4*>(sm + oWgT + ((size_t)(H + k) * H + j0 + i) * 4);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ar[i][c] = fmaf(w.x, xs[c], ar[i][c]);
                az[i][c] = fmaf(w.y, xs[c], az[i][c]);
                ah[i][c] = fmaf(w.z, xs[c], ah[i][c]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const int j = j0 + i;
        const float4 hp4 = *reinterpret_cast<const float4*>(Hp + j * LD + s0);
        const float hp[4] = {hp4.x, hp4.y, hp4.z, hp4.w};
        float r[4], z[4], n[4], h[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            r[c] = sigmoidf_(ar[i][c]);
            z[c] = sigmoidf_(az[i][c]);
            n[c] = tanhf(ai[i][c] + r[c] * ah[i][c]);
            h[c] = (hp[c] - n[c]) * z[c] + n[c];
        }
        *reinterpret_cast<float4*>(Hc + j * LD + s0) = make_float4(h[0], h[1], h[2], h[3]);
        if (STASH) {
            *reinterpret_cast<float4*>(G + (0 * H + j) * LD + s0) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(G + (1 * H + j) * LD + s0) = make_float4(z[0], z[1], z[2], z[3]);
            *reinterpret_cast<float4*>(G + (2 * H + j) * LD + s0) = make_float4(n[0], n[1], n[2], n[3]);
            *reinterpret_cast<float4*>(G + (3 * H + j) * LD + s0) = make_float4(ah[i][0], ah[i][1], ah[i][2], ah[i][3]);
        }
    }
}

// logits of sample s = tid / 4 from relu(Hc): the 4 lanes of a sample each sum 8 units, then two shuffles
__device__ __forceinline__ void logits_of(const float* __restrict__ sm, const float* __restrict__ Hc, float (&z)[NA]) {
    const int s = threadIdx.x >> 2, part = threadIdx.x & 3;
#pragma unroll
    for (int c = 0; c < NA; ++c) z[c] = 0.0f;
#pragma unroll
    for (int i = 0; i < H / 4; ++i) {
        const int j = 4 * i + part;
        const float h = fmaxf(Hc[j * LD + s], 0.0f);
        const float4 w = *reinterpret_cast<const float4*>(sm + oW2T + j * 8);
        const float w4 = sm[oW2T + j * 8 + 4];
        z[0] = fmaf(w.x, h, z[0]); z[1] = fmaf(w.y, h, z[1]); z[2] = fmaf(w.z, h, z[2]);
        z[3] = fmaf(w.w, h, z[3]); z[4] = fmaf(w4, h, z[4]);
    }
#pragma unroll
    for (int c = 0; c < NA; ++c) {
        z[c] += __shfl_xor_sync(0xffffffffu, z[c], 1);
        z[c] += __shfl_xor_sync(0xffffffffu, z[c], 2);
        z[c] += sm[oB2 + c];
    }
}

// this thread's 4 consecutive samples (s0..s0+3 of the tile at b0) of one global row of B floats <-> shared memory
__device__ __forceinline__ void row4_store(float* __restrict__ grow, int b0, int s0, int valid, int B, float4 v) {
    float* dst = grow + b0 + s0;
    if (s0 + 3 < valid && ((B & 3) == 0)) {
        *reinterpret_cast<float4*>(dst) = v;
    } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (s0 + c < valid) dst[c] = e[c];
    }
}
__device__ __forceinline__ float4 row4_load(const float* __restrict__ grow, int b0, int s0, int valid, int B) {
    const float* src = grow + b0 + s0;
    if (s0 + 3 < valid && ((B & 3) == 0)) return *reinterpret_cast<const float4*>(src);
    float e[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (s0 + c < valid) e[c] = src[c];
    return make_float4(e[0], e[1], e[2], e[3]);
}

// debug timeline (clock64) of CTA 0, first tile: slots 0-3 = forward step 1, slots 4-15 = backward step 1
__device__ long long g_gru_tl[16];
#define GTL(slot, cond) do { if (blockIdx.x == 0 && u == 0 && i == 1 && tid == 0 && (cond)) g_gru_tl[slot] = clock64(); } while (0)

template <int NU, bool STASH>
__global__ void __launch_bounds__(NTMAX / NU, 1) tbptt_chunk_kernel(ChunkArgs a) {
    constexpr int NT = NTMAX / NU;
    extern __shared__ __align__(128) float sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oBar);
    const int tid = threadIdx.x;
    const int tiles_b = (a.B + M - 1) / M;
    const int units = a.N * tiles_b;
    const int nsteps = a.t1 - a.t0;
    const GruLayout& L = a.L;

    // shared-memory zeroing first: under launch chaining it runs while the kernel in front (an Adam step) still does
    for (int i = tid; i < 2 * KIN * LD; i += NT) sm[oX + i] = 0.0f;
    for (int i = tid; i < PMAXG; i += NT) sm[oDW + i] = 0.0f;
    for (int i = tid; i < 2 * G3 * GLD; i += NT) sm[oDG + i] = 0.0f;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    pdl_wait_then_trigger();
    load_weights<NT>(sm, a);
    __syncthreads();

    float* dW = sm + oDW;
    float st[PolicyHead::NSTAT];
#pragma unroll
    for (int k = 0; k < PolicyHead::NSTAT; ++k) st[k] = 0.0f;

    const int sg = tid & 15, og = tid >> 4;
    const int s0 = 4 * sg, j0 = NU * og;
    float* Hp = sm + oHp;
    float* X1 = sm + oX1;
    float* G = sm + oG;
    float* Hc = sm + oHc;
    float* DH = sm + oDH;
    float* Z = sm + oZ;

    int it = 0;                                   // global step counter: X buffer = it & 1, barrier phase = (it >> 1) & 1
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int g = u / tiles_b, b0 = (u - g * tiles_b) * M;
        const int valid = min(M, a.B - b0);
        // step list of this tile: pass 1 t0..t1-1, pass 2 t1-1..t0.  Step i's input rows are issued one step ahead.
        auto t_of = [&](int i) { return i < nsteps ? a.t0 + i : a.t1 - 1 - (i - nsteps); };
        issue_x(sm + oX + (it & 1) * KIN * LD, &bars[it & 1], a, t_of((a.passes & 1) ? 0 : nsteps), g, b0);
        // chunk-start hidden state: zeros at the start of an epoch (LSTM:558), else h_seq[t0]
        for (int i = tid; i < H * M; i += NT) {
            const int j = i / M, s = i - j * M;
            float v = 0.0f;
            if (a.t0 > 0 && s < valid) v = a.h_seq[(((size_t)a.t0 * a.N + g) * H + j) * a.B + b0 + s];
            Hp[j * LD + s] = v;
        }
        __syncthreads();

        // ------------------------------------------------------------------ pass 1: hidden states
        for (int i = 0; i < ((a.passes & 1) ? nsteps : 0); ++i, ++it) {
            const int t = a.t0 + i;
            if (i + 1 < nsteps || (a.passes & 2))
                issue_x(sm + oX + ((it + 1) & 1) * KIN * LD, &bars[(it + 1) & 1], a, t_of(i + 1), g, b0);
            const float* X = sm + oX + (it & 1) * KIN * LD;
            GTL(0, true);
            mbar_wait(&bars[it & 1], (it >> 1) & 1);
            fc1<NU>(sm, X, a.in_rows, g, X1);
            __syncthreads();
            GTL(1, true);
            gru_cell<NU, STASH>(sm, X1, Hp, Hc, G);
            GTL(2, true);
            // this thread's 2 x 4 block of h_{t+1} -> h_seq[t+1] and becomes Hp of the next step
            __syncthreads();                       // every read of Hp is done
#pragma unroll
            for (int q = 0; q < NU; ++q) {
                const int j = j0 + q;
                const float4 h = *reinterpret_cast<const float4*>(Hc + j * LD + s0);
                *reinterpret_cast<float4*>(Hp + j * LD + s0) = h;
                row4_store(a.h_seq + (((size_t)(t + 1) * a.N + g) * H + j) * a.B, b0, s0, valid, a.B, h);
                if (STASH) {   // the blocks this thread itself computed: x1 and the four gate rows of its two units
                    float* slab = a.stash + ((size_t)t * a.N + g) * (5 * H) * a.B;
                    float4 v[5];
                    v[0] = *reinterpret_cast<const float4*>(X1 + j * LD + s0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k + 1] = *reinterpret_cast<const float4*>(G + (k * H + j) * LD + s0);
#pragma unroll
                    for (int k = 0; k < 5; ++k) row4_store(slab + (size_t)(k * H + j) * a.B, b0, s0, valid, a.B, v[k]);
                }
            }
            __syncthreads();
            GTL(3, true);
        }

        // ------------------------------------------------------------------ pass 2: recompute + backward
        for (int i = tid; i < H * LD; i += NT) DH[i] = 0.0f;
        // Hp currently holds h_{t1}; the first backward step needs h_{t1-1}: reload below like every other step
        for (int i = 0; i < ((a.passes & 2) ? nsteps : 0); ++i, ++it) {
            const int t = a.t1 - 1 - i;
            GTL(4, true);
            if (i + 1 < nsteps)
                issue_x(sm + oX + ((it + 1) & 1) * KIN * LD, &bars[(it + 1) & 1], a, t_of(nsteps + i + 1), g, b0);
            // Hp = h_seq[t] (zeros for t == 0): this thread's own 2 x 4 block, written by this CTA in pass 1
            // (t > t0) or by an earlier launch (t == t0)
            {
                // all global loads first (independent, one L2 round trip), then the shared-memory stores: interleaved,
                // the compiler must assume the generic-pointer stores alias the next load and serialises 14 round trips
                float4 hp4[NU], hc4[NU], st4[NU][5];
#pragma unroll
                for (int q = 0; q < NU; ++q) {
                    const int j = j0 + q;
                    hp4[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (t > 0) hp4[q] = row4_load(a.h_seq + (((size_t)t * a.N + g) * H + j) * a.B, b0, s0, valid, a.B);
                    if (STASH) {   // what pass 1 left for this step: h_{t+1}, x1 and the gates (this thread's own blocks)
                        hc4[q] = row4_load(a.h_seq + (((size_t)(t + 1) * a.N + g) * H + j) * a.B, b0, s0, valid, a.B);
                        const float* slab = a.stash + ((size_t)t * a.N + g) * (5 * H) * a.B;
#pragma unroll
                        for (int k = 0; k < 5; ++k) st4[q][k] = row4_load(slab + (size_t)(k * H + j) * a.B, b0, s0, valid, a.B);
                    }
                }
#pragma unroll
                for (int q = 0; q < NU; ++q) {
                    const int j = j0 + q;
                    *reinterpret_cast<float4*>(Hp + j * LD + s0) = hp4[q];
                    if (STASH) {
                        *reinterpret_cast<float4*>(Hc + j * LD + s0) = hc4[q];
                        *reinterpret_cast<float4*>(X1 + j * LD + s0) = st4[q][0];
#pragma unroll
                        for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(G + (k * H + j) * LD + s0) = st4[q][k + 1];
                    }
                }
            }
            const float* X = sm + oX + (it & 1) * KIN * LD;
            mbar_wait(&bars[it & 1], (it >> 1) & 1);
            if (!STASH) {
                fc1<NU>(sm, X, a.in_rows, g, X1);
                __syncthreads();
                gru_cell<NU, true>(sm, X1, Hp, Hc, G);
            }
            __syncthreads();
            GTL(5, true);

            // head: logits of this step, loss terms / statistics / dlogits (LSTM:574-593, 628-638)
            {
                float z[NA], dz[NA];
                if (tid < 4 * M) logits_of(sm, Hc, z);          // warp-uniform: 64 samples x 4 lanes = the first 8 warps
                const int s = tid >> 2;
                if ((tid & 3) == 0 && tid < 4 * M) {
                    PolicyHead::apply(a.head, z, t, g, b0 + s, a.N, a.B, (b0 + s) < a.B, true, dz, st);
#pragma unroll
                    for (int c = 0; c < NA; ++c) Z[c * LD + s] = dz[c];
                }
            }
            __syncthreads();
            GTL(6, true);

            // (a) dW2 += dz relu(h')^T, db2 ; (b) gate gradients in place of the stash, dh carry
            if (tid < NA * H) {
                const int c = tid >> 5, j = tid & 31;
                float acc = 0.0f;
#pragma unroll 4
                for (int q = 0; q < NQ; ++q) {
                    const float4 h = *reinterpret_cast<const float4*>(Hc + j * LD + 4 * q);
                    const float4 d = *reinterpret_cast<const float4*>(Z + c * LD + 4 * q);
                    acc = fmaf(d.x, fmaxf(h.x, 0.0f), acc); acc = fmaf(d.y, fmaxf(h.y, 0.0f), acc);
                    acc = fmaf(d.z, fmaxf(h.z, 0.0f), acc); acc = fmaf(d.w, fmaxf(h.w, 0.0f), acc);
                }
                dW[L.w2 + c * H + j] += acc;
            } else if (tid < NA * H + NA) {
                const int c = tid - NA * H;
                float acc = 0.0f;
                for (int q = 0; q < NQ; ++q) {
                    const float4 d = *reinterpret_cast<const float4*>(Z + c * LD + 4 * q);
                    acc += (d.x + d.y) + (d.z + d.w);
                }
                dW[L.b2 + c] += acc;
            }
            {
                float4 dzv[NA];
#pragma unroll
                for (int c = 0; c < NA; ++c) dzv[c] = *reinterpret_cast<const float4*>(Z + c * LD + s0);
#pragma unroll
                for (int q = 0; q < NU; ++q) {
                    const int j = j0 + q;
                    const float4 w = *reinterpret_cast<const float4*>(sm + oW2T + j * 8);
                    const float w4 = sm[oW2T + j * 8 + 4];
                    const float4 hc4 = *reinterpret_cast<const float4*>(Hc + j * LD + s0);
                    const float4 hp4 = *reinterpret_cast<const float4*>(Hp + j * LD + s0);
                    const float4 dh4 = *reinterpret_cast<const float4*>(DH + j * LD + s0);
                    const float4 r4 = *reinterpret_cast<const float4*>(G + (0 * H + j) * LD + s0);
                    const float4 z4 = *reinterpret_cast<const float4*>(G + (1 * H + j) * LD + s0);
                    const float4 n4 = *reinterpret_cast<const float4*>(G + (2 * H + j) * LD + s0);
                    const float4 g4 = *reinterpret_cast<const float4*>(G + (3 * H + j) * LD + s0);
                    const float hc[4] = {hc4.x, hc4.y, hc4.z, hc4.w}, hp[4] = {hp4.x, hp4.y, hp4.z, hp4.w};
                    const float dhc[4] = {dh4.x, dh4.y, dh4.z, dh4.w};
                    const float r[4] = {r4.x, r4.y, r4.z, r4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
                    const float n[4] = {n4.x, n4.y, n4.z, n4.w}, gh[4] = {g4.x, g4.y, g4.z, g4.w};
                    const float d0[4] = {dzv[0].x, dzv[0].y, dzv[0].z, dzv[0].w};
                    const float d1[4] = {dzv[1].x, dzv[1].y, dzv[1].z, dzv[1].w};
                    const float d2[4] = {dzv[2].x, dzv[2].y, dzv[2].z, dzv[2].w};
                    const float d3[4] = {dzv[3].x, dzv[3].y, dzv[3].z, dzv[3].w};
                    const float d4[4] = {dzv[4].x, dzv[4].y, dzv[4].z, dzv[4].w};
                    float o_r[4], o_z[4], o_n[4], o_h[4], o_dh[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float up = w.x * d0[c];
                        up = fmaf(w.y, d1[c], up); up = fmaf(w.z, d2[c], up); up = fmaf(w.w, d3[c], up); up = fmaf(w4, d4[c], up);
                        const float dh = dhc[c] + (hc[c] > 0.0f ? up : 0.0f);
                        const float dn = dh * (1.0f - zz[c]);
                        const float dzg = dh * (hp[c] - n[c]);
                        o_dh[c] = dh * zz[c];
                        const float dan = dn * (1.0f - n[c] * n[c]);
                        o_n[c] = dan;
                        o_h[c] = dan * r[c];
                        o_r[c] = dan * gh[c] * (r[c] * (1.0f - r[c]));
                        o_z[c] = dzg * (zz[c] * (1.0f - zz[c]));
                    }
                    *reinterpret_cast<float4*>(G + (0 * H + j) * LD + s0) = make_float4(o_r[0], o_r[1], o_r[2], o_r[3]);
                    *reinterpret_cast<float4*>(G + (1 * H + j) * LD + s0) = make_float4(o_z[0], o_z[1], o_z[2], o_z[3]);
                    *reinterpret_cast<float4*>(G + (2 * H + j) * LD + s0) = make_float4(o_n[0], o_n[1], o_n[2], o_n[3]);
                    *reinterpret_cast<float4*>(G + (3 * H + j) * LD + s0) = make_float4(o_h[0], o_h[1], o_h[2], o_h[3]);
                    *reinterpret_cast<float4*>(DH + j * LD + s0) = make_float4(o_dh[0], o_dh[1], o_dh[2], o_dh[3]);
                }
            }
            __syncthreads();
            GTL(7, true);

            // (c) dWih += [da_r, da_z, da_n] x1^T ; dWhh += [da_r, da_z, da_hn] h^T and the bias-row sums.
            //     Meanwhile the previous step's inputs are pulled into L2 (they are read at its start).
            if (STASH && i + 1 < nsteps) {
#pragma unroll
                for (int q = 0; q < NU; ++q) {
                    const int j = j0 + q;
                    const float* slab = a.stash + ((size_t)(t - 1) * a.N + g) * (5 * H) * a.B;
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(slab + (size_t)(k * H + j) * a.B + b0 + s0));
                    if (t > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.h_seq + (((size_t)(t - 1) * a.N + g) * H + j) * a.B + b0 + s0));
                }
            }
            if (tid < 128) {
                // 128 register patches of 6 rows (the r, z, n rows of two units) x 8 columns over all 64 samples, fixed
                // order.  The stage is bound by shared-memory operand traffic, not by threads: 3 x 8 patches on all 256
                // threads, or the samples split over two thread halves, measured the same 7 700 - 8 900 cycles.
                const int mi = tid >> 6, p = tid & 63;
                const int jg = p >> 2, kg = p & 3;                      // units jg, jg + 16; columns kg + 4 b
                const float* Xs = mi == 0 ? X1 : Hp;
                float acc[6][8];
#pragma unroll
                for (int aa = 0; aa < 6; ++aa)
#pragma unroll
                    for (int bb = 0; bb < 8; ++bb) acc[aa][bb] = 0.0f;
                // G rows: da_r (0..31), da_z (32..63), da_n (64..95, ih) / da_hn (96..127, hh)
                const int gn = mi == 1 ? 3 * H : 2 * H;
                int grow[6];
#pragma unroll
                for (int aa = 0; aa < 6; ++aa) grow[aa] = jg + 16 * (aa & 1) + ((aa >> 1) == 2 ? gn : (aa >> 1) * H);
#pragma unroll 2
                for (int q = 0; q < NQ; ++q) {
                    float4 d[6], x[8];
#pragma unroll
                    for (int aa = 0; aa < 6; ++aa) d[aa] = *reinterpret_cast<const float4*>(G + grow[aa] * LD + 4 * q);
#pragma unroll
                    for (int bb = 0; bb < 8; ++bb) x[bb] = *reinterpret_cast<const float4*>(Xs + (kg + 4 * bb) * LD + 4 * q);
#pragma unroll
                    for (int aa = 0; aa < 6; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 8; ++bb) {
                            acc[aa][bb] = fmaf(d[aa].x, x[bb].x, acc[aa][bb]);
                            acc[aa][bb] = fmaf(d[aa].y, x[bb].y, acc[aa][bb]);
                            acc[aa][bb] = fmaf(d[aa].z, x[bb].z, acc[aa][bb]);
                            acc[aa][bb] = fmaf(d[aa].w, x[bb].w, acc[aa][bb]);
                        }
                }
                float* dst = sm + oDG + mi * G3 * GLD;
#pragma unroll
                for (int aa = 0; aa < 6; ++aa)
#pragma unroll
                    for (int bb = 0; bb < 8; ++bb)
                        dst[(jg + 16 * (aa & 1) + (aa >> 1) * H) * GLD + kg + 4 * bb] += acc[aa][bb];
            } else if (tid < 128 + 4 * H) {
                // dbih rows 0..95 (dbhh rows 0..63 are the same sums), dbhh rows 64..95 from da_hn
                const int row = tid - 128;
                float sum = 0.0f;
                for (int q = 0; q < NQ; ++q) {
                    const float4 d = *reinterpret_cast<const float4*>(G + row * LD + 4 * q);
                    sum += (d.x + d.y) + (d.z + d.w);
                }
                if (row < G3) dW[L.bih + row] += sum;
                if (row < 2 * H) dW[L.bhh + row] += sum;
                if (row >= G3) dW[L.bhh + row - H] += sum;
            }
            __syncthreads();
            GTL(8, true);

            // (d) dx1 = (Wih^T da_i) . relu'(x1) in place ; dh carry += Whh^T da_h
            {
                float ax[NU][4], ahh[NU][4];
#pragma unroll
                for (int q = 0; q < NU; ++q)
#pragma unroll
                    for (int c = 0; c < 4; ++c) { ax[q][c] = 0.0f; ahh[q][c] = 0.0f; }
                // packed backward weights: float4 per (gate row, output pair) = (Wih[row][2p], Wih[row][2p+1], Whh[..][2p], Whh[..][2p+1])
                const int pair = (NU * og) >> 1, odd = (NU * og) & 1;
#pragma unroll 8
                for (int row = 0; row < 2 * H; ++row) {              // r and z gates feed both
                    const float4 d = *reinterpret_cast<const float4*>(G + row * LD + s0);
                    const float4 wv = *reinterpret_cast<const float4*>(sm + oWih + (row * (H / 2) + pair) * 4);
                    const float wi[2] = {NU == 2 ? wv.x : (odd ? wv.y : wv.x), wv.y};
                    const float wh[2] = {NU == 2 ? wv.z : (odd ? wv.w : wv.z), wv.w};
                    const float ds[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int q = 0; q < NU; ++q)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            ax[q][c] = fmaf(wi[q], ds[c], ax[q][c]);
                            ahh[q][c] = fmaf(wh[q], ds[c], ahh[q][c]);
                        }
                }
#pragma unroll 8
                for (int row = 2 * H; row < G3; ++row) {
                    const float4 di = *reinterpret_cast<const float4*>(G + row * LD + s0);         // da_n
                    const float4 dh = *reinterpret_cast<const float4*>(G + (row + H) * LD + s0);   // da_hn
                    const float4 wv = *reinterpret_cast<const float4*>(sm + oWih + (row * (H / 2) + pair) * 4);
                    const float wi[2] = {NU == 2 ? wv.x : (odd ? wv.y : wv.x), wv.y};
                    const float wh[2] = {NU == 2 ? wv.z : (odd ? wv.w : wv.z), wv.w};
                    const float dis[4] = {di.x, di.y, di.z, di.w}, dhs[4] = {dh.x, dh.y, dh.z, dh.w};
#pragma unroll
                    for (int q = 0; q < NU; ++q)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            ax[q][c] = fmaf(wi[q], dis[c], ax[q][c]);
                            ahh[q][c] = fmaf(wh[q], dhs[c], ahh[q][c]);
                        }
                }
#pragma unroll
                for (int q = 0; q < NU; ++q) {
                    const int j = j0 + q;
                    float4* px = reinterpret_cast<float4*>(X1 + j * LD + s0);
                    const float4 x = *px;
                    *px = make_float4(x.x > 0.0f ? ax[q][0] : 0.0f, x.y > 0.0f ? ax[q][1] : 0.0f,
                                      x.z > 0.0f ? ax[q][2] : 0.0f, x.w > 0.0f ? ax[q][3] : 0.0f);
                    float4* pd = reinterpret_cast<float4*>(DH + j * LD + s0);
                    float4 d = *pd;
                    d.x += ahh[q][0]; d.y += ahh[q][1]; d.z += ahh[q][2]; d.w += ahh[q][3];
                    *pd = d;
                }
            }
            __syncthreads();
            GTL(9, true);

            // (e) dW1 += dx1 x^T, db1 (+ folded id column): 48 patches of 4 x 4, samples split 4 ways, fixed-order combine
            {
                float* scr = sm + oScr;
                if (tid < 192) {
                    const int q4 = tid / 48, p = tid - q4 * 48;
                    const int jg = p / 6, kg = p - jg * 6;               // rows jg + 8 a, columns kg + 6 b (< 24)
                    float acc[4][4];
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 4; ++bb) acc[aa][bb] = 0.0f;
#pragma unroll
                    for (int qq = 0; qq < NQ / 4; ++qq) {
                        const int q = q4 * (NQ / 4) + qq;
                        float4 d[4], x[4];
#pragma unroll
                        for (int aa = 0; aa < 4; ++aa) d[aa] = *reinterpret_cast<const float4*>(X1 + (jg + 8 * aa) * LD + 4 * q);
#pragma unroll
                        for (int bb = 0; bb < 4; ++bb) x[bb] = *reinterpret_cast<const float4*>(X + (kg + 6 * bb) * LD + 4 * q);
#pragma unroll
                        for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                            for (int bb = 0; bb < 4; ++bb) {
                                acc[aa][bb] = fmaf(d[aa].x, x[bb].x, acc[aa][bb]);
                                acc[aa][bb] = fmaf(d[aa].y, x[bb].y, acc[aa][bb]);
                                acc[aa][bb] = fmaf(d[aa].z, x[bb].z, acc[aa][bb]);
                                acc[aa][bb] = fmaf(d[aa].w, x[bb].w, acc[aa][bb]);
                            }
                    }
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 4; ++bb) scr[(q4 * H + jg + 8 * aa) * KIN + kg + 6 * bb] = acc[aa][bb];
                } else if (tid < 192 + H) {
                    const int j = tid - 192;
                    float acc = 0.0f;
                    for (int q = 0; q < NQ; ++q) {
                        const float4 d = *reinterpret_cast<const float4*>(X1 + j * LD + 4 * q);
                        acc += (d.x + d.y) + (d.z + d.w);
                    }
                    dW[L.b1 + j] += acc;
                    if (a.fold_ids) dW[L.w1 + j * L.in + a.in_rows + g] += acc;
                }
                __syncthreads();
                for (int i2 = tid; i2 < H * KIN; i2 += NT) {
                    const int j = i2 / KIN, k = i2 - j * KIN;
                    if (k < a.in_rows)
                        dW[L.w1 + j * L.in + k] += (scr[(0 * H + j) * KIN + k] + scr[(1 * H + j) * KIN + k]) +
                                                   (scr[(2 * H + j) * KIN + k] + scr[(3 * H + j) * KIN + k]);
                }
            }
            __syncthreads();
            GTL(10, true);
        }
    }

    // per-CTA partial: gradients + statistics (fixed-order reduction follows in reduce_partials_kernel)
    float* out = a.partials + (size_t)blockIdx.x * (L.count + CMARL_N_STATS);
    for (int i = tid; i < L.count; i += NT) {
        float v = dW[i];
        if (i >= L.wih && i < L.bih) {            // dWih | dWhh live in the padded accumulators
            const int k = i - L.wih, mi = k / (G3 * H), rc = k - mi * (G3 * H);
            v = sm[oDG + mi * G3 * GLD + (rc / H) * GLD + (rc % H)];
        }
        out[i] = v;
    }
    float* red = sm + oRed;
#pragma unroll
    for (int k = 0; k < PolicyHead::NSTAT; ++k) {
        const float v = warp_sum_f(st[k]);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            float s = 0.0f;
            for (int w = 0; w < NT / 32; ++w) s += red[w];
            out[L.count + k] = s;
        }
    }
    if (tid == 0)
        for (int k = PolicyHead::NSTAT; k < CMARL_N_STATS; ++k) out[L.count + k] = 0.0f;
}

// ---- K2 (recurrent) alone: one thread per (agent, env); weights read through the read-only path ----------------
struct ActGruArgs {
    const float* params;
    GruLayout L;
    const float* obs;         // [N][O][B]
    const float* h_in;        // [N][H][B] or null
    const uint8_t* avail;     // [N][A][B] or null
    const float* noise;       // [N][A][B]
    int32_t* actions;         // [N][B]
    float* logp;              // [N][B]
    float* logits;            // [N][A][B] or null
    float* h_out;             // [N][H][B]
    int N, B;
};

__global__ void __launch_bounds__(128) actor_act_gru_kernel(ActGruArgs a) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.N * a.B) return;
    const int n = idx / a.B, b = idx - n * a.B;
    const GruLayout& L = a.L;
    const float* P = a.params;
    const int O = L.in;
    float x1[H], h[H];
#pragma unroll
    for (int j = 0; j < H; ++j) x1[j] = __ldcg(P + L.b1 + j);
    for (int k = 0; k < O; ++k) {
        const float xk = a.obs[((size_t)n * O + k) * a.B + b];
#pragma unroll
        for (int j = 0; j < H; ++j) x1[j] = fmaf(__ldcg(P + L.w1 + j * O + k), xk, x1[j]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
        x1[j] = fmaxf(x1[j], 0.0f);
        h[j] = a.h_in ? a.h_in[((size_t)n * H + j) * a.B + b] : 0.0f;
    }
    float z[NA];
#pragma unroll
    for (int c = 0; c < NA; ++c) z[c] = __ldcg(P + L.b2 + c);
    for (int j = 0; j < H; ++j) {
        float ar = __ldcg(P + L.bih + j) + __ldcg(P + L.bhh + j);
        float az = __ldcg(P + L.bih + H + j) + __ldcg(P + L.bhh + H + j);
        float ai = __ldcg(P + L.bih + 2 * H + j), ah = __ldcg(P + L.bhh + 2 * H + j);
#pragma unroll
        for (int k = 0; k < H; ++k) {
            ar = fmaf(__ldcg(P + L.wih + (0 * H + j) * H + k), x1[k], ar);
            az = fmaf(__ldcg(P + L.wih + (1 * H + j) * H + k), x1[k], az);
            ai = fmaf(__ldcg(P + L.wih + (2 * H + j) * H + k), x1[k], ai);
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
            ar = fmaf(__ldcg(P + L.whh + (0 * H + j) * H + k), h[k], ar);
            az = fmaf(__ldcg(P + L.whh + (1 * H + j) * H + k), h[k], az);
            ah = fmaf(__ldcg(P + L.whh + (2 * H + j) * H + k), h[k], ah);
        }
        const float r = sigmoidf_(ar), zz = sigmoidf_(az);
        const float nn = tanhf(ai + r * ah);
        // h[j] is read by later units' dot products: keep the old vector intact, emit the new one directly
        float hj = 0.0f;
#pragma unroll
        for (int k = 0; k < H; ++k) hj = (k == j) ? h[k] : hj;
        const float hn = (hj - nn) * zz + nn;
        a.h_out[((size_t)n * H + j) * a.B + b] = hn;
        const float hr = fmaxf(hn, 0.0f);
#pragma unroll
        for (int c = 0; c < NA; ++c) z[c] = fmaf(__ldcg(P + L.w2 + c * H + j), hr, z[c]);
    }
    float q[NA];
#pragma unroll
    for (int c = 0; c < NA; ++c) {
        if (a.avail && !a.avail[((size_t)n * NA + c) * a.B + b]) z[c] = -1e9f;     // LSTM:182-183
        q[c] = a.noise[((size_t)n * NA + c) * a.B + b];
        if (a.logits) a.logits[((size_t)n * NA + c) * a.B + b] = z[c];
    }
    // Categorical(logits=z).sample() as the exponential race + log_prob (LSTM:172-174)
    float mx = z[0];
#pragma unroll
    for (int c = 1; c < NA; ++c) mx = fmaxf(mx, z[c]);
    float se = 0.0f;
#pragma unroll
    for (int c = 0; c < NA; ++c) se += expf(z[c] - mx);
    const float lse = mx + logf(se);
    float l[NA], p[NA];
    float mx2 = -INFINITY;
#pragma unroll
    for (int c = 0; c < NA; ++c) { l[c] = z[c] - lse; mx2 = fmaxf(mx2, l[c]); }
    float se2 = 0.0f;
#pragma unroll
    for (int c = 0; c < NA; ++c) { p[c] = expf(l[c] - mx2); se2 += p[c]; }
    float best = -1.0f, lp = l[0];
    int action = 0;
#pragma unroll
    for (int c = 0; c < NA; ++c) {
        const float r = (p[c] / se2) / q[c];
        if (r > best) { best = r; action = c; lp = l[c]; }
    }
    a.actions[(size_t)n * a.B + b] = action;
    a.logp[(size_t)n * a.B + b] = lp;
}

}  // namespace gru

extern "C" int cmarl_debug_gru_timeline(long long* out_host16) {
    return (int)cudaMemcpyFromSymbol(out_host16, gru::g_gru_tl, sizeof(long long) * 16);
}

int cmarl_gru_setup(cmarl_ctx* ctx) {
    (void)ctx;
    int e = 0;
#define SETK(NU, ST) if (!e) e = cmarl_check_cuda(cudaFuncSetAttribute(gru::tbptt_chunk_kernel<NU, ST>, \
        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gru::SMEM_BYTES), "cudaFuncSetAttribute(tbptt_chunk_kernel)");
    SETK(1, false) SETK(1, true) SETK(2, false) SETK(2, true)
#undef SETK
    if (!e) e = cmarl_tc_gru_setup();
    return e;
}

extern "C" int cmarl_tbptt_chunk_grads(cmarl_ctx* ctx, const float* actor_params, const float* state, const float* obs,
                                       const int32_t* actions, const float* logp_old, const float* adv,
                                       const uint8_t* mask, const uint8_t* avail, double clip, double ent_coef,
                                       int32_t t0, int32_t t1, float* h_seq, float* stash, float* grads_out,
                                       void* workspace, void* stream) {
    CMARL_ARG(ctx && actor_params && actions && logp_old && adv && h_seq && grads_out && workspace, "null argument");
    CMARL_ARG(ctx->cfg.actor_recurrent, "context was not created with actor_recurrent = 1");
    CMARL_ARG(state || obs, "state or obs required");
    const cmarl_config& c = ctx->cfg;
    CMARL_ARG(0 <= t0 && t0 < t1 && t1 <= c.n_steps, "chunk must satisfy 0 <= t0 < t1 <= n_steps");
    cudaStream_t st = as_stream(stream);
    gru::ChunkArgs a;
    a.params = actor_params; a.L = ctx->gru;
    a.T = c.n_steps; a.N = c.n_agents; a.B = c.n_envs; a.t0 = t0; a.t1 = t1;
    if (obs) {
        a.x = obs; a.in_rows = c.obs_dim; a.fold_ids = 0;
        a.stride_t = (size_t)c.n_agents * c.obs_dim * c.n_envs; a.stride_g = (size_t)c.obs_dim * c.n_envs;
    } else {
        a.x = state; a.in_rows = CMARL_RAW_OBS; a.fold_ids = c.obs_dim > CMARL_RAW_OBS;
        a.stride_t = (size_t)c.state_dim * c.n_envs; a.stride_g = (size_t)CMARL_RAW_OBS * c.n_envs;
    }
    a.h_seq = h_seq;
    a.stash = stash;
    a.partials = reinterpret_cast<float*>(workspace);
    a.head.actions = actions; a.head.logp_old = logp_old; a.head.adv = adv; a.head.mask = mask; a.head.avail = avail;
    a.head.V = ctx->n_heads; a.head.A = c.n_actions;
    a.head.clip = (float)clip; a.head.ent_coef = (float)ent_coef; a.head.inv_groups = 1.0f / (float)c.n_agents;
    a.passes = 3; a.flush = 32;      // weight-gradient accumulators leave TMEM at the end of every tile (<= 25 steps)
    // Which kernels: the tcgen05 pair of tc_gru.cu (tensor cores on and a gate stash given: it is how the two kernels meet),
    // else the fp32 FFMA kernel below.  CMARL_TBPTT = ffma | tc | tcfwd | tcbwd mixes them (cross-checks, measurements).
    // (read at every call: a host-side getenv is noise next to a launch, and tests switch modes within one process)
    const int mode_env = [] {
        const char* v = getenv("CMARL_TBPTT");
        if (!v) return -1;
        return !strcmp(v, "ffma") ? 0 : !strcmp(v, "tc") ? 3 : !strcmp(v, "tcfwd") ? 1 : !strcmp(v, "tcbwd") ? 2 : -1;
    }();
    const int flush_env = [] { const char* v = getenv("CMARL_TC_GRU_FLUSH"); const int n = v ? atoi(v) : 0; return n >= 1 && n <= 64 ? n : 0; }();
    if (flush_env) a.flush = flush_env;
    int tc_which = (ctx->use_tc && stash) ? 3 : 0;
    if (mode_env >= 0 && stash) tc_which = mode_env == 2 ? 3 : mode_env;     // (the tcgen05 backward needs the tcgen05 forward's dlogits)
    {
        // workspace map of the tcgen05 pair (cmarl_workspace_bytes): [0, sm_count) actor rows: backward partials; second half
        // of the actor block: forward partials; behind the actor and critic blocks: dlogits
        const size_t arow = (size_t)ctx->actor.count + CMARL_N_STATS, crow = (size_t)ctx->critic.count + CMARL_N_STATS;
        float* ws = reinterpret_cast<float*>(workspace);
        a.fwd_partials = ws + (size_t)ctx->sm_count * arow;
        a.dlogits = ws + (size_t)2 * ctx->sm_count * (arow + crow);
        const int units_tc = c.n_agents * ceil_div(c.n_envs, 128);
        a.grid_fwd = units_tc < 2 * ctx->sm_count ? units_tc : 2 * ctx->sm_count;
    }
    const int units = c.n_agents * ceil_div(c.n_envs, gru::M);
    int grid = units < ctx->sm_count ? units : ctx->sm_count;
    {
        KernelTimer kt(ctx, K_TBPTT, st);
        if (tc_which & 1) { const int e = cmarl_tc_gru_launch(ctx, a, 1, nullptr, st); if (e) return e; }
        if (tc_which != 3) {
            // units per thread: 2 -> 256 threads (8 warps / SM, the default), 1 -> 512 threads (16 warps / SM).  Measured at
            // 8 192 envs: 0.475 vs 0.533 ms per chunk -- twice the warps leave the gate GEMM at the same 7 900 cycles per
            // step and slow the dx1/dh stage down (11.7 k vs 7.4 k cycles): the stages are not latency-bound.
            static const int nu = [] { const char* v = getenv("CMARL_TBPTT_NU"); return (v && v[0] == '1') ? 1 : 2; }();
            a.passes = tc_which == 1 ? 2 : tc_which == 2 ? 1 : 3;
            cudaError_t ce;
            if (nu == 2) {
                const dim3 block(gru::NTMAX / 2);
                ce = stash ? cmarl_launch(ctx, gru::tbptt_chunk_kernel<2, true>, dim3(grid), block, gru::SMEM_BYTES, st, a)
                           : cmarl_launch(ctx, gru::tbptt_chunk_kernel<2, false>, dim3(grid), block, gru::SMEM_BYTES, st, a);
            } else {
                const dim3 block(gru::NTMAX);
                ce = stash ? cmarl_launch(ctx, gru::tbptt_chunk_kernel<1, true>, dim3(grid), block, gru::SMEM_BYTES, st, a)
                           : cmarl_launch(ctx, gru::tbptt_chunk_kernel<1, false>, dim3(grid), block, gru::SMEM_BYTES, st, a);
            }
            CMARL_CUDA(ce);
        }
        if (tc_which & 2) { const int e = cmarl_tc_gru_launch(ctx, a, 2, &grid, st); if (e) return e; }
    }
    return cmarl_reduce_one_net(ctx, a.partials, grid, ctx->gru.count, nullptr, 0, 0, (float)c.n_agents, grads_out, st);
}

extern "C" int cmarl_actor_act_recurrent(cmarl_ctx* ctx, const float* actor_params, const float* obs, const float* h_in,
                                         const uint8_t* avail, const float* noise, int32_t* actions, float* logp,
                                         float* logits_out, float* h_out, void* stream) {
    CMARL_ARG(ctx && actor_params && obs && noise && actions && logp && h_out, "null argument");
    CMARL_ARG(ctx->cfg.actor_recurrent, "context was not created with actor_recurrent = 1");
    gru::ActGruArgs a;
    a.params = actor_params; a.L = ctx->gru; a.obs = obs; a.h_in = h_in; a.avail = avail; a.noise = noise;
    a.actions = actions; a.logp = logp; a.logits = logits_out; a.h_out = h_out;
    a.N = ctx->cfg.n_agents; a.B = ctx->cfg.n_envs;
    cudaStream_t st = as_stream(stream);
    {
        KernelTimer kt(ctx, K_ACT, st);
        gru::actor_act_gru_kernel<<<ceil_div(a.N * a.B, 128), 128, 0, st>>>(a);
    }
    return cmarl_check_cuda(cudaGetLastError(), "actor_act_gru_kernel");
}
