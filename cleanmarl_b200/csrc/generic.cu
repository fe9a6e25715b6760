// Layered ("generic") kernels: every shape the reference's CLI accepts that the fused kernels are not built for --
// any number of hidden layers (MME:160-171, 186-196: num_layer hidden->hidden blocks), any hidden width <= 256, and
// simple_spread with N agents / L landmarks (1 <= N, L <= 8; PettingZoo simple_spread_v3(N)).  Same entry points, same
// device layouts, same contracts (unnormalised gradient sums + CMARL_N_STATS statistics) as the fused path; the entries in
// rollout.cu / chain.cu / exact.cu route here when cmarl_ctx.generic is set.  Plain fp32 FFMA, deterministic (every
// reduction has a fixed order).  Compiled with -fmad=false (float64 env physics follows numpy's rounding); the GEMMs use
// explicit fmaf.
//
//   rollout      gen_rollout_kernel: one launch for the T steps; block = 32 envs x N agents, thread = (env, agent):
//                observation from the shared-memory env state, the whole MLP per thread (weights in shared memory when
//                they fit), exponential-race sample, then one thread per env integrates World.step (MME:408-453)
//   update       per net, per chunk of time steps: gather X0, one tiled GEMM per Linear layer (bias + ReLU epilogue),
//                per-sample head (heads.cuh: the same code as the fused chains), backward GEMMs (ReLU' epilogue), weight
//                gradients as split-S GEMMs + fixed-order reduction into the flat gradient vector (MME:527-582)
//   Adam         gen_sqsum_kernel (per-tensor sums of squares) + gen_adam_kernel (norm of norms, clip, Adam; MME:584-594)
#include <stdlib.h>

#include "common.cuh"
#include "spread.cuh"
#include "sample.cuh"
#include "chain.cuh"
#include "heads.cuh"

using chain::PolicyHead;
using chain::PolicyHeadArgs;
using chain::ValueHead;
using chain::ValueHeadArgs;

namespace {

constexpr int GE = 32;                       // envs per rollout block
constexpr int MAXN = 8;                      // agents / landmarks
constexpr int NACT = sample::NACT;

struct GenNetDev {                           // GenNet by value for kernels
    int n_lin;
    int dims[CMARL_GEN_MAX_LIN + 1];
    int w_off[CMARL_GEN_MAX_LIN], b_off[CMARL_GEN_MAX_LIN];
    int count;
};
static GenNetDev to_dev(const GenNet& g) {
    GenNetDev d;
    d.n_lin = g.n_lin; d.count = g.count;
    for (int i = 0; i <= CMARL_GEN_MAX_LIN; ++i) d.dims[i] = g.dims[i];
    for (int i = 0; i < CMARL_GEN_MAX_LIN; ++i) { d.w_off[i] = g.w_off[i]; d.b_off[i] = g.b_off[i]; }
    return d;
}

// ------------------------------------------------------------------------------------------------
// simple_spread(N, L): the same operation order as spread.cuh / oracle/spread.py, loops instead of the N = 3 unrolling
// ------------------------------------------------------------------------------------------------
// raw observation of agent n: vel, pos, landmarks - pos, other agents - pos (index order), 2 zeros per other agent
__device__ __forceinline__ void gen_observe(int n, int N, int L, const double* p, const double* v, const double* lm, float* o) {
    const double px = p[2 * n], py = p[2 * n + 1];
    o[0] = (float)v[2 * n]; o[1] = (float)v[2 * n + 1];
    o[2] = (float)px; o[3] = (float)py;
    for (int l = 0; l < L; ++l) {
        o[4 + 2 * l] = (float)(lm[2 * l] - px);
        o[5 + 2 * l] = (float)(lm[2 * l + 1] - py);
    }
    int c = 4 + 2 * L;
    for (int j = 0; j < N; ++j) {
        if (j == n) continue;
        o[c] = (float)(p[2 * j] - px); o[c + 1] = (float)(p[2 * j + 1] - py);
        c += 2;
    }
    for (int j = 0; j < 2 * (N - 1); ++j) o[c + j] = 0.0f;
}

// World.step for one env: p, v updated in place; returns agent 0's reward (pettingzoo_wrapper.py:66)
__device__ __forceinline__ double gen_world_step(int N, int L, double* p, double* v, const double* lm, const int* act) {
    double fx[MAXN], fy[MAXN];
    for (int n = 0; n < N; ++n) {
        double ux = 0.0, uy = 0.0;
        const int a = act[n];
        if (a == 1) ux = -1.0;
        if (a == 2) ux = +1.0;
        if (a == 3) uy = -1.0;
        if (a == 4) uy = +1.0;
        fx[n] = ux * spread::SENSITIVITY + 0.0;
        fy[n] = uy * spread::SENSITIVITY + 0.0;
    }
    for (int a = 0; a < N; ++a)
        for (int b = a + 1; b < N; ++b) {
            double gx, gy;
            spread::pair_force(p[2 * a], p[2 * a + 1], p[2 * b], p[2 * b + 1], gx, gy);
            fx[a] = gx + fx[a]; fy[a] = gy + fy[a];
            fx[b] = -gx + fx[b]; fy[b] = -gy + fy[b];
        }
    for (int n = 0; n < N; ++n) spread::integrate(p[2 * n], p[2 * n + 1], v[2 * n], v[2 * n + 1], fx[n], fy[n]);
    double g = 0.0;
    for (int l = 0; l < L; ++l) {
        double m = spread::dist2d(p[0], p[1], lm[2 * l], lm[2 * l + 1]);
        for (int a = 1; a < N; ++a) m = fmin(m, spread::dist2d(p[2 * a], p[2 * a + 1], lm[2 * l], lm[2 * l + 1]));
        g = g - m;
    }
    double loc = 0.0;
    for (int j = 1; j < N; ++j) loc = loc - 1.0 * (spread::dist2d(p[2 * j], p[2 * j + 1], p[0], p[1]) < spread::DIST_MIN ? 1.0 : 0.0);
    return g * (1 - spread::LOCAL_RATIO) + loc * spread::LOCAL_RATIO;
}

__global__ void __launch_bounds__(256) gen_env_reset_kernel(double* __restrict__ env, int B, int N, int L, uint64_t seed,
                                                            uint64_t episode, const uint64_t* __restrict__ episode_dev) {
    pdl_wait_then_trigger();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (episode_dev) episode = *episode_dev;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    // reset_world order: agent positions, then landmark positions; entity i draws Philox counter 0xE0000000 + i (the same
    // draws as env_reset_kernel for N = L = 3)
    for (int i = 0; i < N + L; ++i) {
        const Philox4 r = philox4x32_10((uint32_t)b, 0xE0000000u + i, (uint32_t)episode, (uint32_t)(episode >> 32), k0, k1);
        const double x = -1.0 + 2.0 * u64_to_unit(r.x, r.y), y = -1.0 + 2.0 * u64_to_unit(r.z, r.w);
        const int row = i < N ? 2 * i : 4 * N + 2 * (i - N);
        env[(size_t)row * B + b] = x;
        env[(size_t)(row + 1) * B + b] = y;
    }
    for (int i = 0; i < 2 * N; ++i) env[(size_t)(2 * N + i) * B + b] = 0.0;
}

__global__ void __launch_bounds__(128) gen_env_step_kernel(double* __restrict__ env, const int32_t* __restrict__ actions,
                                                           float* __restrict__ state_out, float* __restrict__ reward_out,
                                                           int B, int N, int L, int R) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double p[2 * MAXN], v[2 * MAXN], lm[2 * MAXN];
    for (int i = 0; i < 2 * N; ++i) { p[i] = env[(size_t)i * B + b]; v[i] = env[(size_t)(2 * N + i) * B + b]; }
    for (int i = 0; i < 2 * L; ++i) lm[i] = env[(size_t)(4 * N + i) * B + b];
    if (actions) {
        int act[MAXN];
        for (int n = 0; n < N; ++n) act[n] = actions[(size_t)n * B + b];
        const double r = gen_world_step(N, L, p, v, lm, act);
        for (int i = 0; i < 2 * N; ++i) { env[(size_t)i * B + b] = p[i]; env[(size_t)(2 * N + i) * B + b] = v[i]; }
        if (reward_out) reward_out[b] = (float)r;
    }
    if (state_out) {
        float x[4 + 2 * MAXN + 4 * (MAXN - 1)];
        for (int n = 0; n < N; ++n) {
            gen_observe(n, N, L, p, v, lm, x);
            for (int k = 0; k < R; ++k) state_out[(size_t)(n * R + k) * B + b] = x[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The whole MLP for one sample in one thread: activations ping-pong in (L1-cached) local memory, weights are read at
// warp-uniform addresses (shared memory when the network fits, else global / L1), four outputs at a time.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gen_mlp_thread(const GenNetDev& net, const float* __restrict__ W, float* h0, float* h1, float* z) {
    float* in = h0;
    float* out = h1;
    for (int l = 0; l < net.n_lin; ++l) {
        const int K = net.dims[l], J = net.dims[l + 1];
        const float* w = W + net.w_off[l];
        const float* bias = W + net.b_off[l];
        const bool last = l == net.n_lin - 1;
        float* dst = last ? z : out;
        int j = 0;
        for (; j + 3 < J; j += 4) {
            float a0 = bias[j], a1 = bias[j + 1], a2 = bias[j + 2], a3 = bias[j + 3];
            const float* w0 = w + (size_t)j * K;
            for (int k = 0; k < K; ++k) {
                const float x = in[k];
                a0 = fmaf(x, w0[k], a0); a1 = fmaf(x, w0[K + k], a1);
                a2 = fmaf(x, w0[2 * K + k], a2); a3 = fmaf(x, w0[3 * K + k], a3);
            }
            if (!last) { a0 = fmaxf(a0, 0.0f); a1 = fmaxf(a1, 0.0f); a2 = fmaxf(a2, 0.0f); a3 = fmaxf(a3, 0.0f); }
            dst[j] = a0; dst[j + 1] = a1; dst[j + 2] = a2; dst[j + 3] = a3;
        }
        for (; j < J; ++j) {
            float a0 = bias[j];
            const float* w0 = w + (size_t)j * K;
            for (int k = 0; k < K; ++k) a0 = fmaf(in[k], w0[k], a0);
            dst[j] = last ? a0 : fmaxf(a0, 0.0f);
        }
        float* t = in; in = out; out = t;
    }
}

struct GenRolloutArgs {
    const float* actor;
    double* env;            // [4N + 2L][B]
    const float* noise;     // [T][N][A][B] or null
    uint64_t seed, episode;
    const uint64_t* episode_dev;
    float* state;           // [T][N R][B]
    float* obs;             // [T][N][O][B] or null
    int32_t* actions;       // [T][N][B]
    float* logp;            // [T][N][B]
    float* reward;          // [T][B]
    double* ep_return;      // [B] or null
    int T, B, N, L, R, O;
    int w_in_smem;
    GenNetDev net;
};

__global__ void __launch_bounds__(GE * MAXN) gen_rollout_kernel(GenRolloutArgs a) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const int N = a.N, L = a.L, R = a.R, B = a.B;
    const int rows = 4 * N + 2 * L;
    double* es = reinterpret_cast<double*>(gsm);                        // [rows][GE]
    int* acts = reinterpret_cast<int*>(gsm + (size_t)rows * GE * 8);     // [N][GE]
    float* wsm = reinterpret_cast<float*>(gsm + (size_t)rows * GE * 8 + (size_t)N * GE * 4);
    const int tid = threadIdx.x, e = tid % GE, n = tid / GE;            // lane = env
    const int b = blockIdx.x * GE + e;
    const bool live = b < B;
    pdl_wait_then_trigger();
    for (int i = tid; i < rows * GE; i += blockDim.x) {
        const int r = i / GE, c = i - r * GE;
        const int bb = blockIdx.x * GE + c;
        es[i] = bb < B ? a.env[(size_t)r * B + bb] : 0.0;
    }
    const float* W = a.actor;
    if (a.w_in_smem) {
        for (int i = tid; i < a.net.count; i += blockDim.x) wsm[i] = __ldcg(a.actor + i);
        W = wsm;
    }
    __syncthreads();
    const uint64_t episode = a.episode_dev ? *a.episode_dev : a.episode;
    double ep_acc = 0.0;
    float h0[CMARL_GEN_MAX_DIM], h1[CMARL_GEN_MAX_DIM];
    for (int t = 0; t < a.T; ++t) {
        {   // observation before the action (MME:426-430) -> buffers, network input (raw + one-hot id)
            double p[2 * MAXN], v[2 * MAXN], lm[2 * MAXN];
            for (int i = 0; i < 2 * N; ++i) { p[i] = es[i * GE + e]; v[i] = es[(2 * N + i) * GE + e]; }
            for (int i = 0; i < 2 * L; ++i) lm[i] = es[(4 * N + i) * GE + e];
            gen_observe(n, N, L, p, v, lm, h0);
            for (int k = R; k < a.O; ++k) h0[k] = (k - R == n) ? 1.0f : 0.0f;
            if (live) {
                for (int k = 0; k < R; ++k) a.state[((size_t)t * N * R + n * R + k) * B + b] = h0[k];
                if (a.obs)
                    for (int k = 0; k < a.O; ++k) a.obs[(((size_t)t * N + n) * a.O + k) * B + b] = h0[k];
            }
        }
        float z[NACT];
        {
            float zz[8];
            gen_mlp_thread(a.net, W, h0, h1, zz);
#pragma unroll
            for (int k = 0; k < NACT; ++k) z[k] = zz[k];
        }
        float q[NACT];
        if (a.noise) {
#pragma unroll
            for (int k = 0; k < NACT; ++k) q[k] = live ? a.noise[(((size_t)t * N + n) * NACT + k) * B + b] : 1.0f;
        } else {
            sample::philox_exp5(a.seed, episode, (uint32_t)t, (uint32_t)n, (uint32_t)b, q);
        }
        int action; float lp;
        sample::race_sample(z, q, action, lp);
        acts[n * GE + e] = action;
        if (live) {
            a.actions[((size_t)t * N + n) * B + b] = action;
            a.logp[((size_t)t * N + n) * B + b] = lp;
        }
        __syncthreads();
        if (n == 0) {   // World.step of env e
            double p[2 * MAXN], v[2 * MAXN], lm[2 * MAXN];
            int act[MAXN];
            for (int i = 0; i < 2 * N; ++i) { p[i] = es[i * GE + e]; v[i] = es[(2 * N + i) * GE + e]; }
            for (int i = 0; i < 2 * L; ++i) lm[i] = es[(4 * N + i) * GE + e];
            for (int i = 0; i < N; ++i) act[i] = acts[i * GE + e];
            const double r = gen_world_step(N, L, p, v, lm, act);
            for (int i = 0; i < 2 * N; ++i) { es[i * GE + e] = p[i]; es[(2 * N + i) * GE + e] = v[i]; }
            ep_acc += r;
            if (live) a.reward[(size_t)t * B + b] = (float)r;
        }
        __syncthreads();
    }
    if (n == 0 && live && a.ep_return) a.ep_return[b] = ep_acc;
    for (int i = tid; i < 4 * N * GE; i += blockDim.x) {
        const int r = i / GE, c = i - r * GE;
        const int bb = blockIdx.x * GE + c;
        if (bb < B) a.env[(size_t)r * B + bb] = es[i];
    }
}

struct GenActArgs {
    const float* actor; const float* obs; const uint8_t* avail; const float* noise;
    int32_t* actions; float* logp; float* logits;
    int B, N, O;
    GenNetDev net;
};
__global__ void __launch_bounds__(128) gen_actor_act_kernel(GenActArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N * a.B) return;
    const int n = i / a.B, b = i - n * a.B;
    float h0[CMARL_GEN_MAX_DIM], h1[CMARL_GEN_MAX_DIM], zz[8];
    for (int k = 0; k < a.O; ++k) h0[k] = a.obs[((size_t)n * a.O + k) * a.B + b];
    gen_mlp_thread(a.net, a.actor, h0, h1, zz);
    float z[NACT], q[NACT];
#pragma unroll
    for (int k = 0; k < NACT; ++k) {
        z[k] = zz[k];
        if (a.avail && !a.avail[((size_t)n * NACT + k) * a.B + b]) z[k] = -1e9f;     // MME:182
        q[k] = a.noise[((size_t)n * NACT + k) * a.B + b];
        if (a.logits) a.logits[((size_t)n * NACT + k) * a.B + b] = z[k];
    }
    int action; float lp;
    sample::race_sample(z, q, action, lp);
    a.actions[(size_t)n * a.B + b] = action;
    a.logp[(size_t)n * a.B + b] = lp;
}

// ------------------------------------------------------------------------------------------------
// Update path.  Samples of a chunk: s = ((t - t0) * G + g) * nb + b; activations feature-major [rows][S].
// ------------------------------------------------------------------------------------------------
// X0[k][s]: rows k < in_rows from x (+ one-hot id rows of agent group g when the ids are not stored)
__global__ void __launch_bounds__(256) gen_gather_kernel(const float* __restrict__ x, size_t stride_t, size_t stride_g, int ld,
                                                         int in_rows, int id_rows, int G, int nb, int t0, int S,
                                                         float* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int b = s % nb, r = s / nb, g = r % G, t = t0 + r / G;
    const float* base = x + (size_t)t * stride_t + (size_t)g * stride_g + b;
    for (int k = 0; k < in_rows; ++k) out[(size_t)k * S + s] = __ldcg(base + (size_t)k * ld);
    for (int k = 0; k < id_rows; ++k) out[(size_t)(in_rows + k) * S + s] = (k == g) ? 1.0f : 0.0f;
}

// C[M][N] = epi(A * Bm): A(m, k) = Aw[m * a_rs + k * a_cs] (a weight matrix or its transpose), Bm [K][N], C [M][N], N = samples.
//   EPI 0: relu(acc + bias[m])   EPI 1: acc + bias[m]   EPI 2: mask[m][n] > 0 ? acc : 0   (ReLU' through the layer's own output)
constexpr int TM = 64, TN = 64, TK = 16;
template <int EPI>
__global__ void __launch_bounds__(256) gen_gemm_kernel(const float* __restrict__ Aw, int a_rs, int a_cs, const float* __restrict__ Bm,
                                                       const float* __restrict__ bias, const float* __restrict__ mask,
                                                       float* __restrict__ C, int M, int N, int K) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;          // 16 x 16 threads, 4 x 4 outputs each
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int i = tid; i < TK * TM; i += 256) {
            const int kk = i / TM, mm = i - kk * TM;
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? __ldcg(Aw + (size_t)m * a_rs + (size_t)k * a_cs) : 0.0f;
        }
        for (int i = tid; i < TK * TN; i += 256) {
            const int kk = i / TN, nn = i - kk * TN;
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < K) ? Bm[(size_t)k * N + n] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][4 * ty]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][4 * tx]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + 4 * ty + i;
        if (m >= M) continue;
        const float bm = (EPI == 2) ? 0.0f : bias[m];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 4 * tx + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (EPI == 0) v = fmaxf(v + bm, 0.0f);
            if (EPI == 1) v = v + bm;
            if (EPI == 2) v = mask[(size_t)m * N + n] > 0.0f ? v : 0.0f;
            C[(size_t)m * N + n] = v;
        }
    }
}

// Weight gradient, split over the samples: part[split][j][k] = sum_{s in split} dPre[j][s] * Xin[k][s]  (k == K: ones -> bias)
__global__ void __launch_bounds__(256) gen_dw_kernel(const float* __restrict__ dPre, const float* __restrict__ Xin, int J, int K,
                                                     int S, int s_per, float* __restrict__ part) {
    __shared__ float Ds[TK][TM + 4];       // [s][j]
    __shared__ float Xs[TK][TN + 4];       // [s][k]
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int j0 = blockIdx.y * TM, k0 = blockIdx.z * TN;
    const int s_lo = blockIdx.x * s_per, s_hi = min(S, s_lo + s_per);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int s0 = s_lo; s0 < s_hi; s0 += TK) {
        for (int i = tid; i < TK * TM; i += 256) {
            const int jj = i / TK, ss = i - jj * TK;                    // s fastest: coalesced rows
            const int j = j0 + jj, s = s0 + ss;
            Ds[ss][jj] = (j < J && s < s_hi) ? dPre[(size_t)j * S + s] : 0.0f;
        }
        for (int i = tid; i < TK * TN; i += 256) {
            const int kk = i / TK, ss = i - kk * TK;
            const int k = k0 + kk, s = s0 + ss;
            Xs[ss][kk] = (s < s_hi) ? (k < K ? Xin[(size_t)k * S + s] : (k == K ? 1.0f : 0.0f)) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int ss = 0; ss < TK; ++ss) {
            const float4 dv = *reinterpret_cast<const float4*>(&Ds[ss][4 * ty]);
            const float4 xv = *reinterpret_cast<const float4*>(&Xs[ss][4 * tx]);
            const float d4[4] = {dv.x, dv.y, dv.z, dv.w}, x4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d4[i], x4[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = part + (size_t)blockIdx.x * J * (K + 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = j0 + 4 * ty + i;
        if (j >= J) continue;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int k = k0 + 4 * tx + jj;
            if (k <= K) out[(size_t)j * (K + 1) + k] = acc[i][jj];
        }
    }
}

// fixed-order sum over the splits -> gW [J][K], gb [J] of the flat gradient vector (first chunk stores, later chunks add)
__global__ void __launch_bounds__(256) gen_dw_reduce_kernel(const float* __restrict__ part, int nsplit, int J, int K,
                                                            float* __restrict__ gW, float* __restrict__ gb, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J * (K + 1)) return;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    const size_t stride = (size_t)J * (K + 1);
    int c = 0;
    for (; c + 3 < nsplit; c += 4) {
        a0 += part[c * stride + i]; a1 += part[(c + 1) * stride + i];
        a2 += part[(c + 2) * stride + i]; a3 += part[(c + 3) * stride + i];
    }
    for (; c < nsplit; ++c) a0 += part[c * stride + i];
    const float v = (a0 + a1) + (a2 + a3);
    const int j = i / (K + 1), k = i - j * (K + 1);
    float* dst = k < K ? gW + (size_t)j * K + k : gb + j;
    *dst = accumulate ? *dst + v : v;
}

// Per-sample heads (the code of heads.cuh): Z [OUT][S] -> dZ [OUT][S], statistics as per-block partial sums [blocks][8]
template <class Head>
__global__ void __launch_bounds__(256) gen_head_kernel(typename Head::Args ha, const float* __restrict__ Z, float* __restrict__ dZ,
                                                       int G, int nb, int ld, int t0, int S, float* __restrict__ stat_part) {
    constexpr int OUT = Head::OUT;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    float st[Head::NSTAT];
#pragma unroll
    for (int k = 0; k < Head::NSTAT; ++k) st[k] = 0.0f;
    if (s < S) {
        const int b = s % nb, r = s / nb, g = r % G, t = t0 + r / G;
        float z[OUT], dz[OUT];
#pragma unroll
        for (int a = 0; a < OUT; ++a) z[a] = Z[(size_t)a * S + s];
        Head::apply(ha, z, t, g, b, G, ld, true, true, dz, st);
#pragma unroll
        for (int a = 0; a < OUT; ++a) dZ[(size_t)a * S + s] = dz[a];
    }
    __shared__ float red[8][Head::NSTAT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < Head::NSTAT; ++k) {
        const float v = chain::warp_sum_f(st[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < CMARL_N_STATS) {
        float v = 0.0f;
        if (threadIdx.x < Head::NSTAT)
            for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        stat_part[(size_t)blockIdx.x * CMARL_N_STATS + threadIdx.x] = v;
    }
}

// forward-only value head: values[t][g][b] <- Z (contiguous when nb == ld: the output GEMM writes there directly)
// statistics of one net: fixed-order sums of the per-block partials (the count in double) -> the 8 statistics slots
//   policy: st 0 loss 1 entropy 2 kl 3 clipfrac 4 samples -> out 0, 2, 3, 4, 5 (count / G);  value: st 0 loss (1 samples) -> out 1
__global__ void gen_stats_reduce_kernel(const float* __restrict__ part, int nblocks, int is_policy, float n_groups,
                                        float* __restrict__ out_stats, int accumulate) {
    const int k = threadIdx.x;
    if (k >= CMARL_N_STATS) return;
    double d = 0.0;
    for (int c = 0; c < nblocks; ++c) d += (double)part[(size_t)c * CMARL_N_STATS + k];
    int dst = -1;
    double v = d;
    if (is_policy) {
        if (k == 0) dst = 0;
        else if (k >= 1 && k <= 3) dst = k + 1;
        else if (k == 4) { dst = 5; v = d / (double)n_groups; }
    } else if (k == 0) dst = 1;
    if (dst >= 0) out_stats[dst] = accumulate ? out_stats[dst] + (float)v : (float)v;
}

__global__ void gen_zero_stats_kernel(float* __restrict__ out_stats) {
    if (threadIdx.x < CMARL_N_STATS) out_stats[threadIdx.x] = 0.0f;
}

// ------------------------------------------------------------------------------------------------
// Adam for any tensor list: same arithmetic as clip_adam_kernel (exact.cu), two launches.
// ------------------------------------------------------------------------------------------------
constexpr int GEN_MAX_TENSORS = 4 * CMARL_GEN_MAX_LIN;
struct GenAdamArgs {
    float* params; const float* grads; float* m; float* v; float* stats_out;
    int32_t* step_dev; unsigned int* ticket; float* tsq;
    int step, n_tensors, n_actor_tensors, P;
    int tensor_off[GEN_MAX_TENSORS + 1];
    double lr[2], beta1, beta2, eps, max_norm, wd[2];
    float extra_div; int raw_stats;
};

// one block per tensor: sum of squares of g / count in a fixed order
__global__ void __launch_bounds__(256) gen_sqsum_kernel(GenAdamArgs a) {
    const int k = blockIdx.x;
    const float count = a.grads[a.P + 5] * a.extra_div;
    const int lo = a.tensor_off[k], hi = a.tensor_off[k + 1];
    float s = 0.0f;
    for (int i = lo + threadIdx.x; i < hi; i += 256) {
        const float g = a.grads[i] / count;
        s += g * g;
    }
    __shared__ float red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w >= 1; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.tsq[k] = red[0];
}

__global__ void __launch_bounds__(256) gen_adam_kernel(GenAdamArgs a) {
    __shared__ double bc_sh[2];
    __shared__ float norm_sh[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) {
        int step = a.step;
        if (a.step_dev) {
            step = *reinterpret_cast<volatile int32_t*>(a.step_dev) + 1;
            __threadfence();
            const unsigned t = atomicAdd(a.ticket, 1u);
            if (t == gridDim.x - 1) { *a.ticket = 0; *a.step_dev = step; }
        }
        double p1 = 1.0, p2 = 1.0, b1 = a.beta1, b2 = a.beta2;
        for (unsigned e = (unsigned)step; e; e >>= 1) {
            if (e & 1u) { p1 *= b1; p2 *= b2; }
            b1 *= b1; b2 *= b2;
        }
        bc_sh[0] = 1.0 - p1;
        bc_sh[1] = sqrt(1.0 - p2);
        // norm of the per-tensor norms (norm_d, MME:221-224)
        for (int net = 0; net < 2; ++net) {
            float s = 0.0f;
            const int k0 = net == 0 ? 0 : a.n_actor_tensors, k1 = net == 0 ? a.n_actor_tensors : a.n_tensors;
            for (int k = k0; k < k1; ++k) { const float tn = sqrtf(a.tsq[k]); s += tn * tn; }
            norm_sh[net] = sqrtf(s);
        }
    }
    __syncthreads();
    const float n_valid = a.grads[a.P + 5];
    const float count = n_valid * a.extra_div;
    float coef[2] = {1.0f, 1.0f};
    if (a.max_norm > 0.0) {
#pragma unroll
        for (int net = 0; net < 2; ++net) {
            const float c = (float)a.max_norm / (norm_sh[net] + 1e-6f);
            coef[net] = c < 1.0f ? c : 1.0f;
        }
    }
    if (i < a.P) {
        const int actor_end = a.tensor_off[a.n_actor_tensors];
        const int net = i < actor_end ? 0 : 1;
        const double bc1 = bc_sh[0];
        const float bc2_sqrt = (float)bc_sh[1];
        const float w1 = (float)(1.0 - a.beta1), b2 = (float)a.beta2, w2 = (float)(1.0 - a.beta2), eps = (float)a.eps;
        const float nss = (float)(-(a.lr[net] / bc1));
        float gi = a.grads[i] / count;
        if (a.max_norm > 0.0) gi = gi * coef[net];
        float m = a.m[i], v = a.v[i];
        m = fmaf(w1, gi - m, m);
        v = v * b2;
        v = v + (w2 * gi) * gi;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        float pw = a.params[i];
        if (a.wd[net] != 0.0) pw = pw * (float)(1.0 - a.lr[net] * a.wd[net]);
        a.params[i] = pw + (nss * m) / denom;
        a.m[i] = m;
        a.v[i] = v;
    }
    if (i == 0 && a.stats_out) {
        for (int k = 0; k < 5; ++k) a.stats_out[k] = a.grads[a.P + k] / count;
        a.stats_out[5] = norm_sh[0];
        a.stats_out[6] = norm_sh[1];
        a.stats_out[7] = count;
    }
}

}  // namespace

// ================================================================================================
// Host side
// ================================================================================================
static int gen_rows(const cmarl_ctx* ctx) { return 4 * ctx->cfg.n_agents + 2 * ctx->n_landmarks; }

int cmarl_gen_env_reset(cmarl_ctx* ctx, double* env, uint64_t seed, uint64_t episode, cudaStream_t st) {
    const int B = ctx->cfg.n_envs;
    KernelTimer kt(ctx, K_RESET, st);
    return cmarl_check_cuda(cmarl_launch(ctx, gen_env_reset_kernel, dim3(ceil_div(B, 256)), dim3(256), 0, st, env, B, ctx->cfg.n_agents,
                                         ctx->n_landmarks, seed, episode, (const uint64_t*)ctx->episode_dev),
                            "gen_env_reset_kernel");
}

int cmarl_gen_env_step(cmarl_ctx* ctx, double* env, const int32_t* actions, float* state_out, float* reward_out, cudaStream_t st) {
    const int B = ctx->cfg.n_envs;
    KernelTimer kt(ctx, K_ENVSTEP, st);
    gen_env_step_kernel<<<ceil_div(B, 128), 128, 0, st>>>(env, actions, state_out, reward_out, B, ctx->cfg.n_agents, ctx->n_landmarks,
                                                          ctx->raw_obs);
    return cmarl_check_cuda(cudaGetLastError(), "gen_env_step_kernel");
}

static size_t gen_rollout_smem(const cmarl_ctx* ctx, bool with_weights) {
    return (size_t)gen_rows(ctx) * GE * 8 + (size_t)ctx->cfg.n_agents * GE * 4 + (with_weights ? (size_t)ctx->gactor.count * 4 : 0);
}

int cmarl_gen_setup(cmarl_ctx* ctx) {
    const bool fits = gen_rollout_smem(ctx, true) <= 200 * 1024;
    return cmarl_check_cuda(cudaFuncSetAttribute(gen_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)gen_rollout_smem(ctx, fits)),
                            "cudaFuncSetAttribute(gen_rollout_kernel)");
}

int cmarl_gen_rollout(cmarl_ctx* ctx, const float* actor_params, double* env, const float* noise, uint64_t seed, uint64_t episode,
                      float* state, float* obs, int32_t* actions, float* logp, float* reward, double* ep_return, cudaStream_t st) {
    GenRolloutArgs a;
    a.actor = actor_params; a.env = env; a.noise = noise; a.seed = seed; a.episode = episode; a.episode_dev = ctx->episode_dev;
    a.state = state; a.obs = obs; a.actions = actions; a.logp = logp; a.reward = reward; a.ep_return = ep_return;
    a.T = ctx->cfg.n_steps; a.B = ctx->cfg.n_envs; a.N = ctx->cfg.n_agents; a.L = ctx->n_landmarks; a.R = ctx->raw_obs;
    a.O = ctx->cfg.obs_dim;
    a.w_in_smem = gen_rollout_smem(ctx, true) <= 200 * 1024;
    a.net = to_dev(ctx->gactor);
    KernelTimer kt(ctx, K_ROLLOUT, st);
    return cmarl_check_cuda(cmarl_launch(ctx, gen_rollout_kernel, dim3(ceil_div(a.B, GE)), dim3(GE * a.N), gen_rollout_smem(ctx, a.w_in_smem),
                                         st, a),
                            "gen_rollout_kernel");
}

int cmarl_gen_actor_act(cmarl_ctx* ctx, const float* actor_params, const float* obs, const uint8_t* avail, const float* noise,
                        int32_t* actions, float* logp, float* logits_out, cudaStream_t st) {
    GenActArgs a;
    a.actor = actor_params; a.obs = obs; a.avail = avail; a.noise = noise; a.actions = actions; a.logp = logp; a.logits = logits_out;
    a.B = ctx->cfg.n_envs; a.N = ctx->cfg.n_agents; a.O = ctx->cfg.obs_dim;
    a.net = to_dev(ctx->gactor);
    KernelTimer kt(ctx, K_ACT, st);
    gen_actor_act_kernel<<<ceil_div(a.N * a.B, 128), 128, 0, st>>>(a);
    return cmarl_check_cuda(cudaGetLastError(), "gen_actor_act_kernel");
}

// ---- workspace ---------------------------------------------------------------------------------------------------------
constexpr int GEN_CHUNK_SAMPLES = 1 << 18;
constexpr int GEN_MAX_SPLIT = 128;
struct GenPlan {
    int nt_chunk;          // time steps per chunk
    size_t S_max;          // samples of the largest chunk
    size_t act_floats;     // all activations of one net for a chunk
    size_t grad_floats;    // dPre ping-pong
    size_t part_floats;    // weight-gradient partials
    size_t stat_floats;    // head statistics partials
};
static GenPlan gen_plan(const GenNet& net, int T, int G, int nb) {
    GenPlan p;
    const size_t per_t = (size_t)G * nb;
    p.nt_chunk = (int)(GEN_CHUNK_SAMPLES / per_t);
    if (p.nt_chunk < 1) p.nt_chunk = 1;
    if (p.nt_chunk > T) p.nt_chunk = T;
    p.S_max = per_t * p.nt_chunk;
    size_t rows = 0;
    int wmax = 0, jk = 0;
    for (int l = 0; l <= net.n_lin; ++l) { rows += net.dims[l]; if (l > 0 && net.dims[l] > wmax) wmax = net.dims[l]; }
    for (int l = 0; l < net.n_lin; ++l) { const int v = net.dims[l + 1] * (net.dims[l] + 1); if (v > jk) jk = v; }
    p.act_floats = rows * p.S_max;
    p.grad_floats = (size_t)2 * wmax * p.S_max;
    p.part_floats = (size_t)GEN_MAX_SPLIT * jk;
    p.stat_floats = (size_t)ceil_div((int)p.S_max, 256) * CMARL_N_STATS;
    return p;
}
static size_t plan_floats(const GenPlan& p) { return p.act_floats + p.grad_floats + p.part_floats + p.stat_floats + 64; }

size_t cmarl_gen_workspace_bytes(const cmarl_ctx* ctx) {
    const cmarl_config& c = ctx->cfg;
    const GenPlan pa = gen_plan(ctx->gactor, c.n_steps, c.n_agents, c.n_envs);
    const GenPlan pc = gen_plan(ctx->gcritic, c.n_steps, c.critic_on_obs ? c.n_agents : 1, c.n_envs);
    const size_t a = plan_floats(pa), b = plan_floats(pc);
    return (a > b ? a : b) * sizeof(float);
}

struct GenInput {            // where X0 comes from
    const float* x; size_t stride_t, stride_g; int in_rows, id_rows, G;
};

// forward (+ backward when `train`) of one net over all time steps in chunks
template <class Head>
static int gen_run_net(cmarl_ctx* ctx, const GenNet& net, const float* params, const GenInput& in, const typename Head::Args& ha,
                       bool train, int nb, float* ws, float* g_net, float* out_stats, bool is_policy, float* values_out,
                       cudaStream_t st) {
    const cmarl_config& c = ctx->cfg;
    const int T = c.n_steps, ld = c.n_envs, G = in.G;
    const GenPlan plan = gen_plan(net, T, G, nb);
    float* acts = ws;
    float* gbuf = acts + plan.act_floats;
    float* part = gbuf + plan.grad_floats;
    float* spart = part + plan.part_floats;
    int chunk = 0;
    for (int t0 = 0; t0 < T; t0 += plan.nt_chunk, ++chunk) {
        const int nt = (t0 + plan.nt_chunk <= T) ? plan.nt_chunk : T - t0;
        const int S = nt * G * nb;
        // activations A_0 .. A_{n_lin}
        float* A[CMARL_GEN_MAX_LIN + 1];
        {
            float* p = acts;
            for (int l = 0; l <= net.n_lin; ++l) { A[l] = p; p += (size_t)net.dims[l] * S; }
        }
        gen_gather_kernel<<<ceil_div(S, 256), 256, 0, st>>>(in.x, in.stride_t, in.stride_g, ld, in.in_rows, in.id_rows, G, nb, t0, S, A[0]);
        for (int l = 0; l < net.n_lin; ++l) {
            const int K = net.dims[l], J = net.dims[l + 1];
            const dim3 grid(ceil_div(S, TN), ceil_div(J, TM));
            float* Cout = A[l + 1];
            if (l == net.n_lin - 1) {
                if (!train && values_out && nb == ld) Cout = values_out + (size_t)t0 * G * ld;      // values [T][G][B]
                gen_gemm_kernel<1><<<grid, 256, 0, st>>>(params + net.w_off[l], K, 1, A[l], params + net.b_off[l], nullptr, Cout, J, S, K);
            } else {
                gen_gemm_kernel<0><<<grid, 256, 0, st>>>(params + net.w_off[l], K, 1, A[l], params + net.b_off[l], nullptr, Cout, J, S, K);
            }
        }
        if (!train) continue;
        // head: Z -> dZ (+ statistics)
        float* dcur = gbuf;
        float* dnext = gbuf + (size_t)plan.grad_floats / 2;
        const int hb = ceil_div(S, 256);
        gen_head_kernel<Head><<<hb, 256, 0, st>>>(ha, A[net.n_lin], dcur, G, nb, ld, t0, S, spart);
        gen_stats_reduce_kernel<<<1, 32, 0, st>>>(spart, hb, is_policy ? 1 : 0, (float)G, out_stats, chunk > 0 ? 1 : 0);
        for (int l = net.n_lin - 1; l >= 0; --l) {
            const int K = net.dims[l], J = net.dims[l + 1];
            int nsplit = ceil_div(S, 2048);
            if (nsplit > GEN_MAX_SPLIT) nsplit = GEN_MAX_SPLIT;
            int s_per = ceil_div(S, nsplit);
            s_per = ceil_div(s_per, TK) * TK;
            nsplit = ceil_div(S, s_per);
            gen_dw_kernel<<<dim3(nsplit, ceil_div(J, TM), ceil_div(K + 1, TN)), 256, 0, st>>>(dcur, A[l], J, K, S, s_per, part);
            gen_dw_reduce_kernel<<<ceil_div(J * (K + 1), 256), 256, 0, st>>>(part, nsplit, J, K, g_net + net.w_off[l], g_net + net.b_off[l],
                                                                            chunk > 0 ? 1 : 0);
            if (l > 0) {   // dPre_{l-1} = (W_l^T dPre_l) . relu'(A_l)
                gen_gemm_kernel<2><<<dim3(ceil_div(S, TN), ceil_div(K, TM)), 256, 0, st>>>(params + net.w_off[l], 1, K, dcur, nullptr, A[l], dnext,
                                                                                          K, S, J);
                float* t = dcur; dcur = dnext; dnext = t;
            }
        }
    }
    return cmarl_check_cuda(cudaGetLastError(), "generic net kernels");
}

static GenInput gen_actor_input(const cmarl_ctx* ctx, const float* state, const float* obs) {
    const cmarl_config& c = ctx->cfg;
    GenInput in;
    in.G = c.n_agents;
    if (obs) { in.x = obs; in.stride_t = (size_t)c.n_agents * c.obs_dim * c.n_envs; in.stride_g = (size_t)c.obs_dim * c.n_envs; in.in_rows = c.obs_dim; in.id_rows = 0; }
    else { in.x = state; in.stride_t = (size_t)c.state_dim * c.n_envs; in.stride_g = (size_t)ctx->raw_obs * c.n_envs; in.in_rows = ctx->raw_obs; in.id_rows = c.obs_dim - ctx->raw_obs; }
    return in;
}
static GenInput gen_critic_input(const cmarl_ctx* ctx, const float* state, const float* obs) {
    const cmarl_config& c = ctx->cfg;
    if (c.critic_on_obs) return gen_actor_input(ctx, state, obs);
    GenInput in;
    in.G = 1; in.x = state; in.stride_t = (size_t)c.state_dim * c.n_envs; in.stride_g = 0; in.in_rows = c.state_dim; in.id_rows = 0;
    return in;
}

int cmarl_gen_critic_values(cmarl_ctx* ctx, const float* critic_params, const float* state, const float* obs, float* values,
                            void* workspace, cudaStream_t st) {
    ValueHeadArgs ha;
    ha.returns = nullptr; ha.mask = nullptr; ha.values_out = values; ha.inv_heads = 1.0f / (float)ctx->n_heads; ha.values_old = nullptr; ha.vclip = 0.0f;
    KernelTimer kt(ctx, K_CRITIC, st);
    return gen_run_net<ValueHead>(ctx, ctx->gcritic, critic_params, gen_critic_input(ctx, state, obs), ha, false, ctx->cfg.n_envs,
                                  reinterpret_cast<float*>(workspace), nullptr, nullptr, false, values, st);
}

int cmarl_gen_ppo_epoch_grads(cmarl_ctx* ctx, const float* params, const float* state, const float* obs, const int32_t* actions,
                              const float* logp_old, const float* adv, const float* returns, const float* values_old,
                              const uint8_t* mask, const uint8_t* avail, double clip, double ent_coef, double value_clip,
                              int env_count, float* grads_out, void* workspace, cudaStream_t st) {
    const cmarl_config& c = ctx->cfg;
    const int Pa = ctx->gactor.count, Pc = ctx->gcritic.count;
    float* ws = reinterpret_cast<float*>(workspace);
    float* stats = grads_out + Pa + Pc;
    gen_zero_stats_kernel<<<1, 32, 0, st>>>(stats);
    PolicyHeadArgs pa;
    pa.actions = actions; pa.logp_old = logp_old; pa.adv = adv; pa.mask = mask; pa.avail = avail;
    pa.V = ctx->n_heads; pa.A = c.n_actions;
    pa.clip = (float)clip; pa.ent_coef = (float)ent_coef; pa.inv_groups = 1.0f / (float)c.n_agents;
    int e;
    {
        KernelTimer kt(ctx, K_PPO_ACTOR, st);
        e = gen_run_net<PolicyHead>(ctx, ctx->gactor, params, gen_actor_input(ctx, state, obs), pa, true, env_count, ws, grads_out, stats, true,
                                    nullptr, st);
    }
    if (e) return e;
    ValueHeadArgs va;
    va.returns = returns; va.mask = mask; va.values_out = nullptr; va.inv_heads = 1.0f / (float)ctx->n_heads;
    va.values_old = values_old; va.vclip = value_clip > 0.0 ? (float)value_clip : 0.0f;
    {
        KernelTimer kt(ctx, K_PPO_CRITIC, st);
        e = gen_run_net<ValueHead>(ctx, ctx->gcritic, params + Pa, gen_critic_input(ctx, state, obs), va, true, env_count, ws, grads_out + Pa,
                                   stats, false, nullptr, st);
    }
    return e;
}

int cmarl_gen_clip_adam_step(cmarl_ctx* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t step,
                             int32_t* step_dev, double lr_actor, double lr_critic, double beta1, double beta2, double eps,
                             double max_norm, float* stats_out, cudaStream_t st) {
    GenAdamArgs a;
    a.params = params; a.grads = grads; a.m = exp_avg; a.v = exp_avg_sq; a.stats_out = stats_out;
    a.step_dev = step_dev; a.step = step; a.ticket = ctx->dev_words + CMARL_DW_ADAM_TICKET; a.tsq = ctx->dev_floats;
    int k = 0, base = 0;
    const GenNet* nets[2] = {&ctx->gactor, &ctx->gcritic};
    for (int n = 0; n < 2; ++n) {
        for (int l = 0; l < nets[n]->n_lin; ++l) {
            a.tensor_off[k++] = base + nets[n]->w_off[l];
            a.tensor_off[k++] = base + nets[n]->b_off[l];
        }
        if (n == 0) a.n_actor_tensors = k;
        base += nets[n]->count;
    }
    a.tensor_off[k] = base;
    a.n_tensors = k; a.P = base;
    a.lr[0] = lr_actor; a.lr[1] = lr_critic; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
    a.wd[0] = ctx->weight_decay[0]; a.wd[1] = ctx->weight_decay[1];
    a.extra_div = 1.0f; a.raw_stats = 0;
    KernelTimer kt(ctx, K_ADAM, st);
    gen_sqsum_kernel<<<a.n_tensors, 256, 0, st>>>(a);
    gen_adam_kernel<<<ceil_div(a.P, 256), 256, 0, st>>>(a);
    return cmarl_check_cuda(cudaGetLastError(), "generic Adam kernels");
}
