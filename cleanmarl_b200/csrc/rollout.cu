// K1 + K2 + K3: device-resident rollout (simple_spread_v3 step + Actor.act + buffer stores) and the
// stand-alone Actor.act / env reset entries.  Compiled with -fmad=false (float64 physics must follow
// the oracle's rounding); the MLP uses explicit fmaf.
//
// The env state (18 doubles per env) lives in shared memory for the whole episode; the T steps run inside
// ONE launch (no host round trip per step, MME:408-453).  Thread mapping of the rollout: see rollout_kernel;
// the stand-alone Actor.act entry keeps one thread per (env, agent).
#include "common.cuh"
#include "spread.cuh"
#include "sample.cuh"
#include "tc_ptx.cuh"
#include "tc_tile.cuh"

namespace {

constexpr int EPB = 32;            // envs per CTA
constexpr int NAG = 3;             // agents
constexpr int NACT = sample::NACT;
using sample::race_sample;
using sample::philox_exp5;
constexpr int RT = EPB * NAG;      // threads per CTA

template <int H>
struct ActorSmem {
    // float offsets
    static constexpr int oW1T = 0;                        // [21][H] in-major
    static constexpr int oB1 = oW1T + 21 * H;             // [4][H]  b1 (+ folded id column per agent)
    static constexpr int oW2T = oB1 + 4 * H;              // [H][H]  in-major
    static constexpr int oB2 = oW2T + H * H;              // [H]
    static constexpr int oW3T = oB2 + H;                  // [H][8]
    static constexpr int oB3 = oW3T + H * 8;              // [8]
    static constexpr int oAct = oB3 + 8;                  // [H][RT] per-thread activation column
    static constexpr int oEnd = oAct + H * RT;
};

template <int H>
__device__ void load_actor(float* sm, const float* __restrict__ P, int O, bool fold, int nthreads) {
    using S = ActorSmem<H>;
    const float* W1 = P;
    const float* b1 = W1 + H * O;
    const float* W2 = b1 + H;
    const float* b2 = W2 + H * H;
    const float* W3 = b2 + H;
    const float* b3 = W3 + NACT * H;
    for (int i = threadIdx.x; i < 21 * H; i += nthreads) {
        const int k = i / H, j = i - k * H;
        sm[S::oW1T + i] = (k < O) ? W1[j * O + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < 4 * H; i += nthreads) {
        const int g = i / H, j = i - g * H;
        float v = b1[j];
        if (fold && g < NAG) v += W1[j * O + CMARL_RAW_OBS + g];
        sm[S::oB1 + i] = v;
    }
    for (int i = threadIdx.x; i < H * H; i += nthreads) {
        const int j = i / H, k = i - j * H;
        sm[S::oW2T + k * H + j] = W2[i];
    }
    for (int i = threadIdx.x; i < H; i += nthreads) sm[S::oB2 + i] = b2[i];
    for (int i = threadIdx.x; i < H * 8; i += nthreads) {
        const int j = i / 8, a = i - j * 8;
        sm[S::oW3T + i] = (a < NACT) ? W3[a * H + j] : 0.0f;
    }
    if (threadIdx.x < 8) sm[S::oB3 + threadIdx.x] = (threadIdx.x < NACT) ? b3[threadIdx.x] : 0.0f;
}

// Actor.logits (MME:178-183) for one agent: x[KX] -> z[5].  Weights are broadcast LDS.128 reads;
// hidden activations round-trip through this thread's private smem column (stride = nthreads).
template <int H, int KX>
__device__ __forceinline__ void actor_mlp(const float (&x)[KX], const float* __restrict__ sm, const float* b1,
                                          float* __restrict__ col, int stride, float (&z)[NACT]) {
    using S = ActorSmem<H>;
    float acc[H];
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = b1[j];
#pragma unroll
    for (int k = 0; k < KX; ++k) {
        const float xk = x[k];
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(sm + S::oW1T + k * H + j);
            acc[j] = fmaf(xk, w.x, acc[j]); acc[j + 1] = fmaf(xk, w.y, acc[j + 1]);
            acc[j + 2] = fmaf(xk, w.z, acc[j + 2]); acc[j + 3] = fmaf(xk, w.w, acc[j + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < H; ++j) col[j * stride] = fmaxf(acc[j], 0.0f);
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = sm[S::oB2 + j];
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
        const float hk = col[k * stride];
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(sm + S::oW2T + k * H + j);
            acc[j] = fmaf(hk, w.x, acc[j]); acc[j + 1] = fmaf(hk, w.y, acc[j + 1]);
            acc[j + 2] = fmaf(hk, w.z, acc[j + 2]); acc[j + 3] = fmaf(hk, w.w, acc[j + 3]);
        }
    }
#pragma unroll
    for (int a = 0; a < NACT; ++a) z[a] = sm[S::oB3 + a];
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const float hk = fmaxf(acc[k], 0.0f);
        const float4 w = *reinterpret_cast<const float4*>(sm + S::oW3T + k * 8);
        const float w4 = sm[S::oW3T + k * 8 + 4];
        z[0] = fmaf(hk, w.x, z[0]); z[1] = fmaf(hk, w.y, z[1]); z[2] = fmaf(hk, w.z, z[2]);
        z[3] = fmaf(hk, w.w, z[3]); z[4] = fmaf(hk, w4, z[4]);
    }
}

struct RolloutArgs {
    const float* actor;
    double* env;            // [18][B]
    const float* noise;     // [T][N][A][B] or null
    uint64_t seed, episode;
    const uint64_t* episode_dev;   // optional device-resident episode counter (CUDA-graph replay): used instead of `episode`
    float* state;           // [T][54][B]
    float* obs;             // [T][N][O][B] or null
    int32_t* actions;       // [T][N][B]
    float* logp;            // [T][N][B]
    float* reward;          // [T][B]
    double* ep_return;      // [B] or null
    int T, B, O;
    int dbg;                // timeline experiments (CMARL_ROLLOUT_DBG, rollout_tc_kernel): 2 = no buffer stores of the observation
};

// Rollout kernel.  CTA = 32 envs = 12 warps; warp w = (agent n = w / 4, quarter qq = w % 4), lane = env.
// The four warps of an agent each own H/4 hidden units of both hidden layers, so every lane of a warp multiplies
// by the SAME weight: weights are warp-uniform LDS.128 broadcasts (one wavefront per four weights -- the
// 128 B/clk shared-memory path then sustains the full FFMA rate, which it cannot when every lane needs its own
// weights), and the 4-way split gives 4x more warps with 4x shorter dependent chains than one thread per
// (env, agent): B = 4096 is only 12 288 samples per step, a latency problem.  (Measured alternatives, all
// 0.18-0.21 ms: thread per sample 0.20, four lanes per sample with per-lane weights 0.18 (LSU-bound),
// constant-bank weights 0.19-0.21 (7.7 KB of weights thrash the uniform cache: ~20 cycles per FFMA).)
// Hidden activations cross the four warps through shared memory (conflict-free, lane = env).  Per agent the
// quarter-0 warp samples the action, the quarter-2 warp prepares the race noise meanwhile, and the quarter-1
// warps run the float64 physics lane-dense ((pair, env), then (agent, env)).
// The env state (18 doubles per env) lives in shared memory for the whole episode; T steps in ONE launch.
constexpr int REPB = 32;                 // envs per CTA
constexpr int NQ = 4;                    // warps per agent
constexpr int RTHREADS = REPB * NAG * NQ;   // 384 actor threads
// + one warp per agent pair that evaluates the float64 contact force (exp / log1p / sqrt, ~2 000 cycles) WHILE the actor
// warps run the network: the force depends only on the positions, not on the action, so it leaves the critical path
// of a step (obs -> layers -> sample -> integrate).  (The recurrent variant fits the 15 warps at 128 registers per
// thread with 48 bytes of spills: 0.563 -> 0.512 ms at 8 192 envs.)
constexpr int RPHYS = 3;
template <bool GRU> struct RolloutThreads { static constexpr int N = RTHREADS + 32 * RPHYS; };
constexpr int W1LD = 16;                 // layer-1 rows padded to 16 inputs (14 non-zero observation entries)

// debug timeline of CTA 0, step 10 (clock64): slots 0-7 warp (0,0) [sampler], 8-15 warp (0,1) [physics]
__device__ long long g_roll_tl[16 + 32 + 32 + 8 + 32];   // + 8: kernel-level stamps of CTA 0 (entry, predecessor complete, set-up done, steps done, exit)
#define KTL(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_roll_tl[80 + slot] = clock_here(); } while (0)
__device__ long long g_roll_tl_unused_;      // + 32: issue stamps of the MMAs, + 32: arrival of every warp at the two block barriers (rollout_tc_kernel)
// (a volatile asm with a memory clobber: the plain clock64() was hoisted across bar.sync by the compiler)
__device__ __forceinline__ long long clock_here() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory"); return c; }
#define RTL(slot, cond) do { if (blockIdx.x == 0 && t == 10 && e == 0 && (cond)) g_roll_tl[slot] = clock_here(); } while (0)

__device__ __forceinline__ void agent_bar(int n) { asm volatile("bar.sync %0, 128;" ::"r"(1 + n) : "memory"); }
// the per-step block barriers: the actor warps and the contact-force warps arrive from DIFFERENT instructions (warp
// specialisation), so the barrier names its thread count instead of being a __syncthreads()
template <int NT> __device__ __forceinline__ void block_bar() { asm volatile("bar.sync 5, %0;" ::"n"(NT) : "memory"); }   // ids 1-3: agent_bar, 4: physics warps

// reward of agent 0 (pettingzoo_wrapper.py:66) from the distance table: d[3 l + a] = |agent a - landmark l|,
// d[9], d[10] = |agent 1 - agent 0|, |agent 2 - agent 0|; same operation order as spread::reward_agent0
__device__ __forceinline__ double reward_from_table(const double* d) {
    double g = 0.0;
#pragma unroll
    for (int l = 0; l < 3; ++l) g = g - fmin(fmin(d[3 * l], d[3 * l + 1]), d[3 * l + 2]);
    double loc = 0.0;
    loc = loc - 1.0 * (d[9] < spread::DIST_MIN ? 1.0 : 0.0);
    loc = loc - 1.0 * (d[10] < spread::DIST_MIN ? 1.0 : 0.0);
    return g * (1 - spread::LOCAL_RATIO) + loc * spread::LOCAL_RATIO;
}

// GRU = true: the recurrent actor of mappo_lstm_multienvs.py:162-184 (fc1 + GRUCell + fc2); the hidden state of every
// (agent, env) lives in shared memory for the whole episode (zeros at t = 0, mappo_lstm_multienvs.py:406), double
// buffered because each of an agent's four warps reads all H old values and writes its own H/4 new ones.
template <int H, int O, bool GRU>
struct RolloutLayout {
    // parameter offsets in the flat actor vector (torch order)
    static constexpr int pW1 = 0, pB1 = pW1 + H * O;
    static constexpr int pW2 = pB1 + H, pB2 = pW2 + H * H;                                    // MLP: layer 2
    static constexpr int pWih = pB1 + H, pWhh = pWih + 3 * H * H, pBih = pWhh + 3 * H * H, pBhh = pBih + 3 * H;   // GRU
    static constexpr int pW3 = GRU ? pBhh + 3 * H : pB2 + H, pB3 = pW3 + NACT * H;            // output layer (fc2)
    // shared-memory image (float offsets)
    static constexpr int sW1 = 0, sB1 = sW1 + H * W1LD;
    static constexpr int sW2 = sB1 + NAG * H, sB2 = sW2 + H * H;                              // MLP
    static constexpr int sWih = sB1 + NAG * H, sWhh = sWih + 3 * H * H, sBg = sWhh + 3 * H * H;   // GRU: native [3H][H] x2, bias [H][4]
    static constexpr int sW3 = GRU ? sBg + 4 * H : sB2 + H, sB3 = sW3 + NACT * H, sEnd = sB3 + 8;
    static constexpr int nHS = GRU ? 2 * NAG * H * REPB : 0;                                  // hidden state, double buffered
    static constexpr int floats = sEnd + NAG * H * REPB + NAG * NQ * NACT * REPB + NAG * NACT * REPB + nHS;
};

template <int H, int O, bool GRU>
__global__ void __launch_bounds__(RolloutThreads<GRU>::N) rollout_kernel(RolloutArgs a) {
    constexpr int NTHR = RolloutThreads<GRU>::N;
    constexpr bool PHYSW = NTHR > RTHREADS;                   // dedicated contact-force warps present
    using RL = RolloutLayout<H, O, GRU>;
    constexpr int pW1 = RL::pW1, pB1 = RL::pB1, pW2 = RL::pW2, pB2 = RL::pB2, pW3 = RL::pW3, pB3 = RL::pB3;
    constexpr int sW1 = RL::sW1, sB1 = RL::sB1, sW2 = RL::sW2, sB2 = RL::sB2, sW3 = RL::sW3, sB3 = RL::sB3, sEnd = RL::sEnd;
    constexpr bool FOLD = O > CMARL_RAW_OBS;                   // one-hot agent ids appended to the observation
    constexpr int JL = H / NQ;                                // hidden units per warp
    extern __shared__ __align__(16) float dyn[];
    float* sw = dyn;                                          // W1[H][16] | b1 (+ id column) [3][H] | W2[H][H] | b2 | W3[5][H] | b3
    float (*hx)[H][REPB] = reinterpret_cast<float (*)[H][REPB]>(dyn + sEnd);                       // [NAG] layer-1 activations of an agent's four warps
    float (*zp)[NQ][NACT][REPB] = reinterpret_cast<float (*)[NQ][NACT][REPB]>(dyn + sEnd + NAG * H * REPB);   // [NAG] partial logits of the four warps
    float (*qs)[NACT][REPB] = reinterpret_cast<float (*)[NACT][REPB]>(dyn + sEnd + NAG * H * REPB + NAG * NQ * NACT * REPB);   // [NAG] race noise prepared by the quarter-2 warp
    float (*hs)[NAG][H][REPB] = reinterpret_cast<float (*)[NAG][H][REPB]>(dyn + sEnd + NAG * H * REPB + NAG * NQ * NACT * REPB + NAG * NACT * REPB);   // [2] GRU hidden state
    __shared__ double es[18][REPB];
    __shared__ int acts[NAG][REPB];
    __shared__ double pf[3][REPB][2];                         // contact force of pair (0,1), (0,2), (1,2) on its first entity
    __shared__ double rd[REPB][12];                           // distance table of the team reward
    const int tid = threadIdx.x;
    const int w = tid >> 5, e = tid & 31;                     // lane = env within the CTA
    const int n = w >> 2, qq = w & 3;                         // agent, quarter (warp-uniform)
    const int b = blockIdx.x * REPB + e;
    const bool live = b < a.B;
    const int B = a.B;
    const int j0 = qq * JL;

    pdl_wait_then_trigger();
    CMARL_STRIDED(i, 18 * REPB, NTHR) {
        const int r = i / REPB, c = i - r * REPB;
        const int bb = blockIdx.x * REPB + c;
        es[r][c] = (bb < B) ? a.env[(size_t)r * B + bb] : 0.0;
    }
    {
        const float* __restrict__ P = a.actor;
        CMARL_STRIDED(i, H * W1LD, NTHR) {
            const int j = i / W1LD, k = i - j * W1LD;
            sw[sW1 + i] = (k < CMARL_RAW_OBS - 4) ? __ldcg(P + pW1 + j * O + k) : 0.0f;
        }
        CMARL_STRIDED(i, NAG * H, NTHR) {
            const int g = i / H, j = i - g * H;
            sw[sB1 + i] = __ldcg(P + pB1 + j) + (FOLD ? __ldcg(P + pW1 + j * O + CMARL_RAW_OBS + g) : 0.0f);   // one-hot id column of agent g
        }
        if (GRU) {
            CMARL_STRIDED(i, 3 * H * H, NTHR) { sw[RL::sWih + i] = __ldcg(P + RL::pWih + i); sw[RL::sWhh + i] = __ldcg(P + RL::pWhh + i); }
            CMARL_STRIDED(j, H, NTHR) {     // bir + bhr, biz + bhz, bin, bhn
                sw[RL::sBg + 4 * j + 0] = __ldcg(P + RL::pBih + j) + __ldcg(P + RL::pBhh + j);
                sw[RL::sBg + 4 * j + 1] = __ldcg(P + RL::pBih + H + j) + __ldcg(P + RL::pBhh + H + j);
                sw[RL::sBg + 4 * j + 2] = __ldcg(P + RL::pBih + 2 * H + j);
                sw[RL::sBg + 4 * j + 3] = __ldcg(P + RL::pBhh + 2 * H + j);
            }
            for (int i = tid; i < 2 * NAG * H * REPB; i += NTHR) (&hs[0][0][0][0])[i] = 0.0f;   // h = None
        } else {
            CMARL_STRIDED(i, H * H, NTHR) sw[sW2 + i] = __ldcg(P + pW2 + i);
            CMARL_STRIDED(i, H, NTHR) sw[sB2 + i] = __ldcg(P + pB2 + i);
        }
        CMARL_STRIDED(i, NACT * H, NTHR) sw[sW3 + i] = __ldcg(P + pW3 + i);
        if (tid < 8) sw[sB3 + tid] = tid < NACT ? __ldcg(P + pB3 + tid) : 0.0f;
    }
    __syncthreads();
    double ep_acc = 0.0;                                     // threads 0..REPB-1: episode return of env tid
    const uint64_t episode = a.episode_dev ? *a.episode_dev : a.episode;

    if (PHYSW && w >= NAG * NQ) {
        // contact-force warps: pair p = (0,1), (0,2), (1,2) of every env, each evaluated ONCE per step from the
        // positions the integration of the previous step left in shared memory; two block barriers per step, like
        // the actor warps below
        const int p = w - NAG * NQ;
        const int ia = (p == 2) ? 1 : 0, ib = (p == 0) ? 1 : 2;
        for (int t = 0; t < a.T; ++t) {
            double gx, gy;
            spread::pair_force(es[2 * ia][e], es[2 * ia + 1][e], es[2 * ib][e], es[2 * ib + 1][e], gx, gy);
            pf[p][e][0] = gx; pf[p][e][1] = gy;
            block_bar<NTHR>();
            block_bar<NTHR>();
        }
    } else
    for (int t = 0; t < a.T; ++t) {
        // ---- observation before the action (what the reference stores, MME:426-430): vel, pos, landmarks - pos,
        //      other agents - pos, 4 zeros (spread::observe); dynamic indices go to shared memory ------------------
        RTL(0, w == 0); RTL(8, w == 1);
        const double opx = es[2 * n][e], opy = es[2 * n + 1][e], ovx = es[6 + 2 * n][e], ovy = es[6 + 2 * n + 1][e];
        const int oj0 = (n == 0) ? 1 : 0, oj1 = (n == 2) ? 1 : 2;          // the other two agents, index order
        float x[CMARL_RAW_OBS];
        x[0] = (float)ovx; x[1] = (float)ovy; x[2] = (float)opx; x[3] = (float)opy;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            x[4 + 2 * l] = (float)(es[12 + 2 * l][e] - opx);
            x[5 + 2 * l] = (float)(es[13 + 2 * l][e] - opy);
        }
        x[10] = (float)(es[2 * oj0][e] - opx); x[11] = (float)(es[2 * oj0 + 1][e] - opy);
        x[12] = (float)(es[2 * oj1][e] - opx); x[13] = (float)(es[2 * oj1 + 1][e] - opy);
        x[14] = 0.0f; x[15] = 0.0f; x[16] = 0.0f; x[17] = 0.0f;
        if (live) {   // the agent's four warps share the stores (row k by warp k % 4)
#pragma unroll
            for (int k = 0; k < CMARL_RAW_OBS; ++k)
                if ((k & 3) == qq) __stcs(a.state + ((size_t)t * 54 + n * CMARL_RAW_OBS + k) * B + b, x[k]);
            if (a.obs) {
                float* o = a.obs + ((size_t)t * NAG + n) * O * B + b;
#pragma unroll
                for (int k = 0; k < CMARL_RAW_OBS; ++k)
                    if ((k & 3) == qq) __stcs(o + (size_t)k * B, x[k]);
                if (FOLD && qq < NAG) __stcs(o + (size_t)(CMARL_RAW_OBS + qq) * B, qq == n ? 1.0f : 0.0f);
            }
        }
        RTL(1, w == 0);
        // ---- Actor.logits (MME:178-183): layer 1, this warp's JL hidden units; weights = constant-bank operands --
        {   // JL independent accumulator chains advance together (k outer, j inner)
            float acc[JL];
#pragma unroll
            for (int i = 0; i < JL; ++i) acc[i] = sw[sB1 + n * H + j0 + i];
#pragma unroll
            for (int k4 = 0; k4 < W1LD; k4 += 4) {                           // x[14..17] == 0; padded weights are 0
#pragma unroll
                for (int i = 0; i < JL; ++i) {
                    const float4 wv = *reinterpret_cast<const float4*>(sw + sW1 + (j0 + i) * W1LD + k4);   // warp-uniform address
                    acc[i] = fmaf(x[k4], wv.x, acc[i]); acc[i] = fmaf(x[k4 + 1], wv.y, acc[i]);
                    acc[i] = fmaf(x[k4 + 2], wv.z, acc[i]); acc[i] = fmaf(x[k4 + 3], wv.w, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < JL; ++i) hx[n][j0 + i][e] = fmaxf(acc[i], 0.0f);
        }
        if (qq == 2) {   // meanwhile: the race noise of this (t, agent, env)
            float q[NACT];
            if (a.noise) {
#pragma unroll
                for (int k = 0; k < NACT; ++k)
                    q[k] = live ? __ldcs(a.noise + (((size_t)t * NAG + n) * NACT + k) * B + b) : 1.0f;
            } else {
                philox_exp5(a.seed, episode, (uint32_t)t, (uint32_t)n, (uint32_t)b, q);
            }
#pragma unroll
            for (int k = 0; k < NACT; ++k) qs[n][k][e] = q[k];
        }
        RTL(2, w == 0);
        agent_bar(n);
        RTL(3, w == 0);
        // ---- layer 2 (all H inputs, this warp's JL outputs) and this warp's share of the output layer --------------
        float h1[H];
#pragma unroll
        for (int k = 0; k < H; ++k) h1[k] = hx[n][k][e];
        float z[NACT];
#pragma unroll
        for (int k = 0; k < NACT; ++k) z[k] = 0.0f;
        if (GRU) {
            // GRUCell (gate order r, z, n; ATen gru_cell): this warp's JL units; x1 = h1, old hidden state from hs[t & 1]
            const int cur = t & 1;
            float ar[JL], az[JL], ai[JL], ah[JL];
#pragma unroll
            for (int i = 0; i < JL; ++i) {
                const float4 bg = *reinterpret_cast<const float4*>(sw + RL::sBg + 4 * (j0 + i));
                ar[i] = bg.x; az[i] = bg.y; ai[i] = bg.z; ah[i] = bg.w;
            }
#pragma unroll
            for (int k4 = 0; k4 < H; k4 += 4) {
#pragma unroll
                for (int i = 0; i < JL; ++i) {
                    const float4 wr = *reinterpret_cast<const float4*>(sw + RL::sWih + (0 * H + j0 + i) * H + k4);   // warp-uniform
                    const float4 wz = *reinterpret_cast<const float4*>(sw + RL::sWih + (1 * H + j0 + i) * H + k4);
                    const float4 wn = *reinterpret_cast<const float4*>(sw + RL::sWih + (2 * H + j0 + i) * H + k4);
                    ar[i] = fmaf(h1[k4], wr.x, ar[i]); ar[i] = fmaf(h1[k4 + 1], wr.y, ar[i]);
                    ar[i] = fmaf(h1[k4 + 2], wr.z, ar[i]); ar[i] = fmaf(h1[k4 + 3], wr.w, ar[i]);
                    az[i] = fmaf(h1[k4], wz.x, az[i]); az[i] = fmaf(h1[k4 + 1], wz.y, az[i]);
                    az[i] = fmaf(h1[k4 + 2], wz.z, az[i]); az[i] = fmaf(h1[k4 + 3], wz.w, az[i]);
                    ai[i] = fmaf(h1[k4], wn.x, ai[i]); ai[i] = fmaf(h1[k4 + 1], wn.y, ai[i]);
                    ai[i] = fmaf(h1[k4 + 2], wn.z, ai[i]); ai[i] = fmaf(h1[k4 + 3], wn.w, ai[i]);
                }
            }
#pragma unroll
            for (int k4 = 0; k4 < H; k4 += 4) {
                const float p0 = hs[cur][n][k4][e], p1 = hs[cur][n][k4 + 1][e], p2 = hs[cur][n][k4 + 2][e], p3 = hs[cur][n][k4 + 3][e];
#pragma unroll
                for (int i = 0; i < JL; ++i) {
                    const float4 wr = *reinterpret_cast<const float4*>(sw + RL::sWhh + (0 * H + j0 + i) * H + k4);
                    const float4 wz = *reinterpret_cast<const float4*>(sw + RL::sWhh + (1 * H + j0 + i) * H + k4);
                    const float4 wn = *reinterpret_cast<const float4*>(sw + RL::sWhh + (2 * H + j0 + i) * H + k4);
                    ar[i] = fmaf(p0, wr.x, ar[i]); ar[i] = fmaf(p1, wr.y, ar[i]); ar[i] = fmaf(p2, wr.z, ar[i]); ar[i] = fmaf(p3, wr.w, ar[i]);
                    az[i] = fmaf(p0, wz.x, az[i]); az[i] = fmaf(p1, wz.y, az[i]); az[i] = fmaf(p2, wz.z, az[i]); az[i] = fmaf(p3, wz.w, az[i]);
                    ah[i] = fmaf(p0, wn.x, ah[i]); ah[i] = fmaf(p1, wn.y, ah[i]); ah[i] = fmaf(p2, wn.z, ah[i]); ah[i] = fmaf(p3, wn.w, ah[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < JL; ++i) {
                const float r = 1.0f / (1.0f + expf(-ar[i])), zg = 1.0f / (1.0f + expf(-az[i]));
                const float nn = tanhf(ai[i] + r * ah[i]);
                const float hn = (hs[cur][n][j0 + i][e] - nn) * zg + nn;
                hs[cur ^ 1][n][j0 + i][e] = hn;
                const float h2 = fmaxf(hn, 0.0f);
#pragma unroll
                for (int k = 0; k < NACT; ++k) z[k] = fmaf(h2, sw[sW3 + k * H + j0 + i], z[k]);
            }
        } else {
            float acc[JL];
#pragma unroll
            for (int i = 0; i < JL; ++i) acc[i] = sw[sB2 + j0 + i];
#pragma unroll
            for (int k4 = 0; k4 < H; k4 += 4) {
#pragma unroll
                for (int i = 0; i < JL; ++i) {
                    const float4 wv = *reinterpret_cast<const float4*>(sw + sW2 + (j0 + i) * H + k4);      // warp-uniform address
                    acc[i] = fmaf(h1[k4], wv.x, acc[i]); acc[i] = fmaf(h1[k4 + 1], wv.y, acc[i]);
                    acc[i] = fmaf(h1[k4 + 2], wv.z, acc[i]); acc[i] = fmaf(h1[k4 + 3], wv.w, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < JL; ++i) {
                const float h2 = fmaxf(acc[i], 0.0f);
#pragma unroll
                for (int k = 0; k < NACT; ++k) z[k] = fmaf(h2, sw[sW3 + k * H + j0 + i], z[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < NACT; ++k) zp[n][qq][k][e] = z[k];
        RTL(4, w == 0);
        agent_bar(n);
        RTL(5, w == 0);
        int action = 0;
        if (qq == 0) {
            // ---- Categorical sample: quarter sums in fixed order, then the exponential race ------------------------
            float q[NACT];
#pragma unroll
            for (int k = 0; k < NACT; ++k) {
                z[k] = ((zp[n][0][k][e] + zp[n][1][k][e]) + (zp[n][2][k][e] + zp[n][3][k][e])) + sw[sB3 + k];
                q[k] = qs[n][k][e];
            }
            float lp;
            race_sample(z, q, action, lp);
            acts[n][e] = action;
            if (live) {
                __stcs(a.actions + ((size_t)t * NAG + n) * B + b, action);
                __stcs(a.logp + ((size_t)t * NAG + n) * B + b, lp);
            }
        }
        RTL(6, w == 0); RTL(9, w == 1);
        block_bar<NTHR>();
        RTL(10, w == 1);
        // ---- physics (World.step), lane-dense float64 on the quarter-1 warps ----------------------------------------
        // (a) last step's team reward from the distance table (written after the previous state update)
        if (t > 0 && tid < REPB) {
            const double r = reward_from_table(rd[tid]);
            ep_acc += r;
            if (live) __stcs(a.reward + (size_t)(t - 1) * B + b, (float)r);
        }
        if (qq == 1) {
            if (!PHYSW) {
                // (b) the 3 contact pairs of every env, each evaluated ONCE: warp (n, 1) takes pair n: (0,1), (0,2), (1,2)
                const int ia = (n == 2) ? 1 : 0, ib = (n == 0) ? 1 : 2;
                double gx, gy;
                spread::pair_force(es[2 * ia][e], es[2 * ia + 1][e], es[2 * ib][e], es[2 * ib + 1][e], gx, gy);
                pf[n][e][0] = gx; pf[n][e][1] = gy;
                RTL(11, w == 1);
                asm volatile("bar.sync 4, 96;" ::: "memory");               // the three physics warps
            }
            RTL(12, w == 1);
            // (c) integration of agent n; forces added in the reference's pair order (0,1),(0,2),(1,2)
            const int act = acts[n][e];
            double ux = 0.0, uy = 0.0;
            if (act == 1) ux = -1.0;
            if (act == 2) ux = +1.0;
            if (act == 3) uy = -1.0;
            if (act == 4) uy = +1.0;
            double fx = ux * spread::SENSITIVITY + 0.0;
            double fy = uy * spread::SENSITIVITY + 0.0;
            const int p1 = (n == 2) ? 1 : 0, p2 = (n == 0) ? 1 : 2;     // the agent's first / second pair
            const bool neg1 = (n != 0), neg2 = (n == 2);                // agent is entity b of that pair
            const double g1x = pf[p1][e][0], g1y = pf[p1][e][1], g2x = pf[p2][e][0], g2y = pf[p2][e][1];
            fx = (neg1 ? -g1x : g1x) + fx; fy = (neg1 ? -g1y : g1y) + fy;
            fx = (neg2 ? -g2x : g2x) + fx; fy = (neg2 ? -g2y : g2y) + fy;
            double px = opx, py = opy, vx = ovx, vy = ovy;
            spread::integrate(px, py, vx, vy, fx, fy);
            // every reader of the old state in this phase is a physics warp and passed the bar.sync above
            es[2 * n][e] = px; es[2 * n + 1][e] = py;
            es[6 + 2 * n][e] = vx; es[6 + 2 * n + 1][e] = vy;
            RTL(13, w == 1);
        }
        block_bar<NTHR>();
        RTL(14, w == 1);
        // (d) distance table of the new state for the team reward: (task, env); tasks 0-8: agent a to landmark l
        //     (task = 3 l + a), 9-10: agents 1, 2 to agent 0; consumed after the next barrier
        if (w < 11) {
            const int task = w;
            int ea, eb;                                                // rows of es holding the two points
            if (task < 9) { ea = 2 * (task % 3); eb = 12 + 2 * (task / 3); }
            else { ea = 2 * (task - 8); eb = 0; }
            rd[e][task] = spread::dist2d(es[ea][e], es[ea + 1][e], es[eb][e], es[eb + 1][e]);
        }
        RTL(15, w == 1); RTL(7, w == 0);
    }
    __syncthreads();
    if (tid < REPB) {
        const double r = reward_from_table(rd[tid]);
        ep_acc += r;
        if (live) {
            __stcs(a.reward + (size_t)(a.T - 1) * B + b, (float)r);
            if (a.ep_return) a.ep_return[b] = ep_acc;
        }
    }
    for (int i = tid; i < 12 * REPB; i += NTHR) {
        const int r = i / REPB, c = i - r * REPB;
        const int bb = blockIdx.x * REPB + c;
        if (bb < B) a.env[(size_t)r * B + bb] = es[r][c];
    }
}

// ------------------------------------------------------------------------------------------------
// Rollout with layer 2 of the actor on the tensor cores (MLP actor, H = 32 -- the reference's default -- or 64).
//   The 96 samples of a CTA's step (3 agents x 32 envs) are the rows of ONE M = 128 tile: layer 1 stays on the CUDA cores
//   (K = 14: four warps per agent, H / 4 hidden units each, FFMA2) and writes relu(h1), split into tf32 hi / lo, straight
//   into the K-major A images in shared memory (row 32 n + e; one 16-byte chunk per four units: conflict-free STS.128); a
//   dedicated issue warp multiplies by the hi / lo images of W2 -- pass 0: A_hi x [W2_hi ; W2_lo] (one N = 2H operand),
//   pass 1: A_lo x W2_hi on the first H columns: 2 H / 8 tcgen05.mma of 128 x (2H | H) x 8 -- into 2H TMEM columns.
//   The epilogue (b2, relu, output layer) is bound to the TMEM lane quadrants: warp w reads quadrant w % 4, so the rows of
//   agent a are finished by the warps (n', qq = a) -- one of each agent's four -- in column groups of 12 / 12 / 8 (H = 64:
//   24 / 24 / 16).  Warp (0, a) decides agent a's race in the log domain (argmax_a z_a - log q_a) and publishes the action;
//   warp (1, a) repeats the decision and, behind the block barrier, evaluates the log-probability; warp (2, a) draws the
//   first Philox block of the next step's noise under the MMAs; the warps (n, 3) = warps 12-14, which own no TMEM quadrant
//   with rows in it, integrate agent n.  Warps 3, 7, 11 are the contact-force warps: pair forces, the distance table of the
//   team reward for the state the previous step left, the second Philox block, and (warp 3) the reward itself -- on the
//   issue warp's scheduler, off the critical path of a step.  Two 480-thread block barriers per step (behind the sampling,
//   behind the integration), met by these 15 warps at the same two instructions; the issue warp meets them through the two
//   mbarriers only.  DESIGN.md 3 has the measurements that shaped this; profiles/rollout_timeline_r2.txt the timelines.
// ------------------------------------------------------------------------------------------------
// log of the Exp(1) race noise of (t, agent n, env b), supplied by the caller or drawn with Philox: values 0..3 (first Philox
// block) and value 4 (second block)
__device__ __forceinline__ void log_noise_block0(const RolloutArgs& a, uint64_t episode, int t, int n, int b, bool live, float (&lq)[4]) {
    if (a.noise) {
#pragma unroll
        for (int k = 0; k < 4; ++k) lq[k] = live ? __logf(__ldcs(a.noise + (((size_t)t * NAG + n) * NACT + k) * a.B + b)) : 0.0f;
    } else {
        sample::philox_logexp_block0(a.seed, episode, (uint32_t)t, (uint32_t)n, (uint32_t)b, lq);
    }
}
__device__ __forceinline__ float log_noise_block1(const RolloutArgs& a, uint64_t episode, int t, int n, int b, bool live) {
    if (a.noise) return live ? __logf(__ldcs(a.noise + (((size_t)t * NAG + n) * NACT + 4) * a.B + b)) : 0.0f;
    return sample::philox_logexp_block1(a.seed, episode, (uint32_t)t, (uint32_t)n, (uint32_t)b);
}

namespace tcroll {
constexpr int NCG = 3;                                   // column groups of the epilogue (one per warp reading a TMEM quadrant)
constexpr int NTHR = RTHREADS + 32 * RPHYS + 32;         // 12 actor warps, 3 contact-force warps, the issue warp
constexpr int NBAR = RTHREADS + 32 * RPHYS;              // the per-step block barriers leave the issue warp out (it meets the others through mbarriers only)
constexpr int LBO = 128;                                 // next chunk of 4 k
template <int H_>
struct L {
    static constexpr int H = H_;
    static constexpr int NC0 = H == 64 ? 24 : 12;        // columns per epilogue group: 24 / 24 / 16 (H = 64), 12 / 12 / 8 (H = 32)
    static constexpr int A_BYTES = 128 * H * 4;          // one A image (hi or lo): 128 rows x H k, K-major core matrices
    static constexpr int B_BYTES = H * H * 4;            // one W2 image
    static constexpr int SBO = (H / 4) * 128;            // next group of 8 rows
    // shared memory (bytes)
    static constexpr int oBar = 0;                       // full, done mbarriers + the TMEM base
    static constexpr int oAh = 128, oAl = oAh + A_BYTES;
    static constexpr int oBh = oAl + A_BYTES, oBl = oBh + B_BYTES;
    static constexpr int oW1 = oBl + B_BYTES;            // f32 [16][H]: input-major, so that one LDS.128 holds four units' weights of one input
    static constexpr int oB1 = oW1 + H * W1LD * 4;       // f32 [3][H] b1 (+ the agent's folded id column)
    static constexpr int oB2 = oB1 + NAG * H * 4;        // f32 [H]
    static constexpr int oW3 = oB2 + H * 4;              // f32 [H][8] (5 used)
    static constexpr int oB3 = oW3 + H * 8 * 4;          // f32 [8]
    static constexpr int oZp = oB3 + 32;                 // f32 [NAG][NCG][NACT][32] partial logits
    static constexpr int oQs = oZp + NAG * NCG * NACT * REPB * 4;   // f32 [2][NAG][NACT][32] log race noise, by step parity
    static constexpr int smem_bytes = oQs + 2 * NAG * NACT * REPB * 4;
};
template <int NT> __device__ __forceinline__ void bar_named(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }
// two IEEE fp32 FMAs in one instruction (FFMA2): the kernel is bound by instruction issue, not by the FMA pipe; each half
// rounds like fmaf, so results are bit-identical to the scalar form
__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ra) : "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
// b2, relu and the output layer on NC accumulator columns starting at column j0
// (v: hi hi + lo hi products, u: hi lo products)
template <int NC>
__device__ __forceinline__ void headn(const uint32_t (&v)[NC], const uint32_t (&u)[NC], int j0, const float* sB2f, const float* sW3f,
                                      float (&z)[NACT]) {
    float2 z01 = make_float2(z[0], z[1]), z23 = make_float2(z[2], z[3]);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int j = j0 + i;
        const float h2 = fmaxf((__uint_as_float(v[i]) + __uint_as_float(u[i])) + sB2f[j], 0.0f);
        const float4 wv = *reinterpret_cast<const float4*>(sW3f + j * 8);
        z01 = ffma2(h2, make_float2(wv.x, wv.y), z01);
        z23 = ffma2(h2, make_float2(wv.z, wv.w), z23);
        z[4] = fmaf(h2, sW3f[j * 8 + 4], z[4]);
    }
    z[0] = z01.x; z[1] = z01.y; z[2] = z23.x; z[3] = z23.y;
}
}  // namespace tcroll

template <int H, int O>
__global__ void __launch_bounds__(tcroll::NTHR) rollout_tc_kernel(RolloutArgs a) {
    using namespace tcroll;
    using C = L<H>;
    constexpr int A_BYTES = C::A_BYTES, SBO = C::SBO, NC0 = C::NC0;
    constexpr int oBar = C::oBar, oAh = C::oAh, oAl = C::oAl, oBh = C::oBh, oBl = C::oBl, oW1 = C::oW1, oB1 = C::oB1, oB2 = C::oB2,
                  oW3 = C::oW3, oB3 = C::oB3, oZp = C::oZp, oQs = C::oQs;
    using RL = RolloutLayout<H, O, false>;
    constexpr bool FOLD = O > CMARL_RAW_OBS;
    constexpr int JL = H / NQ;
    extern __shared__ __align__(128) uint8_t smb[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smb + oBar);            // [0] full (384 arrivals), [1] done (commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smb + oBar + 16);
    const float* sW1f = reinterpret_cast<const float*>(smb + oW1);
    const float* sB1f = reinterpret_cast<const float*>(smb + oB1);
    const float* sB2f = reinterpret_cast<const float*>(smb + oB2);
    const float* sW3f = reinterpret_cast<const float*>(smb + oW3);
    const float* sB3f = reinterpret_cast<const float*>(smb + oB3);
    float (*zp)[NCG][NACT][REPB] = reinterpret_cast<float (*)[NCG][NACT][REPB]>(smb + oZp);
    float (*qs)[NAG][NACT][REPB] = reinterpret_cast<float (*)[NAG][NACT][REPB]>(smb + oQs);    // [step parity] log race noise
    __shared__ double es[18][REPB];
    __shared__ int acts[NAG][REPB];
    __shared__ double pf[3][REPB][2];
    __shared__ double rd[REPB][12];
    const int tid = threadIdx.x;
    const int w = tid >> 5, e = tid & 31;
    // roles: warps 3, 7, 11 are the contact-force warps (pair p = w / 4), warps 12-14 the actor warps (n = w - 12, qq = 3):
    // the float64 physics then shares its scheduler (warp id % 4 == 3) only with the issue warp, and the four schedulers
    // carry about the same number of instructions per step
    const bool physw = w < NAG * NQ && (w & 3) == 3;
    const int n = w < NAG * NQ ? w >> 2 : w - NAG * NQ, qq = w < NAG * NQ ? w & 3 : 3;
    const int b = blockIdx.x * REPB + e;
    const bool live = b < a.B;
    const int B = a.B;
    const int j0 = qq * JL;

    KTL(0);
    // ---- set-up: what touches no global memory first (runs under the tail of the launch in front) ----------------------
    for (int i = tid * 16; i < 2 * A_BYTES; i += NTHR * 16) *reinterpret_cast<uint4*>(smb + oAh + i) = make_uint4(0, 0, 0, 0);   // rows 96..127 stay zero
    if (tid == 0) {
        tc::mbar_init(&bars[0], RTHREADS);
        tc::mbar_init(&bars[1], 1);
        tc::fence_mbar_init();
    }
    if (w == 15) tc::tmem_alloc(tmem_slot, 2 * H);       // D = A_hi W2_hi^T + A_lo W2_hi^T | A_hi W2_lo^T
    KTL(1);
    pdl_wait_then_trigger();
    KTL(2);
    const uint64_t episode = a.episode_dev ? *a.episode_dev : a.episode;
    CMARL_STRIDED(i, 18 * REPB, NTHR) {
        const int r = i / REPB, c = i - r * REPB;
        const int bb = blockIdx.x * REPB + c;
        es[r][c] = (bb < B) ? a.env[(size_t)r * B + bb] : 0.0;
    }
    {
        const float* __restrict__ P = a.actor;
        float* fw = reinterpret_cast<float*>(smb);
        CMARL_STRIDED(i, H * W1LD, NTHR) {
            const int k = i / H, j = i - k * H;
            fw[oW1 / 4 + i] = (k < CMARL_RAW_OBS - 4) ? __ldcg(P + RL::pW1 + j * O + k) : 0.0f;
        }
        CMARL_STRIDED(i, NAG * H, NTHR) {
            const int g = i / H, j = i - g * H;
            fw[oB1 / 4 + i] = __ldcg(P + RL::pB1 + j) + (FOLD ? __ldcg(P + RL::pW1 + j * O + CMARL_RAW_OBS + g) : 0.0f);
        }
        {   // W2 [out j][in k] -> B images (N = j, K = k), every load of a thread in flight before the first split
            constexpr int NR = H * H / NTHR;
            static_assert(NR * NTHR == H * H, "W2 elements per thread");
            float wv[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) wv[r] = __ldcg(P + RL::pW2 + tid + r * NTHR);
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int i = tid + r * NTHR, j = i / H, k = i - j * H;
                float hi, lo;
                tc::split_tf32(wv[r], hi, lo);
                const int o = tctile::kmaj(j, k, H);
                *reinterpret_cast<float*>(smb + oBh + o) = hi;
                *reinterpret_cast<float*>(smb + oBl + o) = lo;
            }
        }
        CMARL_STRIDED(i, H, NTHR) fw[oB2 / 4 + i] = __ldcg(P + RL::pB2 + i);
        CMARL_STRIDED(i, H * 8, NTHR) {
            const int j = i >> 3, k = i & 7;
            fw[oW3 / 4 + i] = k < NACT ? __ldcg(P + RL::pW3 + k * H + j) : 0.0f;
        }
        if (tid < 8) fw[oB3 / 4 + tid] = tid < NACT ? __ldcg(P + RL::pB3 + tid) : 0.0f;
    }
    if (physw) {     // the race noise of step 0 (step t + 1 is drawn during step t, below)
        float q[4];
        log_noise_block0(a, episode, 0, w >> 2, b, live, q);
#pragma unroll
        for (int k = 0; k < 4; ++k) qs[0][w >> 2][k][e] = q[k];
        qs[0][w >> 2][4][e] = log_noise_block1(a, episode, 0, w >> 2, b, live);
    }
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    KTL(3);
    const uint32_t tmem = *tmem_slot;
    double ep_acc = 0.0;                                     // warp 3: episode return of env e

    // The issue warp runs its own loop (it meets the others through the two mbarriers only); the contact-force and the actor
    // warps share ONE step loop, so that they meet the step's two block barriers at the same instructions
    // (compute-sanitizer's synccheck reports warps arriving at a barrier from different instructions as divergent; with the
    // issue warp inside that loop too, the compiler no longer kept the MMA descriptors in uniform registers: 116 instead of
    // 50 cycles per MMA).
    if (w == 15) {
        // ================================ MMA issue warp ==================================================
        const uint32_t sbase = tc::smem_u32(smb);
        // pass 0: A_hi x [W2_hi ; W2_lo] (the lo image directly follows the hi image: ONE N = 2H operand, A read once) ->
        // columns 0..H-1 = hi hi, H..2H-1 = hi lo; pass 1: A_lo x W2_hi added to columns 0..H-1.  2 x H/8 MMAs instead of
        // 3 x H/8: an MMA of this size costs ~50 cycles whatever its N (measured: issue stamps of the timeline tool)
        constexpr uint32_t idesc2 = tc::make_idesc_tf32(128, 2 * H, 0, 0), idesc1 = tc::make_idesc_tf32(128, H, 0, 0);
        const bool leader = tc::elect_one();
        for (int t = 0; t < a.T; ++t) {
            tctile::acquire(&bars[0], (uint32_t)(t & 1));
            RTL(11, true);
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                uint64_t da = tc::make_smem_desc(sbase + (pass == 0 ? oAh : oAl), LBO, SBO, 0);
                uint64_t db = tc::make_smem_desc(sbase + oBh, LBO, SBO, 0);
                const uint32_t idesc = pass == 0 ? idesc2 : idesc1;
#pragma unroll 2
                for (int ks = 0; ks < H / 8; ++ks) {
                    if (leader) tc::mma_tf32(tmem, da, db, idesc, (uint32_t)(pass | ks));
                    if (blockIdx.x == 0 && t == 10 && e == 0) g_roll_tl[16 + pass * (H / 8) + ks] = clock_here();
                    da += (uint64_t)((2 * LBO) >> 4);
                    db += (uint64_t)((2 * LBO) >> 4);
                }
            }
            if (leader) tc::mma_commit(&bars[1]);
            RTL(15, true);
            __syncwarp();
        }
    } else {
    const int p = w >> 2;                                    // contact-force warp: pair p = (0,1), (0,2), (1,2) of every env
    const int ia = (p == 2) ? 1 : 0, ib = (p == 0) ? 1 : 2;
    const bool logpw = w < NAG * NQ && n == 1 && qq < 3;     // warp (1, a): log-probability of agent a's action, off the critical path
    for (int t = 0; t < a.T; ++t) {
        double opx = 0.0, opy = 0.0, ovx = 0.0, ovy = 0.0;   // actor warps: this agent's state before the step
        float zl[NACT] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, zsel = 0.0f;   // warps (0, a), (1, a): the logits of this lane's sample
        int action = 0;
        if (physw) {
            // ================================ contact-force warps =============================================
            // pair p of every env from the positions the previous step left, and (t > 0) that state's distance table for
            // the team reward: tasks 0-8 agent a to landmark l (task = 3 l + a), 9-10 agents 1, 2 to agent 0
            if (t > 0) {
#pragma unroll
                for (int task = p; task < 11; task += 3) {
                    int ea, eb;
                    if (task < 9) { ea = 2 * (task % 3); eb = 12 + 2 * (task / 3); }
                    else { ea = 2 * (task - 8); eb = 0; }
                    rd[e][task] = spread::dist2d(es[ea][e], es[ea + 1][e], es[eb][e], es[eb + 1][e]);
                }
            }
            double gx, gy;
            spread::pair_force(es[2 * ia][e], es[2 * ia + 1][e], es[2 * ib][e], es[2 * ib + 1][e], gx, gy);
            pf[p][e][0] = gx; pf[p][e][1] = gy;
            // the second Philox block of agent p's race noise for the NEXT step (see the actor warps)
            if (t + 1 < a.T) qs[(t + 1) & 1][p][4][e] = log_noise_block1(a, episode, t + 1, p, b, live);
        } else {
            // ================================ actor warps =====================================================
            // ---- observation before the action (MME:426-430), as in rollout_kernel ---------------------------------------
            RTL(0, w == 0); RTL(8, w == 1);
            opx = es[2 * n][e]; opy = es[2 * n + 1][e]; ovx = es[6 + 2 * n][e]; ovy = es[6 + 2 * n + 1][e];
            const int oj0 = (n == 0) ? 1 : 0, oj1 = (n == 2) ? 1 : 2;
            float x[CMARL_RAW_OBS];
            x[0] = (float)ovx; x[1] = (float)ovy; x[2] = (float)opx; x[3] = (float)opy;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                x[4 + 2 * l] = (float)(es[12 + 2 * l][e] - opx);
                x[5 + 2 * l] = (float)(es[13 + 2 * l][e] - opy);
            }
            x[10] = (float)(es[2 * oj0][e] - opx); x[11] = (float)(es[2 * oj0 + 1][e] - opy);
            x[12] = (float)(es[2 * oj1][e] - opx); x[13] = (float)(es[2 * oj1 + 1][e] - opy);
            x[14] = 0.0f; x[15] = 0.0f; x[16] = 0.0f; x[17] = 0.0f;
            // ---- layer 1 (this warp's JL units) -> relu -> tf32 hi / lo -> A images, row 32 n + e --------------------------
            {
                // pairs of units advance together (FFMA2); per unit the inputs are added in ascending order, as before
                float2 acc2[JL / 2];
#pragma unroll
                for (int i = 0; i < JL / 2; ++i) acc2[i] = *reinterpret_cast<const float2*>(sB1f + n * H + j0 + 2 * i);
#pragma unroll
                for (int k = 0; k < CMARL_RAW_OBS - 4; ++k) {                       // x[14..17] == 0
#pragma unroll
                    for (int c = 0; c < JL / 4; ++c) {
                        const float4 wv = *reinterpret_cast<const float4*>(sW1f + k * H + j0 + 4 * c);   // warp-uniform address
                        acc2[2 * c] = ffma2(x[k], make_float2(wv.x, wv.y), acc2[2 * c]);
                        acc2[2 * c + 1] = ffma2(x[k], make_float2(wv.z, wv.w), acc2[2 * c + 1]);
                    }
                }
                float acc[JL];
#pragma unroll
                for (int i = 0; i < JL / 2; ++i) { acc[2 * i] = acc2[i].x; acc[2 * i + 1] = acc2[i].y; }
                const int r = 32 * n + e;
                uint8_t* row = smb + (r >> 3) * SBO + (r & 7) * 16 + (j0 >> 2) * LBO;
#pragma unroll
                for (int c = 0; c < JL / 4; ++c) {
                    float4 hi, lo;
                    tc::split_tf32(fmaxf(acc[4 * c], 0.0f), hi.x, lo.x);
                    tc::split_tf32(fmaxf(acc[4 * c + 1], 0.0f), hi.y, lo.y);
                    tc::split_tf32(fmaxf(acc[4 * c + 2], 0.0f), hi.z, lo.z);
                    tc::split_tf32(fmaxf(acc[4 * c + 3], 0.0f), hi.w, lo.w);
                    *reinterpret_cast<float4*>(row + oAh + c * LBO) = hi;
                    *reinterpret_cast<float4*>(row + oAl + c * LBO) = lo;
                }
            }
            tctile::publish(&bars[0]);
            RTL(2, w == 0);
            // ---- under the MMAs: the buffer stores of the observation (row k by warp k % 4) ------------------------------------
            if (live && !(a.dbg & 2)) {
#pragma unroll
                for (int k = 0; k < CMARL_RAW_OBS; ++k)
                    if ((k & 3) == qq) __stcs(a.state + ((size_t)t * 54 + n * CMARL_RAW_OBS + k) * B + b, x[k]);
                if (a.obs) {
                    float* o = a.obs + ((size_t)t * NAG + n) * O * B + b;
#pragma unroll
                    for (int k = 0; k < CMARL_RAW_OBS; ++k)
                        if ((k & 3) == qq) __stcs(o + (size_t)k * B, x[k]);
                    if (FOLD && qq < NAG) __stcs(o + (size_t)(CMARL_RAW_OBS + qq) * B, qq == n ? 1.0f : 0.0f);
                }
            }
            RTL(1, w == 0);
            // the log race noise of the NEXT step goes to the other half of the double-buffered table: warp (2, a) draws the first
            // Philox block of agent a (four values) while the MMAs run -- at H = 32 its share of the epilogue is the shortest --,
            // the contact-force warp a the second block behind its physics.  (Measured alternatives, timeline tool: on the
            // warps (n, 3) under the MMAs the draws starved the issue warp on the same scheduler -- 1 450 instead of 680 cycles
            // for 12 MMAs --; all of it on the contact-force warps or on warp (2, a) made that warp the last at the barrier.)
            if (n == 2 && qq < 3 && t + 1 < a.T) {
                float lqn[4];
                log_noise_block0(a, episode, t + 1, qq, b, live, lqn);
#pragma unroll
                for (int k = 0; k < 4; ++k) qs[(t + 1) & 1][qq][k][e] = lqn[k];
            }
            if (qq < 3) {
                // ---- epilogue of agent qq's rows (TMEM quadrant qq), column group n: b2, relu, output layer -----------------
                const int c0 = n * NC0;                          // warp-uniform
                float z[NACT];
#pragma unroll
                for (int k = 0; k < NACT; ++k) z[k] = 0.0f;
                tctile::acquire(&bars[1], (uint32_t)(t & 1));
                RTL(3, w == 0);
                const uint32_t ta = tmem + ((uint32_t)(32 * qq) << 16) + (uint32_t)c0;
                uint32_t v0[8], u0[8];
                tc::tmem_ld8(ta, v0); tc::tmem_ld8(ta + H, u0);
                if constexpr (H == 64) {       // 24 / 24 / 16 columns
                    uint32_t v1[8], u1[8], v2[8], u2[8];
                    tc::tmem_ld8(ta + 8, v1); tc::tmem_ld8(ta + H + 8, u1);
                    if (n < 2) { tc::tmem_ld8(ta + 16, v2); tc::tmem_ld8(ta + H + 16, u2); }
                    tc::tmem_wait_ld();
                    headn<8>(v0, u0, c0, sB2f, sW3f, z); headn<8>(v1, u1, c0 + 8, sB2f, sW3f, z);
                    if (n < 2) headn<8>(v2, u2, c0 + 16, sB2f, sW3f, z);
                } else {                       // 12 / 12 / 8 columns
                    uint32_t v1[4], u1[4];
                    if (n < 2) { tc::tmem_ld4(ta + 8, v1); tc::tmem_ld4(ta + H + 8, u1); }
                    tc::tmem_wait_ld();
                    headn<8>(v0, u0, c0, sB2f, sW3f, z);
                    if (n < 2) headn<4>(v1, u1, c0 + 8, sB2f, sW3f, z);
                }
#pragma unroll
                for (int k = 0; k < NACT; ++k) zp[qq][n][k][e] = z[k];
                RTL(4, w == 0);
                bar_named<96>(1 + qq);
                RTL(5, w == 0);
                if (n < 2) {
                    // ---- Categorical sample of agent qq: column-group sums in fixed order, then the race in the log domain.
                    //      Warp (0, qq) publishes the action -- the only thing the physics waits for --; warp (1, qq) evaluates
                    //      the same decision from the same operands and, behind the barrier, while the physics runs, the
                    //      log-probability, and writes both to the buffer ---------------------------------------------------------
                    float lq[NACT];
#pragma unroll
                    for (int k = 0; k < NACT; ++k) {
                        zl[k] = ((zp[qq][0][k][e] + zp[qq][1][k][e]) + zp[qq][2][k][e]) + sB3f[k];
                        lq[k] = qs[t & 1][qq][k][e];
                    }
                    sample::race_action_log(zl, lq, action, zsel);
                    if (n == 0) acts[qq][e] = action;
                    RTL(6, w == 0);
                }
            }
        }
        RTL(9, w == 1);
        RTL(48 + w, true);
        bar_named<NBAR>(5);
        RTL(10, w == 12);
        if (physw) {
            // last step's team reward from the distance table (complete: every contact-force warp is behind the barrier);
            // this scheduler is idle while the actor warps integrate
            if (t > 0 && w == 3) {
                const double r = reward_from_table(rd[e]);
                ep_acc += r;
                if (live) __stcs(a.reward + (size_t)(t - 1) * B + b, (float)r);
            }
        } else {
            if (logpw) {
                const float lp = sample::race_logp_fast(zl, zsel);
                if (live) {
                    __stcs(a.actions + ((size_t)t * NAG + qq) * B + b, action);
                    __stcs(a.logp + ((size_t)t * NAG + qq) * B + b, lp);
                }
            }
            // ---- physics (World.step) -----------------------------------------------------------------------------------
            if (qq == 3) {
                RTL(12, w == 12);
                // (c) integration of agent n by warp (n, 3) -- it owns no TMEM quadrant with rows in it, so it is free when the
                //     action arrives; forces added in the reference's pair order (0,1),(0,2),(1,2)
                const int act = acts[n][e];
                double ux = 0.0, uy = 0.0;
                if (act == 1) ux = -1.0;
                if (act == 2) ux = +1.0;
                if (act == 3) uy = -1.0;
                if (act == 4) uy = +1.0;
                double fx = ux * spread::SENSITIVITY + 0.0;
                double fy = uy * spread::SENSITIVITY + 0.0;
                const int p1 = (n == 2) ? 1 : 0, p2 = (n == 0) ? 1 : 2;
                const bool neg1 = (n != 0), neg2 = (n == 2);
                const double g1x = pf[p1][e][0], g1y = pf[p1][e][1], g2x = pf[p2][e][0], g2y = pf[p2][e][1];
                fx = (neg1 ? -g1x : g1x) + fx; fy = (neg1 ? -g1y : g1y) + fy;
                fx = (neg2 ? -g2x : g2x) + fx; fy = (neg2 ? -g2y : g2y) + fy;
                double px = opx, py = opy, vx = ovx, vy = ovy;
                spread::integrate(px, py, vx, vy, fx, fy);
                // every reader of the old state (the other warps' observations, the pair forces) is behind the barrier above
                es[2 * n][e] = px; es[2 * n + 1][e] = py;
                es[6 + 2 * n][e] = vx; es[6 + 2 * n + 1][e] = vy;
                RTL(13, w == 12);
            }
        }
        RTL(64 + w, true);
        bar_named<NBAR>(6);
        RTL(14, w == 12); RTL(7, w == 0);
        if (blockIdx.x == 0 && tid == 0 && t < 32) g_roll_tl[88 + t] = clock_here();     // end of every step, warp 0
    }
    }
    if (physw) {   // the distance table of the final state
#pragma unroll
        for (int task = w >> 2; task < 11; task += 3) {
            int ea, eb;
            if (task < 9) { ea = 2 * (task % 3); eb = 12 + 2 * (task / 3); }
            else { ea = 2 * (task - 8); eb = 0; }
            rd[e][task] = spread::dist2d(es[ea][e], es[ea + 1][e], es[eb][e], es[eb + 1][e]);
        }
    }
    KTL(4);
    tc::tcgen05_fence_before();
    __syncthreads();
    if (w == 15) tc::tmem_dealloc(tmem, 2 * H);
    if (w == 3) {
        const double r = reward_from_table(rd[e]);
        ep_acc += r;
        if (live) {
            __stcs(a.reward + (size_t)(a.T - 1) * B + b, (float)r);
            if (a.ep_return) a.ep_return[b] = ep_acc;
        }
    }
    for (int i = tid; i < 12 * REPB; i += NTHR) {
        const int r = i / REPB, c = i - r * REPB;
        const int bb = blockIdx.x * REPB + c;
        if (bb < B) a.env[(size_t)r * B + bb] = es[r][c];
    }
    KTL(5);
}

// ---- K2 alone ------------------------------------------------------------------------------
struct ActArgs {
    const float* actor;
    const float* obs;         // [N][O][B]
    const uint8_t* avail;     // [N][A][B] or null
    const float* noise;       // [N][A][B]
    int32_t* actions;         // [N][B]
    float* logp;              // [N][B]
    float* logits;            // [N][A][B] or null
    int B, O;
};

template <int H>
__global__ void __launch_bounds__(RT) actor_act_kernel(ActArgs a) {
    extern __shared__ __align__(16) float smf[];
    using S = ActorSmem<H>;
    const int tid = threadIdx.x;
    const int n = tid / EPB, e = tid - n * EPB;
    const int b = blockIdx.x * EPB + e;
    const bool live = b < a.B;
    load_actor<H>(smf, a.actor, a.O, false, RT);
    __syncthreads();
    float x[21];
#pragma unroll
    for (int k = 0; k < 21; ++k) x[k] = (live && k < a.O) ? a.obs[((size_t)n * a.O + k) * a.B + b] : 0.0f;
    float z[NACT];
    actor_mlp<H, 21>(x, smf, smf + S::oB1 + 3 * H, smf + S::oAct + tid, RT, z);
    if (!live) return;
    float q[NACT];
#pragma unroll
    for (int k = 0; k < NACT; ++k) {
        if (a.avail && !a.avail[((size_t)n * NACT + k) * a.B + b]) z[k] = -1e9f;     // MME:182
        q[k] = a.noise[((size_t)n * NACT + k) * a.B + b];
        if (a.logits) a.logits[((size_t)n * NACT + k) * a.B + b] = z[k];
    }
    int action; float lp;
    race_sample(z, q, action, lp);
    a.actions[(size_t)n * a.B + b] = action;
    a.logp[(size_t)n * a.B + b] = lp;
}

// ---- K1 reset ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) env_reset_kernel(double* __restrict__ env, int B, uint64_t seed, uint64_t episode,
                                                        const uint64_t* __restrict__ episode_dev) {
    pdl_wait_then_trigger();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (episode_dev) episode = __ldcg(reinterpret_cast<const unsigned long long*>(episode_dev));   // written by episode_advance_kernel: coherent load
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    // 12 uniforms in reset_world order: agent positions (x,y) x3, then landmark positions x3
    double u[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const Philox4 r = philox4x32_10((uint32_t)b, 0xE0000000u + i, (uint32_t)episode, (uint32_t)(episode >> 32), k0, k1);
        u[2 * i] = u64_to_unit(r.x, r.y);
        u[2 * i + 1] = u64_to_unit(r.z, r.w);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        env[(size_t)i * B + b] = -1.0 + 2.0 * u[i];          // agent p_pos ~ U(-1, 1)
        env[(size_t)(6 + i) * B + b] = 0.0;                  // p_vel = 0
        env[(size_t)(12 + i) * B + b] = -1.0 + 2.0 * u[6 + i];   // landmark p_pos
    }
}

// ---- K1 alone: observe / single step, one thread per env -------------------------------------
__global__ void __launch_bounds__(128) env_step_kernel(double* __restrict__ env, const int32_t* __restrict__ actions,
                                                       float* __restrict__ state_out, float* __restrict__ reward_out,
                                                       int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double p[6], v[6], lm[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        p[i] = env[(size_t)i * B + b]; v[i] = env[(size_t)(6 + i) * B + b]; lm[i] = env[(size_t)(12 + i) * B + b];
    }
    if (actions) {
        double np_[6], nv[6];
#pragma unroll
        for (int n = 0; n < NAG; ++n) {
            double fx, fy;
            spread::agent_force(n, p, actions[(size_t)n * B + b], fx, fy);
            double px = p[2 * n], py = p[2 * n + 1], vx = v[2 * n], vy = v[2 * n + 1];
            spread::integrate(px, py, vx, vy, fx, fy);
            np_[2 * n] = px; np_[2 * n + 1] = py; nv[2 * n] = vx; nv[2 * n + 1] = vy;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            p[i] = np_[i]; v[i] = nv[i];
            env[(size_t)i * B + b] = p[i]; env[(size_t)(6 + i) * B + b] = v[i];
        }
        if (reward_out) reward_out[b] = (float)spread::reward_agent0(p, lm);
    }
    if (state_out) {
#pragma unroll
        for (int n = 0; n < NAG; ++n) {
            float x[CMARL_RAW_OBS];
            spread::observe(n, p, v, lm, x);
#pragma unroll
            for (int k = 0; k < CMARL_RAW_OBS; ++k) state_out[(size_t)(n * CMARL_RAW_OBS + k) * B + b] = x[k];
        }
    }
}

template <int H>
size_t actor_smem_bytes() { return (size_t)ActorSmem<H>::oEnd * sizeof(float); }

}  // namespace

extern "C" int cmarl_debug_rollout_timeline(long long* out_host16) {
    return (int)cudaMemcpyFromSymbol(out_host16, g_roll_tl, sizeof(long long) * 16);
}
extern "C" int cmarl_debug_rollout_timeline_mma(long long* out_host32) {
    return (int)cudaMemcpyFromSymbol(out_host32, g_roll_tl, sizeof(long long) * 104, sizeof(long long) * 16);
}

// generic.cu: the layered kernels behind the same entries when cmarl_ctx.generic is set
int cmarl_gen_env_reset(cmarl_ctx* ctx, double* env, uint64_t seed, uint64_t episode, cudaStream_t st);
int cmarl_gen_env_step(cmarl_ctx* ctx, double* env, const int32_t* actions, float* state_out, float* reward_out, cudaStream_t st);
int cmarl_gen_rollout(cmarl_ctx* ctx, const float* actor_params, double* env, const float* noise, uint64_t seed, uint64_t episode,
                      float* state, float* obs, int32_t* actions, float* logp, float* reward, double* ep_return, cudaStream_t st);
int cmarl_gen_actor_act(cmarl_ctx* ctx, const float* actor_params, const float* obs, const uint8_t* avail, const float* noise,
                        int32_t* actions, float* logp, float* logits_out, cudaStream_t st);

__global__ void episode_advance_kernel(uint64_t* e) { pdl_wait_then_trigger(); *e += 1; }

// shared-memory opt-ins, once per context (not inside the launch path: keeps cmarl_rollout CUDA-graph capturable)
int cmarl_rollout_setup(cmarl_ctx* ctx) {
    (void)ctx;
    const size_t smem64 = (size_t)(64 * W1LD + NAG * 64 + 64 * 64 + 64 + NACT * 64 + 8 + NAG * 64 * REPB + NAG * NQ * NACT * REPB +
                                   NAG * NACT * REPB) * sizeof(float);
    CMARL_CUDA(cudaFuncSetAttribute(rollout_kernel<64, 21, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
    CMARL_CUDA(cudaFuncSetAttribute(rollout_kernel<64, 18, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
    CMARL_CUDA(cudaFuncSetAttribute(rollout_kernel<32, 21, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(RolloutLayout<32, 21, true>::floats * sizeof(float))));
    CMARL_CUDA(cudaFuncSetAttribute(rollout_kernel<32, 18, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(RolloutLayout<32, 18, true>::floats * sizeof(float))));
    CMARL_CUDA(cudaFuncSetAttribute(actor_act_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)actor_smem_bytes<64>()));
    CMARL_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<64, 21>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcroll::L<64>::smem_bytes));
    CMARL_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<64, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcroll::L<64>::smem_bytes));
    CMARL_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<32, 21>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcroll::L<32>::smem_bytes));
    CMARL_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<32, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcroll::L<32>::smem_bytes));
    return 0;
}

extern "C" int cmarl_ctx_set_episode_counter(cmarl_ctx* ctx, uint64_t* episode_dev) {
    CMARL_ARG(ctx, "null ctx");
    ctx->episode_dev = episode_dev;
    return 0;
}

extern "C" int cmarl_episode_advance(cmarl_ctx* ctx, void* stream) {
    CMARL_ARG(ctx && ctx->episode_dev, "no device episode counter set (cmarl_ctx_set_episode_counter)");
    ctx->launches++;
    return cmarl_check_cuda(cmarl_launch(ctx, episode_advance_kernel, dim3(1), dim3(1), 0, as_stream(stream), ctx->episode_dev),
                            "episode_advance_kernel");
}

extern "C" int cmarl_env_reset(cmarl_ctx* ctx, double* env, uint64_t seed, uint64_t episode, void* stream) {
    CMARL_ARG(ctx && env, "null argument");
    if (ctx->generic) return cmarl_gen_env_reset(ctx, env, seed, episode, as_stream(stream));
    const int B = ctx->cfg.n_envs;
    {
        KernelTimer kt(ctx, K_RESET, as_stream(stream));
        CMARL_CUDA(cmarl_launch(ctx, env_reset_kernel, dim3(ceil_div(B, 256)), dim3(256), 0, as_stream(stream), env, B, seed, episode,
                                (const uint64_t*)ctx->episode_dev));
    }
    return 0;
}

extern "C" int cmarl_env_observe(cmarl_ctx* ctx, const double* env, float* state_out, void* stream) {
    CMARL_ARG(ctx && env && state_out, "null argument");
    if (ctx->generic) return cmarl_gen_env_step(ctx, const_cast<double*>(env), nullptr, state_out, nullptr, as_stream(stream));
    const int B = ctx->cfg.n_envs;
    {
        KernelTimer kt(ctx, K_ENVSTEP, as_stream(stream));
        env_step_kernel<<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(const_cast<double*>(env), nullptr, state_out,
                                                                        nullptr, B);
    }
    return cmarl_check_cuda(cudaGetLastError(), "env_step_kernel(observe)");
}

extern "C" int cmarl_env_step(cmarl_ctx* ctx, double* env, const int32_t* actions, float* state_out,
                              float* reward_out, void* stream) {
    CMARL_ARG(ctx && env && actions, "null argument");
    if (ctx->generic) return cmarl_gen_env_step(ctx, env, actions, state_out, reward_out, as_stream(stream));
    const int B = ctx->cfg.n_envs;
    {
        KernelTimer kt(ctx, K_ENVSTEP, as_stream(stream));
        env_step_kernel<<<ceil_div(B, 128), 128, 0, as_stream(stream)>>>(env, actions, state_out, reward_out, B);
    }
    return cmarl_check_cuda(cudaGetLastError(), "env_step_kernel");
}

extern "C" int cmarl_rollout(cmarl_ctx* ctx, const float* actor_params, double* env, const float* noise,
                             uint64_t seed, uint64_t episode, float* state, float* obs, int32_t* actions,
                             float* logp, float* reward, double* ep_return, void* stream) {
    CMARL_ARG(ctx && actor_params && env && state && actions && logp && reward, "null argument");
    if (ctx->generic)
        return cmarl_gen_rollout(ctx, actor_params, env, noise, seed, episode, state, obs, actions, logp, reward, ep_return, as_stream(stream));
    RolloutArgs a;
    a.actor = actor_params; a.env = env; a.noise = noise; a.seed = seed; a.episode = episode;
    a.episode_dev = ctx->episode_dev;
    a.state = state; a.obs = obs; a.actions = actions; a.logp = logp; a.reward = reward; a.ep_return = ep_return;
    a.T = ctx->cfg.n_steps; a.B = ctx->cfg.n_envs; a.O = ctx->cfg.obs_dim;
    { const char* v = getenv("CMARL_ROLLOUT_DBG"); a.dbg = v ? atoi(v) : 0; }
    const int grid = ceil_div(a.B, REPB);
    cudaStream_t st = as_stream(stream);
    const bool ids = a.O > CMARL_RAW_OBS;
    const int H = ctx->cfg.actor_hidden;
    if (ctx->cfg.actor_recurrent) {
        constexpr size_t smem21 = (size_t)RolloutLayout<32, 21, true>::floats * sizeof(float);
        constexpr size_t smem18 = (size_t)RolloutLayout<32, 18, true>::floats * sizeof(float);
        KernelTimer kt(ctx, K_ROLLOUT, st);
        return cmarl_check_cuda(ids ? cmarl_launch(ctx, rollout_kernel<32, 21, true>, dim3(grid), dim3(RolloutThreads<true>::N), smem21, st, a)
                                    : cmarl_launch(ctx, rollout_kernel<32, 18, true>, dim3(grid), dim3(RolloutThreads<true>::N), smem18, st, a),
                                "rollout_kernel (recurrent)");
    }
    const size_t smem = (size_t)(H * W1LD + NAG * H + H * H + H + NACT * H + 8 + NAG * H * REPB + NAG * NQ * NACT * REPB +
                                 NAG * NACT * REPB) * sizeof(float);
    const dim3 block(RolloutThreads<false>::N);
    KernelTimer kt(ctx, K_ROLLOUT, st);
    {
        // layer 2 on the tensor cores (default); CMARL_ROLLOUT=ffma selects the CUDA-core kernel (read at every call: tests
        // compare the two within one process)
        const char* v = getenv("CMARL_ROLLOUT");
        if (!(v && v[0] == 'f')) {
            const dim3 tb(tcroll::NTHR);
            cudaError_t err;
            if (H == 32) err = ids ? cmarl_launch(ctx, rollout_tc_kernel<32, 21>, dim3(grid), tb, (size_t)tcroll::L<32>::smem_bytes, st, a)
                                   : cmarl_launch(ctx, rollout_tc_kernel<32, 18>, dim3(grid), tb, (size_t)tcroll::L<32>::smem_bytes, st, a);
            else         err = ids ? cmarl_launch(ctx, rollout_tc_kernel<64, 21>, dim3(grid), tb, (size_t)tcroll::L<64>::smem_bytes, st, a)
                                   : cmarl_launch(ctx, rollout_tc_kernel<64, 18>, dim3(grid), tb, (size_t)tcroll::L<64>::smem_bytes, st, a);
            return cmarl_check_cuda(err, "rollout_tc_kernel");
        }
    }
    if (H == 32)
        return cmarl_check_cuda(ids ? cmarl_launch(ctx, rollout_kernel<32, 21, false>, dim3(grid), block, smem, st, a)
                                    : cmarl_launch(ctx, rollout_kernel<32, 18, false>, dim3(grid), block, smem, st, a), "rollout_kernel");
    return cmarl_check_cuda(ids ? cmarl_launch(ctx, rollout_kernel<64, 21, false>, dim3(grid), block, smem, st, a)
                                : cmarl_launch(ctx, rollout_kernel<64, 18, false>, dim3(grid), block, smem, st, a), "rollout_kernel");
}

extern "C" int cmarl_actor_act(cmarl_ctx* ctx, const float* actor_params, const float* obs, const uint8_t* avail,
                               const float* noise, int32_t* actions, float* logp, float* logits_out, void* stream) {
    CMARL_ARG(ctx && actor_params && obs && noise && actions && logp, "null argument");
    if (ctx->generic) return cmarl_gen_actor_act(ctx, actor_params, obs, avail, noise, actions, logp, logits_out, as_stream(stream));
    ActArgs a;
    a.actor = actor_params; a.obs = obs; a.avail = avail; a.noise = noise; a.actions = actions; a.logp = logp;
    a.logits = logits_out; a.B = ctx->cfg.n_envs; a.O = ctx->cfg.obs_dim;
    const int grid = ceil_div(a.B, EPB);
    cudaStream_t st = as_stream(stream);
    if (ctx->cfg.actor_hidden == 32) {
        KernelTimer kt(ctx, K_ACT, st);
        actor_act_kernel<32><<<grid, RT, actor_smem_bytes<32>(), st>>>(a);
    } else {
        KernelTimer kt(ctx, K_ACT, st);
        actor_act_kernel<64><<<grid, RT, actor_smem_bytes<64>(), st>>>(a);
    }
    return cmarl_check_cuda(cudaGetLastError(), "actor_act_kernel");
}
