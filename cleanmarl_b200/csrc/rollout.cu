// K1 + K2 + K3: device-resident rollout (simple_spread_v3 step + Actor.act + buffer stores) and the
// stand-alone Actor.act / env reset entries.  Compiled with -fmad=false (float64 physics must follow
// the oracle's rounding); the MLP uses explicit fmaf.
//
// Thread mapping: one thread per (env, agent); a CTA owns EPB = 32 envs as 3 warps, warp n = agent n,
// so every global access is a 128-B coalesced row of the [..][B] layout and the agent index is
// warp-uniform.  The env state (18 doubles per env) lives in shared memory for the whole episode;
// the T steps run inside ONE launch (no host round trip per step, MME:408-453).
#include "common.cuh"
#include "spread.cuh"

namespace {

constexpr int EPB = 32;            // envs per CTA
constexpr int NAG = 3;             // agents
constexpr int NACT = 5;
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

// Categorical(logits=z).sample() as the exponential race torch.multinomial runs on CPU
// (argmax(probs / q), first maximum wins) + log_prob of the drawn action (MME:174-176).
__device__ __forceinline__ void race_sample(const float (&z)[NACT], const float (&q)[NACT], int& action, float& logp) {
    float mx = z[0];
#pragma unroll
    for (int a = 1; a < NACT; ++a) mx = fmaxf(mx, z[a]);
    float se = 0.0f;
#pragma unroll
    for (int a = 0; a < NACT; ++a) se += expf(z[a] - mx);
    const float lse = mx + logf(se);
    float l[NACT], p[NACT];
    float mx2 = -INFINITY;
#pragma unroll
    for (int a = 0; a < NACT; ++a) { l[a] = z[a] - lse; mx2 = fmaxf(mx2, l[a]); }
    float se2 = 0.0f;
#pragma unroll
    for (int a = 0; a < NACT; ++a) { p[a] = expf(l[a] - mx2); se2 += p[a]; }
    float best = -1.0f;
    action = 0;
    logp = l[0];
#pragma unroll
    for (int a = 0; a < NACT; ++a) {
        const float r = (p[a] / se2) / q[a];
        if (r > best) { best = r; action = a; logp = l[a]; }
    }
}

__device__ __forceinline__ void philox_exp5(uint64_t seed, uint64_t episode, uint32_t t, uint32_t n, uint32_t b,
                                            float (&q)[NACT]) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const Philox4 r0 = philox4x32_10(b, t * 8u + n, (uint32_t)episode, (uint32_t)(episode >> 32) ^ 0x51u, k0, k1);
    const Philox4 r1 = philox4x32_10(b, t * 8u + n, (uint32_t)episode, (uint32_t)(episode >> 32) ^ 0xA3u, k0, k1);
    q[0] = -logf(u32_to_unit_open0(r0.x)); q[1] = -logf(u32_to_unit_open0(r0.y));
    q[2] = -logf(u32_to_unit_open0(r0.z)); q[3] = -logf(u32_to_unit_open0(r0.w));
    q[4] = -logf(u32_to_unit_open0(r1.x));
#pragma unroll
    for (int a = 0; a < NACT; ++a) q[a] = fmaxf(q[a], 1e-30f);
}

struct RolloutArgs {
    const float* actor;
    double* env;            // [18][B]
    const float* noise;     // [T][N][A][B] or null
    uint64_t seed, episode;
    float* state;           // [T][54][B]
    float* obs;             // [T][N][O][B] or null
    int32_t* actions;       // [T][N][B]
    float* logp;            // [T][N][B]
    float* reward;          // [T][B]
    double* ep_return;      // [B] or null
    int T, B, O;
};

template <int H>
__global__ void __launch_bounds__(RT) rollout_kernel(RolloutArgs a) {
    extern __shared__ __align__(16) float smf[];
    using S = ActorSmem<H>;
    __shared__ double es[18][EPB];
    __shared__ int acts[NAG][EPB];
    const int tid = threadIdx.x;
    const int n = tid / EPB, e = tid - n * EPB;          // warp n handles agent n
    const int b = blockIdx.x * EPB + e;
    const bool live = b < a.B;
    const int B = a.B;
    const bool fold = a.O > CMARL_RAW_OBS;

    load_actor<H>(smf, a.actor, a.O, fold, RT);
    for (int i = tid; i < 18 * EPB; i += RT) {
        const int r = i / EPB, c = i - r * EPB;
        const int bb = blockIdx.x * EPB + c;
        es[r][c] = (bb < B) ? a.env[(size_t)r * B + bb] : 0.0;
    }
    __syncthreads();

    float* col = smf + S::oAct + tid;
    const float* b1 = smf + S::oB1 + (fold ? n : 3) * H;   // row 3 holds the plain bias
    double ep_ret = 0.0;

    for (int t = 0; t < a.T; ++t) {
        double p[6], v[6], lm[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { p[i] = es[i][e]; v[i] = es[6 + i][e]; lm[i] = es[12 + i][e]; }
        // observation before the action (what the reference stores, MME:426-430)
        float x[CMARL_RAW_OBS];
        spread::observe(n, p, v, lm, x);
        if (live) {
#pragma unroll
            for (int k = 0; k < CMARL_RAW_OBS; ++k)
                __stcs(a.state + ((size_t)t * 54 + n * CMARL_RAW_OBS + k) * B + b, x[k]);
            if (a.obs) {
                float* o = a.obs + ((size_t)t * NAG + n) * a.O * B + b;
#pragma unroll
                for (int k = 0; k < CMARL_RAW_OBS; ++k) __stcs(o + (size_t)k * B, x[k]);
                if (fold)
                    for (int m = 0; m < NAG; ++m) __stcs(o + (size_t)(CMARL_RAW_OBS + m) * B, m == n ? 1.0f : 0.0f);
            }
        }
        float z[NACT];
        actor_mlp<H, CMARL_RAW_OBS>(x, smf, b1, col, RT, z);
        float q[NACT];
        if (a.noise) {
#pragma unroll
            for (int k = 0; k < NACT; ++k)
                q[k] = live ? __ldcs(a.noise + (((size_t)t * NAG + n) * NACT + k) * B + b) : 1.0f;
        } else {
            philox_exp5(a.seed, a.episode, (uint32_t)t, (uint32_t)n, (uint32_t)b, q);
        }
        int action; float lp;
        race_sample(z, q, action, lp);
        acts[n][e] = action;
        if (live) {
            __stcs(a.actions + ((size_t)t * NAG + n) * B + b, action);
            __stcs(a.logp + ((size_t)t * NAG + n) * B + b, lp);
        }
        __syncthreads();
        // physics: this thread integrates agent n (World.step), forces in the reference's pair order
        double fx, fy;
        spread::agent_force(n, p, acts[n][e], fx, fy);
        double px = p[2 * n], py = p[2 * n + 1], vx = v[2 * n], vy = v[2 * n + 1];
        spread::integrate(px, py, vx, vy, fx, fy);
        es[2 * n][e] = px; es[2 * n + 1][e] = py;
        es[6 + 2 * n][e] = vx; es[6 + 2 * n + 1][e] = vy;
        __syncthreads();
        if (n == 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) p[i] = es[i][e];
            const double r = spread::reward_agent0(p, lm);
            ep_ret += r;
            if (live) __stcs(a.reward + (size_t)t * B + b, (float)r);
        }
    }
    __syncthreads();
    for (int i = tid; i < 12 * EPB; i += RT) {
        const int r = i / EPB, c = i - r * EPB;
        const int bb = blockIdx.x * EPB + c;
        if (bb < B) a.env[(size_t)r * B + bb] = es[r][c];
    }
    if (n == 0 && live && a.ep_return) a.ep_return[b] = ep_ret;
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
__global__ void __launch_bounds__(256) env_reset_kernel(double* __restrict__ env, int B, uint64_t seed, uint64_t episode) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
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

extern "C" int cmarl_env_reset(cmarl_ctx* ctx, double* env, uint64_t seed, uint64_t episode, void* stream) {
    CMARL_ARG(ctx && env, "null argument");
    const int B = ctx->cfg.n_envs;
    {
        KernelTimer kt(ctx, K_RESET, as_stream(stream));
        env_reset_kernel<<<ceil_div(B, 256), 256, 0, as_stream(stream)>>>(env, B, seed, episode);
    }
    return cmarl_check_cuda(cudaGetLastError(), "env_reset_kernel");
}

extern "C" int cmarl_env_observe(cmarl_ctx* ctx, const double* env, float* state_out, void* stream) {
    CMARL_ARG(ctx && env && state_out, "null argument");
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
    RolloutArgs a;
    a.actor = actor_params; a.env = env; a.noise = noise; a.seed = seed; a.episode = episode;
    a.state = state; a.obs = obs; a.actions = actions; a.logp = logp; a.reward = reward; a.ep_return = ep_return;
    a.T = ctx->cfg.n_steps; a.B = ctx->cfg.n_envs; a.O = ctx->cfg.obs_dim;
    const int grid = ceil_div(a.B, EPB);
    cudaStream_t st = as_stream(stream);
    if (ctx->cfg.actor_hidden == 32) {
        KernelTimer kt(ctx, K_ROLLOUT, st);
        rollout_kernel<32><<<grid, RT, actor_smem_bytes<32>(), st>>>(a);
    } else {
        CMARL_CUDA(cudaFuncSetAttribute(rollout_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)actor_smem_bytes<64>()));
        KernelTimer kt(ctx, K_ROLLOUT, st);
        rollout_kernel<64><<<grid, RT, actor_smem_bytes<64>(), st>>>(a);
    }
    return cmarl_check_cuda(cudaGetLastError(), "rollout_kernel");
}

extern "C" int cmarl_actor_act(cmarl_ctx* ctx, const float* actor_params, const float* obs, const uint8_t* avail,
                               const float* noise, int32_t* actions, float* logp, float* logits_out, void* stream) {
    CMARL_ARG(ctx && actor_params && obs && noise && actions && logp, "null argument");
    ActArgs a;
    a.actor = actor_params; a.obs = obs; a.avail = avail; a.noise = noise; a.actions = actions; a.logp = logp;
    a.logits = logits_out; a.B = ctx->cfg.n_envs; a.O = ctx->cfg.obs_dim;
    const int grid = ceil_div(a.B, EPB);
    cudaStream_t st = as_stream(stream);
    if (ctx->cfg.actor_hidden == 32) {
        KernelTimer kt(ctx, K_ACT, st);
        actor_act_kernel<32><<<grid, RT, actor_smem_bytes<32>(), st>>>(a);
    } else {
        CMARL_CUDA(cudaFuncSetAttribute(actor_act_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)actor_smem_bytes<64>()));
        KernelTimer kt(ctx, K_ACT, st);
        actor_act_kernel<64><<<grid, RT, actor_smem_bytes<64>(), st>>>(a);
    }
    return cmarl_check_cuda(cudaGetLastError(), "actor_act_kernel");
}
