// Context, error reporting and configuration validation of libcmarl_b200.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void cmarl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cmarl_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    cmarl_set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

int cmarl_chain_setup(cmarl_ctx* ctx);   // chain.cu: shared-memory attributes + grid sizes
int cmarl_rollout_setup(cmarl_ctx* ctx); // rollout.cu: shared-memory attributes
int cmarl_gru_setup(cmarl_ctx* ctx);     // gru.cu: shared-memory attributes of the recurrent kernels
int cmarl_gen_setup(cmarl_ctx* ctx);     // generic.cu: shared-memory attribute of the layered rollout kernel
size_t cmarl_gen_workspace_bytes(const cmarl_ctx* ctx);

extern "C" int cmarl_version(void) { return CMARL_VERSION; }
extern "C" const char* cmarl_last_error(void) { return g_err; }

extern "C" int cmarl_ctx_create(const cmarl_config* cfg, cmarl_ctx** out) {
    CMARL_ARG(cfg && out, "null argument");
    *out = nullptr;
    CMARL_ARG(cfg->n_envs >= 1 && cfg->n_steps >= 1, "n_envs and n_steps must be positive");
    const int N = cfg->n_agents, L = cfg->n_landmarks > 0 ? cfg->n_landmarks : cfg->n_agents;
    CMARL_ARG(N >= 1 && N <= 8 && L >= 1 && L <= 8, "simple_spread: 1 <= n_agents, n_landmarks <= 8");
    const int R = 4 + 2 * L + 4 * (N - 1);      // vel, pos, landmarks - pos, others - pos, 2 silent comm slots per other agent
    CMARL_ARG(cfg->n_actions == 5, "simple_spread_v3 has 5 discrete actions (only A=5 is built)");
    CMARL_ARG(cfg->state_dim == N * R, "state_dim must be n_agents * (4 + 2 L + 4 (N - 1))  (54 for N = L = 3)");
    CMARL_ARG(cfg->obs_dim == R || cfg->obs_dim == R + N, "obs_dim must be the raw width (no ids) or raw + n_agents (agent ids)");
    CMARL_ARG(cfg->actor_layers >= 1 && cfg->actor_layers <= CMARL_GEN_MAX_LIN - 2 && cfg->critic_layers >= 1 &&
                  cfg->critic_layers <= CMARL_GEN_MAX_LIN - 2, "*_num_layers must be in [1, 6]");
    CMARL_ARG(cfg->actor_hidden >= 1 && cfg->actor_hidden <= CMARL_GEN_MAX_DIM && cfg->critic_hidden >= 1 &&
                  cfg->critic_hidden <= CMARL_GEN_MAX_DIM, "*_hidden_dim must be in [1, 256]");
    CMARL_ARG(cfg->critic_on_obs == 0 || cfg->critic_on_obs == 1, "critic_on_obs must be 0 or 1");
    CMARL_ARG(cfg->actor_recurrent == 0 || cfg->actor_recurrent == 1, "actor_recurrent must be 0 or 1");
    // the fused kernels' problem: N = L = 3, one hidden->hidden block, hidden 32 or 64; everything else is layered (generic.cu)
    const bool fused = N == 3 && L == 3 && cfg->actor_layers == 1 && cfg->critic_layers == 1 &&
                       (cfg->actor_hidden == 32 || cfg->actor_hidden == 64) && (cfg->critic_hidden == 32 || cfg->critic_hidden == 64);
    CMARL_ARG(!cfg->actor_recurrent || (fused && cfg->actor_hidden == 32),
              "the recurrent actor is built for the reference's default shapes only (3 agents, actor_hidden_dim 32, critic_num_layers 1)");
    int ndev = 0;
    CMARL_CUDA(cudaGetDeviceCount(&ndev));
    CMARL_ARG(cfg->device >= 0 && cfg->device < ndev, "no such CUDA device (there is no CPU fallback)");
    CMARL_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CMARL_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        cmarl_set_error("cmarl_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                        cfg->device, prop.major, prop.minor);
        return -2;
    }
    cmarl_ctx* ctx = (cmarl_ctx*)calloc(1, sizeof(cmarl_ctx));
    CMARL_ARG(ctx, "out of host memory");
    ctx->cfg = *cfg;
    ctx->generic = fused ? 0 : 1;
    ctx->n_landmarks = L;
    ctx->raw_obs = R;
    ctx->n_heads = cfg->critic_on_obs ? cfg->n_agents : 1;
    ctx->critic_in = cfg->critic_on_obs ? cfg->obs_dim : cfg->state_dim;
    ctx->actor.set(cfg->obs_dim, cfg->actor_hidden, cfg->n_actions);
    ctx->critic.set(ctx->critic_in, cfg->critic_hidden, 1);
    ctx->gru.set(cfg->obs_dim, cfg->actor_hidden, cfg->n_actions);
    if (cfg->actor_recurrent) ctx->actor.count = ctx->gru.count;
    ctx->gactor.set(cfg->obs_dim, cfg->actor_hidden, cfg->actor_layers, cfg->n_actions);
    ctx->gcritic.set(ctx->critic_in, cfg->critic_hidden, cfg->critic_layers, 1);
    if (ctx->generic) { ctx->actor.count = ctx->gactor.count; ctx->critic.count = ctx->gcritic.count; }
    ctx->sm_count = prop.multiProcessorCount;
    if (!ctx->generic && ctx->actor.count + ctx->critic.count + CMARL_N_STATS > CMARL_MAX_PARAMS) {
        cmarl_set_error("cmarl_ctx_create: %d parameters (limit %d)", ctx->actor.count + ctx->critic.count, CMARL_MAX_PARAMS - CMARL_N_STATS);
        free(ctx);
        return -1;
    }
    {
        int me = cmarl_check_cuda(cudaMalloc(&ctx->dev_words, CMARL_DEV_WORDS * sizeof(unsigned int)), "cudaMalloc(dev_words)");
        if (!me) me = cmarl_check_cuda(cudaMemset(ctx->dev_words, 0, CMARL_DEV_WORDS * sizeof(unsigned int)), "cudaMemset(dev_words)");
        if (!me) me = cmarl_check_cuda(cudaMalloc(&ctx->dev_floats, CMARL_DEV_FLOATS * sizeof(float)), "cudaMalloc(dev_floats)");
        if (me) { cudaFree(ctx->dev_words); free(ctx); return me; }
    }
    int e = 0;
    if (ctx->generic) {
        // the layered kernels keep their activations in a scratch block owned by the context (the entry points of the
        // fused path that have no workspace argument -- cmarl_critic_values -- need it too)
        e = cmarl_gen_setup(ctx);
        if (!e) e = cmarl_check_cuda(cudaMalloc(&ctx->gen_ws, cmarl_gen_workspace_bytes(ctx)), "cudaMalloc(generic workspace)");
    } else {
        e = cmarl_chain_setup(ctx);
        if (!e) e = cmarl_gru_setup(ctx);
        if (!e) e = cmarl_rollout_setup(ctx);
    }
    if (e) { cudaFree(ctx->gen_ws); cudaFree(ctx->dev_floats); cudaFree(ctx->dev_words); free(ctx); return e; }
    *out = ctx;
    return 0;
}

static void timing_drain(cmarl_ctx* ctx) {
    cmarl_timing* t = ctx->timing;
    if (!t) return;
    cudaDeviceSynchronize();
    for (int k = 0; k < CMARL_NK; ++k) {
        for (int i = 0; i < t->n[k]; ++i) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, t->ev[k][i][0], t->ev[k][i][1]) == cudaSuccess) {
                t->sum_ms[k] += ms;
                t->count[k] += 1;
            }
        }
        t->n[k] = 0;
    }
}

void cmarl_time_begin(cmarl_ctx* ctx, int id, cudaStream_t st) {
    cmarl_timing* t = ctx->timing;
    if (t->n[id] == CMARL_TIMING_POOL) timing_drain(ctx);
    cudaEvent_t* e = t->ev[id][t->n[id]];
    if (!e[0]) { cudaEventCreate(&e[0]); cudaEventCreate(&e[1]); }
    cudaEventRecord(e[0], st);
}

void cmarl_time_end(cmarl_ctx* ctx, int id, cudaStream_t st) {
    cmarl_timing* t = ctx->timing;
    cudaEventRecord(t->ev[id][t->n[id]][1], st);
    t->n[id]++;
}

extern "C" int cmarl_timing_enable(cmarl_ctx* ctx, int on) {
    CMARL_ARG(ctx, "null ctx");
    if (on && !ctx->timing) {
        ctx->timing = (cmarl_timing*)calloc(1, sizeof(cmarl_timing));
        CMARL_ARG(ctx->timing, "out of host memory");
    }
    if (!on && ctx->timing_on) timing_drain(ctx);
    ctx->timing_on = on ? 1 : 0;
    return 0;
}

extern "C" int cmarl_timing_read(cmarl_ctx* ctx, double* sum_ms, int64_t* launches) {
    CMARL_ARG(ctx && sum_ms && launches, "null argument");
    for (int k = 0; k < CMARL_NK; ++k) { sum_ms[k] = 0.0; launches[k] = 0; }
    if (!ctx->timing) return 0;
    timing_drain(ctx);
    for (int k = 0; k < CMARL_NK; ++k) {
        sum_ms[k] = ctx->timing->sum_ms[k];
        launches[k] = ctx->timing->count[k];
        ctx->timing->sum_ms[k] = 0.0;
        ctx->timing->count[k] = 0;
    }
    return 0;
}

extern "C" const char* cmarl_kernel_name(int id) {
    static const char* names[CMARL_NK] = {"env_reset", "env_step", "rollout", "actor_act", "critic_values",
                                          "td_lambda_scan", "normalize", "ppo_actor_chain", "ppo_critic_chain",
                                          "ppo_reduce_partials", "clip_adam", "ppo_tbptt_chunk"};
    return (id >= 0 && id < CMARL_NK) ? names[id] : "?";
}

extern "C" int cmarl_comm_detach(cmarl_ctx* ctx);

extern "C" int cmarl_ctx_destroy(cmarl_ctx* ctx) {
    cmarl_comm_detach(ctx);
    if (ctx && ctx->timing) {
        for (int k = 0; k < CMARL_NK; ++k)
            for (int i = 0; i < CMARL_TIMING_POOL; ++i)
                for (int j = 0; j < 2; ++j)
                    if (ctx->timing->ev[k][i][j]) cudaEventDestroy(ctx->timing->ev[k][i][j]);
        free(ctx->timing);
    }
    if (ctx && ctx->dev_words) cudaFree(ctx->dev_words);
    if (ctx && ctx->dev_floats) cudaFree(ctx->dev_floats);
    if (ctx && ctx->gen_ws) cudaFree(ctx->gen_ws);
    free(ctx);
    return 0;
}

extern "C" int cmarl_ctx_set_tensor_cores(cmarl_ctx* ctx, int on) {
    CMARL_ARG(ctx, "null ctx");
    ctx->use_tc = on ? 1 : 0;
    return 0;
}

extern "C" int cmarl_ctx_set_launch_chaining(cmarl_ctx* ctx, int on) {
    CMARL_ARG(ctx, "null ctx");
    ctx->launch_chaining = on ? 1 : 0;
    return 0;
}

extern "C" int cmarl_ctx_set_weight_decay(cmarl_ctx* ctx, double actor_wd, double critic_wd) {
    CMARL_ARG(ctx, "null ctx");
    CMARL_ARG(actor_wd >= 0.0 && critic_wd >= 0.0, "weight decay must be >= 0");
    ctx->weight_decay[0] = actor_wd;
    ctx->weight_decay[1] = critic_wd;
    return 0;
}

extern "C" int cmarl_actor_param_count(const cmarl_ctx* ctx) { return ctx ? ctx->actor.count : -1; }
extern "C" int cmarl_critic_param_count(const cmarl_ctx* ctx) { return ctx ? ctx->critic.count : -1; }
extern "C" int cmarl_value_heads(const cmarl_ctx* ctx) { return ctx ? ctx->n_heads : -1; }
extern "C" int cmarl_launch_count(const cmarl_ctx* ctx) { return ctx ? ctx->launches : -1; }
