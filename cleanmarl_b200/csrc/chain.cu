// K4 (batched critic forward) and K7 (PPO epoch: fused forward + backward of actor and critic).
// See chain.cuh for the tile pipeline.  Reference arithmetic: MME:527-582 (loss), MME:178-200 (nets).
#include <stdlib.h>

#include "chain.cuh"
#include "heads.cuh"
#include "reduce.cuh"

namespace chain {

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
template <class C, class Head, bool TRAIN>
__global__ void __launch_bounds__(C::NT, 1)
chain_kernel(NetDesc nd, TileSrc src, typename Head::Args ha, float* __restrict__ partials, int p_net) {
    extern __shared__ __align__(128) float sm[];
    constexpr int H = C::H, M = C::M, LD = C::LD, NT = C::NT, OUT = Head::OUT;
    static_assert(M == NT, "S3/S5 are thread-per-sample");
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::oBar);
    const int tid = threadIdx.x;
    const int tiles_b = (src.nb + M - 1) / M;
    const int units = src.T * src.G * tiles_b;

    load_weights<C>(sm, nd, src.G);
    for (int i = tid; i < 2 * C::KIN * LD; i += NT) sm[C::oX + i] = 0.0f;      // pad rows stay zero
    if (TRAIN)
        for (int i = tid; i < C::PMAX; i += NT) sm[C::oDW + i] = 0.0f;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // gradient accumulators in the torch parameter order of this net
    float* dW1 = sm + C::oDW;
    float* db1 = dW1 + H * nd.in_dim;
    float* dW2 = db1 + H;
    float* db2 = dW2 + H * H;
    float* dW3 = db2 + H;
    float* db3 = dW3 + nd.out_dim * H;
    float* scr_patch = sm + C::oScr;
    float* scr_w3 = sm + C::oScr + C::SCR_PATCH;
    const int kin_pad = ((nd.in_rows + C::PT - 1) / C::PT) * C::PT;

    float st[Head::NSTAT];
#pragma unroll
    for (int k = 0; k < Head::NSTAT; ++k) st[k] = 0.0f;

    int u = blockIdx.x;
    if (u < units) {
        const int bt = u % tiles_b, r = u / tiles_b;
        issue_tile<C>(sm + C::oX, &bars[0], src, nd.in_rows, r / src.G, r % src.G, bt * M);
    }
    __syncthreads();
    for (int it = 0; u < units; u += gridDim.x, ++it) {
        const int cur = it & 1;
        const int un = u + gridDim.x;
        if (un < units) {
            const int bt = un % tiles_b, r = un / tiles_b;
            issue_tile<C>(sm + C::oX + (cur ^ 1) * C::KIN * LD, &bars[cur ^ 1], src, nd.in_rows, r / src.G, r % src.G,
                          bt * M);
        }
        const int bt = u % tiles_b, r = u / tiles_b;
        const int t = r / src.G, g = r % src.G, b0 = bt * M;
        float* X = sm + C::oX + cur * C::KIN * LD;
        float* H1 = sm + C::oH1;
        float* H2 = sm + C::oH2;
        float* Z = sm + C::oZ;
        mbar_wait(&bars[cur], (it >> 1) & 1);

        // S1, S2: hidden layers
        gemm_rows<C, H, false>(X, nd.in_rows, sm + C::oW1T, sm + C::oB1 + g * H, H1);
        __syncthreads();
        gemm_rows<C, H, false>(H1, H, sm + C::oW2T, sm + C::oB2, H2);
        __syncthreads();

        // S3: output layer + head, one thread per sample
        const int s = tid;
        float z[OUT], dz[OUT];
#pragma unroll
        for (int a = 0; a < OUT; ++a) z[a] = sm[C::oB3 + a];
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
            const float h = H2[j * LD + s];
            if (OUT > 1) {
                const float4 w = *reinterpret_cast<const float4*>(sm + C::oW3T + j * OUTP);
                const float w4 = sm[C::oW3T + j * OUTP + 4];
                z[0] = fmaf(w.x, h, z[0]);
                if (OUT > 1) z[OUT > 1 ? 1 : 0] = fmaf(w.y, h, z[OUT > 1 ? 1 : 0]);
                if (OUT > 2) z[OUT > 2 ? 2 : 0] = fmaf(w.z, h, z[OUT > 2 ? 2 : 0]);
                if (OUT > 3) z[OUT > 3 ? 3 : 0] = fmaf(w.w, h, z[OUT > 3 ? 3 : 0]);
                if (OUT > 4) z[OUT > 4 ? 4 : 0] = fmaf(w4, h, z[OUT > 4 ? 4 : 0]);
            } else {
                z[0] = fmaf(sm[C::oW3T + j * OUTP], h, z[0]);
            }
        }
        Head::apply(ha, z, t, g, b0 + s, src.G, src.B, (b0 + s) < src.nb, TRAIN, dz, st);

        if (TRAIN) {
#pragma unroll
            for (int a = 0; a < OUT; ++a) Z[a * LD + s] = dz[a];
            __syncthreads();

            // S4: dW3 += dz H2^T (split over samples, fixed-order combine), db3
            {
                constexpr int NS = NT / H;                 // sample splits
                constexpr int NQ = M / 4;
                const int j = tid % H, q = tid / H;
                float acc[OUT];
#pragma unroll
                for (int a = 0; a < OUT; ++a) acc[a] = 0.0f;
                for (int c = q; c < NQ; c += NS) {
                    const float4 h4 = *reinterpret_cast<const float4*>(H2 + j * LD + 4 * c);
#pragma unroll
                    for (int a = 0; a < OUT; ++a) {
                        const float4 d4 = *reinterpret_cast<const float4*>(Z + a * LD + 4 * c);
                        acc[a] = fmaf(d4.x, h4.x, acc[a]); acc[a] = fmaf(d4.y, h4.y, acc[a]);
                        acc[a] = fmaf(d4.z, h4.z, acc[a]); acc[a] = fmaf(d4.w, h4.w, acc[a]);
                    }
                }
#pragma unroll
                for (int a = 0; a < OUT; ++a) scr_w3[(q * OUTP + a) * H + j] = acc[a];
                __syncthreads();
                for (int i = tid; i < OUT * H; i += NT) {
                    const int a = i / H, jj = i - a * H;
                    float v = 0.0f;
#pragma unroll
                    for (int qq = 0; qq < NS; ++qq) v += scr_w3[(qq * OUTP + a) * H + jj];
                    dW3[a * H + jj] += v;
                }
                if (tid < OUT) {
                    float v = 0.0f;
                    for (int c = 0; c < NQ; ++c) {
                        const float4 d4 = *reinterpret_cast<const float4*>(Z + tid * LD + 4 * c);
                        v += (d4.x + d4.y) + (d4.z + d4.w);
                    }
                    db3[tid] += v;
                }
            }
            __syncthreads();

            // S5: dH2 = (W3^T dz) . relu'(H2), in place (dz still in this thread's registers)
#pragma unroll 4
            for (int j = 0; j < H; ++j) {
                float acc = 0.0f;
                if (OUT > 1) {
                    const float4 w = *reinterpret_cast<const float4*>(sm + C::oW3T + j * OUTP);
                    const float w4 = sm[C::oW3T + j * OUTP + 4];
                    acc = w.x * dz[0];
                    acc = fmaf(w.y, dz[OUT > 1 ? 1 : 0], acc);
                    acc = fmaf(w.z, dz[OUT > 2 ? 2 : 0], acc);
                    acc = fmaf(w.w, dz[OUT > 3 ? 3 : 0], acc);
                    acc = fmaf(w4, dz[OUT > 4 ? 4 : 0], acc);
                } else {
                    acc = sm[C::oW3T + j * OUTP] * dz[0];
                }
                const float h = H2[j * LD + s];
                H2[j * LD + s] = h > 0.0f ? acc : 0.0f;
            }
            __syncthreads();

            // S6: dW2 += dH2 H1^T, db2
            dw_stage<C>(H2, H1, H, H, H, dW2, db2, nullptr, scr_patch);
            __syncthreads();
            // S7: dH1 = (W2^T dH2) . relu'(H1), in place
            gemm_rows<C, H, true>(H2, H, sm + C::oW2, nullptr, H1);
            __syncthreads();
            // S8: dW1 += dH1 x^T, db1 (+ folded id column of this agent)
            dw_stage<C>(H1, X, kin_pad, nd.in_rows, nd.in_dim, dW1, db1,
                        nd.fold_ids ? dW1 + nd.in_rows + g : nullptr, scr_patch);
        }
        __syncthreads();
    }

    if (TRAIN) {
        float* out = partials + (size_t)blockIdx.x * (p_net + CMARL_N_STATS);
        for (int i = tid; i < p_net; i += NT) out[i] = sm[C::oDW + i];
        float* red = sm + C::oRed;
#pragma unroll
        for (int k = 0; k < Head::NSTAT; ++k) {
            const float v = warp_sum_f(st[k]);
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = v;
            __syncthreads();
            if (tid == 0) {
                float a = 0.0f;
                for (int w = 0; w < NT / 32; ++w) a += red[w];
                out[p_net + k] = a;
            }
        }
        if (tid == 0)
            for (int k = Head::NSTAT; k < CMARL_N_STATS; ++k) out[p_net + k] = 0.0f;
    }
}

// Fixed-order sum of the per-CTA partials -> flat gradient vector + statistics (deterministic): reduce.cuh.
// (512-thread CTAs: the 303 CTAs of the default shapes are resident at once -- 4 per SM.)
constexpr int RED_COLS = 32;
__global__ void __launch_bounds__(RED_COLS * RED_GROUPS) reduce_partials_kernel(ReduceArgs r, float* __restrict__ out) {
    pdl_wait_then_trigger();
    bool have; int i;
    const float v = reduce_column<RED_COLS>(r, blockIdx.x, &have, &i);
    if (have) out[i] = v;
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
template <class C, class Head, bool TRAIN>
static int set_attr() {
    const size_t smem = TRAIN ? C::smem_train : C::smem_fwd;
    return cmarl_check_cuda(cudaFuncSetAttribute(chain_kernel<C, Head, TRAIN>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(chain_kernel)");
}

template <class C, class Head, bool TRAIN>
static int launch(const cmarl_ctx* ctx, const NetDesc& nd, const TileSrc& src, const typename Head::Args& ha,
                  float* partials, int p_net, int grid, cudaStream_t st) {
    const size_t smem = TRAIN ? C::smem_train : C::smem_fwd;
    chain_kernel<C, Head, TRAIN><<<grid, C::NT, smem, st>>>(nd, src, ha, partials, p_net);
    return cmarl_check_cuda(cudaGetLastError(), "chain_kernel launch");
}

static int tile_m(int H, int kin) { return (H == 32 && kin == 24) ? 256 : 128; }

static int units_of(const TileSrc& s, int M) { return s.T * s.G * ceil_div(s.nb, M); }

template <class Head, bool TRAIN>
static int dispatch(const cmarl_ctx* ctx, int H, const NetDesc& nd, const TileSrc& src,
                    const typename Head::Args& ha, float* partials, int p_net, int grid, cudaStream_t st) {
    const int kin = nd.in_rows <= 24 ? 24 : 56;
    if (H == 32 && kin == 24) return launch<Cfg<32, 24>, Head, TRAIN>(ctx, nd, src, ha, partials, p_net, grid, st);
    if (H == 32 && kin == 56) return launch<Cfg<32, 56>, Head, TRAIN>(ctx, nd, src, ha, partials, p_net, grid, st);
    if (H == 64 && kin == 24) return launch<Cfg<64, 24>, Head, TRAIN>(ctx, nd, src, ha, partials, p_net, grid, st);
    if (H == 64 && kin == 56) return launch<Cfg<64, 56>, Head, TRAIN>(ctx, nd, src, ha, partials, p_net, grid, st);
    cmarl_set_error("chain dispatch: unsupported hidden=%d in_rows=%d", H, nd.in_rows);
    return -1;
}

// tiles per flush group of the tcgen05 chains (tc_chain.cu); CMARL_TC_FLUSH overrides for measurements
static int tc_flush_tiles() {
    static const int v = [] { const char* e = getenv("CMARL_TC_FLUSH"); const int n = e ? atoi(e) : 0; return n >= 1 && n <= 64 ? n : 4; }();
    return v;
}

static void actor_desc(const cmarl_ctx* ctx, const float* params, const float* state, const float* obs,
                       NetDesc& nd, TileSrc& src) {
    const cmarl_config& c = ctx->cfg;
    nd.params = params;
    nd.in_dim = c.obs_dim;
    nd.out_dim = c.n_actions;
    src.T = c.n_steps; src.G = c.n_agents; src.B = c.n_envs; src.nb = c.n_envs; src.indep = 0; src.flush = tc_flush_tiles();
    if (obs) {
        nd.in_rows = c.obs_dim; nd.fold_ids = 0;
        src.x = obs; src.stride_t = (size_t)c.n_agents * c.obs_dim * c.n_envs; src.stride_g = (size_t)c.obs_dim * c.n_envs;
    } else {
        nd.in_rows = CMARL_RAW_OBS; nd.fold_ids = c.obs_dim > CMARL_RAW_OBS;
        src.x = state; src.stride_t = (size_t)c.state_dim * c.n_envs; src.stride_g = (size_t)CMARL_RAW_OBS * c.n_envs;
    }
}

static void critic_desc(const cmarl_ctx* ctx, const float* params, const float* state, const float* obs,
                        NetDesc& nd, TileSrc& src) {
    const cmarl_config& c = ctx->cfg;
    if (c.critic_on_obs) {
        actor_desc(ctx, params, state, obs, nd, src);
        nd.out_dim = 1;
        return;
    }
    nd.params = params;
    nd.in_rows = c.state_dim; nd.in_dim = c.state_dim; nd.fold_ids = 0; nd.out_dim = 1;
    src.x = state; src.T = c.n_steps; src.G = 1; src.B = c.n_envs; src.nb = c.n_envs; src.indep = 0; src.flush = tc_flush_tiles();
    src.stride_t = (size_t)c.state_dim * c.n_envs; src.stride_g = 0;
}

}  // namespace chain

using namespace chain;

// generic.cu
int cmarl_gen_critic_values(cmarl_ctx* ctx, const float* critic_params, const float* state, const float* obs, float* values,
                            void* workspace, cudaStream_t st);
int cmarl_gen_ppo_epoch_grads(cmarl_ctx* ctx, const float* params, const float* state, const float* obs, const int32_t* actions,
                              const float* logp_old, const float* adv, const float* returns, const float* values_old,
                              const uint8_t* mask, const uint8_t* avail, double clip, double ent_coef, double value_clip,
                              int env_count, float* grads_out, void* workspace, cudaStream_t st);

// tc_chain.cu
int cmarl_tc_setup();
int cmarl_tc_tile();
int cmarl_tc_ctas_per_sm(int H, int in_rows, bool train, int out);
template <class Head, bool TRAIN>
int cmarl_tc_dispatch(const cmarl_ctx* ctx, int H, const NetDesc& nd, const TileSrc& src, const typename Head::Args& ha,
                      float* partials, int p_net, int grid, cudaStream_t st);

template <class Head, bool TRAIN>
static int run_chain(const cmarl_ctx* ctx, int H, const NetDesc& nd, const TileSrc& src, const typename Head::Args& ha,
                     float* partials, int p_net, int* grid_out, cudaStream_t st) {
    const int kin = nd.in_rows <= 24 ? 24 : 56;
    const int m = ctx->use_tc ? cmarl_tc_tile() : tile_m(H, kin);
    const int units = units_of(src, m);
    const int slots = ctx->sm_count * (ctx->use_tc ? cmarl_tc_ctas_per_sm(H, nd.in_rows, TRAIN, Head::OUT) : 1);
    int grid = units < slots ? units : slots;
    {   // diagnostics only (profiles/tools/grad_accuracy.py): fewer persistent CTAs = more tiles per CTA at a given size
        static const int cap = [] { const char* v = getenv("CMARL_DEBUG_GRID_CAP"); return v ? atoi(v) : 0; }();
        if (cap > 0 && grid > cap) grid = cap;
    }
    if (grid_out) *grid_out = grid;
    if (ctx->use_tc && units >= (1 << 24)) {      // tc_chain_kernel decodes tile indices with fast_divmod (exact below 2^24)
        cmarl_set_error("chain: %d tiles in one launch (limit 2^24)", units);
        return -1;
    }
    if (ctx->use_tc) return cmarl_tc_dispatch<Head, TRAIN>(ctx, H, nd, src, ha, partials, p_net, grid, st);
    return dispatch<Head, TRAIN>(ctx, H, nd, src, ha, partials, p_net, grid, st);
}

int cmarl_chain_setup(cmarl_ctx* ctx) {
    int e = 0;
#define SET(Hh, Kk)                                                        \
    if (!e) e = set_attr<Cfg<Hh, Kk>, PolicyHead, true>();                 \
    if (!e) e = set_attr<Cfg<Hh, Kk>, ValueHead, true>();                  \
    if (!e) e = set_attr<Cfg<Hh, Kk>, ValueHead, false>();
    SET(32, 24) SET(32, 56) SET(64, 24) SET(64, 56)
#undef SET
    if (!e) e = cmarl_tc_setup();
    if (e) return e;
    const cmarl_config& c = ctx->cfg;
    NetDesc nd; TileSrc src;
    actor_desc(ctx, nullptr, nullptr, nullptr, nd, src);
    int ua = units_of(src, tile_m(c.actor_hidden, 24));
    critic_desc(ctx, nullptr, nullptr, nullptr, nd, src);
    int uc = units_of(src, tile_m(c.critic_hidden, nd.in_rows <= 24 ? 24 : 56));
    ctx->ppo_grid_actor = ua < ctx->sm_count ? ua : ctx->sm_count;
    ctx->ppo_grid_critic = uc < ctx->sm_count ? uc : ctx->sm_count;
    return 0;
}

extern "C" size_t cmarl_workspace_bytes(const cmarl_ctx* ctx) {
    if (!ctx) return 0;
    if (ctx->generic) return 256;      // the layered kernels use the context's own scratch block
    // the actor grid may be larger when obs is passed explicitly (21 rows still use the 24-row config)
    // up to two persistent CTAs per SM (tensor-core kernels of the 32-wide networks), one partial row per CTA
    // (the recurrent chunk kernel, gru.cu, launches at most one CTA per SM: covered as well)
    const size_t a = (size_t)2 * ctx->sm_count * (ctx->actor.count + CMARL_N_STATS);
    const size_t c = (size_t)2 * ctx->sm_count * (ctx->critic.count + CMARL_N_STATS);
    // recurrent actor on the tensor cores (tc_gru.cu): the forward kernel's partial rows (dW2 | db2 | statistics, one per
    // forward CTA, up to two per SM) live in the second half of the actor block; its dlogits [T][N][8][B] behind both blocks
    const size_t dl = ctx->cfg.actor_recurrent ? (size_t)ctx->cfg.n_steps * ctx->cfg.n_agents * 8 * ctx->cfg.n_envs : 0;
    return (a + c + dl) * sizeof(float);
}

extern "C" int cmarl_critic_values(cmarl_ctx* ctx, const float* critic_params, const float* state, const float* obs,
                                   float* values, void* stream) {
    CMARL_ARG(ctx && critic_params && values, "null argument");
    CMARL_ARG(ctx->cfg.critic_on_obs ? (state || obs) : (state != nullptr), "critic input missing");
    if (ctx->generic) return cmarl_gen_critic_values(ctx, critic_params, state, obs, values, ctx->gen_ws, as_stream(stream));
    NetDesc nd; TileSrc src;
    critic_desc(ctx, critic_params, state, obs, nd, src);
    ValueHeadArgs ha;
    ha.returns = nullptr; ha.mask = nullptr; ha.values_out = values; ha.inv_heads = 1.0f / (float)ctx->n_heads;
    ha.values_old = nullptr; ha.vclip = 0.0f;
    KernelTimer kt(ctx, K_CRITIC, as_stream(stream));
    return run_chain<ValueHead, false>(ctx, ctx->cfg.critic_hidden, nd, src, ha, nullptr, 0, nullptr, as_stream(stream));
}

extern "C" int cmarl_ppo_epoch_grads(cmarl_ctx* ctx, const float* params, const float* state, const float* obs,
                                     const int32_t* actions, const float* logp_old, const float* adv,
                                     const float* returns, const uint8_t* mask, const uint8_t* avail,
                                     double clip, double ent_coef, float* grads_out, void* workspace, void* stream) {
    return cmarl_ppo_epoch_grads_ex(ctx, params, state, obs, actions, logp_old, adv, returns, nullptr, mask, avail, clip,
                                    ent_coef, -1.0, 0, ctx ? ctx->cfg.n_envs : 0, grads_out, workspace, stream);
}

// One minibatch = the contiguous env block [env_begin, env_begin + env_count): every buffer keeps its row stride B, the
// base pointers move by env_begin, and the kernels' bound is the block's env count (TileSrc.nb).
extern "C" int cmarl_ppo_epoch_grads_ex(cmarl_ctx* ctx, const float* params, const float* state, const float* obs,
                                        const int32_t* actions, const float* logp_old, const float* adv,
                                        const float* returns, const float* values_old, const uint8_t* mask,
                                        const uint8_t* avail, double clip, double ent_coef, double value_clip,
                                        int32_t env_begin, int32_t env_count, float* grads_out, void* workspace,
                                        void* stream) {
    CMARL_ARG(ctx && params && actions && logp_old && adv && returns && workspace, "null argument");
    CMARL_ARG(grads_out || !ctx->generic, "grads_out = NULL (partial rows for cmarl_reduce_clip_adam_step) needs the fused kernels");
    CMARL_ARG(!ctx->cfg.actor_recurrent, "recurrent actor: use cmarl_tbptt_chunk_grads + cmarl_critic_epoch_grads");
    CMARL_ARG(state || obs, "state or obs required");
    CMARL_ARG(ctx->cfg.critic_on_obs || state, "MAPPO critic needs state");
    CMARL_ARG(env_begin >= 0 && env_count >= 1 && env_begin + env_count <= ctx->cfg.n_envs, "env range outside [0, n_envs)");
    CMARL_ARG(value_clip <= 0.0 || values_old, "value clipping needs values_old");
    const cmarl_config& c = ctx->cfg;
    cudaStream_t st = as_stream(stream);
    const int Pa = ctx->actor.count, Pc = ctx->critic.count;
    float* part_a = reinterpret_cast<float*>(workspace);
    float* part_c = part_a + (size_t)2 * ctx->sm_count * (Pa + CMARL_N_STATS);
    if (env_begin) {      // every buffer is env-minor: the block starts env_begin elements into each row
        if (state) state += env_begin;
        if (obs) obs += env_begin;
        actions += env_begin; logp_old += env_begin; adv += env_begin; returns += env_begin;
        if (values_old) values_old += env_begin;
        if (mask) mask += env_begin;
        if (avail) avail += env_begin;
    }
    if (ctx->generic)
        return cmarl_gen_ppo_epoch_grads(ctx, params, state, obs, actions, logp_old, adv, returns, values_old, mask, avail, clip,
                                         ent_coef, value_clip, env_count, grads_out, ctx->gen_ws, st);

    NetDesc nda; TileSrc srca;
    actor_desc(ctx, params, state, obs, nda, srca);
    srca.nb = env_count;
    PolicyHeadArgs pa;
    pa.actions = actions; pa.logp_old = logp_old; pa.adv = adv; pa.mask = mask; pa.avail = avail;
    pa.V = ctx->n_heads; pa.A = c.n_actions;
    pa.clip = (float)clip; pa.ent_coef = (float)ent_coef; pa.inv_groups = 1.0f / (float)c.n_agents;
    int grid_a = 0, grid_c = 0;
    int e;
    {
        KernelTimer kt(ctx, K_PPO_ACTOR, st);
        e = run_chain<PolicyHead, true>(ctx, c.actor_hidden, nda, srca, pa, part_a, Pa, &grid_a, st);
    }
    if (e) return e;

    NetDesc ndc; TileSrc srcc;
    critic_desc(ctx, params + Pa, state, obs, ndc, srcc);
    srcc.nb = env_count;
    ValueHeadArgs va;
    va.returns = returns; va.mask = mask; va.values_out = nullptr; va.inv_heads = 1.0f / (float)ctx->n_heads;
    va.values_old = values_old; va.vclip = value_clip > 0.0 ? (float)value_clip : 0.0f;
    {
        // the critic chain reads nothing the actor chain writes (parameters come from the previous Adam step, which the
        // actor chain has waited for): launched as its programmatic dependent with `indep` set it starts on every SM the
        // actor chain's uneven last round of tiles leaves idle.  Not while the per-kernel event timing is on (the
        // bracketing events would separate the two launches anyway).
        static const bool pdl_ok = [] { const char* v = getenv("CMARL_PDL"); return !(v && v[0] == '0'); }();
        srcc.indep = (ctx->use_tc && pdl_ok && !ctx->timing_on) ? 1 : 0;
        KernelTimer kt(ctx, K_PPO_CRITIC, st);
        e = run_chain<ValueHead, true>(ctx, c.critic_hidden, ndc, srcc, va, part_c, Pc, &grid_c, st);
    }
    if (e) return e;

    if (!grads_out) {      // the caller reduces and steps in one launch: cmarl_reduce_clip_adam_step
        ctx->pending_grid_a = grid_a; ctx->pending_grid_c = grid_c;
        return 0;
    }
    const int n_out = Pa + Pc + CMARL_N_STATS;
    {
        KernelTimer kt(ctx, K_PPO_REDUCE, st);
        ReduceArgs ra;
        ra.pa = part_a; ra.grid_a = grid_a; ra.Pa = Pa; ra.pc = part_c; ra.grid_c = grid_c; ra.Pc = Pc;
        ra.n_groups = (float)c.n_agents; ra.count_from_c = 0;
        CMARL_CUDA(cmarl_launch(ctx, reduce_partials_kernel, dim3(ceil_div(n_out, RED_COLS)), dim3(RED_COLS * RED_GROUPS), 0, st, ra, grads_out));
    }
    return 0;
}

// Fixed-order reduction of one network's per-CTA partials (used by the recurrent path, gru.cu): actor-only
// (pc == nullptr) or critic-only (pa == nullptr; the valid count then comes from the value head's sample count).
int cmarl_reduce_one_net(cmarl_ctx* ctx, const float* pa, int grid_a, int Pa, const float* pc, int grid_c, int Pc,
                         float count_div, float* out, cudaStream_t st) {
    const int n_out = Pa + Pc + CMARL_N_STATS;
    {
        KernelTimer kt(ctx, K_PPO_REDUCE, st);
        ReduceArgs ra;
        ra.pa = pa; ra.grid_a = grid_a; ra.Pa = Pa; ra.pc = pc; ra.grid_c = grid_c; ra.Pc = Pc;
        ra.n_groups = count_div; ra.count_from_c = (int)(pa == nullptr);
        CMARL_CUDA(cmarl_launch(ctx, reduce_partials_kernel, dim3(ceil_div(n_out, RED_COLS)), dim3(RED_COLS * RED_GROUPS), 0, st, ra, out));
    }
    return 0;
}

extern "C" int cmarl_critic_epoch_grads(cmarl_ctx* ctx, const float* critic_params, const float* state, const float* obs,
                                        const float* returns, const uint8_t* mask, float* grads_out, void* workspace,
                                        void* stream) {
    CMARL_ARG(ctx && critic_params && returns && grads_out && workspace, "null argument");
    CMARL_ARG(ctx->cfg.critic_on_obs ? (state || obs) : (state != nullptr), "critic input missing");
    CMARL_ARG(!ctx->generic, "the per-network entries serve the recurrent path (default shapes only)");
    cudaStream_t st = as_stream(stream);
    const int Pc = ctx->critic.count;
    float* part_c = reinterpret_cast<float*>(workspace);
    NetDesc ndc; TileSrc srcc;
    critic_desc(ctx, critic_params, state, obs, ndc, srcc);
    ValueHeadArgs va;
    va.returns = returns; va.mask = mask; va.values_out = nullptr; va.inv_heads = 1.0f / (float)ctx->n_heads;
    va.values_old = nullptr; va.vclip = 0.0f;
    int grid_c = 0, e;
    {
        KernelTimer kt(ctx, K_PPO_CRITIC, st);
        e = run_chain<ValueHead, true>(ctx, ctx->cfg.critic_hidden, ndc, srcc, va, part_c, Pc, &grid_c, st);
    }
    if (e) return e;
    // the value head counts one sample per (t, head, b): divide by the number of heads for the (b,t) count
    return cmarl_reduce_one_net(ctx, nullptr, 0, 0, part_c, grid_c, Pc, (float)ctx->n_heads, grads_out, st);
}
