/*
 * cmarl_b200.h -- C ABI of libcmarl_b200.so: the sm_100a kernels behind the MAPPO multi-env
 * training path of CleanMARL (reference: cleanmarl/mappo_multienvs.py, "MME" below).
 *
 * The reference is pure Python and exposes no FFI seam; the entry points
 * below are what a ctypes binding inside the reference script would call to replace, one by one,
 * the inline blocks of its __main__ loop.  Each entry cites the reference lines it replaces.
 *
 * Conventions
 *  - every pointer is a CUDA *device* pointer owned by the caller (torch tensors) unless it is
 *    marked HOST; the library never allocates device memory;
 *  - launches go on the passed cudaStream_t (as void*), no hidden synchronisation;
 *  - return value: 0 ok, < 0 argument error, > 0 cudaError_t; cmarl_last_error() gives the text;
 *  - one context per GPU, single host thread per context (not re-entrant).
 *
 * Device layout ("time-major, feature-major, env-minor": the env index b is always the fastest
 * dimension so a warp touches 128 contiguous bytes; SURVEY.md section 7 "hard parts"):
 *    state    f32 [T][S][B]      raw per-agent observations of the N agents concatenated (S = N*18)
 *    obs      f32 [T][N][O][B]   optional; NULL => rebuilt from state: obs[n][k<18] = state[18n+k],
 *                                obs[n][18+m] = (m == n) one-hot id (pettingzoo_wrapper.py:93-98)
 *    actions  i32 [T][N][B]      logp f32 [T][N][B]     reward f32 [T][B]
 *    mask     u8  [T][B]         NULL => all ones        avail u8 [T][N][A][B]  NULL => all ones
 *    values / returns / adv  f32 [T][V][B], V = 1 (MAPPO, centralised critic; the reference
 *                                broadcasts the scalar to the N agents, MME:484-485) or V = N (IPPO)
 *    noise    f32 [T][N][A][B]   Exp(1) race noise q (Categorical.sample == argmax(p/q))
 *    env      f64 [18][B]        rows 0-5 agent p_pos (x0,y0,x1,y1,x2,y2), 6-11 agent p_vel,
 *                                12-17 landmark p_pos
 *  Parameters are ONE flat f32 vector in the order of torch's module.parameters():
 *    actor  W1[Ha][O] b1[Ha] W2[Ha][Ha] b2[Ha] W3[A][Ha] b3[A]   then
 *    critic W1[Hc][Sin] b1[Hc] W2[Hc][Hc] b2[Hc] W3[1][Hc] b3[1]   (row-major W[out][in], y = x W^T + b)
 *  Recurrent actor (cmarl_config.actor_recurrent = 1; mappo_lstm_multienvs.py:162-184), 7 205 floats:
 *    fc1.W[H][O] fc1.b[H] gru.weight_ih[3H][H] gru.weight_hh[3H][H] gru.bias_ih[3H] gru.bias_hh[3H]
 *    fc2.W[A][H] fc2.b[A]      (torch.nn.GRUCell gate order r, z, n)
 *    hidden  f32 [N][H][B]     one hidden state per (agent, env); h_seq f32 [T+1][N][H][B]: slice t = state BEFORE step t
 */
#ifndef CMARL_B200_H
#define CMARL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMARL_VERSION 104
#define CMARL_N_STATS 8          /* floats appended to the flat gradient vector, see cmarl_ppo_epoch_grads */
#define CMARL_RAW_OBS 18

typedef struct cmarl_ctx cmarl_ctx;

typedef struct cmarl_config {
    int32_t device;          /* CUDA ordinal */
    int32_t n_envs;          /* B on this GPU (MME Args.batch_size, sharded across GPUs) */
    int32_t n_steps;         /* T = 25 for simple_spread_v3 (max_cycles) */
    int32_t n_agents;        /* N (3 in the reference: kwargs = {} at MME:297) */
    int32_t obs_dim;         /* O = 21 with agent ids (MME:26), 18 without */
    int32_t state_dim;       /* S = 54 */
    int32_t n_actions;       /* A = 5 */
    int32_t actor_hidden;    /* MME Args.actor_hidden_dim  (32) */
    int32_t actor_layers;    /* MME Args.actor_num_layers  (1)  */
    int32_t critic_hidden;   /* MME Args.critic_hidden_dim (64; ippo_multienvs.py:34 -> 32) */
    int32_t critic_layers;   /* MME Args.critic_num_layers (1)  */
    int32_t critic_on_obs;   /* 0: MAPPO critic(state) MME:336; 1: IPPO critic(obs) ippo_multienvs.py:336 */
    int32_t actor_recurrent; /* 0: MLP actor MME:160-183; 1: fc1 + GRUCell + fc2, mappo_lstm_multienvs.py:162-184
                                (actor_layers is then ignored, as in the reference; actor_hidden must be 32) */
    int32_t n_landmarks;     /* L; 0 = n_agents (simple_spread_v3(N): N agents, N landmarks) */
} cmarl_config;
/* Shapes.  The fused kernels (rollout_kernel, the tcgen05 / FFMA chain kernels, clip_adam_kernel) are built for the
 * reference's default problem: N = L = 3, *_num_layers = 1, hidden 32 / 64.  Everything else the reference's CLI
 * accepts -- any *_num_layers >= 1 (MME:160-171, 186-196), any hidden width <= 256, 1 <= N, L <= 8 -- runs through the
 * layered kernels of csrc/generic.cu behind the SAME entry points (same layouts with S = N * R, R = 4 + 2 L + 4 (N - 1),
 * O = R (+ N ids); env f64 [4 N + 2 L][B]: agent positions, agent velocities, landmark positions).  The recurrent
 * actor exists for the default shapes only. */

/* -- library ---------------------------------------------------------------------------- */
int cmarl_version(void);
const char* cmarl_last_error(void);

/* Context: validates the configuration (unsupported shapes fail here, loudly), records the device
 * properties and sets the kernels' shared-memory attributes.  Replaces nothing in the reference;
 * it is the handle the entries below share. */
int cmarl_ctx_create(const cmarl_config* cfg, cmarl_ctx** out);
int cmarl_ctx_destroy(cmarl_ctx* ctx);
/* Selects the implementation of the fused MLP chains (cmarl_critic_values, cmarl_ppo_epoch_grads):
 * 1 = tcgen05 tensor cores, kind::tf32 with a 3-term split (fp32-level accuracy, csrc/tc_chain.cu);
 * 0 = fp32 FFMA block GEMMs (csrc/chain.cu).  Both produce the same quantities to the stated tolerances. */
int cmarl_ctx_set_tensor_cores(cmarl_ctx* ctx, int on);
/* Launch chaining (off initially): while on, every kernel this context launches carries CUDA's programmatic-stream-
 * serialization attribute, i.e. it may become resident while the launch in front of it ON THE SAME STREAM is still
 * running and blocks (griddepcontrol.wait) until that launch has completed before touching global memory -- launch
 * latency and kernel prologues then hide under the predecessor.  Results are unchanged.  Contract for the caller:
 * switch it on only between two launches of this library with nothing else enqueued on the stream in between
 * (a training iteration after its first kernel, MME:379-594), and off again before enqueuing foreign work.
 * Ignored while cmarl_timing_enable is on. */
int cmarl_ctx_set_launch_chaining(cmarl_ctx* ctx, int on);
/* Decoupled weight decay for cmarl_clip_adam_step / cmarl_adam_step_net: torch.optim.AdamW's param.mul_(1 - lr * wd)
 * before the Adam update (--optimizer AdamW, the default of ippo_lstm_multienvs.py:38; torch default wd = 0.01).
 * 0 (the initial value) = torch.optim.Adam. */
int cmarl_ctx_set_weight_decay(cmarl_ctx* ctx, double actor_weight_decay, double critic_weight_decay);
int cmarl_actor_param_count(const cmarl_ctx* ctx);    /* 1 925 for the default shapes */
int cmarl_critic_param_count(const cmarl_ctx* ctx);   /* 7 745 */
int cmarl_value_heads(const cmarl_ctx* ctx);          /* V */
size_t cmarl_workspace_bytes(const cmarl_ctx* ctx);   /* scratch for cmarl_ppo_epoch_grads */
int cmarl_launch_count(const cmarl_ctx* ctx);         /* kernels launched through this ctx so far */

/* Per-kernel device timing (measurement aid for bench.py, off by default): when enabled every launch
 * is bracketed by a cudaEvent pair on its own stream.  cmarl_timing_read synchronises the device and
 * returns, per kernel id, the summed duration in ms and the number of launches since the last read. */
#define CMARL_NK 12
int cmarl_timing_enable(cmarl_ctx* ctx, int on);
int cmarl_timing_read(cmarl_ctx* ctx, double* sum_ms /* HOST [CMARL_NK] */, int64_t* launches /* HOST [CMARL_NK] */);
const char* cmarl_kernel_name(int id);

/* -- K1: env reset.  Replaces the ("reset", None) round trip MME:393-401 -> env_worker MME:250-255
 * -> PettingZooWrapper.reset pettingzoo_wrapper.py:32-38 -> simple_spread reset_world: agent
 * positions then landmark positions ~ U(-1,1), velocities 0.  Draws come from Philox4x32-10
 * keyed by (seed, episode); callers that need given start positions write `env` themselves. */
int cmarl_env_reset(cmarl_ctx* ctx, double* env, uint64_t seed, uint64_t episode, void* stream);

/* Device-resident episode counter (optional): when set, cmarl_env_reset and cmarl_rollout key their Philox draws by
 * *episode_dev instead of their by-value `episode` argument, and cmarl_episode_advance increments it on the stream --
 * so a whole training iteration is a fixed launch sequence that a CUDA graph can replay (the reference draws fresh
 * resets / samples every iteration, MME:393-401, 410-414).  NULL restores the by-value behaviour. */
int cmarl_ctx_set_episode_counter(cmarl_ctx* ctx, uint64_t* episode_dev);
int cmarl_episode_advance(cmarl_ctx* ctx, void* stream);

/* -- K1 alone: the env duck-type one call at a time (cleanmarl/env/common_interface.py:5-23).
 * cmarl_env_observe: raw observations of the current state, state_out f32 [S][B]  (get_state /
 * the obs returned by reset, pettingzoo_wrapper.py:32-38, 93-98).
 * cmarl_env_step: one World.step for given actions i32 [N][B] (PettingZooWrapper.step,
 * pettingzoo_wrapper.py:44-66): updates env, writes the new raw observations and agent 0's reward. */
int cmarl_env_observe(cmarl_ctx* ctx, const double* env, float* state_out, void* stream);
int cmarl_env_step(cmarl_ctx* ctx, double* env, const int32_t* actions, float* state_out, float* reward_out,
                   void* stream);

/* -- K1+K2+K3: one rollout of T lock-step steps for the B envs, entirely on the device.
 * Replaces the whole `while len(alive_envs) > 0` loop MME:408-453: Actor.act (MME:172-183,
 * Categorical sample + log_prob), the ("step", action) round trip (MME:415-417, env_worker
 * MME:268-281, PettingZooWrapper.step pettingzoo_wrapper.py:44-66, MPE World.step), and the
 * per-step buffer appends (MME:425-434) / RolloutBuffer.add (MME:103-107).
 *   noise   NULL => q drawn on the device (Philox, keyed by seed/episode); else the given q is used
 *           so that actions are a deterministic function of (params, env, q)
 *   state, actions, logp, reward   outputs in the device layout above
 *   obs     optional output [T][N][O][B] (NULL to skip; K7 rebuilds it from state)
 *   ep_return f64 [B]   sum over t of the team reward (MME:433), optional
 * `env` holds the start state on entry and the final state on exit.
 * MLP actor (hidden 32 / 64): rollout_tc_kernel, layer 2 on tcgen05 (3xTF32) and the exponential race decided in the log
 * domain (argmax_a z_a - log q_a == argmax_a softmax(z)_a / q_a up to races closer than a few ulp); the environment
 * variable CMARL_ROLLOUT=ffma selects the CUDA-core kernel with the reference's form of the race (read at every call). */
int cmarl_rollout(cmarl_ctx* ctx, const float* actor_params, double* env, const float* noise,
                  uint64_t seed, uint64_t episode,
                  float* state, float* obs, int32_t* actions, float* logp, float* reward,
                  double* ep_return, void* stream);

/* -- K2 alone: Actor.act on given observations (MME:172-176), used by the parity tests and the
 * eval loop.  obs [N][O][B] (one time step), avail u8 [N][A][B] or NULL, noise [N][A][B]. */
int cmarl_actor_act(cmarl_ctx* ctx, const float* actor_params, const float* obs, const uint8_t* avail,
                    const float* noise, int32_t* actions, float* logp, float* logits_out, void* stream);

/* -- K4: batched critic forward.  Replaces the 2*B*T batch-of-1 calls critic(x=b_states[ep, t])
 * of MME:495,502 (IPPO: b_obs, ippo_multienvs.py:495,503) by one pass.
 *   critic_in = state [T][S][B] (MAPPO) or obs [T][N][O][B] / NULL=>from state (IPPO) */
int cmarl_critic_values(cmarl_ctx* ctx, const float* critic_params, const float* state, const float* obs,
                        float* values, void* stream);

/* -- K5: TD(lambda) return / advantage scan, the recurrence of MME:484-504 in the reference's
 * operation order (fp32, no FMA): R_t = r_t + g*(l*R_{t+1} + (1-l)*V_{t+1}), V_{T_ep} := 0,
 * A_t = R_t - V_t.  Entries with mask == 0 are written as 0 (MME:484-485 zero-init). */
int cmarl_td_lambda(cmarl_ctx* ctx, const float* values, const float* reward, const uint8_t* mask,
                    double gamma, double lambda, float* returns, float* adv, void* stream);

/* -- K6: the optional normalisations.  mode 0: reward, (r-mean)/(std+1e-6) over masked entries,
 * unbiased std (MME:143-146); mode 1: advantages / returns, (x-mean)/std of the agent-mean over
 * masked (b,t), unbiased, no eps (MME:505-512).  x is [T][V][B] (V=1 for mode 0), in place.
 * stats_io f64 [4]: (sum, sumsq, count, flag).  phase 0 computes the local sums into stats_io,
 * phase 1 applies them -- a multi-GPU caller all-reduces stats_io[0..2] in between. */
int cmarl_normalize(cmarl_ctx* ctx, float* x, int32_t n_heads, const uint8_t* mask, int32_t mode,
                    int32_t phase, double* stats_io, void* stream);

/* -- K7: one PPO epoch's loss + gradients, forward and backward fused (MME:522-582).
 * Outputs UNNORMALISED sums over this GPU's envs so that shards add: grads_out is
 * f32 [Pa + Pc + CMARL_N_STATS]:
 *   [0,Pa)      d/d(actor params)  of  sum_{b,t} mean_n( -min(A r, A clamp(r)) - ent_coef * H )
 *   [Pa,Pa+Pc)  d/d(critic params) of  sum_{b,t} mean_v (V - R)^2
 *   stats: [0] actor loss sum [1] critic loss sum [2] entropy sum [3] kl sum [4] clip-fraction sum
 *          [5] number of valid (b,t) = b_mask.sum() [6],[7] reserved (0)
 * All become the reference's values after division by stats[5] (MME:572-576), which
 * cmarl_clip_adam_step does after the caller's all-reduce. */
int cmarl_ppo_epoch_grads(cmarl_ctx* ctx, const float* params, const float* state, const float* obs,
                          const int32_t* actions, const float* logp_old, const float* adv,
                          const float* returns, const uint8_t* mask, const uint8_t* avail,
                          double clip, double ent_coef, float* grads_out, void* workspace, void* stream);
/* The same with the two options BASELINE.json's north_star names and the reference does NOT have (both
 * default off, cmarl_ppo_epoch_grads == this entry with value_clip <= 0 and the env range [0, n_envs)):
 *   minibatches   the sums run over the contiguous env block [env_begin, env_begin + env_count) only (envs are i.i.d.,
 *                 so a block is a random minibatch); the caller steps the optimizer once per block
 *   value_clip    > 0: critic loss max((V - R)^2, (V_old + clamp(V - V_old, -c, c) - R)^2) with V_old = values_old
 *                 f32 [T][V][B], the critic's values at rollout time (cmarl_critic_values); <= 0: MME:554-558 as is */
int cmarl_ppo_epoch_grads_ex(cmarl_ctx* ctx, const float* params, const float* state, const float* obs,
                             const int32_t* actions, const float* logp_old, const float* adv,
                             const float* returns, const float* values_old, const uint8_t* mask,
                             const uint8_t* avail, double clip, double ent_coef, double value_clip,
                             int32_t env_begin, int32_t env_count, float* grads_out, void* workspace, void* stream);

/* -- K8: grad scaling, grad norms, optional clipping and the Adam step for both networks
 * (MME:584-594; norm_d MME:221-224; torch.optim.Adam single-tensor defaults amsgrad=False, wd=0).
 *   grads  the (all-reduced) output of cmarl_ppo_epoch_grads
 *   step   1-based Adam step count (used when step_dev == NULL); step_dev: optional device
 *          counter holding the number of steps taken so far -- the kernel uses *step_dev + 1 and
 *          increments it, so the launch is CUDA-graph replayable;  max_norm <= 0 => no clipping
 *          (MME:62-63, 586).  Hyper-parameters are doubles because the reference derives the
 *          bias corrections in Python floats before rounding to fp32 (torch/optim/adam.py).
 *   stats_out f32 [8]: actor_loss, critic_loss, entropy, kl, clip_frac, actor_grad_norm,
 *                      critic_grad_norm, n_valid  -- the scalars logged at MME:597-612 */
int cmarl_clip_adam_step(cmarl_ctx* ctx, float* params, const float* grads, float* exp_avg,
                         float* exp_avg_sq, int32_t step, int32_t* step_dev, double lr_actor,
                         double lr_critic, double beta1, double beta2, double eps, double max_norm,
                         float* stats_out, void* stream);
/* The fixed-order reduction of cmarl_ppo_epoch_grads and this step in ONE launch: call cmarl_ppo_epoch_grads_ex with
 * grads_out = NULL (the chain kernels' per-CTA partial rows then stay in `workspace`), then this entry with the same
 * workspace; grads_out receives the reduced sums (the same values cmarl_ppo_epoch_grads would have written) and the step
 * follows as in cmarl_clip_adam_step.  With a peer-memory exchange attached the sums go to the peers as they are formed.
 * Fused kernels only (default shapes, MLP actor). */
int cmarl_reduce_clip_adam_step(cmarl_ctx* ctx, const void* workspace, float* params, float* grads_out, float* exp_avg,
                                float* exp_avg_sq, int32_t step, int32_t* step_dev, double lr_actor, double lr_critic,
                                double beta1, double beta2, double eps, double max_norm, float* stats_out, void* stream);


/* ======================================================================================================
 * Recurrent-actor path (BASELINE config 4): cleanmarl/mappo_lstm_multienvs.py ("LSTM" below).
 * cmarl_rollout, cmarl_critic_values, cmarl_td_lambda and cmarl_normalize serve it unchanged (a context
 * created with actor_recurrent = 1 runs the GRU actor inside cmarl_rollout, hidden state 0 at episode start,
 * LSTM:406); the entries below replace what differs.
 * ====================================================================================================== */

/* -- K2 (recurrent) alone: Actor.act(x, h, avail) LSTM:170-184 for one time step; used by the parity tests and
 * the eval loop (LSTM:683-705).  obs [N][O][B]; h_in [N][H][B] or NULL (= zeros, LSTM:178-179); avail u8
 * [N][A][B] or NULL; noise [N][A][B]; h_out [N][H][B] (may alias h_in); logits_out [N][A][B] optional. */
int cmarl_actor_act_recurrent(cmarl_ctx* ctx, const float* actor_params, const float* obs, const float* h_in,
                              const uint8_t* avail, const float* noise, int32_t* actions, float* logp,
                              float* logits_out, float* h_out, void* stream);

/* -- K7a: actor loss + gradients of ONE truncated-BPTT chunk, steps [t0, t1) (LSTM:563-607): forward through the
 * chunk from the detached hidden state h_seq[t0], per-step clipped-PPO / entropy head (LSTM:574-593), backward
 * through time inside the chunk.  Output = UNNORMALISED sums so that shards add:
 *   grads_out f32 [Pa + CMARL_N_STATS]:  d/d(actor params) of sum_{t in chunk} sum_b mean_n(-min(A r, A clamp r) - ent H),
 *   stats [0] that loss sum [2] entropy sum [3] kl sum [4] clip-fraction sum [5] valid (b,t) in the chunk, others 0.
 * The reference's division by (n_valid_chunk * T_chunk) (LSTM:605-607) happens in cmarl_adam_step_net after the
 * caller's all-reduce.  h_seq f32 [T+1][N][H][B] is caller-owned scratch that carries the hidden state from chunk
 * to chunk inside one epoch: t0 == 0 starts from zeros (LSTM:558, the kernel ignores slice 0); on return slices
 * t0+1..t1 hold the hidden states after each step (computed with the weights of THIS call, LSTM:620 detach).
 * stash: optional caller-owned scratch f32 [T][N][5H][B] (x1 and the gates r, z, n, W_hn h + b_hn of every step).
 *   With a stash (and the tensor cores on, the default) the chunk runs as the two tcgen05 kernels of csrc/tc_gru.cu --
 *   forward (hidden states, stash, head: dlogits into `workspace`) and backward through time -- which meet through h_seq,
 *   the stash and the workspace (cmarl_workspace_bytes of a recurrent context covers the dlogits buffer f32 [T][N][8][B]).
 *   Without one, or with cmarl_ctx_set_tensor_cores(ctx, 0), the fp32 FFMA kernel of csrc/gru.cu runs (one launch; NULL
 *   selects its recompute variant, which gives the same results as its stash variant bit for bit).  The two
 *   implementations agree to the tolerance both are held to against the reference (gradients <= 2e-5 of the chunk
 *   gradient's max, hidden states <= 2e-6). */
int cmarl_tbptt_chunk_grads(cmarl_ctx* ctx, const float* actor_params, const float* state, const float* obs,
                            const int32_t* actions, const float* logp_old, const float* adv, const uint8_t* mask,
                            const uint8_t* avail, double clip, double ent_coef, int32_t t0, int32_t t1,
                            float* h_seq, float* stash, float* grads_out, void* workspace, void* stream);

/* -- K7b: critic loss + gradients of one epoch alone (LSTM:621-626, 646-649; the critic is stepped once per
 * epoch while the actor is stepped once per chunk).  grads_out f32 [Pc + CMARL_N_STATS]: unnormalised sums,
 * stats [1] critic loss sum, [5] valid (b,t). */
int cmarl_critic_epoch_grads(cmarl_ctx* ctx, const float* critic_params, const float* state, const float* obs,
                             const float* returns, const uint8_t* mask, float* grads_out, void* workspace,
                             void* stream);

/* -- K8 for ONE network (net 0 = actor, 1 = critic): g = grads / (stats[5] * extra_div), norm of per-tensor norms,
 * optional clip, Adam (LSTM:605-619 with extra_div = T_chunk; LSTM:646-655 with extra_div = 1).  params/exp_avg/
 * exp_avg_sq are that network's own flat vectors; step_dev as in cmarl_clip_adam_step.
 *   stats_out f32 [8]: the five input sums [0..4] UNdivided, [5] this step's gradient norm, [6] valid count, [7] 0
 *   (the host adds the chunk sums and divides by b_mask.sum(), LSTM:640-644). */
int cmarl_adam_step_net(cmarl_ctx* ctx, int32_t net, float* params, const float* grads, float* exp_avg,
                        float* exp_avg_sq, int32_t step, int32_t* step_dev, double lr, double beta1, double beta2,
                        double eps, double max_norm, double extra_div, float* stats_out, void* stream);

/* ======================================================================================================
 * C1: gradient exchange across the GPUs of one box (one process per GPU), over peer memory.
 * The reference has no multi-GPU path; SURVEY 8e shards the envs and sums the unnormalised gradient sums once per
 * epoch.  Instead of a separate collective between K7 and K8, the Adam kernel itself exchanges the 38.7 KB over
 * NVLink: every rank pushes its sums into the receive buffers of all peers (blocks mapped through CUDA IPC), flags
 * the peers, waits for their flags and adds all ranks' rows in rank order -- so every rank forms bit-identical totals, and a multi-GPU
 * iteration stays a fixed launch sequence (CUDA-graph replayable, no NCCL call in the loop).
 *   cmarl_comm_create   allocates this rank's block (the only device allocation the library makes: IPC export needs a
 *                       whole cudaMalloc block) and returns its 64-byte cudaIpcMemHandle_t
 *   cmarl_comm_attach   handles = [world][64] bytes, every rank's handle in rank order (exchanged by the caller, e.g.
 *                       torch.distributed.all_gather_object); from then on cmarl_clip_adam_step / cmarl_adam_step_net
 *                       treat `grads` as this rank's LOCAL sums and step with the global ones
 *   cmarl_comm_detach   unmaps / frees (also done by cmarl_ctx_destroy)
 * All ranks must issue the same sequence of Adam calls.  A peer that never arrives traps the kernel after about a minute
 * of polling (the call then reports a launch failure) instead of hanging the GPU. */
size_t cmarl_comm_bytes(void);
int cmarl_comm_create(cmarl_ctx* ctx, uint8_t* handle_out /* HOST [64] */);
int cmarl_comm_attach(cmarl_ctx* ctx, int32_t rank, int32_t world, const uint8_t* handles /* HOST [world][64] */);
int cmarl_comm_detach(cmarl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CMARL_B200_H */
