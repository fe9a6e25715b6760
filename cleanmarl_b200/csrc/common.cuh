// Shared host/device helpers for libcmarl_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cmarl_b200.h"

#define CMARL_MAX_AGENTS 8
#define CMARL_MAX_ACTIONS 8

// Offsets of one 2-hidden-layer MLP inside the flat parameter vector (torch parameters() order).
struct NetLayout {
    int in, hid, out;
    int w1, b1, w2, b2, w3, b3, count;
    __host__ __device__ void set(int in_, int hid_, int out_) {
        in = in_; hid = hid_; out = out_;
        w1 = 0;
        b1 = w1 + hid * in;
        w2 = b1 + hid;
        b2 = w2 + hid * hid;
        w3 = b2 + hid;
        b3 = w3 + out * hid;
        count = b3 + out;
    }
};

// Offsets of the recurrent actor (fc1 + GRUCell + fc2, mappo_lstm_multienvs.py:162-184) in torch parameters() order.
struct GruLayout {
    int in, hid, out;
    int w1, b1, wih, whh, bih, bhh, w2, b2, count;
    __host__ __device__ void set(int in_, int hid_, int out_) {
        in = in_; hid = hid_; out = out_;
        w1 = 0;
        b1 = w1 + hid * in;
        wih = b1 + hid;
        whh = wih + 3 * hid * hid;
        bih = whh + 3 * hid * hid;
        bhh = bih + 3 * hid;
        w2 = bhh + 3 * hid;
        b2 = w2 + out * hid;
        count = b2 + out;
    }
};

// ------------------------------------------------------------------------------------------
// Peer-memory gradient exchange (multi-GPU, one process per GPU): every rank owns one cudaMalloc'ed, IPC-exported
// block of CMARL_COMM_CHANNELS channels; peers map it (cudaIpcOpenMemHandle) and the Adam kernel reads the other
// ranks' gradient sums that the peers pushed into it over NVLink (comm.cu, exact.cu).
// ------------------------------------------------------------------------------------------
constexpr int CMARL_MAX_RANKS = 8;
constexpr int CMARL_COMM_CHANNELS = 2;            // 0: combined / actor steps, 1: critic steps of the per-network path
constexpr int CMARL_COMM_SLOT_FLOATS = 16384;     // >= P + CMARL_N_STATS
struct CommChannel {
    // RECEIVE buffers: slots[parity][r][i] is written by rank r as ONE 64-bit word {flag = call number + 1, fp32 value}
    // (posted NVLink stores; data and flag arrive together, like NCCL's LL protocol) and polled locally
    unsigned long long slots[2][CMARL_MAX_RANKS][CMARL_COMM_SLOT_FLOATS];
    unsigned long long seq;                                 // local: exchanges completed on this channel
    unsigned int ticket_pub;                                // local: CTAs that have read seq
    unsigned int pad[29];
};
struct cmarl_comm {
    int rank, world;                                        // world <= 1: no exchange
    CommChannel* base[CMARL_MAX_RANKS];                     // base[r] = rank r's block mapped into this process
    void* own;                                              // this rank's allocation (cudaFree at detach)
};

// Generic ("layered") networks: any number of hidden layers / hidden width / agent count the reference's CLI accepts
// beyond the shapes the fused kernels are built for (csrc/generic.cu).  n_lin Linear layers, dims[0] -> ... -> dims[n_lin];
// parameters in torch parameters() order (W_l [dims[l+1]][dims[l]] then b_l).
constexpr int CMARL_GEN_MAX_LIN = 8;        // hidden layers + 2 (MME:160-171: num_layer hidden->hidden blocks)
constexpr int CMARL_GEN_MAX_DIM = 256;      // widest layer
struct GenNet {
    int n_lin;
    int dims[CMARL_GEN_MAX_LIN + 1];
    int w_off[CMARL_GEN_MAX_LIN], b_off[CMARL_GEN_MAX_LIN];
    int count;
    void set(int in, int hid, int hidden_layers, int out) {
        n_lin = hidden_layers + 2;
        dims[0] = in;
        for (int l = 1; l < n_lin; ++l) dims[l] = hid;
        dims[n_lin] = out;
        int o = 0;
        for (int l = 0; l < n_lin; ++l) {
            w_off[l] = o; o += dims[l + 1] * dims[l];
            b_off[l] = o; o += dims[l + 1];
        }
        count = o;
    }
};

enum KernelId { K_RESET = 0, K_ENVSTEP, K_ROLLOUT, K_ACT, K_CRITIC, K_TD, K_NORM, K_PPO_ACTOR, K_PPO_CRITIC,
                K_PPO_REDUCE, K_ADAM, K_TBPTT };
static_assert(K_TBPTT + 1 == CMARL_NK, "kernel id table");
constexpr int CMARL_TIMING_POOL = 256;

struct cmarl_timing {
    cudaEvent_t ev[CMARL_NK][CMARL_TIMING_POOL][2];
    int n[CMARL_NK];
    double sum_ms[CMARL_NK];
    long long count[CMARL_NK];
};

struct cmarl_ctx {
    cmarl_config cfg;
    NetLayout actor, critic;
    GruLayout gru;      // recurrent actor (cfg.actor_recurrent): then actor.count == gru.count
    int generic;        // 1: shapes outside the fused kernels' set -> every entry runs the layered kernels of generic.cu
    int n_landmarks;    // L (= n_agents in simple_spread_v3)
    int raw_obs;        // R = 4 + 2 L + 4 (N - 1): vel, pos, landmarks - pos, others - pos, 2 silent comm slots per other
    GenNet gactor, gcritic;
    int n_heads;        // V
    int critic_in;      // S (MAPPO) or O (IPPO)
    int sm_count;
    int ppo_grid_actor, ppo_grid_critic;
    int launches;
    int pending_grid_a, pending_grid_c;   // partial rows left in the workspace by cmarl_ppo_epoch_grads_ex(grads_out = NULL)
    int timing_on;
    int use_tc;         // 1: tcgen05 (3xTF32) chain kernels, 0: fp32 FFMA chain kernels
    cmarl_comm comm;         // peer-memory gradient exchange (world <= 1: off)
    double weight_decay[2];  // actor, critic: decoupled weight decay of the Adam entries (AdamW); 0 = plain Adam
    uint64_t* episode_dev;   // optional device episode counter for the Philox draws (CUDA-graph replay)
    int launch_chaining;     // 1: launches carry the programmatic-stream-serialization attribute (cmarl_ctx_set_launch_chaining)
    cmarl_timing* timing;
    void* gen_ws;            // generic mode: scratch block of the layered kernels (activations, partials), owned by the context
    float* dev_floats;       // CMARL_DEV_FLOATS device floats owned by the context (generic Adam: per-tensor sums of squares)
    unsigned int* dev_words; // CMARL_DEV_WORDS zero-initialised device words owned by the context (tickets of the kernels' last-CTA protocols)
};
enum { CMARL_DW_ADAM_TICKET = 0, CMARL_DW_ADAM_BARRIER = 1 /* 4 words: count, generation of the Adam CTAs' barrier; count, generation of the fused kernel's */, CMARL_DEV_WORDS = 64 };
enum { CMARL_DF_GEN_TSQ = 0 /* 64 floats: generic Adam, per-tensor sums of squares */, CMARL_DF_ADAM_TSQ = 64 /* 32 x 12 */,
       CMARL_DEV_FLOATS = 512 };
constexpr int CMARL_MAX_PARAMS = 16384;     // per context on the fused path (clip_adam_kernel: <= 16 co-resident CTAs x 1024 threads; = CMARL_COMM_SLOT_FLOATS)

void cmarl_time_begin(cmarl_ctx* ctx, int id, cudaStream_t st);
void cmarl_time_end(cmarl_ctx* ctx, int id, cudaStream_t st);

// RAII bracket used at every launch site
struct KernelTimer {
    cmarl_ctx* ctx; int id; cudaStream_t st;
    KernelTimer(cmarl_ctx* c, int i, cudaStream_t s) : ctx(c), id(i), st(s) { ctx->launches++; if (ctx->timing_on) cmarl_time_begin(ctx, id, st); }
    ~KernelTimer() { if (ctx->timing_on) cmarl_time_end(ctx, id, st); }
};

void cmarl_set_error(const char* fmt, ...);
int cmarl_check_cuda(cudaError_t e, const char* what);

#define CMARL_CUDA(call)                                         \
    do {                                                         \
        int _e = cmarl_check_cuda((call), #call);                \
        if (_e) return _e;                                       \
    } while (0)

#define CMARL_ARG(cond, msg)                                     \
    do {                                                         \
        if (!(cond)) {                                           \
            cmarl_set_error("%s: %s", __func__, msg);            \
            return -1;                                           \
        }                                                        \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (launch chaining, cmarl_ctx_set_launch_chaining).  Every kernel of the library begins
// with pdl_wait_then_trigger(): `griddepcontrol.wait` blocks until the grid in front of it on the stream has completed
// and flushed (immediately when the launch carried no programmatic attribute), `griddepcontrol.launch_dependents`
// then lets the NEXT launch become resident.  Waiting BEFORE triggering makes completion transitive along the chain:
// by the time a kernel's dependent starts, everything in front of that kernel has completed, and the dependent's own
// wait covers the kernel itself -- so only launch latency and pre-wait prologues (shared-memory zeroing, barrier
// init, TMEM allocation) overlap, never a read with the write it depends on.
// ------------------------------------------------------------------------------------------
// NON-COHERENT LOADS ARE NOT ALLOWED on anything a predecessor kernel writes (parameters, rollout buffers, partial rows,
// gradients): a programmatic dependent becomes resident while its primary still runs, so its lifetime overlaps the writes,
// and ld.global.nc (`__ldg`, or what the compiler infers from `const T* __restrict__`) may then return a line the SM cached
// during an EARLIER launch -- griddepcontrol.wait orders coherent loads only.  Measured round 2: the actor chain of epoch 2,
// launched as a dependent of the Adam step of epoch 1, read parameters of epoch 0 from the read-only cache (gradient
// sums off by 1e-4 relative; profiles/tools/chain_diag2.py); the same mechanism was behind round 1's unexplained loss of
// bit-identity.  Such data is read with `__ldcg` (L2, coherent); it is streamed once per kernel, so nothing is lost.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_then_trigger() { pdl_wait(); pdl_trigger(); }

// `for (i = threadIdx.x; i < n; i += nt)` with a compile-time trip count, fully unrolled and predicated: straight-line
// code, so the global loads of ALL iterations are in flight together.  (As a plain loop the compiler keeps one load in
// flight per iteration: a kernel prologue that copies a few thousand parameters then costs one L2 round trip per
// iteration -- measured ~12 us in front of the critic chain's first tile.)
#define CMARL_STRIDED(i, n, nt)                                                                                      \
    _Pragma("unroll") for (int i##_r = 0, i = threadIdx.x; i##_r < ((n) + (nt) - 1) / (nt); ++i##_r, i += (nt)) if (i < (n))

// Launch `kernel` on `st`, as a programmatic dependent of the launch in front of it when `pdl` is set.
template <class... KArgs, class... Args>
static inline cudaError_t cmarl_launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                           cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// ... while launch chaining is on (and the per-kernel event timing is off: the bracketing events would separate the
// launches anyway)
static inline bool cmarl_chained(const cmarl_ctx* ctx) { return ctx && ctx->launch_chaining && !ctx->timing_on; }
template <class... KArgs, class... Args>
static inline cudaError_t cmarl_launch(const cmarl_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t st, Args... args) {
    return cmarl_launch_pdl(cmarl_chained(ctx), kernel, grid, block, smem, st, args...);
}
#endif

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011) -- counter-based, so a draw is a pure function of
// (seed, episode, t, agent-row, lane) and shards/replays reproduce it.
// ------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ inline void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c0, hi0, lo0);
        philox_mulhilo(M1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// uniform in (0,1]: never 0 so -log(u) is finite
__host__ __device__ inline float u32_to_unit_open0(uint32_t v) { return ((float)(v >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// uniform double in [0,1) from 53 random bits
__host__ __device__ inline double u64_to_unit(uint32_t hi, uint32_t lo) {
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}
