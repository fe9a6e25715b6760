// K5 TD(lambda) scan, K6 normalisation, K8 clip + Adam.
// Compiled with -fmad=false: these kernels reproduce the reference's fp32 operation order
// (separate ATen ops => separately rounded mul/add), FMA only where ATen itself fuses (lerp).
#include "common.cuh"
#include "reduce.cuh"

// ---------------------------------------------------------------------------------------- K5
// One thread per (value head v, env b); sequential in t (the recurrence is order-sensitive in
// fp32: SURVEY.md 0.4), parallel and coalesced over b.  Algorithmic traffic: r 4 B + V 4 B in,
// R 4 B + A 4 B out per (t, v, b) -- 16 B per env-step for MAPPO (V = 1).
template <int UNROLL>
__global__ void __launch_bounds__(256) td_lambda_kernel(const float* __restrict__ values,
                                                        const float* __restrict__ reward,
                                                        const uint8_t* __restrict__ mask,
                                                        float* __restrict__ returns,
                                                        float* __restrict__ adv,
                                                        int T, int V, int B, float g, float l, float oml) {
    pdl_wait_then_trigger();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= V * B) return;
    const int v = idx / B, b = idx - v * B;
    const size_t strideT = (size_t)V * B;
    const float* vp = values + (size_t)v * B + b;
    float* rp = returns + (size_t)v * B + b;
    float* ap = adv + (size_t)v * B + b;
    float last = 0.0f, vnext = 0.0f;
    bool next_live = false;
    int t = T - 1;
    // main part: UNROLL steps at a time, all loads of a chunk issued before the dependent chain
    for (; t >= UNROLL - 1; t -= UNROLL) {
        float vv[UNROLL], rr[UNROLL];
        uint8_t mm[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int tt = t - u;
            vv[u] = __ldcg(vp + (size_t)tt * strideT);
            rr[u] = __ldcg(reward + (size_t)tt * B + b);
            mm[u] = mask ? __ldcg(mask + (size_t)tt * B + b) : (uint8_t)1;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int tt = t - u;
            const bool live = mm[u] != 0;
            float R = 0.0f, A = 0.0f;
            if (live) {
                const float nv = next_live ? vnext : 0.0f;
                R = rr[u] + g * (l * last + oml * nv);
                A = R - vv[u];
                last = R;
            }
            __stcs(rp + (size_t)tt * strideT, R);
            __stcs(ap + (size_t)tt * strideT, A);
            next_live = live;
            vnext = vv[u];
        }
    }
    for (; t >= 0; --t) {
        const float vt = __ldcg(vp + (size_t)t * strideT);
        const float rt = __ldcg(reward + (size_t)t * B + b);
        const bool live = mask ? (__ldcg(mask + (size_t)t * B + b) != 0) : true;
        float R = 0.0f, A = 0.0f;
        if (live) {
            const float nv = next_live ? vnext : 0.0f;
            R = rt + g * (l * last + oml * nv);
            A = R - vt;
            last = R;
        }
        __stcs(rp + (size_t)t * strideT, R);
        __stcs(ap + (size_t)t * strideT, A);
        next_live = live;
        vnext = vt;
    }
}

extern "C" int cmarl_td_lambda(cmarl_ctx* ctx, const float* values, const float* reward, const uint8_t* mask,
                               double gamma, double lambda, float* returns, float* adv, void* stream) {
    CMARL_ARG(ctx && values && reward && returns && adv, "null argument");
    const int T = ctx->cfg.n_steps, V = ctx->n_heads, B = ctx->cfg.n_envs;
    const int n = V * B;
    // python-float coefficients rounded to fp32 when they meet an fp32 tensor (MME:496-501)
    const float g = (float)gamma, l = (float)lambda, oml = (float)(1.0 - lambda);
    {
        KernelTimer kt(ctx, K_TD, as_stream(stream));
        // Small batches (the trainer's 4 096 envs) are a latency chain: all T loads of a thread in flight at once (chunks of 25
        // steps) and 64-thread CTAs on every SM instead of five chunk round trips on 16 SMs (10.7 -> ~5 us); the large stand-alone
        // scan is HBM-bound and keeps the 5-step chunks in 256-thread CTAs.  Same arithmetic in the same order either way.
        if (n <= 65536 && T >= 25)
            CMARL_CUDA(cmarl_launch(ctx, td_lambda_kernel<25>, dim3(ceil_div(n, 64)), dim3(64), 0, as_stream(stream), values, reward,
                                    mask, returns, adv, T, V, B, g, l, oml));
        else
            CMARL_CUDA(cmarl_launch(ctx, td_lambda_kernel<5>, dim3(ceil_div(n, 256)), dim3(256), 0, as_stream(stream), values, reward,
                                    mask, returns, adv, T, V, B, g, l, oml));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------- K6
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// phase 0: (sum, sumsq, count) of the head-mean over masked (t,b), accumulated in fp64
__global__ void __launch_bounds__(256) norm_stats_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                                         int T, int V, int B, double* __restrict__ stats) {
    double s = 0.0, ss = 0.0, c = 0.0;
    const size_t n = (size_t)T * B;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        const size_t t = i / B, b = i - t * B;
        float m = 0.0f;
        for (int v = 0; v < V; ++v) m += x[(t * V + v) * B + b];
        m = m / (float)V;
        s += (double)m;
        ss += (double)m * (double)m;
        c += 1.0;
    }
    __shared__ double sh[3][8];
    s = warp_sum(s); ss = warp_sum(ss); c = warp_sum(c);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][w] = s; sh[1][w] = ss; sh[2][w] = c; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += sh[threadIdx.x][i];
        atomicAdd(&stats[threadIdx.x], a);
    }
}

// phase 1: apply.  mode 0 (reward): masked entries only, std + 1e-6; mode 1: every entry, no eps.
__global__ void __launch_bounds__(256) norm_apply_kernel(float* __restrict__ x, const uint8_t* __restrict__ mask,
                                                         int T, int V, int B, int mode,
                                                         const double* __restrict__ stats) {
    const double n = stats[2];
    const double mean = stats[0] / n;
    double var = (stats[1] - stats[0] * mean) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean;
    const float sd = (float)sqrt(var);
    const float den = mode == 0 ? sd + 1e-6f : sd;
    const size_t total = (size_t)T * V * B;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (mode == 0 && mask) {
            const size_t tv = i / B, b = i - tv * B, t = tv / V;
            if (!mask[t * B + b]) continue;
        }
        x[i] = (x[i] - mu) / den;
    }
}

extern "C" int cmarl_normalize(cmarl_ctx* ctx, float* x, int32_t n_heads, const uint8_t* mask, int32_t mode,
                               int32_t phase, double* stats_io, void* stream) {
    CMARL_ARG(ctx && x && stats_io, "null argument");
    CMARL_ARG(n_heads >= 1 && (mode == 0 || mode == 1) && (phase == 0 || phase == 1), "bad mode/phase/heads");
    const int T = ctx->cfg.n_steps, B = ctx->cfg.n_envs;
    cudaStream_t st = as_stream(stream);
    KernelTimer kt(ctx, K_NORM, st);
    if (phase == 0) {
        CMARL_CUDA(cudaMemsetAsync(stats_io, 0, 4 * sizeof(double), st));
        int grid = ceil_div(T * B, 256);
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        norm_stats_kernel<<<grid, 256, 0, st>>>(x, mask, T, n_heads, B, stats_io);
    } else {
        int grid = ceil_div(T * n_heads * B, 256);
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        norm_apply_kernel<<<grid, 256, 0, st>>>(x, mask, T, n_heads, B, mode, stats_io);
    }
    return cmarl_check_cuda(cudaGetLastError(), "normalize kernel");
}

// ---------------------------------------------------------------------------------------- K8
struct AdamArgs {
    float* params;
    const float* grads;     // [P + 8] unnormalised sums (+stats)
    float* m;
    float* v;
    float* stats_out;
    int32_t* step_dev;
    unsigned int* ticket;   // per-context device word: CTAs that have read *step_dev (launches of one context are serialised)
    unsigned int* barrier;  // per-context device words [2]: arrivals at the kernel's grid barrier, barrier generation
    float* tsq_part;        // per-context scratch [ADAM_MAX_CTAS][12]: per-CTA sums of squares per tensor
    int step;
    int n_tensors;          // 12
    int tensor_off[13];     // prefix offsets of the 12 parameter tensors, [12] = P
    int n_actor_tensors;    // 6
    double lr[2], beta1, beta2, eps, max_norm;
    double wd[2];           // decoupled weight decay (torch.optim.AdamW: param.mul_(1 - lr * weight_decay)); 0 = Adam
    float extra_div;        // grads are divided by stats[5] * extra_div (T_chunk of a truncated-BPTT chunk, else 1)
    int raw_stats;          // 1: stats_out = the five sums undivided, this step's norm, the valid count (single-net mode)
    CommChannel* ch[CMARL_MAX_RANKS];   // peer-memory exchange: channel of every rank (own included), rank order
    int rank, world;        // world <= 1: `grads` already holds the global sums
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_probe_sys(const unsigned long long* p) {
    unsigned long long v;
    asm("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ceil(P / 1024) CTAs of 1024 threads, one parameter per thread.  The norm of the per-tensor norms (norm_d, MME:221-224)
// needs every gradient: each CTA reduces the squares of ITS 1 024 elements per tensor (a warp whose 32 consecutive elements
// lie in one tensor -- all but <= 11 warp-rows -- adds one shuffle-reduced value), publishes <= 12 partial sums, and after a
// grid barrier (the <= 16 CTAs are co-resident) every CTA adds the partials of all CTAs in CTA order: bit-identical norms and
// clip coefficients in every CTA -- and on every rank of a multi-GPU run.  (Round 1 let every CTA reduce ALL gradients
// redundantly: ~3 600 instructions per warp, 12.8 us per launch; this form is one gradient per thread.)
constexpr int ADAM_THREADS = 1024;
constexpr int ADAM_MAX_CTAS = 32;        // = CMARL_MAX_PARAMS / 512 (the fused kernel's CTAs hold 512 parameters); all co-resident (grid barrier)
constexpr int ADAM_MAX_TENSORS = 12;

__device__ __forceinline__ int tensor_of(const AdamArgs& a, int i) {
    int k = 0;
#pragma unroll
    for (int q = 1; q < ADAM_MAX_TENSORS; ++q) k += (i >= a.tensor_off[q]) ? 1 : 0;
    return k;
}

// `cta` of `ncta` Adam CTAs (the whole grid of clip_adam_kernel; the first CTAs of reduce_clip_adam_kernel).  `pushed`: this
// rank's sums are already on their way to the peers (posted by the reduction phase of the fused kernel).
// NT threads per CTA (1 024 stand-alone, 512 in the fused kernel).  The sums of squares are formed per 512 parameters (16
// warps in order) and combined pairwise, (h0 + h1) per 1 024 parameters and those in order, so that both forms give the
// same bits.
template <bool XCHG, int NT>
__device__ __forceinline__ void adam_body(const AdamArgs& a, const int cta, const int ncta, const bool pushed) {
    constexpr int ADAM_THREADS = NT;
    __shared__ float wsum[ADAM_THREADS / 32][ADAM_MAX_TENSORS];
    __shared__ float tnorm[ADAM_MAX_TENSORS];
    __shared__ double bc_sh[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.tensor_off[12];
    const int mine = cta * ADAM_THREADS + tid;                 // the parameter this thread updates
    constexpr bool xchg = XCHG;       // peer-memory exchange compiled in only for multi-GPU launches
    int par = 0;
    unsigned int tag = 0;
    if (xchg) {
        // ---- peer-memory all-reduce, fused (NCCL-LL style): every rank stores {call number + 1, value} words into its
        //      row of EVERY rank's receive buffer (posted writes over NVLink: one one-way latency, no fence, no separate
        //      flag), and every reader polls its LOCAL copy until the tag matches.  A pull over NVLink measured 46 us per
        //      exchange, push + fence + flags 19 us, NCCL's all-reduce 13 us.
        __shared__ unsigned long long seq_sh;
        CommChannel* me = a.ch[a.rank];
        if (tid == 0) seq_sh = *reinterpret_cast<volatile unsigned long long*>(&me->seq);
        __syncthreads();
        const unsigned long long seq = seq_sh;
        par = (int)(seq & 1ull);
        tag = (unsigned int)(seq + 1ull);
        if (tid == 0) {                          // every CTA has read seq once its ticket is in: the last one advances it
            const unsigned t = atomicAdd(&me->ticket_pub, 1u);
            if (t == (unsigned)ncta - 1) { me->ticket_pub = 0; me->seq = seq + 1; }
        }
        if (!pushed && mine < P) {
            const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(a.grads[mine]);
            for (int r = 0; r < a.world; ++r) st_relaxed_sys(&a.ch[r]->slots[par][a.rank][mine], w);
        }
        if (!pushed && cta == 0 && tid < CMARL_N_STATS) {
            const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(a.grads[P + tid]);
            for (int r = 0; r < a.world; ++r) st_relaxed_sys(&a.ch[r]->slots[par][a.rank][P + tid], w);
        }
    }
    // gradient sum i over the ranks in rank order (every rank computes the identical value); local when world <= 1
    auto G = [&](int i) -> float {
        if (!xchg) return __ldcg(a.grads + i);      // (fused kernel: written by other CTAs of this launch)
        const unsigned long long* row = &a.ch[a.rank]->slots[par][0][i];
        float v[CMARL_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < CMARL_MAX_RANKS; ++r) {
            v[r] = 0.0f;
            if (r < a.world) {
                // first probe: a schedulable (non-volatile) load, so the probes of one thread overlap; a probe that
                // comes back without the tag falls into the ordered polling loop
                unsigned long long w = ld_probe_sys(row + (size_t)r * CMARL_COMM_SLOT_FLOATS);
                unsigned spins = 0;
                while ((unsigned int)(w >> 32) != tag) {
                    if (++spins > (1u << 28)) __trap();          // ~1 min of polling: a missing peer must surface as an error, not as a hang
                    w = ld_relaxed_sys(row + (size_t)r * CMARL_COMM_SLOT_FLOATS);
                }
                v[r] = __uint_as_float((unsigned int)w);
            }
        }
        float sum = 0.0f;
#pragma unroll
        for (int r = 0; r < CMARL_MAX_RANKS; ++r) sum += v[r];      // rank order; absent ranks add +0
        return sum;
    };
    // every global load is issued up front (one L2 round trip)
    const float g_mine = mine < P ? G(mine) : 0.0f;
    const float pm = mine < P ? a.m[mine] : 0.0f;
    const float pv = mine < P ? a.v[mine] : 0.0f;
    const float pp = mine < P ? a.params[mine] : 0.0f;
    const float n_valid = G(P + 5);
    const float count = n_valid * a.extra_div;
    if (tid == ADAM_THREADS - 1) {      // a thread of the last warp: the first warps finish the reductions below
        // the step count is read by this ONE thread per CTA; the last CTA to have read it publishes the new value
        int step = a.step;
        if (a.step_dev) {
            step = *reinterpret_cast<volatile int32_t*>(a.step_dev) + 1;
            __threadfence();
            const unsigned t = atomicAdd(a.ticket, 1u);
            if (t == (unsigned)ncta - 1) { *a.ticket = 0; *a.step_dev = step; }
        }
        // beta^step by repeated squaring (<= 2 log2(step) fp64 multiplies, within a few ulp of pow(): no float32-visible
        // difference in bc1, sqrt(bc2) or lr / bc1 for step <= 10^6)
        double p1 = 1.0, p2 = 1.0, b1 = a.beta1, b2 = a.beta2;
        for (unsigned e = (unsigned)step; e; e >>= 1) {
            if (e & 1u) { p1 *= b1; p2 *= b2; }
            b1 *= b1; b2 *= b2;
        }
        bc_sh[0] = 1.0 - p1;
        bc_sh[1] = sqrt(1.0 - p2);
    }
    if (lane < ADAM_MAX_TENSORS) wsum[warp][lane] = 0.0f;
    __syncwarp();
    // sums of squares of g / count per tensor over this CTA's elements
    {
        const int i0 = cta * ADAM_THREADS + warp * 32;          // this warp's 32 consecutive elements
        if (i0 < P) {                                           // warp-uniform
            const float gs = g_mine / count;
            const float sq = gs * gs;                           // 0 beyond P (g = 0)
            const int k_lo = tensor_of(a, i0), k_hi = tensor_of(a, min(i0 + 31, P - 1));
            if (k_lo == k_hi) {
                const float t = warp_sum(sq);
                if (lane == 0) wsum[warp][k_lo] = t;
            } else {
                const int k = tensor_of(a, min(i0 + lane, P - 1));
                for (int q = k_lo; q <= k_hi; ++q) {            // a tensor boundary inside the warp: masked sums
                    const float t = warp_sum(k == q ? sq : 0.0f);
                    if (lane == 0) wsum[warp][q] = t;
                }
            }
        }
    }
    __syncthreads();
    constexpr int HALVES = NT / 512;                 // 512-parameter halves per CTA
    const int nhalf = HALVES * ncta;                 // partial sums over all CTAs: [half][tensor]
    if (tid < ADAM_MAX_TENSORS) {
        float h[HALVES];
#pragma unroll
        for (int q = 0; q < HALVES; ++q) {
            float s = 0.0f;
            for (int w = 16 * q; w < 16 * q + 16; ++w) s += wsum[w][tid];
            h[q] = s;
        }
        if (ncta > 1) {
#pragma unroll
            for (int q = 0; q < HALVES; ++q) a.tsq_part[(cta * HALVES + q) * ADAM_MAX_TENSORS + tid] = h[q];
        } else {
            tnorm[tid] = sqrtf(HALVES == 2 ? h[0] + h[HALVES - 1] : h[0]);       // one 1 024-parameter block (or less)
        }
    }
    if (ncta > 1) {
        // grid barrier (generation word + arrival count, both per context; the count is back at 0 when the launch ends, so
        // launches with different grid sizes can follow each other; launches of one context are serialised on its stream)
        __syncthreads();
        if (tid == 0) {
            volatile unsigned* gen = reinterpret_cast<volatile unsigned*>(a.barrier + 1);
            const unsigned g0 = *gen;
            __threadfence();
            if (atomicAdd(a.barrier, 1u) == (unsigned)ncta - 1) {
                *reinterpret_cast<volatile unsigned*>(a.barrier) = 0u;
                __threadfence();
                atomicAdd(a.barrier + 1, 1u);
            } else {
                unsigned spins = 0;
                while (*gen == g0)
                    if (++spins > (1u << 30)) __trap();     // a protocol bug must surface as a launch failure, never as a hang
            }
            __threadfence();
        }
        __syncthreads();
        if (tid < ADAM_MAX_TENSORS) {
            float s = 0.0f;
            for (int c = 0; c < nhalf; c += 2) {          // pairs of halves = 1 024-parameter blocks, in order
                const float h0 = *reinterpret_cast<volatile float*>(&a.tsq_part[c * ADAM_MAX_TENSORS + tid]);
                const float h1 = c + 1 < nhalf ? *reinterpret_cast<volatile float*>(&a.tsq_part[(c + 1) * ADAM_MAX_TENSORS + tid]) : 0.0f;
                s += h0 + h1;
            }
            tnorm[tid] = sqrtf(s);
        }
    }
    __syncthreads();
    float net_norm[2], coef[2];
#pragma unroll
    for (int net = 0; net < 2; ++net) {
        float s = 0.0f;
        const int k0 = net == 0 ? 0 : a.n_actor_tensors, k1 = net == 0 ? a.n_actor_tensors : a.n_tensors;
        for (int k = k0; k < k1; ++k) s += tnorm[k] * tnorm[k];
        net_norm[net] = sqrtf(s);
        coef[net] = 1.0f;
        if (a.max_norm > 0.0) {
            // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
            const float c = (float)a.max_norm / (net_norm[net] + 1e-6f);
            coef[net] = c < 1.0f ? c : 1.0f;
        }
    }
    // bias corrections in double, as torch's python-float arithmetic (_single_tensor_adam)
    const double bc1 = bc_sh[0];
    const float bc2_sqrt = (float)bc_sh[1];
    const float w1 = (float)(1.0 - a.beta1);
    const float b2 = (float)a.beta2, w2 = (float)(1.0 - a.beta2);
    const float eps = (float)a.eps;
    const int actor_end = a.tensor_off[a.n_actor_tensors];
    if (mine < P) {
        const int net = mine < actor_end ? 0 : 1;
        const float nss = (float)(-(a.lr[net] / bc1));
        float gi = g_mine / count;
        if (a.max_norm > 0.0) gi = gi * coef[net];
        float m = pm, v = pv;
        m = fmaf(w1, gi - m, m);                // exp_avg.lerp_(grad, 1 - beta1): ATen's lerp is an fma
        v = v * b2;                             // exp_avg_sq.mul_(beta2)
        v = v + (w2 * gi) * gi;                 //            .addcmul_(grad, grad, value=1 - beta2)
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        float pw = pp;
        if (a.wd[net] != 0.0) pw = pw * (float)(1.0 - a.lr[net] * a.wd[net]);   // AdamW: param.mul_(1 - lr * weight_decay)
        a.params[mine] = pw + (nss * m) / denom;   // param.addcdiv_(m, denom, value=-step_size)
        a.m[mine] = m;
        a.v[mine] = v;
    }
    if (cta == 0 && tid == 0) {
        if (a.stats_out && a.raw_stats) {
            for (int k = 0; k < 5; ++k) a.stats_out[k] = G(P + k);
            a.stats_out[5] = net_norm[0];
            a.stats_out[6] = n_valid;
            a.stats_out[7] = 0.0f;
        } else if (a.stats_out) {
            for (int k = 0; k < 5; ++k) a.stats_out[k] = G(P + k) / count;
            a.stats_out[5] = net_norm[0];
            a.stats_out[6] = net_norm[1];
            a.stats_out[7] = count;
        }
    }
}

template <bool XCHG>
__global__ void __launch_bounds__(ADAM_THREADS, 1) clip_adam_kernel(AdamArgs a) {
    pdl_wait_then_trigger();
    adam_body<XCHG, ADAM_THREADS>(a, blockIdx.x, gridDim.x, false);
}

// The fixed-order reduction of the chain kernels' per-CTA partial rows (reduce.cuh; the same per-column order as
// reduce_partials_kernel) and the Adam step in ONE launch: every CTA reduces 64 columns, posts them to the peers right away
// (multi-GPU) and arrives at a barrier; the first `n_adam` CTAs wait for all arrivals and run the Adam body, the others
// exit.  (Only the Adam CTAs ever spin, so the grid need not be co-resident as a whole.)  512-thread CTAs of 32 columns like
// reduce_partials_kernel (four per SM: the 303 CTAs of the default shapes are one wave); a first version with 1 024-thread
// CTAs of 64 columns (one per SM, 152 CTAs = two waves) was slower than the two separate launches (0.651 vs 0.638 ms).
constexpr int FUSED_THREADS = 512;
constexpr int FUSED_COLS = FUSED_THREADS / chain::RED_GROUPS;      // 32
template <bool XCHG>
__global__ void __launch_bounds__(FUSED_THREADS, 2) reduce_clip_adam_kernel(chain::ReduceArgs r, float* __restrict__ grads_out, AdamArgs a,
                                                                           int n_adam) {
    __shared__ unsigned long long seq_sh;
    __shared__ unsigned gen_sh;
    const int tid = threadIdx.x;
    pdl_wait_then_trigger();
    if (tid == 0) {
        gen_sh = *reinterpret_cast<volatile unsigned*>(a.barrier + 3);
        if (XCHG) seq_sh = *reinterpret_cast<volatile unsigned long long*>(&a.ch[a.rank]->seq);
    }
    bool have; int i;
    const float v = chain::reduce_column<FUSED_COLS>(r, blockIdx.x, &have, &i);      // (contains block barriers)
    if (have) {
        grads_out[i] = v;
        if (XCHG) {
            const unsigned long long seq = seq_sh;
            const int par = (int)(seq & 1ull);
            const unsigned long long w = ((unsigned long long)(unsigned int)(seq + 1ull) << 32) | (unsigned long long)__float_as_uint(v);
            for (int rk = 0; rk < a.world; ++rk) st_relaxed_sys(&a.ch[rk]->slots[par][a.rank][i], w);
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.barrier + 2, 1u) == gridDim.x - 1) {
            *reinterpret_cast<volatile unsigned*>(a.barrier + 2) = 0u;
            __threadfence();
            atomicAdd(a.barrier + 3, 1u);
        } else if ((int)blockIdx.x < n_adam) {
            volatile unsigned* gen = reinterpret_cast<volatile unsigned*>(a.barrier + 3);
            unsigned spins = 0;
            while (*gen == gen_sh)
                if (++spins > (1u << 30)) __trap();
        }
        __threadfence();
    }
    if ((int)blockIdx.x >= n_adam) return;
    __syncthreads();
    adam_body<XCHG, FUSED_THREADS>(a, blockIdx.x, n_adam, true);
}

static void fill_comm(const cmarl_ctx* ctx, AdamArgs& a, int channel) {
    a.rank = ctx->comm.rank;
    a.world = ctx->comm.world;
    for (int r = 0; r < CMARL_MAX_RANKS; ++r) a.ch[r] = (r < ctx->comm.world && ctx->comm.base[r]) ? ctx->comm.base[r] + channel : nullptr;
}

// generic.cu
int cmarl_gen_clip_adam_step(cmarl_ctx* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t step,
                             int32_t* step_dev, double lr_actor, double lr_critic, double beta1, double beta2, double eps,
                             double max_norm, float* stats_out, cudaStream_t st);

static int clip_adam_impl(cmarl_ctx* ctx, const void* fused_workspace, float* params, const float* grads, float* grads_rw, float* exp_avg,
                          float* exp_avg_sq, int32_t step, int32_t* step_dev, double lr_actor,
                          double lr_critic, double beta1, double beta2, double eps, double max_norm,
                          float* stats_out, void* stream) {
    CMARL_ARG(ctx && params && grads && exp_avg && exp_avg_sq, "null argument");
    CMARL_ARG(step_dev || step >= 1, "step must be >= 1");
    CMARL_ARG(!ctx->cfg.actor_recurrent, "recurrent actor: use cmarl_adam_step_net");
    if (ctx->generic)
        return cmarl_gen_clip_adam_step(ctx, params, grads, exp_avg, exp_avg_sq, step, step_dev, lr_actor, lr_critic, beta1, beta2, eps,
                                        max_norm, stats_out, as_stream(stream));
    CMARL_ARG(ctx->actor.count + ctx->critic.count <= CMARL_MAX_PARAMS, "too many parameters for clip_adam_kernel");
    AdamArgs a;
    a.params = params; a.grads = grads; a.m = exp_avg; a.v = exp_avg_sq; a.stats_out = stats_out;
    a.step_dev = step_dev; a.step = step; a.ticket = ctx->dev_words + CMARL_DW_ADAM_TICKET;
    a.barrier = ctx->dev_words + CMARL_DW_ADAM_BARRIER; a.tsq_part = ctx->dev_floats + CMARL_DF_ADAM_TSQ;
    a.n_tensors = 12; a.n_actor_tensors = 6;
    const NetLayout* nets[2] = {&ctx->actor, &ctx->critic};
    int base = 0, k = 0;
    for (int n = 0; n < 2; ++n) {
        const NetLayout& L = *nets[n];
        const int offs[6] = {L.w1, L.b1, L.w2, L.b2, L.w3, L.b3};
        for (int j = 0; j < 6; ++j) a.tensor_off[k++] = base + offs[j];
        base += L.count;
    }
    a.tensor_off[12] = base;
    a.lr[0] = lr_actor; a.lr[1] = lr_critic; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
    a.extra_div = 1.0f; a.raw_stats = 0;
    a.wd[0] = ctx->weight_decay[0]; a.wd[1] = ctx->weight_decay[1];
    fill_comm(ctx, a, 0);
    const dim3 grid(ceil_div(a.tensor_off[12], ADAM_THREADS)), block(ADAM_THREADS);
    if (fused_workspace) {
        const dim3 fblock(FUSED_THREADS);
        const int n_adam = ceil_div(a.tensor_off[12], FUSED_THREADS);
        // reduction of the partial rows cmarl_ppo_epoch_grads_ex(grads_out = NULL) left in the workspace + Adam, one launch
        const int Pa = ctx->actor.count, Pc = ctx->critic.count;
        chain::ReduceArgs r;
        r.pa = reinterpret_cast<const float*>(fused_workspace); r.grid_a = ctx->pending_grid_a; r.Pa = Pa;
        r.pc = r.pa + (size_t)2 * ctx->sm_count * (Pa + CMARL_N_STATS); r.grid_c = ctx->pending_grid_c; r.Pc = Pc;
        r.n_groups = (float)ctx->cfg.n_agents; r.count_from_c = 0;
        const dim3 fgrid(ceil_div(Pa + Pc + CMARL_N_STATS, FUSED_COLS));
        KernelTimer kt(ctx, K_ADAM, as_stream(stream));
        CMARL_CUDA(a.world > 1 ? cmarl_launch(ctx, reduce_clip_adam_kernel<true>, fgrid, fblock, 0, as_stream(stream), r, grads_rw, a, n_adam)
                               : cmarl_launch(ctx, reduce_clip_adam_kernel<false>, fgrid, fblock, 0, as_stream(stream), r, grads_rw, a, n_adam));
        return 0;
    }
    {
        KernelTimer kt(ctx, K_ADAM, as_stream(stream));
        CMARL_CUDA(a.world > 1 ? cmarl_launch(ctx, clip_adam_kernel<true>, grid, block, 0, as_stream(stream), a)
                               : cmarl_launch(ctx, clip_adam_kernel<false>, grid, block, 0, as_stream(stream), a));
    }
    return 0;
}

extern "C" int cmarl_clip_adam_step(cmarl_ctx* ctx, float* params, const float* grads, float* exp_avg,
                                    float* exp_avg_sq, int32_t step, int32_t* step_dev, double lr_actor,
                                    double lr_critic, double beta1, double beta2, double eps, double max_norm,
                                    float* stats_out, void* stream) {
    return clip_adam_impl(ctx, nullptr, params, grads, nullptr, exp_avg, exp_avg_sq, step, step_dev, lr_actor, lr_critic, beta1, beta2,
                          eps, max_norm, stats_out, stream);
}

extern "C" int cmarl_reduce_clip_adam_step(cmarl_ctx* ctx, const void* workspace, float* params, float* grads_out, float* exp_avg,
                                           float* exp_avg_sq, int32_t step, int32_t* step_dev, double lr_actor,
                                           double lr_critic, double beta1, double beta2, double eps, double max_norm,
                                           float* stats_out, void* stream) {
    CMARL_ARG(ctx && workspace && grads_out, "null argument");
    CMARL_ARG(!ctx->generic, "layered shapes: use cmarl_ppo_epoch_grads + cmarl_clip_adam_step");
    CMARL_ARG(ctx->pending_grid_a > 0 && ctx->pending_grid_c > 0, "call cmarl_ppo_epoch_grads_ex with grads_out = NULL first");
    const int e = clip_adam_impl(ctx, workspace, params, grads_out, grads_out, exp_avg, exp_avg_sq, step, step_dev, lr_actor, lr_critic,
                                 beta1, beta2, eps, max_norm, stats_out, stream);
    ctx->pending_grid_a = ctx->pending_grid_c = 0;
    return e;
}

// One network at a time (recurrent path: the actor is stepped once per truncated-BPTT chunk, the critic once
// per epoch; mappo_lstm_multienvs.py:605-619, 646-655).  Same kernel: every tensor belongs to "net 0".
extern "C" int cmarl_adam_step_net(cmarl_ctx* ctx, int32_t net, float* params, const float* grads, float* exp_avg,
                                   float* exp_avg_sq, int32_t step, int32_t* step_dev, double lr, double beta1,
                                   double beta2, double eps, double max_norm, double extra_div, float* stats_out,
                                   void* stream) {
    CMARL_ARG(ctx && params && grads && exp_avg && exp_avg_sq, "null argument");
    CMARL_ARG(net == 0 || net == 1, "net must be 0 (actor) or 1 (critic)");
    CMARL_ARG(step_dev || step >= 1, "step must be >= 1");
    CMARL_ARG(extra_div >= 1.0, "extra_div must be >= 1");
    CMARL_ARG(!ctx->generic, "the per-network entries serve the recurrent path (default shapes only)");
    AdamArgs a;
    a.params = params; a.grads = grads; a.m = exp_avg; a.v = exp_avg_sq; a.stats_out = stats_out;
    a.step_dev = step_dev; a.step = step; a.ticket = ctx->dev_words + CMARL_DW_ADAM_TICKET;
    a.barrier = ctx->dev_words + CMARL_DW_ADAM_BARRIER; a.tsq_part = ctx->dev_floats + CMARL_DF_ADAM_TSQ;
    int k = 0;
    if (net == 0 && ctx->cfg.actor_recurrent) {
        const GruLayout& L = ctx->gru;
        const int offs[8] = {L.w1, L.b1, L.wih, L.whh, L.bih, L.bhh, L.w2, L.b2};
        for (int j = 0; j < 8; ++j) a.tensor_off[k++] = offs[j];
        a.tensor_off[k] = L.count;
    } else {
        const NetLayout& L = net == 0 ? ctx->actor : ctx->critic;
        const int offs[6] = {L.w1, L.b1, L.w2, L.b2, L.w3, L.b3};
        for (int j = 0; j < 6; ++j) a.tensor_off[k++] = offs[j];
        a.tensor_off[k] = L.count;
    }
    for (int j = k + 1; j < 13; ++j) a.tensor_off[j] = a.tensor_off[k];
    a.n_tensors = k; a.n_actor_tensors = k;
    CMARL_ARG(a.tensor_off[k] <= CMARL_MAX_PARAMS, "too many parameters for clip_adam_kernel");
    a.lr[0] = lr; a.lr[1] = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
    a.extra_div = (float)extra_div; a.raw_stats = 1;
    a.wd[0] = a.wd[1] = ctx->weight_decay[net];
    fill_comm(ctx, a, net);
    {
        KernelTimer kt(ctx, K_ADAM, as_stream(stream));
        const dim3 grid(ceil_div(a.tensor_off[12], ADAM_THREADS)), block(ADAM_THREADS);
        CMARL_CUDA(a.world > 1 ? cmarl_launch(ctx, clip_adam_kernel<true>, grid, block, 0, as_stream(stream), a)
                               : cmarl_launch(ctx, clip_adam_kernel<false>, grid, block, 0, as_stream(stream), a));
    }
    return 0;
}
