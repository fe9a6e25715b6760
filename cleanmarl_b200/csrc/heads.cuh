// Per-sample heads shared by the FFMA chain (chain.cu) and the tcgen05 chain (tc_chain.cu):
// what happens to the network output z of one sample (loss terms, statistics, dz).
#pragma once

#include "chain.cuh"

namespace chain {

// ------------------------------------------------------------------------------------------------
// Heads: what happens to the network output z of one sample
// ------------------------------------------------------------------------------------------------
// SFU = true (tc_gru.cu): exp / log / the softmax divisions on the special-function unit (ex2.approx, lg2.approx, rcp.approx:
// <= 2 ulp each on the argument ranges that occur: logits - max in [-1e9, 0], a sum of exponentials in [1, 5], a log-ratio of a
// few tenths) instead of the libm forms, which are ~3x the instructions of the whole head.
template <bool SFU>
struct PolicyHeadT {
    static __device__ __forceinline__ float exp_(float x) { return SFU ? __expf(x) : expf(x); }
    static __device__ __forceinline__ float log_(float x) { return SFU ? __logf(x) : logf(x); }
    static __device__ __forceinline__ float div_(float a, float b) { return SFU ? __fdividef(a, b) : a / b; }
    static constexpr int OUT = 5;
    static constexpr int NSTAT = 5;     // loss, entropy, kl, clip fraction, valid samples
    using Args = PolicyHeadArgs;
    // MME:530-551, 561-570 for one (b, t, agent) sample; all sums carry the 1/N of `.mean(dim=-1)`.
    // Per-sample inputs, loadable ahead of the network output (the loads' latency then hides behind the GEMMs).
    struct In {
        bool live;          // in range and mask == 1
        uint32_t unavail;   // bit a set: action a is masked out (MME:182)
        int act;
        float logp_old, adv;
    };
    __device__ static __forceinline__ In load(const Args& h, int t, int g, int b, int G, int B, bool inb) {
        In in;
        in.live = false; in.unavail = 0u; in.act = 0; in.logp_old = 0.0f; in.adv = 0.0f;
        if (!inb) return in;
        const size_t tb = (size_t)t * B + b;
        if (h.mask && !h.mask[tb]) return in;
        in.live = true;
        const size_t tgb = ((size_t)t * G + g) * B + b;
        if (h.avail) {
#pragma unroll
            for (int a = 0; a < OUT; ++a)
                if (!h.avail[(((size_t)t * G + g) * h.A + a) * B + b]) in.unavail |= 1u << a;
        }
        in.act = h.actions[tgb];
        in.logp_old = h.logp_old[tgb];
        in.adv = h.adv[h.V == 1 ? tb : tgb];
        return in;
    }
    __device__ static __forceinline__ void apply(const Args& h, float (&z)[OUT], int t, int g, int b, int G, int B,
                                                 bool inb, bool train, float (&dz)[OUT], float (&st)[NSTAT]) {
        const In in = load(h, t, g, b, G, B, inb);
        compute(h, in, z, train, dz, st);
    }
    __device__ static __forceinline__ void compute(const Args& h, const In& in, float (&z)[OUT], bool train,
                                                   float (&dz)[OUT], float (&st)[NSTAT]) {
#pragma unroll
        for (int a = 0; a < OUT; ++a) dz[a] = 0.0f;
        if (!in.live) return;
#pragma unroll
        for (int a = 0; a < OUT; ++a)
            if (in.unavail & (1u << a)) z[a] = -1e9f;   // masked_fill, MME:182
        // Categorical(logits=z): logits = z - logsumexp(z); probs = softmax(logits)
        float mx = z[0];
#pragma unroll
        for (int a = 1; a < OUT; ++a) mx = fmaxf(mx, z[a]);
        float se = 0.0f;
#pragma unroll
        for (int a = 0; a < OUT; ++a) se += exp_(z[a] - mx);
        const float lse = mx + log_(se);
        float l[OUT], p[OUT];
        float mx2 = -INFINITY;
#pragma unroll
        for (int a = 0; a < OUT; ++a) { l[a] = z[a] - lse; mx2 = fmaxf(mx2, l[a]); }
        float se2 = 0.0f;
#pragma unroll
        for (int a = 0; a < OUT; ++a) { p[a] = exp_(l[a] - mx2); se2 += p[a]; }
        float ent = 0.0f;
#pragma unroll
        for (int a = 0; a < OUT; ++a) { p[a] = div_(p[a], se2); ent -= l[a] * p[a]; }
        const int act = in.act;
        float logp = l[0];
#pragma unroll
        for (int a = 1; a < OUT; ++a) logp = (act == a) ? l[a] : logp;
        const float log_ratio = logp - in.logp_old;
        const float ratio = exp_(log_ratio);
        const float A = in.adv;
        const float lo = 1.0f - h.clip, hi = 1.0f + h.clip;
        const float pg1 = A * ratio;
        const float pg2 = A * fminf(fmaxf(ratio, lo), hi);
        const float pg = fminf(pg1, pg2);
        const float w = h.inv_groups;
        st[0] += w * (-pg - h.ent_coef * ent);
        st[1] += w * ent;
        st[2] += w * ((ratio - 1.0f) - log_ratio);
        st[3] += (fabsf(ratio - 1.0f) > h.clip) ? w : 0.0f;
        st[4] += 1.0f;
        if (!train) return;
        // d(-min(pg1,pg2))/d(ratio): clamp passes the gradient inside [lo,hi] (ties of torch.min split
        // 1/2 + 1/2 and recombine); outside, only the unclipped branch carries one.
        const bool inside = (ratio >= lo) && (ratio <= hi);
        const float dmin = (inside || pg1 < pg2) ? A : ((pg1 == pg2) ? 0.5f * A : 0.0f);
        const float dlogp = -w * dmin * ratio;
        const float we = w * h.ent_coef;
#pragma unroll
        for (int a = 0; a < OUT; ++a) {
            const float onehot = (act == a) ? 1.0f : 0.0f;
            dz[a] = dlogp * (onehot - p[a]) + we * p[a] * (l[a] + ent);
        }
#pragma unroll
        for (int a = 0; a < OUT; ++a)
            if (in.unavail & (1u << a)) dz[a] = 0.0f;
    }
};
using PolicyHead = PolicyHeadT<false>;

struct ValueHead {
    static constexpr int OUT = 1;
    static constexpr int NSTAT = 2;     // loss, valid samples
    using Args = ValueHeadArgs;
    // MME:554-558: sum_env mean_agent (V - R)^2 ; forward-only mode just stores V (MME:495,502).
    struct In {
        bool inb, live;     // in range; in range and mask == 1
        float ret, vold;
        float* vout;        // forward mode: where V goes
    };
    __device__ static __forceinline__ In load(const Args& h, int t, int g, int b, int G, int B, bool inb) {
        In in;
        in.inb = inb; in.live = false; in.ret = 0.0f; in.vold = 0.0f; in.vout = nullptr;
        if (!inb) return in;
        const size_t tgb = ((size_t)t * G + g) * B + b;
        if (h.values_out) { in.vout = h.values_out + tgb; return in; }
        if (h.mask && !h.mask[(size_t)t * B + b]) return in;
        in.live = true;
        in.ret = h.returns[tgb];
        if (h.vclip > 0.0f) in.vold = h.values_old[tgb];
        return in;
    }
    __device__ static __forceinline__ void apply(const Args& h, float (&z)[OUT], int t, int g, int b, int G, int B,
                                                 bool inb, bool train, float (&dz)[OUT], float (&st)[NSTAT]) {
        const In in = load(h, t, g, b, G, B, inb);
        compute(h, in, z, train, dz, st);
    }
    __device__ static __forceinline__ void compute(const Args& h, const In& in, float (&z)[OUT], bool train,
                                                   float (&dz)[OUT], float (&st)[NSTAT]) {
        dz[0] = 0.0f;
        if (!in.inb) return;
        if (!train) {
            *in.vout = z[0];
            return;
        }
        if (!in.live) return;
        const float diff = z[0] - in.ret;
        st[1] += 1.0f;
        if (h.vclip > 0.0f) {
            // beyond the reference (default off): max((V - R)^2, (V_old + clamp(V - V_old, +-c) - R)^2); the clipped branch
            // carries a gradient only while V is inside the clip range (torch.clamp / torch.max autograd; ties 1/2 + 1/2)
            const float dv = z[0] - in.vold;
            const float dvc = fminf(fmaxf(dv, -h.vclip), h.vclip);
            const float diffc = (in.vold + dvc) - in.ret;
            const float l1 = diff * diff, l2 = diffc * diffc;
            const bool inside = (dv >= -h.vclip) && (dv <= h.vclip);
            st[0] += h.inv_heads * fmaxf(l1, l2);
            const float g1 = 2.0f * diff, g2 = inside ? 2.0f * diffc : 0.0f;
            dz[0] = h.inv_heads * (l1 > l2 ? g1 : (l1 == l2 ? 0.5f * (g1 + g2) : g2));
            return;
        }
        st[0] += h.inv_heads * diff * diff;
        dz[0] = h.inv_heads * 2.0f * diff;
    }
};

}  // namespace chain
