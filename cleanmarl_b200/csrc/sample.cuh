// Categorical sampling shared by the rollout kernels (rollout.cu, generic.cu): the exponential race of torch.multinomial
// on CPU and the Philox-drawn Exp(1) noise.
#pragma once

#include "common.cuh"

namespace sample {

constexpr int NACT = 5;

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

// The same draw for a kernel whose critical path is the ACTION (rollout_tc_kernel): argmax_a p_a / q_a = argmax_a (z_a - log q_a)
// -- the softmax normalisations of race_sample are a common positive factor and log is monotone; log q_a comes with the
// noise, off the critical path, so the decision is five subtractions and a compare chain (the two forms can differ only
// for races closer than a few ulp).  log_prob(action) = z_a - logsumexp(z) with the sum accumulated in race_sample's
// order: bit-identical to race_sample's logp for the same action.
__device__ __forceinline__ void race_action_log(const float (&z)[NACT], const float (&lq)[NACT], int& action, float& zsel) {
    float best = -INFINITY;
    action = 0;
    zsel = z[0];
#pragma unroll
    for (int a = 0; a < NACT; ++a) {
        const float r = z[a] - lq[a];
        if (r > best) { best = r; action = a; zsel = z[a]; }
    }
}
__device__ __forceinline__ float race_logp(const float (&z)[NACT], float zsel) {
    float mx = z[0];
#pragma unroll
    for (int a = 1; a < NACT; ++a) mx = fmaxf(mx, z[a]);
    float se = 0.0f;
#pragma unroll
    for (int a = 0; a < NACT; ++a) se += expf(z[a] - mx);
    return zsel - (mx + logf(se));
}

// the same on the SFU (ex2.approx / lg2.approx, ~1e-6 absolute): rollout_tc_kernel, whose log-probability warp would
// otherwise be the last at the step's closing barrier; the training chains evaluate the new log-probabilities the same way
__device__ __forceinline__ float race_logp_fast(const float (&z)[NACT], float zsel) {
    float mx = z[0];
#pragma unroll
    for (int a = 1; a < NACT; ++a) mx = fmaxf(mx, z[a]);
    float se = 0.0f;
#pragma unroll
    for (int a = 0; a < NACT; ++a) se += __expf(z[a] - mx);
    return zsel - (mx + __logf(se));
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

// log of the same Exp(1) noise for the log-domain race (rollout_tc_kernel), split by Philox block so that two idle slots of
// a step can share the work; same counters, keys and uniforms as philox_exp5, the two logarithms on the SFU (lg2.approx:
// the race only needs log q up to a perturbation far below the spacing of the logits; 25 instead of 50 instructions per value)
__device__ __forceinline__ float fast_log_exp1(uint32_t r) {
    const float q = fmaxf(-__logf(u32_to_unit_open0(r)), 1e-30f);
    return __logf(q);
}
__device__ __forceinline__ void philox_logexp_block0(uint64_t seed, uint64_t episode, uint32_t t, uint32_t n, uint32_t b, float (&lq)[4]) {
    const Philox4 r0 = philox4x32_10(b, t * 8u + n, (uint32_t)episode, (uint32_t)(episode >> 32) ^ 0x51u, (uint32_t)seed, (uint32_t)(seed >> 32));
    lq[0] = fast_log_exp1(r0.x); lq[1] = fast_log_exp1(r0.y); lq[2] = fast_log_exp1(r0.z); lq[3] = fast_log_exp1(r0.w);
}
__device__ __forceinline__ float philox_logexp_block1(uint64_t seed, uint64_t episode, uint32_t t, uint32_t n, uint32_t b) {
    const Philox4 r1 = philox4x32_10(b, t * 8u + n, (uint32_t)episode, (uint32_t)(episode >> 32) ^ 0xA3u, (uint32_t)seed, (uint32_t)(seed >> 32));
    return fast_log_exp1(r1.x);
}

}  // namespace sample
