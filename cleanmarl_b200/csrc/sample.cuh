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

}  // namespace sample
