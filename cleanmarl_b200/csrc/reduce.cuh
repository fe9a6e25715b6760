// Fixed-order sum of the per-CTA partial rows of the chain kernels -> flat gradient vector + statistics (deterministic).
// Shared by reduce_partials_kernel (chain.cu) and the fused reduce + Adam kernel (exact.cu).
#pragma once

#include "common.cuh"

namespace chain {

// COLS output columns per CTA (a warp reads 128 contiguous bytes of one partial row); 16 thread groups each sum 1/16 of
// the partial rows with 4 independent accumulators, then a fixed-order tree over the groups -- the per-column order of
// the additions does not depend on COLS.  Returns the column's total in the threads of group 0 (`*have` set there).
constexpr int RED_GROUPS = 16;
struct ReduceArgs {
    const float* pa; int grid_a, Pa;
    const float* pc; int grid_c, Pc;
    float n_groups; int count_from_c;
};
template <int COLS>
__device__ __forceinline__ float reduce_column(const ReduceArgs& r, int cta, bool* have, int* index) {
    __shared__ float part[RED_GROUPS][COLS + 1];
    __shared__ double dpart[RED_GROUPS];
    const int col_l = threadIdx.x % COLS, grp = threadIdx.x / COLS;
    const int i = cta * COLS + col_l;
    const int Pa = r.Pa, Pc = r.Pc, P = Pa + Pc;
    const bool valid = i < P + CMARL_N_STATS;
    const float* src = r.pa;
    int n = 0, stride = 1, col = 0;
    bool zero = false;
    if (valid) {
        if (i < Pa) { src = r.pa; n = r.grid_a; stride = Pa + CMARL_N_STATS; col = i; }
        else if (i < P) { src = r.pc; n = r.grid_c; stride = Pc + CMARL_N_STATS; col = i - Pa; }
        else {
            const int k = i - P;   // out stats: 0 actor loss 1 critic loss 2 entropy 3 kl 4 clipfrac 5 n_valid(b,t)
            if (k == 1) { src = r.pc; n = r.grid_c; stride = Pc + CMARL_N_STATS; col = Pc + 0; }
            else if (k == 5 && r.count_from_c) { src = r.pc; n = r.grid_c; stride = Pc + CMARL_N_STATS; col = Pc + 1; }   // ValueHead stat 1
            else if (k <= 5) { src = r.pa; n = r.grid_a; stride = Pa + CMARL_N_STATS; col = Pa + (k == 0 ? 0 : k - 1); }
            else zero = true;
        }
    }
    const int per = (n + RED_GROUPS - 1) / RED_GROUPS;
    const int c0 = grp * per, c1 = min(n, c0 + per);
    const bool is_count = valid && (i == P + 5);
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    double d = 0.0;
    if (valid && !zero) {
        int c = c0;
        if (is_count) {
            for (; c < c1; ++c) d += (double)__ldcg(src + (size_t)c * stride + col);   // can exceed 2^24: sum in double
        } else {
            for (; c + 3 < c1; c += 4) {
                a0 += __ldcg(src + (size_t)c * stride + col);
                a1 += __ldcg(src + (size_t)(c + 1) * stride + col);
                a2 += __ldcg(src + (size_t)(c + 2) * stride + col);
                a3 += __ldcg(src + (size_t)(c + 3) * stride + col);
            }
            for (; c < c1; ++c) a0 += __ldcg(src + (size_t)c * stride + col);
        }
    }
    part[grp][col_l] = (a0 + a1) + (a2 + a3);
    if (is_count) dpart[grp] = d;
    __syncthreads();
    // fixed-order pairwise tree over the groups
    for (int w = RED_GROUPS / 2; w >= 1; w >>= 1) {
        if (grp < w) {
            part[grp][col_l] += part[grp + w][col_l];
            if (is_count) dpart[grp] += dpart[grp + w];
        }
        __syncthreads();
    }
    *have = grp == 0 && valid;
    *index = i;
    if (zero) return 0.0f;
    if (is_count) return (float)(dpart[0] / (double)r.n_groups);
    return part[0][col_l];
}

}  // namespace chain
