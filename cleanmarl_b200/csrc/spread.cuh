// Device simple_spread_v3 (PettingZoo 1.25 MPE) physics in float64.
// Replaces PettingZooWrapper.step (cleanmarl/env/pettingzoo_wrapper.py:44-66) and the simulator
// behind it.  The operation order follows oracle/spread.py (the CPU restatement) exactly; the file
// is compiled with -fmad=false so no multiply-add is contracted.  PettingZoo itself is not
// available to check against ("parity unpinned"), hence the integration-order switch.
#pragma once

#ifndef CMARL_SPREAD_POS_FIRST
#define CMARL_SPREAD_POS_FIRST 1     // p_pos += p_vel*dt before the velocity update (PettingZoo >= 1.24)
#endif

namespace spread {

constexpr double DT = 0.1;
constexpr double DAMPING = 0.25;
constexpr double CONTACT_FORCE = 1e2;
constexpr double CONTACT_MARGIN = 1e-3;
constexpr double DIST_MIN = 0.15 + 0.15;
constexpr double SENSITIVITY = 5.0;
constexpr double LOCAL_RATIO = 0.5;

// np.logaddexp(0, v) with numpy's branch structure (npy_logaddexp)
__device__ __forceinline__ double logaddexp0(double v) {
    if (v == 0.0) return 0.0 + 0.6931471805599453094;
    const double tmp = 0.0 - v;
    if (tmp > 0.0) return 0.0 + log1p(exp(-tmp));
    if (tmp <= 0.0) return v + log1p(exp(tmp));
    return tmp;   // NaN
}

// contact force on entity a of the pair (a, b), a < b: 100 * d / dist * penetration, d = p_a - p_b
__device__ __forceinline__ void pair_force(double ax, double ay, double bx, double by, double& fx, double& fy) {
    const double dx = ax - bx, dy = ay - by;
    const double dist = sqrt(dx * dx + dy * dy);
    const double pen = logaddexp0(-(dist - DIST_MIN) / CONTACT_MARGIN) * CONTACT_MARGIN;
    fx = CONTACT_FORCE * dx / dist * pen;
    fy = CONTACT_FORCE * dy / dist * pen;
}

__device__ __forceinline__ double dist2d(double ax, double ay, double bx, double by) {
    const double dx = ax - bx, dy = ay - by;
    return sqrt(dx * dx + dy * dy);
}

// Total force on agent n given all positions p[6] = (x0,y0,x1,y1,x2,y2) and its own action.
// Accumulation order = apply_action_force then the pair loop (0,1),(0,2),(1,2) of World.step.
__device__ __forceinline__ void agent_force(int n, const double* p, int action, double& fx, double& fy) {
    double ux = 0.0, uy = 0.0;
    if (action == 1) ux = -1.0;
    if (action == 2) ux = +1.0;
    if (action == 3) uy = -1.0;
    if (action == 4) uy = +1.0;
    fx = ux * SENSITIVITY + 0.0;
    fy = uy * SENSITIVITY + 0.0;
    double gx, gy;
    if (n == 0) {
        pair_force(p[0], p[1], p[2], p[3], gx, gy); fx = gx + fx; fy = gy + fy;
        pair_force(p[0], p[1], p[4], p[5], gx, gy); fx = gx + fx; fy = gy + fy;
    } else if (n == 1) {
        pair_force(p[0], p[1], p[2], p[3], gx, gy); fx = -gx + fx; fy = -gy + fy;
        pair_force(p[2], p[3], p[4], p[5], gx, gy); fx = gx + fx; fy = gy + fy;
    } else {
        pair_force(p[0], p[1], p[4], p[5], gx, gy); fx = -gx + fx; fy = -gy + fy;
        pair_force(p[2], p[3], p[4], p[5], gx, gy); fx = -gx + fx; fy = -gy + fy;
    }
}

__device__ __forceinline__ void integrate(double& px, double& py, double& vx, double& vy, double fx, double fy) {
#if CMARL_SPREAD_POS_FIRST
    px = px + vx * DT; py = py + vy * DT;
    vx = vx * (1 - DAMPING); vy = vy * (1 - DAMPING);
    vx = vx + (fx / 1.0) * DT; vy = vy + (fy / 1.0) * DT;
#else
    vx = vx * (1 - DAMPING); vy = vy * (1 - DAMPING);
    vx = vx + (fx / 1.0) * DT; vy = vy + (fy / 1.0) * DT;
    px = px + vx * DT; py = py + vy * DT;
#endif
}

// reward of agent 0 after the step: 0.5 * global + 0.5 * local (pettingzoo_wrapper.py:66 keeps agent 0's)
__device__ __forceinline__ double reward_agent0(const double* p, const double* lm) {
    double g = 0.0;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const double d0 = dist2d(p[0], p[1], lm[2 * l], lm[2 * l + 1]);
        const double d1 = dist2d(p[2], p[3], lm[2 * l], lm[2 * l + 1]);
        const double d2 = dist2d(p[4], p[5], lm[2 * l], lm[2 * l + 1]);
        g = g - fmin(fmin(d0, d1), d2);
    }
    double loc = 0.0;
    loc = loc - 1.0 * (dist2d(p[2], p[3], p[0], p[1]) < DIST_MIN ? 1.0 : 0.0);
    loc = loc - 1.0 * (dist2d(p[4], p[5], p[0], p[1]) < DIST_MIN ? 1.0 : 0.0);
    return g * (1 - LOCAL_RATIO) + loc * LOCAL_RATIO;
}

// raw observation of agent n (18 floats): vel, pos, landmarks - pos, other agents - pos, 4 zeros
__device__ __forceinline__ void observe(int n, const double* p, const double* v, const double* lm, float* o) {
    const double px = p[2 * n], py = p[2 * n + 1];
    o[0] = (float)v[2 * n]; o[1] = (float)v[2 * n + 1];
    o[2] = (float)px; o[3] = (float)py;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        o[4 + 2 * l] = (float)(lm[2 * l] - px);
        o[5 + 2 * l] = (float)(lm[2 * l + 1] - py);
    }
    const int j0 = (n == 0) ? 1 : 0, j1 = (n == 2) ? 1 : 2;
    o[10] = (float)(p[2 * j0] - px); o[11] = (float)(p[2 * j0 + 1] - py);
    o[12] = (float)(p[2 * j1] - px); o[13] = (float)(p[2 * j1 + 1] - py);
    o[14] = 0.0f; o[15] = 0.0f; o[16] = 0.0f; o[17] = 0.0f;
}

}  // namespace spread
