"""float64 numpy restatement of PettingZoo 1.25.0 MPE ``simple_spread_v3``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

PARITY UNPINNED: the simulator lives in the third-party package
``pettingzoo==1.25.0`` (reference ``pyproject.toml:24``; modules
``pettingzoo/mpe/simple_spread/simple_spread.py`` (Scenario),
``pettingzoo/mpe/_mpe_utils/core.py`` (World) and ``.../simple_env.py``
(SimpleEnv)).  It is neither vendored under /root/reference nor installed in
this image, and the reference holds no test or golden vector for it, so this
file restates the published algorithm and nothing can check it against the
real library here.  The reference's own call sites that this file serves:
``cleanmarl/env/pettingzoo_wrapper.py:18-20`` (construct/reset), ``:36``
(reset(seed)), ``:47`` (step), ``:23-30`` (spaces), ``:92`` (sample).

Algorithm (N agents, L = N landmarks -- ``simple_spread_v3(N=3)`` in the reference, which passes no env kwargs
(MME:297); every function below takes N / L from the array shapes, so the N != 3 device envs are checked by the
same restatement; everything float64):

* world: dt 0.1, damping 0.25, contact_force 100, contact_margin 1e-3;
  agents size 0.15, mass 1, collide, silent, sensitivity 5; landmarks fixed,
  non-colliding.
* reset: per agent ``p_pos ~ U(-1,1)^2``, ``p_vel = 0``; then per landmark
  ``p_pos ~ U(-1,1)^2`` -- drawn in that order from ``np_random``.
* discrete action a: 0 noop, 1 -x, 2 +x, 3 -y, 4 +y; ``u = +-1 * 5.0``.
* step: ``f_i = u_i``; for agent pairs a < b in lexicographic order ((0,1),(0,2),(1,2) for N = 3):
  ``d = p_a - p_b; dist = sqrt(d.d); pen = logaddexp(0, -(dist-0.3)/k) * k;
  F = 100 * d / dist * pen; f_a += F; f_b -= F``; then per agent
  ``p_pos += p_vel*dt`` (old velocity -- ``INTEGRATE_POS_FIRST``),
  ``p_vel = p_vel*0.75 + f*dt``.
* reward after the step: ``g = -sum_l min_a |p_a - p_l|``;
  ``loc_i = -#{j != i : |p_i-p_j| < 0.3}``; ``r_i = 0.5 g + 0.5 loc_i``.
* obs_i = [vel_i, pos_i, lm_k - pos_i (k < L), pos_j - pos_i (j != i, index
  order), 2 zeros per other agent (their silent communication state)] cast to
  float32 -- 4 + 2 L + 4 (N - 1) numbers, 18 for N = L = 3; truncation when 25
  steps were taken.

``INTEGRATE_POS_FIRST`` mirrors the compile-time switch of the CUDA kernel
(``CMARL_SPREAD_POS_FIRST``): the integration order is the one detail of the
restatement that differs between MPE forks, so it is kept switchable.
"""
from __future__ import annotations

import numpy as np

N_AGENTS = 3
N_LANDMARKS = 3
RAW_OBS = 18
MAX_CYCLES = 25
DT = 0.1
DAMPING = 0.25
CONTACT_FORCE = 1e2
CONTACT_MARGIN = 1e-3
AGENT_SIZE = 0.15
SENSITIVITY = 5.0
LOCAL_RATIO = 0.5
INTEGRATE_POS_FIRST = True



def raw_obs_dim(n_agents: int, n_landmarks: int | None = None) -> int:
    L = n_agents if n_landmarks is None else n_landmarks
    return 4 + 2 * L + 4 * (n_agents - 1)


def action_force(actions: np.ndarray) -> np.ndarray:
    """actions int [B,N] -> u float64 [B,N,2]."""
    a = np.asarray(actions).astype(np.int64)
    u = np.zeros(a.shape + (2,), dtype=np.float64)
    u[..., 0] = np.where(a == 1, -1.0, np.where(a == 2, 1.0, 0.0))
    u[..., 1] = np.where(a == 3, -1.0, np.where(a == 4, 1.0, 0.0))
    return u * SENSITIVITY


def step_batched(pos, vel, lm, actions, pos_first: bool = INTEGRATE_POS_FIRST):
    """One world step for B independent envs.

    pos, vel: float64 [B,N,2]; lm: float64 [B,L,2]; actions: int [B,N].
    Returns (pos', vel', reward[B,N] float64).  Operation order follows the
    scalar algorithm in the module docstring exactly so the result is what a
    per-env Python loop would produce bit for bit.
    """
    pos = np.array(pos, dtype=np.float64, copy=True)
    vel = np.array(vel, dtype=np.float64, copy=True)
    lm = np.asarray(lm, dtype=np.float64)
    force = action_force(actions) + 0.0
    k = CONTACT_MARGIN
    dist_min = AGENT_SIZE + AGENT_SIZE
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        n = pos.shape[1]
        for a, b in ((a, b) for a in range(n) for b in range(a + 1, n)):
            d = pos[:, a] - pos[:, b]
            dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
            pen = np.logaddexp(0.0, -(dist - dist_min) / k) * k
            f = CONTACT_FORCE * d / dist[:, None] * pen[:, None]
            force[:, a] = f + force[:, a]
            force[:, b] = -f + force[:, b]
    if pos_first:
        pos = pos + vel * DT
        vel = vel * (1 - DAMPING)
        vel = vel + (force / 1.0) * DT
    else:
        vel = vel * (1 - DAMPING)
        vel = vel + (force / 1.0) * DT
        pos = pos + vel * DT
    return pos, vel, rewards_batched(pos, lm)


def _dist(a, b):
    d = a - b
    return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])


def rewards_batched(pos, lm):
    """Per-agent reward float64 [B,N] = 0.5*global + 0.5*local."""
    B, n_agents, n_landmarks = pos.shape[0], pos.shape[1], lm.shape[1]
    g = np.zeros(B, dtype=np.float64)
    for l in range(n_landmarks):
        dl = np.stack([_dist(pos[:, a], lm[:, l]) for a in range(n_agents)], axis=0)
        g = g - dl.min(axis=0)
    dist_min = AGENT_SIZE + AGENT_SIZE
    rew = np.zeros((B, n_agents), dtype=np.float64)
    for i in range(n_agents):
        loc = np.zeros(B, dtype=np.float64)
        for j in range(n_agents):
            if j == i:
                continue
            loc = loc - 1.0 * (_dist(pos[:, j], pos[:, i]) < dist_min)
        rew[:, i] = g * (1 - LOCAL_RATIO) + loc * LOCAL_RATIO
    return rew


def observe_batched(pos, vel, lm):
    """Raw observations float32 [B,N,4 + 2 L + 4 (N - 1)] (PettingZoo casts to float32)."""
    B, n_agents, n_landmarks = pos.shape[0], pos.shape[1], lm.shape[1]
    obs = np.zeros((B, n_agents, raw_obs_dim(n_agents, n_landmarks)), dtype=np.float64)
    for i in range(n_agents):
        obs[:, i, 0:2] = vel[:, i]
        obs[:, i, 2:4] = pos[:, i]
        for l in range(n_landmarks):
            obs[:, i, 4 + 2 * l : 6 + 2 * l] = lm[:, l] - pos[:, i]
        c = 4 + 2 * n_landmarks
        for j in range(n_agents):
            if j == i:
                continue
            obs[:, i, c : c + 2] = pos[:, j] - pos[:, i]
            c += 2
    return obs.astype(np.float32)


def rollout_batched(pos0, lm, actions_tbn, pos_first: bool = INTEGRATE_POS_FIRST):
    """Open-loop rollout: actions int [T,B,N] -> dict of [T,...] arrays.

    states[t] is the observation *before* action t (what the reference stores,
    ``mappo_multienvs.py:426-430``); reward[t] is agent 0's reward after it
    (``pettingzoo_wrapper.py:66``).
    """
    T = actions_tbn.shape[0]
    pos = np.array(pos0, dtype=np.float64, copy=True)
    vel = np.zeros_like(pos)
    lm = np.asarray(lm, dtype=np.float64)
    raw, rew = [], []
    for t in range(T):
        raw.append(observe_batched(pos, vel, lm))
        pos, vel, r = step_batched(pos, vel, lm, actions_tbn[t], pos_first)
        rew.append(r[:, 0])
    return {
        "raw_obs": np.stack(raw),                       # [T,B,N,R] f32
        "reward": np.stack(rew),                        # [T,B] f64
        "final_pos": pos,
        "final_vel": vel,
        "final_raw_obs": observe_batched(pos, vel, lm),
    }


class _Discrete:
    def __init__(self, n, rng):
        self.n = n
        self._rng = rng

    def sample(self):
        return int(self._rng.integers(self.n))


class _Box:
    def __init__(self, shape):
        self.shape = shape


class SimpleSpreadParallelEnv:
    """The slice of PettingZoo's ``parallel_env`` API the reference touches."""

    def __init__(self, N=3, local_ratio=0.5, max_cycles=MAX_CYCLES, **_unused):
        assert N >= 1 and local_ratio == LOCAL_RATIO
        self.n_agents = N                               # simple_spread_v3(N): N agents and N landmarks
        self.max_cycles = max_cycles
        self.possible_agents = [f"agent_{i}" for i in range(N)]
        self.agents = list(self.possible_agents)
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
        self._act_rng = np.random.default_rng()
        self._reset_world()

    @property
    def num_agents(self):
        return len(self.agents)

    def action_space(self, agent):
        return _Discrete(5, self._act_rng)

    def observation_space(self, agent):
        return _Box((raw_obs_dim(self.n_agents),))

    def _reset_world(self):
        pos = np.zeros((1, self.n_agents, 2))
        lm = np.zeros((1, self.n_agents, 2))
        for i in range(self.n_agents):
            pos[0, i] = self.np_random.uniform(-1, +1, 2)
        for l in range(self.n_agents):
            lm[0, l] = self.np_random.uniform(-1, +1, 2)
        self.pos, self.vel, self.lm = pos, np.zeros_like(pos), lm
        self.steps = 0

    def _obs_dict(self):
        o = observe_batched(self.pos, self.vel, self.lm)[0]
        return {a: o[i] for i, a in enumerate(self.possible_agents)}

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        self.agents = list(self.possible_agents)
        self._reset_world()
        return self._obs_dict(), {a: {} for a in self.agents}

    def step(self, actions):
        act = np.array([[int(actions[a]) for a in self.possible_agents]])
        self.pos, self.vel, rew = step_batched(self.pos, self.vel, self.lm, act)
        self.steps += 1
        trunc = self.steps >= self.max_cycles
        obs = self._obs_dict()
        rewards = {a: float(rew[0, i]) for i, a in enumerate(self.possible_agents)}
        terms = {a: False for a in self.possible_agents}
        truncs = {a: trunc for a in self.possible_agents}
        infos = {a: {} for a in self.possible_agents}
        if trunc:
            self.agents = []
        return obs, rewards, terms, truncs, infos

    def close(self):
        pass


def parallel_env(**kwargs):
    return SimpleSpreadParallelEnv(**kwargs)
