"""Known-answer and cross-restatement tests of oracle/spread.py (CPU only).

The simulator the reference drives (PettingZoo 1.25.0 ``simple_spread_v3``, reference call sites
``cleanmarl/env/pettingzoo_wrapper.py:18-20, 36, 47``) is not available here, so the oracle is PARITY UNPINNED at this
boundary (oracle/spread.py header).  What can be checked is that the batched numpy restatement computes exactly the
published algorithm: hand-derived answers for single steps, and a second, independent, scalar pure-Python restatement
written from the same description that must agree bit for bit on random trajectories.  The adapter logic the reference
itself owns (``pettingzoo_wrapper.py:51-52, 66, 79-98``) is checked on the stand-in package of oracle/env_stub.
"""
import math
import os
import sys

import numpy as np
import pytest

from oracle import spread as osp


# ---------------------------------------------------------------------------------------------------------------------
# independent scalar restatement (one env, python floats = IEEE double, same operation order as the description)
# ---------------------------------------------------------------------------------------------------------------------
def _logaddexp0(x):
    # log(1 + exp(x)) the way numpy's logaddexp(0, x) evaluates it
    return float(np.logaddexp(0.0, x))


def scalar_step(pos, vel, lm, actions, pos_first=True):
    n = len(pos)
    force = []
    for a in actions:
        ux = -1.0 if a == 1 else (1.0 if a == 2 else 0.0)
        uy = -1.0 if a == 3 else (1.0 if a == 4 else 0.0)
        force.append([ux * 5.0, uy * 5.0])
    for a in range(n):
        for b in range(a + 1, n):
            dx, dy = pos[a][0] - pos[b][0], pos[a][1] - pos[b][1]
            dist = math.sqrt(dx * dx + dy * dy)
            pen = _logaddexp0(-(dist - 0.3) / 1e-3) * 1e-3
            fx, fy = 100.0 * dx / dist * pen, 100.0 * dy / dist * pen
            force[a] = [fx + force[a][0], fy + force[a][1]]
            force[b] = [-fx + force[b][0], -fy + force[b][1]]
    npos, nvel = [], []
    for i in range(n):
        px, py = pos[i]
        vx, vy = vel[i]
        if pos_first:
            px, py = px + vx * 0.1, py + vy * 0.1
            vx, vy = vx * 0.75 + force[i][0] * 0.1, vy * 0.75 + force[i][1] * 0.1
        else:
            vx, vy = vx * 0.75 + force[i][0] * 0.1, vy * 0.75 + force[i][1] * 0.1
            px, py = px + vx * 0.1, py + vy * 0.1
        npos.append([px, py])
        nvel.append([vx, vy])

    def d(p, q):
        ex, ey = p[0] - q[0], p[1] - q[1]
        return math.sqrt(ex * ex + ey * ey)

    g = 0.0
    for l in range(len(lm)):
        g = g - min(d(npos[a], lm[l]) for a in range(n))
    rew = []
    for i in range(n):
        loc = 0.0
        for j in range(n):
            if j != i and d(npos[j], npos[i]) < 0.3:
                loc -= 1.0
        rew.append(g * 0.5 + loc * 0.5)
    return npos, nvel, rew


def scalar_obs(pos, vel, lm, i):
    o = [vel[i][0], vel[i][1], pos[i][0], pos[i][1]]
    for l in range(len(lm)):
        o += [lm[l][0] - pos[i][0], lm[l][1] - pos[i][1]]
    for j in range(len(pos)):
        if j != i:
            o += [pos[j][0] - pos[i][0], pos[j][1] - pos[i][1]]
    return np.asarray(o + [0.0] * (2 * (len(pos) - 1)), dtype=np.float64).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# hand-derived answers
# ---------------------------------------------------------------------------------------------------------------------
FAR = np.array([[[-0.9, -0.9], [0.0, 0.9], [0.9, -0.9]]])           # pairwise distances > 1.8: no contact force
LM = np.array([[[-0.9, -0.8], [0.1, 0.9], [0.5, 0.5]]])


def test_action_table():
    u = osp.action_force(np.array([[0, 1, 2], [3, 4, 0]]))
    assert u.dtype == np.float64
    assert u.tolist() == [[[0, 0], [-5, 0], [5, 0]], [[0, -5], [0, 5], [0, 0]]]


def test_free_motion_two_steps_by_hand():
    vel = np.zeros_like(FAR)
    pos, vel, _ = osp.step_batched(FAR, vel, LM, np.array([[2, 3, 0]]))
    # positions move with the OLD velocity (zero); v = 0 * 0.75 + 5 * 0.1
    assert np.array_equal(pos, FAR)
    assert vel[0].tolist() == [[0.5, 0.0], [0.0, -0.5], [0.0, 0.0]]
    pos, vel, _ = osp.step_batched(pos, vel, LM, np.array([[0, 0, 4]]))
    assert pos[0].tolist() == [[-0.9 + 0.5 * 0.1, -0.9], [0.0, 0.9 + (-0.5) * 0.1], [0.9, -0.9]]
    assert vel[0].tolist() == [[0.5 * 0.75, 0.0], [0.0, -0.5 * 0.75], [0.0, 0.5]]
    # the other integration order (switch kept on both sides: CMARL_SPREAD_POS_FIRST) moves with the NEW velocity
    pos2, vel2, _ = osp.step_batched(FAR, np.zeros_like(FAR), LM, np.array([[2, 3, 0]]), pos_first=False)
    assert vel2[0].tolist() == [[0.5, 0.0], [0.0, -0.5], [0.0, 0.0]]
    assert pos2[0].tolist() == [[-0.9 + 0.5 * 0.1, -0.9], [0.0, 0.9 + (-0.5) * 0.1], [0.9, -0.9]]


def test_contact_force_by_hand():
    # agents 0 and 1 overlap by 0.1 along x, agent 2 far away
    pos = np.array([[[0.1, 0.0], [-0.1, 0.0], [0.9, 0.9]]])
    _, vel, _ = osp.step_batched(pos, np.zeros_like(pos), LM, np.array([[0, 0, 0]]))
    pen = np.logaddexp(0.0, -(0.2 - 0.3) / 1e-3) * 1e-3               # softplus(100) * 1e-3 = 0.1 (+ 4e-47)
    f = 100.0 * 0.2 / 0.2 * pen
    assert abs(pen - 0.1) < 1e-15
    assert vel[0, 0].tolist() == [f * 0.1, 0.0] and vel[0, 1].tolist() == [-f * 0.1, 0.0]
    assert vel[0, 2].tolist() == [0.0, 0.0]
    # exactly at contact distance the soft force is k * log 2 * 100, not zero
    pos = np.array([[[0.15, 0.0], [-0.15, 0.0], [0.9, 0.9]]])
    _, vel, _ = osp.step_batched(pos, np.zeros_like(pos), LM, np.array([[0, 0, 0]]))
    assert vel[0, 0, 0] == pytest.approx(100.0 * math.log(2.0) * 1e-3 * 0.1, rel=1e-12)
    # far apart: exp(-1500) underflows, the force is exactly zero (and no warning leaks)
    _, vel, _ = osp.step_batched(FAR, np.zeros_like(FAR), LM, np.array([[0, 0, 0]]))
    assert not vel.any()


def test_contact_forces_conserve_momentum():
    rng = np.random.default_rng(3)
    pos = rng.uniform(-0.2, 0.2, (64, 3, 2))                          # crowded: every pair in contact range
    vel = rng.uniform(-1, 1, (64, 3, 2))
    _, v1, _ = osp.step_batched(pos, vel, rng.uniform(-1, 1, (64, 3, 2)), np.zeros((64, 3), dtype=np.int64))
    np.testing.assert_allclose(v1.sum(axis=1), 0.75 * vel.sum(axis=1), rtol=0, atol=1e-12)


def test_reward_by_hand():
    pos = np.array([[[0.0, 0.0], [0.2, 0.0], [1.0, 1.0]]])
    lm = np.array([[[0.0, 0.3], [0.2, -0.4], [1.0, 0.0]]])
    r = osp.rewards_batched(pos, lm)
    g = -(0.3 + 0.4 + math.sqrt(0.8 * 0.8 + 0.0))                     # landmark 2: closest is agent 1 at (0.2, 0)
    assert r[0].tolist() == [g * 0.5 + -1.0 * 0.5, g * 0.5 + -1.0 * 0.5, g * 0.5 + 0.0]
    # the collision test is strict (< 0.3)
    pos = np.array([[[0.0, 0.0], [0.3, 0.0], [1.0, 1.0]]])
    r = osp.rewards_batched(pos, lm)
    assert r[0, 0] == r[0, 2]


def test_observation_layout():
    rng = np.random.default_rng(5)
    pos, vel, lm = rng.uniform(-1, 1, (1, 3, 2)), rng.uniform(-1, 1, (1, 3, 2)), rng.uniform(-1, 1, (1, 3, 2))
    o = osp.observe_batched(pos, vel, lm)
    assert o.shape == (1, 3, 18) and o.dtype == np.float32
    for i in range(3):
        assert np.array_equal(o[0, i], scalar_obs(pos[0].tolist(), vel[0].tolist(), lm[0].tolist(), i))
    assert not o[..., 14:].any()                                      # silent agents: 2 x 2 communication zeros
    # agent 1 sees agent 0 first, then agent 2 (index order, itself skipped)
    assert o[0, 1, 10] == np.float32(pos[0, 0, 0] - pos[0, 1, 0]) and o[0, 1, 12] == np.float32(pos[0, 2, 0] - pos[0, 1, 0])


@pytest.mark.parametrize("n_agents", [3, 1, 2, 5])
@pytest.mark.parametrize("pos_first", [True, False])
def test_batched_numpy_equals_scalar_restatement_bit_for_bit(pos_first, n_agents):
    """(N = 3 is the reference's env; the other agent counts are simple_spread_v3(N), served by the layered device path)"""
    rng = np.random.default_rng(11)
    B, T = 16, 25
    pos0 = rng.uniform(-1, 1, (B, n_agents, 2))
    pos0[:4] *= 0.25                                                  # a few crowded envs so that contacts happen
    lm = rng.uniform(-1, 1, (B, n_agents, 2))
    acts = rng.integers(0, 5, (T, B, n_agents))
    ref = osp.rollout_batched(pos0, lm, acts, pos_first=pos_first)
    contacts = 0
    for b in range(B):
        pos, vel = pos0[b].tolist(), [[0.0, 0.0] for _ in range(n_agents)]
        for t in range(T):
            for i in range(n_agents):
                o = scalar_obs(pos, vel, lm[b].tolist(), i)
                assert o.shape == (osp.raw_obs_dim(n_agents),) and np.array_equal(ref["raw_obs"][t, b, i], o)
            pos, vel, rew = scalar_step(pos, vel, lm[b].tolist(), acts[t, b].tolist(), pos_first)
            assert rew[0] == ref["reward"][t, b]
            contacts += any(math.hypot(pos[i][0] - pos[j][0], pos[i][1] - pos[j][1]) < 0.3 for i in range(n_agents) for j in range(i))
        assert pos == ref["final_pos"][b].tolist() and vel == ref["final_vel"][b].tolist()
    assert (contacts > 0) == (n_agents > 1)                           # the collision branch was exercised


def test_reset_draw_order_and_seeding():
    env = osp.parallel_env(N=3, local_ratio=0.5, max_cycles=25)
    obs, infos = env.reset(seed=123)
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(123)))
    agents = [g.uniform(-1, +1, 2) for _ in range(3)]                 # agents first ...
    lms = [g.uniform(-1, +1, 2) for _ in range(3)]                    # ... then landmarks
    assert np.array_equal(env.pos[0], np.stack(agents)) and np.array_equal(env.lm[0], np.stack(lms))
    assert not env.vel.any()
    assert sorted(obs) == ["agent_0", "agent_1", "agent_2"] and obs["agent_0"].dtype == np.float32
    obs2, _ = env.reset(seed=123)
    assert all(np.array_equal(obs[a], obs2[a]) for a in obs)
    obs3, _ = env.reset()                                             # no seed: the stream continues, new layout
    assert not np.array_equal(obs3["agent_0"], obs["agent_0"])


def test_truncation_after_25_steps():
    env = osp.parallel_env(N=3, local_ratio=0.5, max_cycles=25)
    env.reset(seed=1)
    for t in range(25):
        assert env.agents
        _, rew, terms, truncs, _ = env.step({a: 0 for a in env.possible_agents})
        assert not any(terms.values())
        assert all(truncs.values()) == (t == 24)
        assert all(isinstance(v, float) for v in rew.values())
    assert env.agents == []


def test_adapter_contract_of_the_stand_in_package():
    """What the reference's own adapter adds (pettingzoo_wrapper.py:51-52, 66, 79-98), on oracle/env_stub."""
    stub = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "env_stub")
    saved = {k: v for k, v in sys.modules.items() if k == "env" or k.startswith("env.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, stub)
    try:
        from env.pettingzoo_wrapper import PettingZooWrapper
        w = PettingZooWrapper("mpe", "simple_spread_v3", agent_ids=True, N=3, local_ratio=0.5, max_cycles=25)
        assert (w.n_agents, w.get_obs_size(), w.get_state_size(), w.get_action_size()) == (3, 21, 54, 5)
        obs, info = w.reset(seed=9)
        assert obs.shape == (3, 21) and obs.dtype == np.float64 and info == {}      # np.eye concat promotes to float64
        assert np.array_equal(obs[:, 18:], np.eye(3))
        state = w.get_state()
        assert state.shape == (54,) and state.dtype == np.float32
        assert np.array_equal(state, obs[:, :18].astype(np.float32).reshape(-1))     # state = raw obs, ids not included
        assert np.array_equal(np.asarray(w.get_avail_actions()), np.ones((3, 5)))
        inner = w.env
        for t in range(25):
            before = (inner.pos.copy(), inner.vel.copy())
            obs, reward, done, truncated, _ = w.step(np.array([1, 2, 3]))
            _, _, r = osp.step_batched(before[0], before[1], inner.lm, np.array([[1, 2, 3]]))
            assert reward == float(r[0, 0])                                          # agent 0's reward only
            assert done is False and truncated == (t == 24)
        with pytest.raises(RuntimeError):
            PettingZooWrapper("mpe", "simple_tag_v3")
    finally:
        sys.path.remove(stub)
        for k in [k for k in sys.modules if k == "env" or k.startswith("env.")]:
            del sys.modules[k]
        sys.modules.update(saved)
