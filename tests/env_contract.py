"""Shared body of the env duck-type tests (``cleanmarl/env/common_interface.py:5-23`` with a leading env axis): run on the
CPU double of the engine by tests/test_host_cpu.py and on the CUDA library by tests/test_gpu_vecenv.py."""
import numpy as np
import torch

from oracle import spread as osp


def unpack(env_tensor):
    e = env_tensor.cpu().numpy()
    B = e.shape[1]
    return e[0:6].T.reshape(B, 3, 2).copy(), e[6:12].T.reshape(B, 3, 2).copy(), e[12:18].T.reshape(B, 3, 2).copy()


def check_env_duck_type(make_env, B, state_tol, obs_tol):
    """``make_env(seed)`` -> a fresh SpreadVecEnv over B envs.  ``state_tol``: float64 env state vs the oracle after every
    step; ``obs_tol``: float32 observations / rewards."""
    env = make_env(5)
    dev = env.engine.device
    assert (env.n_agents, env.get_obs_size(), env.get_state_size(), env.get_action_size()) == (3, 21, 54, 5)
    obs, info = env.reset()
    assert info == {} and obs.shape == (B, 3, 21) and obs.dtype == torch.float32 and obs.device == dev
    avail = env.get_avail_actions()
    assert avail.shape == (B, 3, 5) and avail.dtype == torch.bool and bool(avail.all())
    pos, vel, lm = unpack(env.env)
    assert (np.abs(pos) <= 1).all() and (np.abs(lm) <= 1).all() and not vel.any()          # reset: U(-1,1), v = 0
    ref = osp.observe_batched(pos, vel, lm)
    assert np.abs(obs[..., :18].cpu().numpy() - ref).max() <= obs_tol
    assert torch.equal(obs[..., 18:].cpu(), torch.eye(3).expand(B, 3, 3))                  # pettingzoo_wrapper.py:97
    state = env.get_state()
    assert state.shape == (B, 54) and torch.equal(state.reshape(B, 3, 18), obs[..., :18])   # :95 raw obs, ids excluded
    a = env.sample()
    assert a.shape == (B, 3) and a.device == dev and int(a.min()) >= 0 and int(a.max()) <= 4

    g = torch.Generator().manual_seed(0)
    contacts = 0
    for t in range(25):
        actions = torch.randint(0, 5, (B, 3), generator=g)
        obs, reward, done, truncated, info = env.step(actions.to(dev))
        pos, vel, rew = osp.step_batched(pos, vel, lm, actions.numpy())
        dpos, dvel, dlm = unpack(env.env)
        assert np.abs(dpos - pos).max() <= state_tol and np.abs(dvel - vel).max() <= state_tol
        assert np.array_equal(dlm, lm)
        assert obs.shape == (B, 3, 21) and np.abs(obs[..., :18].cpu().numpy() - osp.observe_batched(pos, vel, lm)).max() <= obs_tol
        assert reward.shape == (B,) and reward.dtype == torch.float32
        assert np.abs(reward.cpu().numpy() - rew[:, 0].astype(np.float32)).max() <= obs_tol * 10   # agent 0's reward (:66), |r| < 10
        assert done is False and truncated == (t == 24) and info == {}                      # :51-52, max_cycles = 25
        contacts += int((rew[:, 0] != rew[:, 1]).sum())
    assert contacts > 0                                                                    # soft contacts were exercised

    # reset: a new episode draws new positions; the same (seed, episode) reproduces them bit for bit
    first = make_env(5)
    o1, _ = first.reset()
    o2, _ = first.reset()
    assert not torch.equal(o1, o2)
    again = make_env(5)
    o3, _ = again.reset()
    assert torch.equal(o1, o3)
    other, _ = make_env(6).reset()
    assert not torch.equal(o1, other)
    o4, _ = again.reset(seed=6)            # reset(seed) re-keys the draws (pettingzoo_wrapper.py:36)
    assert not torch.equal(o4, o2)
    env.close()
