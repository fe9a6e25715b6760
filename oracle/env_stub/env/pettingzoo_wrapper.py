"""Stand-in for ``cleanmarl/env/pettingzoo_wrapper.py`` without gymnasium/pettingzoo.

Restates the adapter's own (verifiable) logic on top of ``oracle.spread``:
obs = raw per-agent observations with one-hot agent ids appended when
``agent_ids`` (reference ``pettingzoo_wrapper.py:93-98``; the ``np.eye`` concat
makes the array float64), state = the raw observations flattened (``:95``),
reward = agent 0's reward (``:66``), done/truncated = all() over agents
(``:51-52``), avail = ones (``:79-90``).
"""
import numpy as np

from oracle import spread
from .common_interface import CommonInterface


class PettingZooWrapper(CommonInterface):
    def __init__(self, family, env_name, agent_ids=False, **kwargs):
        if (family, env_name) != ("mpe", "simple_spread_v3"):
            raise RuntimeError(f"stub env only provides mpe/simple_spread_v3, got {family}/{env_name}")
        self.env = spread.parallel_env(**kwargs)
        self.env.reset()
        self.n_agents = self.env.num_agents
        self.agents = self.env.agents
        self.n_actions = self.env.action_space(self.agents[0]).n
        self.raw_obs = self.env.observation_space(self.agents[0]).shape[0]
        self.agent_ids = agent_ids

    def process_obs(self, obs):
        obs = np.array([obs[a].flatten() for a in self.agents])
        self.state = obs.reshape(-1)
        if self.agent_ids:
            obs = np.concatenate((obs, np.eye(self.n_agents)), axis=1)
        return obs

    def reset(self, seed=None):
        obs, _ = self.env.reset(seed=seed)
        obs = self.process_obs(obs)
        self.last_obs = obs
        return obs, {}

    def step(self, actions):
        acts = {a: actions[i].item() for i, a in enumerate(self.agents)}
        observations, rewards, dones, truncated, infos = self.env.step(acts)
        obs = self.process_obs(observations)
        rewards = [rewards[a] for a in self.agents]
        done = all(dones[a] for a in self.agents)
        truncated = all(truncated[a] for a in self.agents)
        self.last_obs = obs
        return obs, rewards[0], done, truncated, {}

    def get_obs_size(self):
        return self.raw_obs + self.agent_ids * self.n_agents

    def get_state_size(self):
        return self.raw_obs * self.n_agents

    def get_state(self):
        return self.state

    def get_action_size(self):
        return self.n_actions

    def get_avail_actions(self):
        return np.array([[1] * self.n_actions for _ in range(self.n_agents)])

    def sample(self):
        return [self.env.action_space(a).sample() for a in self.agents]

    def close(self):
        return self.env.close()
