"""GPU test of the env duck-type: ``SpreadVecEnv`` over ``cmarl_env_reset`` / ``cmarl_env_observe`` /
``cmarl_env_step`` (one thread per env, the physics of csrc/spread.cuh that the rollout kernel also uses) against
oracle/spread.py, one call at a time like the reference's ``env.reset() / env.step()`` (pettingzoo_wrapper.py:32-66)."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))

pytestmark = pytest.mark.gpu


def test_env_duck_type_vs_oracle():
    """float64 env state <= 1e-10 after each of 25 steps (device exp / log1p differ from numpy's in the last place and the
    contact force is stiff), float32 observations <= 2e-6, rewards <= 2e-5; ragged env count (one partial block)."""
    import cleanmarl_b200 as cm
    from cleanmarl_b200 import _lib
    from cleanmarl_b200.mappo import SpreadVecEnv
    from env_contract import check_env_duck_type
    _lib.load()
    B = 1000
    check_env_duck_type(lambda seed: SpreadVecEnv(cm.Engine(cm.Shapes(n_envs=B), device=0), agent_ids=True, seed=seed), B,
                        state_tol=1e-10, obs_tol=2e-6)
