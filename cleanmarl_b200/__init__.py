"""cleanmarl_b200 -- B200-native (sm_100a) kernels for CleanMARL's MAPPO multi-env training path.

Scope: the hot path of ``cleanmarl/mappo_multienvs.py`` (rollout on a device-resident
simple_spread_v3, TD(lambda) returns, clipped-PPO epochs, Adam) and its IPPO variant.
Everything numerical runs in ``libcmarl_b200.so`` (hand-written CUDA, C ABI in
``include/cmarl_b200.h``); importing the engine without that library raises.
"""
from .engine import Engine, Shapes  # noqa: F401

__all__ = ["Engine", "Shapes"]
