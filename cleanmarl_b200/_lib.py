"""ctypes binding of libcmarl_b200.so (C ABI declared in include/cmarl_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("CMARL_B200_LIB", PKG / "libcmarl_b200.so"))


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "device", "n_envs", "n_steps", "n_agents", "obs_dim", "state_dim", "n_actions",
        "actor_hidden", "actor_layers", "critic_hidden", "critic_layers", "critic_on_obs", "actor_recurrent", "n_landmarks")]


class CmarlError(RuntimeError):
    pass


_P = C.c_void_p
_SIGNATURES = {
    "cmarl_version": (C.c_int, []),
    "cmarl_last_error": (C.c_char_p, []),
    "cmarl_ctx_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "cmarl_ctx_destroy": (C.c_int, [_P]),
    "cmarl_ctx_set_tensor_cores": (C.c_int, [_P, C.c_int]),
    "cmarl_ctx_set_launch_chaining": (C.c_int, [_P, C.c_int]),
    "cmarl_ctx_set_weight_decay": (C.c_int, [_P, C.c_double, C.c_double]),
    "cmarl_actor_param_count": (C.c_int, [_P]),
    "cmarl_critic_param_count": (C.c_int, [_P]),
    "cmarl_value_heads": (C.c_int, [_P]),
    "cmarl_workspace_bytes": (C.c_size_t, [_P]),
    "cmarl_launch_count": (C.c_int, [_P]),
    "cmarl_timing_enable": (C.c_int, [_P, C.c_int]),
    "cmarl_timing_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "cmarl_kernel_name": (C.c_char_p, [C.c_int]),
    "cmarl_ctx_set_episode_counter": (C.c_int, [_P, _P]),
    "cmarl_episode_advance": (C.c_int, [_P, _P]),
    "cmarl_env_reset": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "cmarl_env_observe": (C.c_int, [_P, _P, _P, _P]),
    "cmarl_env_step": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "cmarl_rollout": (C.c_int, [_P, _P, _P, _P, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P, _P, _P]),
    "cmarl_actor_act": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cmarl_critic_values": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "cmarl_td_lambda": (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_double, _P, _P, _P]),
    "cmarl_normalize": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P]),
    "cmarl_ppo_epoch_grads": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, _P, _P, _P]),
    "cmarl_ppo_epoch_grads_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double,
                                           C.c_int32, C.c_int32, _P, _P, _P]),
    "cmarl_clip_adam_step": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, _P, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_double, C.c_double, _P, _P]),
    "cmarl_reduce_clip_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, _P, C.c_double, C.c_double, C.c_double,
                                              C.c_double, C.c_double, C.c_double, _P, _P]),
    # peer-memory gradient exchange
    "cmarl_comm_bytes": (C.c_size_t, []),
    "cmarl_comm_create": (C.c_int, [_P, _P]),
    "cmarl_comm_attach": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "cmarl_comm_detach": (C.c_int, [_P]),
    # recurrent-actor path (mappo_lstm_multienvs.py)
    "cmarl_actor_act_recurrent": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cmarl_tbptt_chunk_grads": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_int32,
                                          C.c_int32, _P, _P, _P, _P, _P]),
    "cmarl_critic_epoch_grads": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cmarl_adam_step_net": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.c_int32, _P, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_double, C.c_double, _P, _P]),
}

EXPORTS = tuple(_SIGNATURES)
VERSION = 104          # CMARL_VERSION of include/cmarl_b200.h
N_KERNEL_IDS = 12      # CMARL_NK

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and declare every entry point of include/cmarl_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.is_file():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C cleanmarl_b200/csrc`). cleanmarl_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.cmarl_version() != VERSION:
        raise ImportError(f"libcmarl_b200.so version {lib.cmarl_version()} != {VERSION} (stale build)")
    _lib = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load().cmarl_last_error().decode(errors="replace")
        raise CmarlError(f"{what} failed with status {code}: {msg}")
