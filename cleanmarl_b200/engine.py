"""Thin torch-tensor front end of the C ABI: one ``Engine`` per GPU.

PyTorch is the host container only (device memory, streams); every computation is a kernel of
libcmarl_b200.so.  All tensors use the device layout of include/cmarl_b200.h
(time-major, feature-major, env-minor); ``to_device_layout`` / ``to_reference_layout`` convert
from/to the reference's batch-major ``RolloutBuffer.get_batch`` tuple (MME:148-157).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib

N_STATS = 8
RAW_OBS = 18


@dataclass
class Shapes:
    n_envs: int
    n_steps: int = 25
    n_agents: int = 3
    obs_dim: int = 21
    state_dim: int = 54
    n_actions: int = 5
    actor_hidden: int = 32
    actor_layers: int = 1
    critic_hidden: int = 64
    critic_layers: int = 1
    critic_on_obs: bool = False
    actor_recurrent: bool = False     # fc1 + GRUCell + fc2 (mappo_lstm_multienvs.py:162-184)
    n_landmarks: int = 0              # 0 = n_agents (simple_spread_v3(N): N agents, N landmarks)

    @staticmethod
    def spread(n_envs, n_agents=3, agent_ids=True, n_landmarks=0, **kw):
        """Shapes of simple_spread with N agents / L landmarks: raw observation R = 4 + 2 L + 4 (N - 1) (vel, pos, landmarks -
        pos, other agents - pos, 2 silent communication slots per other agent), O = R (+ N ids), S = N R."""
        L = n_landmarks or n_agents
        R = 4 + 2 * L + 4 * (n_agents - 1)
        return Shapes(n_envs=n_envs, n_agents=n_agents, n_landmarks=L, obs_dim=R + (n_agents if agent_ids else 0),
                      state_dim=n_agents * R, **kw)

    @property
    def env_rows(self):
        """rows of the f64 env state [rows][B]: agent positions, agent velocities, landmark positions"""
        return 4 * self.n_agents + 2 * (self.n_landmarks or self.n_agents)


def _ptr(t, dtype, device, name):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if t.device != device:
        raise ValueError(f"{name}: tensor on {t.device}, engine on {device} (no CPU path exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: dtype {t.dtype}, expected {dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, shapes: Shapes, device: int | torch.device = 0, tensor_cores: bool | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("cleanmarl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        self.shapes = shapes
        cfg = _lib.Config(self.device.index, shapes.n_envs, shapes.n_steps, shapes.n_agents, shapes.obs_dim,
                          shapes.state_dim, shapes.n_actions, shapes.actor_hidden, shapes.actor_layers,
                          shapes.critic_hidden, shapes.critic_layers, int(shapes.critic_on_obs),
                          int(shapes.actor_recurrent), int(shapes.n_landmarks))
        h = C.c_void_p()
        _lib.check(self.lib.cmarl_ctx_create(C.byref(cfg), C.byref(h)), "cmarl_ctx_create")
        self._h = h
        if tensor_cores is None:
            tensor_cores = os.environ.get("CMARL_TENSOR_CORES", "1") != "0"
        self.tensor_cores = bool(tensor_cores)
        _lib.check(self.lib.cmarl_ctx_set_tensor_cores(h, int(self.tensor_cores)), "cmarl_ctx_set_tensor_cores")
        self.n_actor = self.lib.cmarl_actor_param_count(h)
        self.n_critic = self.lib.cmarl_critic_param_count(h)
        self.n_params = self.n_actor + self.n_critic
        self.n_heads = self.lib.cmarl_value_heads(h)
        self.workspace = torch.empty(self.lib.cmarl_workspace_bytes(h) // 4, dtype=torch.float32, device=self.device)
        # default shapes, MLP actor: the partial reduction and the Adam step can share one launch (reduce_clip_adam_step)
        self.fused_update = (not shapes.actor_recurrent and shapes.n_agents == 3 and (shapes.n_landmarks or 3) == 3
                             and shapes.actor_layers == 1 and shapes.critic_layers == 1
                             and shapes.actor_hidden in (32, 64) and shapes.critic_hidden in (32, 64)
                             and os.environ.get("CMARL_FUSED_UPDATE", "0") == "1")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cmarl_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    @property
    def launches(self) -> int:
        return self.lib.cmarl_launch_count(self._h)

    def timing(self, on: bool):
        _lib.check(self.lib.cmarl_timing_enable(self._h, int(on)), "cmarl_timing_enable")

    def read_timing(self) -> dict:
        """{kernel name: (total ms, launches)} since the last read (synchronises the device)."""
        nk = _lib.N_KERNEL_IDS
        ms = (C.c_double * nk)()
        cnt = (C.c_int64 * nk)()
        _lib.check(self.lib.cmarl_timing_read(self._h, ms, cnt), "cmarl_timing_read")
        return {self.lib.cmarl_kernel_name(k).decode(): (ms[k], cnt[k]) for k in range(nk) if cnt[k]}

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f(self, t, name):
        return _ptr(t, torch.float32, self.device, name)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def alloc_rollout(self, with_obs=False):
        s = self.shapes
        T, B, N = s.n_steps, s.n_envs, s.n_agents
        buf = {
            "state": self.empty(T, s.state_dim, B),
            "actions": self.empty(T, N, B, dtype=torch.int32),
            "logp": self.empty(T, N, B),
            "reward": self.empty(T, B),
            "ep_return": self.empty(B, dtype=torch.float64),
            "values": self.empty(T, self.n_heads, B),
            "returns": self.empty(T, self.n_heads, B),
            "adv": self.empty(T, self.n_heads, B),
            "obs": self.empty(T, N, s.obs_dim, B) if with_obs else None,
        }
        return buf

    # ------------------------------------------------------------------ entries
    def set_weight_decay(self, actor_wd: float, critic_wd: float):
        """AdamW's decoupled weight decay for the Adam entries (0 = Adam)."""
        _lib.check(self.lib.cmarl_ctx_set_weight_decay(self._h, float(actor_wd), float(critic_wd)),
                   "cmarl_ctx_set_weight_decay")

    def set_launch_chaining(self, on: bool):
        """Programmatic dependent launches between consecutive kernels of this context (cmarl_ctx_set_launch_chaining):
        on only between two launches of the library on one stream, off before foreign work is enqueued."""
        _lib.check(self.lib.cmarl_ctx_set_launch_chaining(self._h, int(bool(on))), "cmarl_ctx_set_launch_chaining")

    def comm_setup(self, rank: int, world: int, group=None) -> bool:
        """Peer-memory gradient exchange: allocate + export this rank's block, gather every rank's IPC handle through
        ``torch.distributed`` and map the peers.  Afterwards the Adam entries exchange the gradients themselves.
        Returns False (after undoing the setup on every rank) if any rank could not map its peers -- e.g. GPUs without
        peer access -- so that the caller can use the NCCL transport instead; every step is a collective decision."""
        import torch.distributed as dist
        ok = 1
        handle = (C.c_uint8 * 64)()
        try:
            _lib.check(self.lib.cmarl_comm_create(self._h, handle), "cmarl_comm_create")
        except _lib.CmarlError as e:
            ok, err = 0, str(e)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        if ok:
            try:
                blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
                _lib.check(self.lib.cmarl_comm_attach(self._h, rank, world, blob), "cmarl_comm_attach")
            except _lib.CmarlError as e:
                ok, err = 0, str(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.lib.cmarl_comm_detach(self._h)
            if not ok:
                import sys
                print(f"cleanmarl_b200: peer-memory exchange unavailable on rank {rank} ({err}); using NCCL", file=sys.stderr)
            return False
        self.comm_world = world
        return True

    def set_episode_counter(self, counter):
        """``counter``: int64 device tensor [1] (or None) -- see cmarl_ctx_set_episode_counter."""
        ptr = None if counter is None else _ptr(counter, torch.int64, self.device, "episode counter")
        self._episode_counter = counter            # keep the tensor alive
        _lib.check(self.lib.cmarl_ctx_set_episode_counter(self._h, ptr), "cmarl_ctx_set_episode_counter")

    def episode_advance(self):
        _lib.check(self.lib.cmarl_episode_advance(self._h, self._stream()), "cmarl_episode_advance")

    def env_reset(self, env, seed: int, episode: int):
        _lib.check(self.lib.cmarl_env_reset(self._h, _ptr(env, torch.float64, self.device, "env"),
                                            seed & (2**64 - 1), episode & (2**64 - 1), self._stream()), "cmarl_env_reset")

    def env_observe(self, env, state_out):
        _lib.check(self.lib.cmarl_env_observe(self._h, _ptr(env, torch.float64, self.device, "env"),
                                              self._f(state_out, "state_out"), self._stream()), "cmarl_env_observe")

    def env_step(self, env, actions, state_out=None, reward_out=None):
        _lib.check(self.lib.cmarl_env_step(self._h, _ptr(env, torch.float64, self.device, "env"),
                                           _ptr(actions, torch.int32, self.device, "actions"),
                                           self._f(state_out, "state_out"), self._f(reward_out, "reward_out"),
                                           self._stream()), "cmarl_env_step")

    def rollout(self, actor_params, env, state, actions, logp, reward, *, noise=None, obs=None, ep_return=None,
                seed=0, episode=0):
        _lib.check(self.lib.cmarl_rollout(
            self._h, self._f(actor_params, "actor_params"), _ptr(env, torch.float64, self.device, "env"),
            self._f(noise, "noise"), seed & (2**64 - 1), episode & (2**64 - 1), self._f(state, "state"),
            self._f(obs, "obs"), _ptr(actions, torch.int32, self.device, "actions"), self._f(logp, "logp"),
            self._f(reward, "reward"), _ptr(ep_return, torch.float64, self.device, "ep_return"), self._stream()),
            "cmarl_rollout")

    def actor_act(self, actor_params, obs, noise, actions, logp, *, avail=None, logits=None):
        _lib.check(self.lib.cmarl_actor_act(
            self._h, self._f(actor_params, "actor_params"), self._f(obs, "obs"),
            _ptr(avail, torch.uint8, self.device, "avail"), self._f(noise, "noise"),
            _ptr(actions, torch.int32, self.device, "actions"), self._f(logp, "logp"), self._f(logits, "logits"),
            self._stream()), "cmarl_actor_act")

    def critic_values(self, critic_params, values, *, state=None, obs=None):
        _lib.check(self.lib.cmarl_critic_values(self._h, self._f(critic_params, "critic_params"),
                                                self._f(state, "state"), self._f(obs, "obs"),
                                                self._f(values, "values"), self._stream()), "cmarl_critic_values")

    def td_lambda(self, values, reward, returns, adv, gamma, lam, *, mask=None):
        _lib.check(self.lib.cmarl_td_lambda(self._h, self._f(values, "values"), self._f(reward, "reward"),
                                            _ptr(mask, torch.uint8, self.device, "mask"), float(gamma), float(lam),
                                            self._f(returns, "returns"), self._f(adv, "adv"), self._stream()),
                   "cmarl_td_lambda")

    def normalize(self, x, n_heads, mode, phase, stats, *, mask=None):
        _lib.check(self.lib.cmarl_normalize(self._h, self._f(x, "x"), n_heads,
                                            _ptr(mask, torch.uint8, self.device, "mask"), mode, phase,
                                            _ptr(stats, torch.float64, self.device, "stats"), self._stream()),
                   "cmarl_normalize")

    def ppo_epoch_grads(self, params, grads, *, state=None, obs=None, actions, logp_old, adv, returns, mask=None,
                        avail=None, clip=0.2, ent_coef=0.001, value_clip=-1.0, values_old=None, env_begin=0,
                        env_count=None):
        """``value_clip`` / ``values_old`` and the env block ``[env_begin, env_begin + env_count)`` are the two default-off
        extensions of cmarl_ppo_epoch_grads_ex (not in the reference); without them this is cmarl_ppo_epoch_grads.
        ``grads=None`` leaves the chain kernels' partial rows in the workspace for ``reduce_clip_adam_step``."""
        if grads is not None and value_clip <= 0 and env_begin == 0 and env_count in (None, self.shapes.n_envs):
            _lib.check(self.lib.cmarl_ppo_epoch_grads(
                self._h, self._f(params, "params"), self._f(state, "state"), self._f(obs, "obs"),
                _ptr(actions, torch.int32, self.device, "actions"), self._f(logp_old, "logp_old"), self._f(adv, "adv"),
                self._f(returns, "returns"), _ptr(mask, torch.uint8, self.device, "mask"),
                _ptr(avail, torch.uint8, self.device, "avail"), float(clip), float(ent_coef), self._f(grads, "grads"),
                C.c_void_p(self.workspace.data_ptr()), self._stream()), "cmarl_ppo_epoch_grads")
            return
        _lib.check(self.lib.cmarl_ppo_epoch_grads_ex(
            self._h, self._f(params, "params"), self._f(state, "state"), self._f(obs, "obs"),
            _ptr(actions, torch.int32, self.device, "actions"), self._f(logp_old, "logp_old"), self._f(adv, "adv"),
            self._f(returns, "returns"), self._f(values_old, "values_old"), _ptr(mask, torch.uint8, self.device, "mask"),
            _ptr(avail, torch.uint8, self.device, "avail"), float(clip), float(ent_coef), float(value_clip),
            int(env_begin), int(self.shapes.n_envs - env_begin if env_count is None else env_count),
            self._f(grads, "grads"), C.c_void_p(self.workspace.data_ptr()), self._stream()), "cmarl_ppo_epoch_grads_ex")

    def reduce_clip_adam_step(self, params, grads, exp_avg, exp_avg_sq, *, step=1, step_dev=None, lr_actor=8e-4,
                              lr_critic=8e-4, beta1=0.9, beta2=0.999, eps=1e-8, max_norm=-1.0, stats_out=None):
        """cmarl_reduce_clip_adam_step: the reduction of the partial rows left by ``ppo_epoch_grads(grads=None)`` and the
        Adam step in one launch; ``grads`` receives the reduced sums."""
        _lib.check(self.lib.cmarl_reduce_clip_adam_step(
            self._h, C.c_void_p(self.workspace.data_ptr()), self._f(params, "params"), self._f(grads, "grads"),
            self._f(exp_avg, "exp_avg"), self._f(exp_avg_sq, "exp_avg_sq"), int(step),
            _ptr(step_dev, torch.int32, self.device, "step_dev"), float(lr_actor), float(lr_critic), float(beta1),
            float(beta2), float(eps), float(max_norm), self._f(stats_out, "stats_out"), self._stream()),
            "cmarl_reduce_clip_adam_step")

    def clip_adam_step(self, params, grads, exp_avg, exp_avg_sq, *, step=1, step_dev=None, lr_actor=8e-4,
                       lr_critic=8e-4, beta1=0.9, beta2=0.999, eps=1e-8, max_norm=-1.0, stats_out=None):
        _lib.check(self.lib.cmarl_clip_adam_step(
            self._h, self._f(params, "params"), self._f(grads, "grads"), self._f(exp_avg, "exp_avg"),
            self._f(exp_avg_sq, "exp_avg_sq"), int(step), _ptr(step_dev, torch.int32, self.device, "step_dev"),
            float(lr_actor), float(lr_critic), float(beta1), float(beta2), float(eps), float(max_norm),
            self._f(stats_out, "stats_out"), self._stream()), "cmarl_clip_adam_step")


    # ------------------------------------------------------------------ recurrent-actor entries
    def alloc_h_seq(self):
        """f32 [T+1][N][H][B]: hidden state before every step of one epoch (slice 0 = zeros, never read)."""
        s = self.shapes
        return torch.zeros(s.n_steps + 1, s.n_agents, s.actor_hidden, s.n_envs, dtype=torch.float32, device=self.device)

    def actor_act_recurrent(self, actor_params, obs, noise, actions, logp, h_out, *, h_in=None, avail=None, logits=None):
        _lib.check(self.lib.cmarl_actor_act_recurrent(
            self._h, self._f(actor_params, "actor_params"), self._f(obs, "obs"), self._f(h_in, "h_in"),
            _ptr(avail, torch.uint8, self.device, "avail"), self._f(noise, "noise"),
            _ptr(actions, torch.int32, self.device, "actions"), self._f(logp, "logp"), self._f(logits, "logits"),
            self._f(h_out, "h_out"), self._stream()), "cmarl_actor_act_recurrent")

    def alloc_gate_stash(self):
        """f32 [T][N][5H][B]: forward-pass gate activations for the backward pass (optional, see cmarl_tbptt_chunk_grads)."""
        s = self.shapes
        return torch.empty(s.n_steps, s.n_agents, 5 * s.actor_hidden, s.n_envs, dtype=torch.float32, device=self.device)

    def tbptt_chunk_grads(self, actor_params, grads, h_seq, t0, t1, *, state=None, obs=None, actions, logp_old, adv,
                          mask=None, avail=None, clip=0.2, ent_coef=0.001, stash=None):
        _lib.check(self.lib.cmarl_tbptt_chunk_grads(
            self._h, self._f(actor_params, "actor_params"), self._f(state, "state"), self._f(obs, "obs"),
            _ptr(actions, torch.int32, self.device, "actions"), self._f(logp_old, "logp_old"), self._f(adv, "adv"),
            _ptr(mask, torch.uint8, self.device, "mask"), _ptr(avail, torch.uint8, self.device, "avail"),
            float(clip), float(ent_coef), int(t0), int(t1), self._f(h_seq, "h_seq"), self._f(stash, "stash"),
            self._f(grads, "grads"),
            C.c_void_p(self.workspace.data_ptr()), self._stream()), "cmarl_tbptt_chunk_grads")

    def critic_epoch_grads(self, critic_params, grads, *, state=None, obs=None, returns, mask=None):
        _lib.check(self.lib.cmarl_critic_epoch_grads(
            self._h, self._f(critic_params, "critic_params"), self._f(state, "state"), self._f(obs, "obs"),
            self._f(returns, "returns"), _ptr(mask, torch.uint8, self.device, "mask"), self._f(grads, "grads"),
            C.c_void_p(self.workspace.data_ptr()), self._stream()), "cmarl_critic_epoch_grads")

    def adam_step_net(self, net, params, grads, exp_avg, exp_avg_sq, *, step=1, step_dev=None, lr=8e-4, beta1=0.9,
                      beta2=0.999, eps=1e-8, max_norm=-1.0, extra_div=1.0, stats_out=None):
        _lib.check(self.lib.cmarl_adam_step_net(
            self._h, int(net), self._f(params, "params"), self._f(grads, "grads"), self._f(exp_avg, "exp_avg"),
            self._f(exp_avg_sq, "exp_avg_sq"), int(step), _ptr(step_dev, torch.int32, self.device, "step_dev"),
            float(lr), float(beta1), float(beta2), float(eps), float(max_norm), float(extra_div),
            self._f(stats_out, "stats_out"), self._stream()), "cmarl_adam_step_net")


# ---------------------------------------------------------------------- layout conversion
def to_device_layout(batch, device, with_obs=True):
    """Reference 8-tuple (MME:148-157, batch-major) -> dict of device-layout tensors."""
    obs, actions, logp, reward, states, avail, done, mask = batch
    out = {
        "state": states.permute(1, 2, 0).contiguous().float().to(device),               # [T][S][B]
        "actions": actions.permute(1, 2, 0).contiguous().to(torch.int32).to(device),    # [T][N][B]
        "logp": logp.permute(1, 2, 0).contiguous().float().to(device),
        "reward": reward.permute(1, 0).contiguous().float().to(device),                 # [T][B]
        "mask": mask.permute(1, 0).contiguous().to(torch.uint8).to(device),
        "avail": avail.permute(1, 2, 3, 0).contiguous().to(torch.uint8).to(device),     # [T][N][A][B]
        "done": done.permute(1, 0).contiguous().float().to(device),
    }
    if with_obs:
        out["obs"] = obs.permute(1, 2, 3, 0).contiguous().float().to(device)            # [T][N][O][B]
    return out


def heads_to_device(x, n_heads, device):
    """[B,T,N] (reference returns/advantages) -> [T][V][B]; V=1 keeps agent 0 (all agents equal, MME:484-485)."""
    x = x.permute(1, 2, 0)
    if n_heads == 1:
        x = x[:, :1]
    return x.contiguous().float().to(device)


def heads_to_reference(x, n_agents):
    """[T][V][B] -> [B,T,N] (broadcast the centralised head to the agents)."""
    T, V, B = x.shape
    x = x.permute(2, 0, 1)
    if V == 1:
        x = x.expand(B, T, n_agents)
    return x.contiguous()


def obs_from_state(state, n_agents=3, agent_ids=True):
    """[T][S][B] -> [T][N][O][B]: raw rows + one-hot ids (pettingzoo_wrapper.py:93-98)."""
    T, S, B = state.shape
    raw = state.reshape(T, n_agents, S // n_agents, B)
    if not agent_ids:
        return raw.contiguous()
    ids = torch.eye(n_agents, device=state.device, dtype=state.dtype)[None, :, :, None].expand(T, n_agents, n_agents, B)
    return torch.cat([raw, ids], dim=2).contiguous()


def to_reference_layout(buf, n_agents=3, n_actions=5, agent_ids=True):
    """Device rollout buffers -> the reference's ``get_batch`` 8-tuple (dtypes of MME:148-157)."""
    state = buf["state"]
    T, S, B = state.shape
    obs = buf.get("obs")
    if obs is None:
        obs = obs_from_state(state, n_agents, agent_ids)
    mask = buf.get("mask")
    avail = buf.get("avail")
    done = buf.get("done")
    return (
        obs.permute(3, 0, 1, 2).contiguous().float(),
        buf["actions"].permute(2, 0, 1).contiguous().long(),
        buf["logp"].permute(2, 0, 1).contiguous().float(),
        buf["reward"].permute(1, 0).contiguous().float(),
        state.permute(2, 0, 1).contiguous().float(),
        (avail.permute(3, 0, 1, 2).contiguous().bool() if avail is not None
         else torch.ones(B, T, n_agents, n_actions, dtype=torch.bool, device=state.device)),
        (done.permute(1, 0).contiguous().float() if done is not None
         else torch.zeros(B, T, device=state.device)),
        (mask.permute(1, 0).contiguous().bool() if mask is not None
         else torch.ones(B, T, dtype=torch.bool, device=state.device)),
    )
