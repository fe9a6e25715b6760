"""Host-side mirror of the reference's in-file objects for the MAPPO / IPPO multi-env path.

Reference: ``cleanmarl/mappo_multienvs.py`` (MME) and ``cleanmarl/ippo_multienvs.py``.
  Args            <- MME:18-79 (same fields, same defaults except ``device``)
  SpreadVecEnv    <- env duck-type ``cleanmarl/env/common_interface.py:5-23`` with a leading env axis,
                     backed by the device simple_spread_v3 kernel instead of one process per env
                     (MME:246-285, 299-319)
  ActorCritic     <- ``Actor`` / ``Critic`` (MME:160-200) as one flat device parameter vector with the
                     reference's initialisation (MME:291-294, 329-339)
  MAPPO           <- the body of the ``while step < total_timesteps`` loop (MME:379-612)

Only orchestration lives here; every number is produced by a kernel of libcmarl_b200.so.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.nn as nn

from .engine import Engine, Shapes, obs_from_state, to_reference_layout


@dataclass
class Args:
    env_type: str = "pz"
    """ Pettingzoo only (the reference also lists smaclite / lbf; not built here) """
    env_name: str = "simple_spread_v3"
    """ Name of the environment"""
    env_family: str = "mpe"
    """ Env family when using pz"""
    agent_ids: bool = True
    """ Include id (one-hot vector) at the agent of the observations"""
    batch_size: int = 3
    """ Number of episodes to collect in each rollout (= number of parallel device envs)"""
    actor_hidden_dim: int = 32
    """ Hidden dimension of actor network"""
    actor_num_layers: int = 1
    """ Number of hidden layers of actor network"""
    critic_hidden_dim: int = 64
    """ Hidden dimension of critic network"""
    critic_num_layers: int = 1
    """ Number of hidden layers of critic network"""
    optimizer: str = "Adam"
    """ The optimizer"""
    learning_rate_actor: float = 0.0008
    """ Learning rate for the actor"""
    learning_rate_critic: float = 0.0008
    """ Learning rate for the critic"""
    total_timesteps: int = 1000000
    """ Total steps in the environment during training"""
    gamma: float = 0.99
    """ Discount factor"""
    td_lambda: float = 0.95
    """ TD(lambda) discount factor"""
    normalize_reward: bool = False
    """ Normalize the rewards if True"""
    normalize_advantage: bool = False
    """ Normalize the advantage if True"""
    normalize_return: bool = False
    """ Normalize the returns if True"""
    epochs: int = 3
    """ Number of training epochs"""
    ppo_clip: float = 0.2
    """ PPO clipping factor """
    entropy_coef: float = 0.001
    """ Entropy coefficient """
    clip_gradients: float = -1
    """ 0< for no clipping and 0> if clipping at clip_gradients"""
    log_every: int = 10
    """ Logging steps """
    eval_steps: int = 50
    """ Evaluate the policy each eval_steps training steps"""
    num_eval_ep: int = 10
    """ Number of evaluation episodes"""
    use_wnb: bool = False
    """ Logging to Weights & Biases if True"""
    wnb_project: str = ""
    """ Weights & Biases project name"""
    wnb_entity: str = ""
    """ Weights & Biases entity name"""
    device: str = "cuda"
    """ Device: cuda only (the reference defaults to cpu; this implementation has no CPU path)"""
    seed: int = 1
    """ Random seed"""
    value_clip: float = -1
    """ (beyond the reference, default off) 0< : PPO2-style clipped value loss with this range; 0> : the reference's MSE"""
    num_minibatches: int = 1
    """ (beyond the reference, default 1 = full batch) optimizer steps per epoch, one per contiguous block of envs"""
    n_agents: int = 3
    """ (beyond the reference, which passes no env kwargs: MME:297) simple_spread_v3(N=n_agents): N agents and N landmarks"""


# the two fields above are the options BASELINE.json's north_star names that the reference does not have
EXTENSION_FIELDS = ("value_clip", "num_minibatches", "n_agents")


@dataclass
class ArgsRecurrent(Args):
    """``Args`` of ``cleanmarl/mappo_lstm_multienvs.py:18-81``: MME's plus ``tbptt``; ``num_eval_ep`` defaults to 5.
    (``actor_num_layers`` is accepted and unused, as in the reference: its recurrent ``Actor`` takes no layer count.)"""
    tbptt: int = 10
    """Chunck size for Truncated Backpropagation Through Time tbptt"""
    num_eval_ep: int = 5
    """ Number of evaluation episodes"""


@dataclass
class ArgsRecurrentIPPO(ArgsRecurrent):
    """``Args`` of ``cleanmarl/ippo_lstm_multienvs.py:18-81``: decentralised critic (hidden 32), AdamW, tbptt 5."""
    critic_hidden_dim: int = 32
    """ Hidden dimension of critic network"""
    optimizer: str = "AdamW"
    """ The optimizer"""
    tbptt: int = 5
    """Chunck size for Truncated Backpropagation Through Time tbptt"""
    num_eval_ep: int = 10
    """ Number of evaluation episodes"""


def validate_args(args: Args):
    """Fail loudly at start-up on anything the device path does not implement (no fallbacks)."""
    if args.env_type != "pz" or args.env_family != "mpe" or args.env_name != "simple_spread_v3":
        raise SystemExit(f"cleanmarl_b200 only implements --env_type pz --env_family mpe --env_name simple_spread_v3 "
                         f"(got {args.env_type}/{args.env_family}/{args.env_name})")
    if not str(args.device).startswith("cuda"):
        raise SystemExit(f"--device {args.device}: cleanmarl_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    if args.optimizer not in ("Adam", "AdamW"):
        raise SystemExit("only --optimizer Adam and AdamW are implemented")
    recurrent = hasattr(args, "tbptt")
    # MLP paths: any *_num_layers >= 1 and hidden width <= 256 (MME:160-171, 186-196); the default shapes run the fused
    # kernels, everything else the layered ones (csrc/generic.cu)
    if not (1 <= args.actor_num_layers <= 6 and 1 <= args.critic_num_layers <= 6):
        raise SystemExit("*_num_layers must be in [1, 6]")
    if not (1 <= args.actor_hidden_dim <= 256 and 1 <= args.critic_hidden_dim <= 256):
        raise SystemExit("*_hidden_dim must be in [1, 256]")
    if not 1 <= args.n_agents <= 8:
        raise SystemExit("--n_agents must be in [1, 8]")
    if recurrent and args.tbptt < 1:
        raise SystemExit("--tbptt must be positive")
    if recurrent and (args.actor_hidden_dim != 32 or args.critic_num_layers != 1 or args.n_agents != 3
                      or args.critic_hidden_dim not in (32, 64)):
        raise SystemExit("the recurrent actor is built for the reference's default shapes only (--actor_hidden_dim 32, "
                         "--critic_num_layers 1, --critic_hidden_dim 32 / 64, 3 agents)")
    if args.batch_size < 1 or args.epochs < 1:
        raise SystemExit("batch_size and epochs must be positive")
    if args.num_minibatches < 1 or args.num_minibatches > args.batch_size:
        raise SystemExit("--num_minibatches must be in [1, batch_size]")
    if recurrent and (args.num_minibatches != 1 or args.value_clip > 0):
        raise SystemExit("--num_minibatches / --value_clip are implemented for the MLP actor paths only")


# ---------------------------------------------------------------------------------------------
class ActorCritic:
    """Flat parameter vector [actor | critic] with the reference's init.

    ``torch.manual_seed(seed)`` followed by the ``nn.Linear`` constructors in the order of
    MME:329-339 (actor layers, then critic layers) reproduces the reference's parameters bit for
    bit; layout = ``module.parameters()`` order (W1, b1, W2, b2, W3, b3 per network).
    """

    def __init__(self, engine: Engine, seed: int | None):
        s = engine.shapes
        if seed is not None:
            torch.manual_seed(seed)
        critic_in = s.obs_dim if s.critic_on_obs else s.state_dim
        dims_a = [s.obs_dim] + [s.actor_hidden] * (s.actor_layers + 1) + [s.n_actions]
        dims_c = [critic_in] + [s.critic_hidden] * (s.critic_layers + 1) + [1]
        flat = []
        if getattr(s, "actor_recurrent", False):
            # mappo_lstm_multienvs.py:165-168: fc1 Linear, GRUCell, fc2 Linear -- constructed (= initialised) in that order
            mods = [nn.Linear(s.obs_dim, s.actor_hidden), nn.GRUCell(s.actor_hidden, s.actor_hidden),
                    nn.Linear(s.actor_hidden, s.n_actions)]
            flat += [p.detach().reshape(-1) for m in mods for p in m.parameters()]
            dims_a = []
        for dims in (dims_a, dims_c):
            for i in range(len(dims) - 1):
                lin = nn.Linear(dims[i], dims[i + 1])
                flat += [lin.weight.detach().reshape(-1), lin.bias.detach().reshape(-1)]
        flat = torch.cat(flat)
        assert flat.numel() == engine.n_params
        self.engine = engine
        self.flat = flat.to(engine.device).contiguous()
        self.n_actor = engine.n_actor

    @property
    def actor(self):
        return self.flat[:self.n_actor]

    @property
    def critic(self):
        return self.flat[self.n_actor:]


class SpreadVecEnv:
    """B device-resident simple_spread_v3 envs behind the reference's env duck-type
    (``env/common_interface.py:5-23``), with a leading env axis and tensors that stay in HBM."""

    def __init__(self, engine: Engine, agent_ids: bool = True, seed: int = 0):
        self.engine = engine
        s = engine.shapes
        self.n_agents = s.n_agents
        self.n_envs = s.n_envs
        self.agent_ids = agent_ids
        self.seed = seed
        self.episode = 0
        self.raw_obs = s.state_dim // s.n_agents
        self.env = engine.empty(s.env_rows, s.n_envs, dtype=torch.float64)
        self._state = engine.empty(s.state_dim, s.n_envs)
        self._reward = engine.empty(s.n_envs)
        self.steps = 0

    def get_obs_size(self):
        return self.raw_obs + self.agent_ids * self.n_agents

    def get_state_size(self):
        return self.raw_obs * self.n_agents

    def get_action_size(self):
        return 5

    def get_avail_actions(self):
        return torch.ones(self.n_envs, self.n_agents, 5, dtype=torch.bool, device=self.engine.device)

    def get_state(self):
        return self._state.t()                                   # [B, 54]

    def _obs(self):
        o = obs_from_state(self._state[None], self.n_agents, self.agent_ids)[0]     # [N][O][B]
        return o.permute(2, 0, 1)                                # [B, N, O]

    def reset(self, seed=None):
        if seed is not None:
            self.seed = seed
        self.engine.env_reset(self.env, self.seed, self.episode)
        self.episode += 1
        self.steps = 0
        self.engine.env_observe(self.env, self._state)
        return self._obs(), {}

    def step(self, actions):
        """actions int [B, N] -> (obs [B,N,O], reward [B], done, truncated, info)"""
        a = actions.to(torch.int32).t().contiguous()
        self.engine.env_step(self.env, a, self._state, self._reward)
        self.steps += 1
        return self._obs(), self._reward.clone(), False, self.steps >= self.engine.shapes.n_steps, {}

    def sample(self):
        return torch.randint(0, 5, (self.n_envs, self.n_agents), device=self.engine.device)

    def close(self):
        pass


# ---------------------------------------------------------------------------------------------
class MAPPO:
    """One trainer per GPU (one process per GPU).  ``iteration()`` = one pass of the reference's
    outer loop: rollout (MME:380-453) -> TD(lambda) (MME:484-512) -> ``epochs`` x [loss + grads
    (MME:522-582), gradient all-reduce across GPUs, clip + Adam (MME:584-594)]."""

    def __init__(self, args: Args, device_index: int = 0, rank: int = 0, world_size: int = 1, ippo: bool = False,
                 process_group=None, engine_factory=Engine, use_graph: bool | None = None):
        """``engine_factory(shapes, device_index)`` builds the per-GPU kernel front end; the default (and only
        product) value is ``Engine`` = libcmarl_b200.so.  tests/ pass a CPU test double there to exercise the
        sharding / all-reduce logic under ``gloo`` without a GPU."""
        validate_args(args)
        if args.batch_size % world_size:
            raise SystemExit(f"--batch_size {args.batch_size} must be divisible by the number of GPUs {world_size}")
        self.args, self.rank, self.world, self.ippo, self.pg = args, rank, world_size, ippo, process_group
        self.B = args.batch_size // world_size
        self.T = 25                                              # simple_spread_v3 max_cycles (kwargs = {}, MME:297)
        if args.batch_size * self.T >= 1 << 24:
            # the valid (b, t) count travels through the gradient exchange as ONE fp32 word (exact below 2^24): beyond
            # that the 1/n gradient scale and every logged mean would be silently rounded
            raise SystemExit(f"--batch_size {args.batch_size}: batch_size * {self.T} steps must stay below 2^24 "
                             f"(the sample count is exchanged as an exact fp32 integer)")
        self.recurrent = hasattr(args, "tbptt")
        shapes = Shapes.spread(self.B, args.n_agents, bool(args.agent_ids), n_steps=self.T,
                               actor_hidden=args.actor_hidden_dim, actor_layers=1 if self.recurrent else args.actor_num_layers,
                               critic_hidden=args.critic_hidden_dim, critic_layers=args.critic_num_layers,
                               critic_on_obs=ippo, actor_recurrent=self.recurrent)
        self.engine = eng = engine_factory(shapes, device_index)
        if args.optimizer == "AdamW":
            eng.set_weight_decay(0.01, 0.01)                     # torch.optim.AdamW default weight_decay
        self.net = ActorCritic(eng, args.seed)                   # identical on every rank (same seed)
        self.exp_avg = torch.zeros_like(self.net.flat)
        self.exp_avg_sq = torch.zeros_like(self.net.flat)
        self.adam_step = torch.zeros(1, dtype=torch.int32, device=eng.device)
        self.grads = eng.empty(eng.n_params + 8)
        # the per-step results a caller reads back (per-env episode returns f64 [B], per-epoch statistics f32 [epochs][8])
        # live in ONE device block so that `results_to_host` is a single D2H copy
        # (+ one f64: the sum of the episode returns over ALL ranks, filled by `stage_scalars` on multi-GPU runs)
        self.n_mb = args.num_minibatches
        if self.B < self.n_mb:
            raise SystemExit(f"--num_minibatches {self.n_mb} exceeds the {self.B} envs of one GPU")
        self.mb_stats = eng.empty(args.epochs, self.n_mb, 8) if self.n_mb > 1 else None
        self.results = torch.zeros(self.B * 8 + args.epochs * 32 + 8, dtype=torch.uint8, device=eng.device)
        self.epoch_stats = self.results[self.B * 8:self.B * 8 + args.epochs * 32].view(torch.float32).view(args.epochs, 8)
        self.return_sum = self.results[self.B * 8 + args.epochs * 32:].view(torch.float64)
        self._host_blocks, self._host_turn = None, 0
        self._eval_ctx = {}
        self.norm_stats = eng.empty(4, dtype=torch.float64)
        self.buf = eng.alloc_rollout()
        self.buf["ep_return"] = self.results[:self.B * 8].view(torch.float64)
        if self.recurrent:
            self.chunks = tbptt_chunks(self.T, args.tbptt)
            self.h_seq = eng.alloc_h_seq()
            # forward-pass gate activations for the backward pass (48 KB per env): skipped with CMARL_GATE_STASH=0
            self.stash = (eng.alloc_gate_stash() if hasattr(eng, "alloc_gate_stash")
                          and os.environ.get("CMARL_GATE_STASH", "1") != "0" else None)
            self.grads_a = eng.empty(eng.n_actor + 8)
            self.grads_c = eng.empty(eng.n_critic + 8)
            self.adam_step_a = torch.zeros(1, dtype=torch.int32, device=eng.device)
            self.chunk_stats = eng.empty(args.epochs, len(self.chunks), 8)
            self.critic_stats = eng.empty(args.epochs, 8)
        self.env = eng.empty(shapes.env_rows, self.B, dtype=torch.float64)
        # every rank draws from its own Philox key so shards are independent
        self.rng_key = (args.seed + 0x9E3779B97F4A7C15 * (rank + 1)) & (2**64 - 1)
        self.episode = 0
        self.step = 0                                            # env steps over all GPUs (MME:435)
        self.training_step = 0
        self.num_episodes = 0
        # One iteration is a fixed sequence of ~15 launches: replaying it as a CUDA graph removes the host launch
        # path from the loop.  Needs the device episode counter; single-GPU only (the NCCL all-reduce stays eager).
        if use_graph is None:
            use_graph = os.environ.get("CMARL_GRAPH", "1") != "0"
        # Multi-GPU gradient exchange: "p2p" = inside the Adam kernel over peer memory (NVLink; keeps the iteration a
        # fixed launch sequence, so the graph replay also works across GPUs), "nccl" = torch.distributed.all_reduce
        # between the gradient kernels and Adam (eager launches: NCCL collectives captured into the graph hung when the
        # bench mixed replays with eager collectives -- measured round 1).  The normalisation statistics (3 doubles,
        # flags off by default) always use the process group.
        self.comm = "none"
        if world_size > 1:
            self.comm = os.environ.get("CMARL_COMM", "p2p") if engine_factory is Engine else "nccl"
            if self.comm == "p2p" and not eng.comm_setup(rank, world_size, process_group):
                self.comm = "nccl"
        norm_flags = args.normalize_reward or args.normalize_advantage or args.normalize_return
        graph_ok = world_size == 1 or (self.comm == "p2p" and not norm_flags)
        self.use_graph = bool(use_graph) and graph_ok and engine_factory is Engine
        # Launch chaining: after the first kernel of an iteration every launch is a programmatic dependent of the one in
        # front of it (launch latency and kernel prologues hide under the predecessor; results unchanged).
        self.chain = os.environ.get("CMARL_LAUNCH_CHAIN", "0") != "0"
        self._graphs = {}
        self._episode_dev = None
        self.launches_per_iteration = None                       # kernel nodes of the captured graph (library launches)

    # -- rollout -------------------------------------------------------------------------------
    def collect(self, env_init=None, noise=None, chain=False):
        """MME:380-458.  ``env_init`` f64 [18][B] / ``noise`` f32 [T][N][A][B] make the rollout a
        function of its inputs (parity, end-to-end benchmark); default: device Philox draws.
        ``chain``: switch launch chaining on behind the first kernel (the caller switches it off again)."""
        eng, buf = self.engine, self.buf
        if env_init is None:
            eng.env_reset(self.env, self.rng_key, self.episode)
            eng.set_launch_chaining(chain)
        elif env_init is not self.env:
            self.env.copy_(env_init, non_blocking=True)
        eng.rollout(self.net.actor, self.env, buf["state"], buf["actions"], buf["logp"], buf["reward"], noise=noise,
                    ep_return=buf["ep_return"], seed=self.rng_key, episode=self.episode)
        eng.set_launch_chaining(chain)
        if self._episode_dev is not None:
            eng.episode_advance()
        self.episode += 1
        self.step += self.B * self.T * self.world
        self.num_episodes += self.B * self.world

    def _allreduce(self, t):
        if self.world > 1:
            torch.distributed.all_reduce(t, group=self.pg)

    def _allreduce_grads(self, t):
        """Unnormalised gradient sums (+ statistics): summed by the Adam kernel itself in p2p mode."""
        if self.world > 1 and self.comm != "p2p":
            torch.distributed.all_reduce(t, group=self.pg)

    def _normalize(self, x, heads, mode):
        eng = self.engine
        eng.normalize(x, heads, mode, 0, self.norm_stats)
        self._allreduce(self.norm_stats)                         # global statistics (MME:143-146, 505-512)
        eng.normalize(x, heads, mode, 1, self.norm_stats)

    # -- returns -------------------------------------------------------------------------------
    def advantages(self):
        """MME:143-146 (reward normalisation happens at collate time) then MME:484-512."""
        eng, buf, a = self.engine, self.buf, self.args
        if a.normalize_reward:
            self._normalize(buf["reward"], 1, 0)
        eng.critic_values(self.net.critic, buf["values"], state=buf["state"])
        eng.td_lambda(buf["values"], buf["reward"], buf["returns"], buf["adv"], a.gamma, a.td_lambda)
        if a.normalize_advantage:
            self._normalize(buf["adv"], eng.n_heads, 1)
        if a.normalize_return:
            self._normalize(buf["returns"], eng.n_heads, 1)

    # -- PPO epochs ----------------------------------------------------------------------------
    def update_recurrent(self):
        """mappo_lstm_multienvs.py:551-664: per epoch, one actor step per truncated-BPTT chunk (gradients of the chunk's
        loss / (n_valid_chunk * T_chunk), hidden state carried detached from chunk to chunk) and one critic step; one
        all-reduce of unnormalised sums before every step, so shards add and replicas stay identical."""
        eng, buf, a = self.engine, self.buf, self.args
        na = eng.n_actor
        p_a, p_c = self.net.flat[:na], self.net.flat[na:]
        m_a, m_c = self.exp_avg[:na], self.exp_avg[na:]
        v_a, v_c = self.exp_avg_sq[:na], self.exp_avg_sq[na:]
        for ep in range(a.epochs):
            for ci, (t0, t1) in enumerate(self.chunks):
                eng.tbptt_chunk_grads(p_a, self.grads_a, self.h_seq, t0, t1, state=buf["state"], actions=buf["actions"],
                                      logp_old=buf["logp"], adv=buf["adv"], clip=a.ppo_clip, ent_coef=a.entropy_coef,
                                      stash=self.stash)
                self._allreduce_grads(self.grads_a)
                eng.adam_step_net(0, p_a, self.grads_a, m_a, v_a, step_dev=self.adam_step_a, lr=a.learning_rate_actor,
                                  max_norm=a.clip_gradients, extra_div=t1 - t0, stats_out=self.chunk_stats[ep, ci])
            eng.critic_epoch_grads(p_c, self.grads_c, state=buf["state"], returns=buf["returns"])   # IPPO: obs rebuilt from state
            self._allreduce_grads(self.grads_c)
            eng.adam_step_net(1, p_c, self.grads_c, m_c, v_c, step_dev=self.adam_step, lr=a.learning_rate_critic,
                              max_norm=a.clip_gradients, stats_out=self.critic_stats[ep])
            self.training_step += 1
        # per-epoch scalars of LSTM:640-662 in the layout of the MLP path's epoch_stats (logging only)
        cs, ks = self.chunk_stats, self.critic_stats
        n = cs[:, :, 6].sum(dim=1)                                  # b_mask.sum()
        es = self.epoch_stats
        es[:, 0] = cs[:, :, 0].sum(dim=1) / n
        es[:, 1] = ks[:, 1] / ks[:, 6]
        es[:, 2:5] = cs[:, :, 2:5].sum(dim=1) / n[:, None]
        es[:, 5] = cs[:, :, 5].mean(dim=1)                          # np.mean(actor_gradient), LSTM:662
        es[:, 6] = ks[:, 5]
        es[:, 7] = n

    def update(self):
        """MME:521-603: one NCCL all-reduce of the flat gradient (+8 statistics) per epoch."""
        if self.recurrent:
            return self.update_recurrent()
        eng, buf, a = self.engine, self.buf, self.args
        ext = {}
        if a.value_clip > 0:                                     # beyond the reference: V at rollout time = buf["values"]
            ext = dict(value_clip=a.value_clip, values_old=buf["values"])
        M = self.n_mb
        for ep in range(a.epochs):
            for mb in range(M):
                if M > 1:                                        # minibatch = contiguous block of this GPU's envs
                    lo, hi = mb * self.B // M, (mb + 1) * self.B // M
                    ext.update(env_begin=lo, env_count=hi - lo)
                # one launch for the partial reduction + Adam unless an NCCL all-reduce has to sit between them
                fused = getattr(eng, "fused_update", False) and (self.world == 1 or self.comm == "p2p")
                eng.ppo_epoch_grads(self.net.flat, None if fused else self.grads, state=buf["state"], actions=buf["actions"],
                                    logp_old=buf["logp"], adv=buf["adv"], returns=buf["returns"], clip=a.ppo_clip,
                                    ent_coef=a.entropy_coef, **ext)
                self._allreduce_grads(self.grads)
                (eng.reduce_clip_adam_step if fused else eng.clip_adam_step)(
                    self.net.flat, self.grads, self.exp_avg, self.exp_avg_sq, step_dev=self.adam_step,
                    lr_actor=a.learning_rate_actor, lr_critic=a.learning_rate_critic, max_norm=a.clip_gradients,
                    stats_out=self.epoch_stats[ep] if M == 1 else self.mb_stats[ep, mb])
                self.training_step += 1
        if M > 1:                                                # logged per epoch: the mean over its optimizer steps
            torch.mean(self.mb_stats, dim=1, out=self.epoch_stats)

    def _iteration_eager(self, env_init=None, noise=None):
        try:
            self.collect(env_init, noise, chain=self.chain)
            self.advantages()
            self.update()
        finally:
            self.engine.set_launch_chaining(False)

    def _counters(self):
        return (self.episode, self.step, self.num_episodes, self.training_step)

    def _capture(self, reset: bool):
        """Capture the launch sequence of one iteration (nothing executes; host counters are restored)."""
        eng = self.engine
        side = torch.cuda.Stream(device=eng.device)
        side.wait_stream(torch.cuda.current_stream(eng.device))
        graph = torch.cuda.CUDAGraph()
        saved = self._counters()
        with torch.cuda.graph(graph, stream=side):
            self._launch_iteration(reset)
        self.episode, self.step, self.num_episodes, self.training_step = saved
        self._graphs[reset] = graph

    def _launch_iteration(self, reset: bool):
        """The kernels of one iteration; ``reset`` False = start states were already written to ``self.env``."""
        eng, buf = self.engine, self.buf
        try:
            if reset:
                eng.env_reset(self.env, self.rng_key, self.episode)
                eng.set_launch_chaining(self.chain)
            eng.rollout(self.net.actor, self.env, buf["state"], buf["actions"], buf["logp"], buf["reward"],
                        ep_return=buf["ep_return"], seed=self.rng_key, episode=self.episode)
            eng.set_launch_chaining(self.chain)
            if self._episode_dev is not None:
                eng.episode_advance()
            self.episode += 1
            self.step += self.B * self.T * self.world
            self.num_episodes += self.B * self.world
            self.advantages()
            self.update()
        finally:
            eng.set_launch_chaining(False)

    def iteration(self, env_init=None, noise=None):
        """One pass of the reference's outer loop.  Device-drawn noise (the default) replays a captured CUDA graph;
        explicit ``noise`` (parity tests) and multi-GPU runs launch eagerly."""
        if not self.use_graph or noise is not None:
            return self._iteration_eager(env_init, noise)
        reset = env_init is None
        if not reset and env_init is not self.env:          # callers may write the start states into `self.env` directly
            self.env.copy_(env_init, non_blocking=True)
        if self._episode_dev is None:
            self._episode_dev = torch.full((1,), self.episode, dtype=torch.int64, device=self.engine.device)
            self.engine.set_episode_counter(self._episode_dev)
        graph = self._graphs.get(reset)
        if graph is None:
            # first use: this iteration runs eagerly (it is the warm-up), the following ones replay the capture
            l0 = self.engine.launches
            self._launch_iteration(reset)
            self.launches_per_iteration = self.engine.launches - l0
            self._capture(reset)
            return
        graph.replay()
        self.episode += 1
        self.step += self.B * self.T * self.world
        self.num_episodes += self.B * self.world
        self.training_step += self.args.epochs * self.n_mb

    def results_to_host(self, pinned: torch.Tensor):
        """One asynchronous D2H copy of this step's results into a pinned uint8 buffer of ``self.results.numel()`` bytes;
        ``split_results`` gives the (episode returns f64 [B], epoch statistics f32 [epochs][8]) views of it."""
        pinned.copy_(self.results, non_blocking=True)

    def split_results(self, host: torch.Tensor):
        n = self.B * 8 + self.args.epochs * 32
        return (host[:self.B * 8].view(torch.float64),
                host[self.B * 8:n].view(torch.float32).view(self.args.epochs, 8))

    # -- the CLI's logging path: ONE D2H copy per iteration, read one iteration late -------------------
    def stage_scalars(self):
        """Enqueue -- without synchronising the host -- the copy of everything the script logs for the iteration just
        launched (per-epoch statistics of MME:597-612, episode returns of MME:454-468) into one of two pinned host
        blocks.  Multi-GPU: the sum of the episode returns is all-reduced on the device first (collective: every rank
        calls this every iteration).  Returns a handle for `read_scalars`."""
        if self.world > 1:
            torch.sum(self.buf["ep_return"], dim=0, keepdim=True, out=self.return_sum)
            self._allreduce(self.return_sum)
        cuda = self.results.is_cuda
        if self._host_blocks is None:
            mk = (lambda: torch.empty(self.results.numel(), dtype=torch.uint8).pin_memory()) if cuda else \
                 (lambda: torch.empty(self.results.numel(), dtype=torch.uint8))
            self._host_blocks = [mk(), mk()]
        host = self._host_blocks[self._host_turn]
        self._host_turn ^= 1
        host.copy_(self.results, non_blocking=True)
        event = None
        if cuda:
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream(self.results.device))
        return host, event

    def read_scalars(self, handle) -> dict:
        """Wait for a `stage_scalars` copy and decode it: the seven train scalars (means over the epochs) and the mean
        episode return over the envs of all ranks."""
        host, event = handle
        if event is not None:
            event.synchronize()
        ret, stats = self.split_results(host)
        s = stats.mean(dim=0)
        keys = ("actor_loss", "critic_loss", "entropy", "kl_divergence", "clipped_ratios", "actor_gradients",
                "critic_gradients")
        out = {k: float(s[i]) for i, k in enumerate(keys)}
        if self.world > 1:
            total = float(host[self.B * 8 + self.args.epochs * 32:].view(torch.float64)[0])
        else:
            total = float(ret.sum())
        out["ep_reward"] = total / (self.B * self.world)
        out["ep_length"] = float(self.T)
        return out

    # -- read-backs (each is one small D2H copy; nothing else synchronises) ---------------------
    def train_scalars(self) -> dict:
        """Means over the epochs of the scalars logged at MME:605-612."""
        s = self.epoch_stats.mean(dim=0).cpu()
        keys = ("actor_loss", "critic_loss", "entropy", "kl_divergence", "clipped_ratios", "actor_gradients",
                "critic_gradients")
        return {k: float(s[i]) for i, k in enumerate(keys)}

    def rollout_scalars(self) -> dict:
        r = self.buf["ep_return"]
        tot = torch.stack([r.sum(), (r * r).sum()])
        self._allreduce(tot)
        n = self.B * self.world
        mean = float(tot[0]) / n
        return {"ep_reward": mean, "ep_length": float(self.T)}

    def get_batch(self):
        """The reference's ``RolloutBuffer.get_batch()`` 8-tuple (MME:148-157) for the last rollout."""
        return to_reference_layout(self.buf, self.engine.shapes.n_agents, 5, bool(self.args.agent_ids))


def tbptt_chunks(T: int, tbptt: int):
    """[(t0, t1)) step ranges after whose last step the reference back-propagates and steps the actor
    (``(t + 1) % tbptt == 0 or t == T - 1``, mappo_lstm_multienvs.py:603)."""
    out, t0 = [], 0
    for t in range(T):
        if ((t + 1) % tbptt == 0) or (t == T - 1):
            out.append((t0, t + 1))
            t0 = t + 1
    return out


def evaluate(trainer: MAPPO, num_episodes: int, seed: int, env_init=None, noise=None):
    """MME:614-644: ``num_eval_ep`` episodes with the *sampling* policy (``actor.act``), run as ``num_eval_ep`` parallel
    device envs on a context kept for the trainer's lifetime.  Returns (mean, population std, mean length) of the
    episode reward -- ``np.mean`` / ``np.std`` / ``np.mean`` of MME:642-644.  ``env_init`` f64 [18][n] / ``noise``
    f32 [T][N][A][n] make the evaluation a function of its inputs (parity test); default: device Philox draws."""
    ctx = trainer._eval_ctx.get(num_episodes)
    if ctx is None:
        import dataclasses
        eng = Engine(dataclasses.replace(trainer.engine.shapes, n_envs=num_episodes), trainer.engine.device.index)
        ctx = trainer._eval_ctx[num_episodes] = (eng, eng.alloc_rollout(),
                                                 eng.empty(eng.shapes.env_rows, num_episodes, dtype=torch.float64))
    eng, buf, env = ctx
    if env_init is None:
        eng.env_reset(env, seed, 0)
    else:
        env.copy_(env_init, non_blocking=True)
    eng.rollout(trainer.net.actor, env, buf["state"], buf["actions"], buf["logp"], buf["reward"], noise=noise,
                ep_return=buf["ep_return"], seed=seed, episode=0)
    r = buf["ep_return"].cpu()
    return float(r.mean()), float(r.std(unbiased=False)), float(trainer.engine.shapes.n_steps)


def init_distributed():
    """One process per GPU (torchrun): returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(local)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local
