"""torch-CPU restatement of the arithmetic of ``cleanmarl/mappo_multienvs.py`` (MME).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows.  The restatement deliberately uses the same torch
CPU operators as the reference (``nn.Linear``, ``Categorical``, autograd,
``torch.optim.Adam``) so that it *is* the reference arithmetic; it is pinned
against the unmodified reference file by ``tests/golden/*.npz`` (see
``tests/golden/gen_golden.py`` and ``tests/test_oracle_golden.py``).

Layouts here are the reference's batch-major ones: obs ``[B,T,N,O]``, states
``[B,T,S]``, actions ``[B,T,N]`` int64, log_probs ``[B,T,N]``, reward ``[B,T]``,
avail ``[B,T,N,A]`` bool, done ``[B,T]``, mask ``[B,T]`` bool (MME:148-157).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions.categorical import Categorical


# --------------------------------------------------------------------------- networks
class MLP(nn.Module):
    """Linear+ReLU stack; ``Actor`` (MME:160-183) and ``Critic`` (MME:186-200) share it.

    Construction order of the ``nn.Linear`` modules equals the reference's, so the
    same ``torch.manual_seed`` yields the same parameters.
    """

    def __init__(self, input_dim: int, hidden_dim: int, num_layer: int, output_dim: int):
        super().__init__()
        dims = [input_dim] + [hidden_dim] * (num_layer + 1) + [output_dim]
        self.linears = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1))

    def forward(self, x):
        for lin in self.linears[:-1]:
            x = F.relu(lin(x))
        return self.linears[-1](x)

    def flat_params(self) -> torch.Tensor:
        return torch.cat([p.detach().reshape(-1) for p in self.parameters()])

    def flat_grads(self) -> torch.Tensor:
        return torch.cat([p.grad.detach().reshape(-1) for p in self.parameters()])

    def load_flat(self, flat: torch.Tensor):
        off = 0
        with torch.no_grad():
            for p in self.parameters():
                n = p.numel()
                p.copy_(flat[off:off + n].reshape(p.shape))
                off += n


def build_networks(seed, obs_dim=21, state_dim=54, n_actions=5, actor_hidden=32, actor_layers=1,
                   critic_hidden=64, critic_layers=1):
    """Seed, then Actor, then Critic -- MME:291-294, 329-339 (order fixes the init)."""
    torch.manual_seed(seed)
    actor = MLP(obs_dim, actor_hidden, actor_layers, n_actions)
    critic = MLP(state_dim, critic_hidden, critic_layers, 1)
    return actor, critic


def actor_logits(actor: MLP, x, avail=None):
    """MME:178-183."""
    z = actor(x)
    if avail is not None:
        z = z.masked_fill(~avail, -1e9)
    return z


def race_sample(logits: torch.Tensor, q: torch.Tensor):
    """``Actor.act`` (MME:172-176) with the exponential noise ``q`` made explicit.

    ``Categorical(logits).sample()`` is ``torch.multinomial(probs, 1, True)`` which on
    CPU is ``argmax(probs / q)``, ``q = empty_like(probs).exponential_(1)``.
    Returns (actions int64, log_prob f32) exactly as the reference would for that ``q``.
    """
    dist = Categorical(logits=logits)
    actions = torch.argmax(dist.probs / q, dim=-1)
    return actions, dist.log_prob(actions)


def draw_race_noise(shape, generator=None) -> torch.Tensor:
    return torch.empty(shape, dtype=torch.float32).exponential_(1, generator=generator)


# --------------------------------------------------------------------------- buffer
def collate(episodes, num_agents, obs_space, state_space, action_space, normalize_reward=False):
    """``RolloutBuffer.add`` + ``get_batch`` (MME:103-157) for a list of episode dicts."""
    eps = [{k: torch.from_numpy(np.stack(v)).float() for k, v in ep.items()} for ep in episodes]
    B = len(eps)
    lengths = [len(ep["obs"]) for ep in eps]
    T = max(lengths)
    obs = torch.zeros(B, T, num_agents, obs_space)
    avail = torch.zeros(B, T, num_agents, action_space)
    actions = torch.zeros(B, T, num_agents)
    log_probs = torch.zeros(B, T, num_agents)
    reward = torch.zeros(B, T)
    states = torch.zeros(B, T, state_space)
    done = torch.zeros(B, T)
    mask = torch.zeros(B, T, dtype=torch.bool)
    for i, ep in enumerate(eps):
        L = lengths[i]
        obs[i, :L] = ep["obs"]
        avail[i, :L] = ep["avail_actions"]
        actions[i, :L] = ep["actions"]
        log_probs[i, :L] = ep["log_prob"]
        reward[i, :L] = ep["reward"]
        states[i, :L] = ep["states"]
        done[i, :L] = ep["done"]
        mask[i, :L] = True
    if normalize_reward:
        reward = normalize_reward_(reward, mask)
    return obs, actions.long(), log_probs, reward, states, avail.bool(), done, mask


def normalize_reward_(reward, mask):
    """MME:143-146 -- unbiased std, eps 1e-6, masked entries only."""
    reward = reward.clone()
    mu = torch.mean(reward[mask])
    std = torch.std(reward[mask])
    reward[mask] = (reward[mask] - mu) / (std + 1e-6)
    return reward


# --------------------------------------------------------------------------- TD(lambda)
@torch.no_grad()
def td_lambda_loop(critic: MLP, critic_in, reward, mask, gamma, lam, n_agents):
    """The reference's double Python loop, MME:484-504 (IPPO: ``ippo_multienvs.py:484-504``).

    ``critic_in`` is ``b_states [B,T,S]`` (MAPPO, critic output ``[1]`` broadcast to the
    N agents) or ``b_obs [B,T,N,O]`` (IPPO, critic output squeezed to ``[N]``).  Batch-of-1
    critic calls, fp32 scalar recurrence with python-float coefficients -- kept as is.
    """
    ippo = critic_in.dim() == 4
    B, T = reward.shape
    ret = torch.zeros(B, T, n_agents)
    adv = torch.zeros(B, T, n_agents)
    val = (lambda x: critic(x).squeeze()) if ippo else critic
    for b in range(B):
        ep_len = int(mask[b].sum())
        last = 0
        for t in reversed(range(ep_len)):
            nv = 0 if t == ep_len - 1 else val(critic_in[b, t + 1])
            ret[b, t] = last = reward[b, t] + gamma * (lam * last + (1 - lam) * nv)
            adv[b, t] = ret[b, t] - val(critic_in[b, t])
    return ret, adv


@torch.no_grad()
def td_lambda_batched(critic: MLP, critic_in, reward, mask, gamma, lam, n_agents):
    """Same recurrence, critic evaluated once on the whole batch and the scan vectorised
    over envs (what ``mappo_jax_multienvs.py:336-417`` does).  Differs from
    ``td_lambda_loop`` only by the critic's batch-of-1 vs batched GEMM rounding
    (<= ~2e-6); used where the loop form would take minutes."""
    ippo = critic_in.dim() == 4
    B, T = reward.shape
    v = critic(critic_in).squeeze(-1)                       # [B,T] or [B,T,N]
    if not ippo:
        v = v.unsqueeze(-1).expand(B, T, n_agents)
    return td_lambda_scan(v, reward, mask, gamma, lam)


@torch.no_grad()
def td_lambda_scan(values, reward, mask, gamma, lam):
    """The scalar recurrence of MME:496-504 on precomputed values ``[B,T,N]``.

    fp32 ops in the reference's order: ``r + g*(l*R + (1-l)*nv)`` with the three
    python floats rounded to fp32 first (a python scalar times an fp32 tensor is an
    fp32 op); bit-exact target for the CUDA scan kernel."""
    B, T, N = values.shape
    g = torch.tensor(gamma, dtype=torch.float32)
    l = torch.tensor(lam, dtype=torch.float32)
    oml = torch.tensor(1 - lam, dtype=torch.float32)
    ep_len = mask.sum(dim=1)                                # [B]
    ret = torch.zeros(B, T, N)
    adv = torch.zeros(B, T, N)
    last = torch.zeros(B, N)
    for t in reversed(range(T)):
        live = (t < ep_len)
        is_last = (t == ep_len - 1)
        if t + 1 < T:
            nv = torch.where(is_last[:, None], torch.zeros(B, N), values[:, t + 1])
        else:
            nv = torch.zeros(B, N)
        r_t = reward[:, t][:, None] + g * (l * last + oml * nv)
        a_t = r_t - values[:, t]
        ret[:, t] = torch.where(live[:, None], r_t, ret[:, t])
        adv[:, t] = torch.where(live[:, None], a_t, adv[:, t])
        last = torch.where(live[:, None], r_t, last)
    return ret, adv


def normalize_masked(x, mask):
    """MME:505-512: ``(x - mu)/sd`` with mu/sd of the agent-mean over masked (b,t); unbiased, no eps."""
    m = x.mean(dim=-1)[mask]
    return (x - m.mean()) / m.std()


# --------------------------------------------------------------------------- PPO epoch
@dataclass
class EpochOut:
    actor_loss: torch.Tensor
    critic_loss: torch.Tensor
    entropy: torch.Tensor
    kl: torch.Tensor
    clipfrac: torch.Tensor


def clipped_value_loss(values, ret, values_old, value_clip):
    """Beyond the reference (default off, BASELINE.json north_star "value-clip"): PPO2's clipped value loss, elementwise
    ``max((V - R)^2, (V_old + clamp(V - V_old, -c, c) - R)^2)``."""
    vc = values_old + torch.clamp(values - values_old, -value_clip, value_clip)
    return torch.max((values - ret) ** 2, (vc - ret) ** 2)


def ppo_epoch_loop(actor, critic, obs, actions, old_logp, critic_in, avail, mask, adv, ret,
                   clip, ent_coef, value_clip=-1.0, values_old=None) -> EpochOut:
    """One epoch's loss accumulation, the reference's loop over t -- MME:522-576
    (IPPO: critic on ``b_obs`` without expand, ``ippo_multienvs.py:554``)."""
    ippo = critic_in.dim() == 4
    n_agents = obs.shape[2]
    actor_loss = critic_loss = entropies = kl = clipped = 0
    for t in range(obs.size(1)):
        m = mask[:, t]
        dist = Categorical(logits=actor_logits(actor, obs[:, t], avail[:, t]))
        logp = dist.log_prob(actions[:, t])
        log_ratio = logp - old_logp[:, t]
        ratio = torch.exp(log_ratio)
        pg1 = adv[:, t] * ratio
        pg2 = adv[:, t] * torch.clamp(ratio, 1 - clip, 1 + clip)
        pg = torch.min(pg1[m], pg2[m]).mean(dim=-1).sum()
        ent = dist.entropy()[m].mean(dim=-1).sum()
        entropies = entropies + ent
        actor_loss = actor_loss + (-pg - ent_coef * ent)
        if ippo:
            values = critic(critic_in[:, t]).squeeze(-1)
        else:
            values = critic(critic_in[:, t]).expand(-1, n_agents)
        if value_clip > 0:
            critic_loss = critic_loss + clipped_value_loss(values[m], ret[:, t][m], values_old[:, t][m], value_clip).mean() * m.sum()
        else:
            critic_loss = critic_loss + F.mse_loss(values[m], ret[:, t][m]) * m.sum()
        kl = kl + ((ratio - 1) - log_ratio)[m].mean(dim=-1).sum()
        clipped = clipped + ((ratio - 1.0).abs() > clip)[m].float().mean(dim=-1).sum()
    n = mask.sum()
    return EpochOut(actor_loss / n, critic_loss / n, entropies / n, kl / n, clipped / n)


def ppo_epoch_flat(actor, critic, obs, actions, old_logp, critic_in, avail, mask, adv, ret,
                   clip, ent_coef, value_clip=-1.0, values_old=None) -> EpochOut:
    """Algebraically the same losses as ``ppo_epoch_loop`` without the loop over t
    (one batched pass; sums reassociated).  For sizes where the loop is too slow."""
    ippo = critic_in.dim() == 4
    n_agents = obs.shape[2]
    n = mask.sum()
    w = mask.float().unsqueeze(-1) / (n_agents * n)
    dist = Categorical(logits=actor_logits(actor, obs, avail))
    logp = dist.log_prob(actions)
    log_ratio = logp - old_logp
    ratio = torch.exp(log_ratio)
    pg = torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - clip, 1 + clip))
    ent = (dist.entropy() * w).sum()
    actor_loss = -(pg * w).sum() - ent_coef * ent
    v = critic(critic_in).squeeze(-1)
    if not ippo:
        v = v.unsqueeze(-1).expand_as(ret)
    if value_clip > 0:
        critic_loss = (clipped_value_loss(v, ret, values_old, value_clip) * w).sum()
    else:
        critic_loss = (((v - ret) ** 2) * w).sum()
    kl = (((ratio - 1) - log_ratio) * w).sum()
    clipped = (((ratio - 1.0).abs() > clip).float() * w).sum()
    return EpochOut(actor_loss, critic_loss, ent, kl, clipped)


def norm_d(grads, d=2):
    """MME:221-224 -- norm of the per-tensor norms."""
    norms = [torch.linalg.vector_norm(g.detach(), d) for g in grads]
    return torch.linalg.vector_norm(torch.tensor(norms), d)


def ppo_update(actor, critic, actor_opt, critic_opt, batch, adv, ret, *, epochs, clip, ent_coef,
               clip_gradients=-1.0, critic_on_obs=False, flat=False, record_grads=False,
               value_clip=-1.0, values_old=None, num_minibatches=1):
    """The training loop MME:521-603: ``epochs`` x (loss, backward x2, grad norms, optional
    clip, Adam step x2).  ``batch`` is the ``get_batch`` 8-tuple.  Returns per-epoch stats
    (one entry per optimizer step).  ``value_clip`` / ``num_minibatches`` are the default-off
    extensions BASELINE.json names (not in the reference): minibatch m = the contiguous env
    block ``[m * B / M, (m + 1) * B / M)``, one optimizer step per block."""
    obs, actions, old_logp, reward, states, avail, done, mask = batch
    critic_in = obs if critic_on_obs else states
    epoch_fn = ppo_epoch_flat if flat else ppo_epoch_loop
    stats = {k: [] for k in ("actor_loss", "critic_loss", "entropy", "kl", "clipfrac",
                             "actor_grad_norm", "critic_grad_norm")}
    grads = []
    B = obs.shape[0]
    blocks = [slice(m * B // num_minibatches, (m + 1) * B // num_minibatches) for m in range(num_minibatches)]
    for sl in [s for _ in range(epochs) for s in blocks]:
        kw = {} if value_clip <= 0 else dict(value_clip=value_clip, values_old=values_old[sl])
        out = epoch_fn(actor, critic, obs[sl], actions[sl], old_logp[sl], critic_in[sl], avail[sl], mask[sl], adv[sl],
                       ret[sl], clip, ent_coef, **kw)
        actor_opt.zero_grad()
        critic_opt.zero_grad()
        out.actor_loss.backward()
        out.critic_loss.backward()
        stats["actor_grad_norm"].append(float(norm_d([p.grad for p in actor.parameters()])))
        stats["critic_grad_norm"].append(float(norm_d([p.grad for p in critic.parameters()])))
        if record_grads:
            grads.append((actor.flat_grads().clone(), critic.flat_grads().clone()))
        if clip_gradients > 0:
            torch.nn.utils.clip_grad_norm_(actor.parameters(), max_norm=clip_gradients)
            torch.nn.utils.clip_grad_norm_(critic.parameters(), max_norm=clip_gradients)
        actor_opt.step()
        critic_opt.step()
        stats["actor_loss"].append(out.actor_loss.item())
        stats["critic_loss"].append(out.critic_loss.item())
        stats["entropy"].append(out.entropy.item())
        stats["kl"].append(out.kl.item())
        stats["clipfrac"].append(float(out.clipfrac))
    if record_grads:
        stats["grads"] = grads
    return stats


def make_optimizers(actor, critic, lr_actor=8e-4, lr_critic=8e-4, name="Adam"):
    """MME:341-343."""
    Opt = getattr(torch.optim, name)
    return Opt(actor.parameters(), lr=lr_actor), Opt(critic.parameters(), lr=lr_critic)


# --------------------------------------------------------------------------- synthetic batch
def synthetic_batch(B, T=25, N=3, A=5, seed=1, actor=None):
    """Seeded synthetic rollout batch of the simple_spread shape (raw obs 18, ids, state 54).

    obs columns 0..17 ~U(-2,2) (velocities ~U(-1.3,1.3)), one-hot ids; state = the three raw
    rows concatenated (``pettingzoo_wrapper.py:95``); reward ~ -U(0,4); actions ~ randint;
    old log-probs = current policy log-prob + N(0, 0.05^2) when ``actor`` is given."""
    g = torch.Generator().manual_seed(seed)
    raw = (torch.rand(B, T, N, 18, generator=g) * 4 - 2)
    raw[..., 0:2] = torch.rand(B, T, N, 2, generator=g) * 2.6 - 1.3
    raw[..., 14:18] = 0.0
    ids = torch.eye(N).expand(B, T, N, N)
    obs = torch.cat([raw, ids], dim=-1).contiguous()
    states = raw.reshape(B, T, N * 18).contiguous()
    reward = -torch.rand(B, T, generator=g) * 4
    actions = torch.randint(0, A, (B, T, N), generator=g)
    avail = torch.ones(B, T, N, A, dtype=torch.bool)
    mask = torch.ones(B, T, dtype=torch.bool)
    done = torch.zeros(B, T)
    if actor is not None:
        with torch.no_grad():
            logp = Categorical(logits=actor_logits(actor, obs, avail)).log_prob(actions)
        logp = logp + 0.05 * torch.randn(B, T, N, generator=g)
    else:
        logp = -torch.rand(B, T, N, generator=g) * 2 - 0.5
    return obs, actions, logp, reward, states, avail, done, mask
