"""torch-CPU restatement of the recurrent-actor delta of ``cleanmarl/mappo_lstm_multienvs.py`` (LSTM below).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Only what differs from ``oracle.mappo`` lives here:

  GRUActor            <- ``Actor`` LSTM:162-184: fc1 = Linear(O,H)+ReLU, GRUCell(H,H), fc2 = ReLU+Linear(H,A)
  rollout_act         <- the ``actor.act(obs, h=alive_h, ...)`` calls of the rollout loop LSTM:406-439
  ppo_update_tbptt    <- the training loop LSTM:551-664: truncated BPTT, one actor Adam step per chunk of
                         ``tbptt`` steps (loss / (n_valid_chunk * T_chunk)), critic stepped once per epoch

The same torch CPU operators as the reference are used (``nn.GRUCell``, ``Categorical``, autograd, Adam), and
the restatement is pinned against the unmodified reference file by ``tests/golden/g8_mappo_lstm.npz`` and
``g1_params.npz`` (``tests/test_oracle_golden.py``).  Layouts are the reference's batch-major ones.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions.categorical import Categorical

from . import mappo as om


class GRUActor(nn.Module):
    """LSTM:162-184.  Module construction order (fc1 Linear, GRUCell, fc2 Linear) equals the reference's, so
    the same ``torch.manual_seed`` gives the same parameters; ``parameters()`` order:
    fc1.W[H,O] fc1.b[H] gru.weight_ih[3H,H] gru.weight_hh[3H,H] gru.bias_ih[3H] gru.bias_hh[3H] fc2.W[A,H] fc2.b[A]
    (gate order r, z, n)."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.fc1 = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU())
        self.gru = nn.GRUCell(hidden_dim, hidden_dim)
        self.fc2 = nn.Sequential(nn.ReLU(), nn.Linear(hidden_dim, output_dim))

    def logits(self, x, h=None, avail_action=None):
        """LSTM:176-184; x [M,O], h [M,H] or None (zeros)."""
        x = self.fc1(x)
        if h is None:
            h = torch.zeros(x.size(0), self.hidden_dim)
        h = self.gru(x, h)
        x = self.fc2(h)
        if avail_action is not None:
            x = x.masked_fill(~avail_action, -1e9)
        return x, h

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


def build_networks(seed, obs_dim=21, state_dim=54, n_actions=5, actor_hidden=32, critic_hidden=64, critic_layers=1):
    """Seed, then Actor, then Critic -- LSTM:291-294, 327-338 (ippo_lstm_multienvs.py: state_dim = obs_dim, hidden 32)."""
    torch.manual_seed(seed)
    actor = GRUActor(obs_dim, actor_hidden, n_actions)
    critic = om.MLP(state_dim, critic_hidden, critic_layers, 1)
    return actor, critic


@torch.no_grad()
def rollout_act(actor: GRUActor, obs, h, avail, q):
    """One ``actor.act`` of the rollout (LSTM:409-426, 172-174) with the exponential race noise explicit.
    obs [B,N,O], h [B*N,H] or None, avail [B,N,A] bool, q [B,N,A] -> actions [B,N] int64, logp [B,N], h' [B*N,H]."""
    B, N, _ = obs.shape
    z, h = actor.logits(obs.reshape(B * N, -1), h, avail.reshape(B * N, -1))
    a, lp = om.race_sample(z, q.reshape(B * N, -1))
    return a.reshape(B, N), lp.reshape(B, N), h, z.reshape(B, N, -1)


def tbptt_chunks(T: int, tbptt: int):
    """[(t0, t1)) ranges at whose last step the reference back-propagates (LSTM:603)."""
    out, t0 = [], 0
    for t in range(T):
        if ((t + 1) % tbptt == 0) or (t == T - 1):
            out.append((t0, t + 1))
            t0 = t + 1
    return out


def ppo_update_tbptt(actor: GRUActor, critic, actor_opt, critic_opt, batch, adv, ret, *, epochs, clip, ent_coef,
                     tbptt=10, clip_gradients=-1.0, record_grads=False, critic_on_obs=False):
    """LSTM:551-664, line for line.  Returns per-epoch statistics (and, optionally, the per-chunk actor
    gradients / per-epoch critic gradients that autograd produced)."""
    obs, actions, old_logp, reward, states, avail, done, mask = batch
    B, T, N, _ = obs.shape
    stats = {k: [] for k in ("actor_loss", "critic_loss", "entropy", "kl", "clipfrac", "actor_grad_norm",
                             "critic_grad_norm")}
    grads = []
    for _ in range(epochs):
        total_actor_loss = 0
        actor_gradient = []
        critic_loss = 0
        entropies = 0
        kl_divergence = 0
        clipped_ratio = 0
        h = None
        truncated = None
        denom = None
        Tc = None
        chunk_grads = []
        for t in range(T):
            m = mask[:, t]
            z, h = actor.logits(obs[:, t].reshape(B * N, -1), h, avail[:, t].reshape(B * N, -1))
            dist = Categorical(logits=z.reshape(B, N, -1))
            logp = dist.log_prob(actions[:, t])
            log_ratio = logp - old_logp[:, t]
            ratio = torch.exp(log_ratio)
            pg1 = adv[:, t] * ratio
            pg2 = adv[:, t] * torch.clamp(ratio, 1 - clip, 1 + clip)
            pg = torch.min(pg1[m], pg2[m]).mean(dim=-1).sum()
            ent = dist.entropy()[m].mean(dim=-1).sum()
            entropies = entropies + ent
            actor_loss = -pg - ent_coef * ent
            total_actor_loss = total_actor_loss + actor_loss
            if truncated is None:
                truncated, denom, Tc = actor_loss, m.sum(), 1
            else:
                truncated = truncated + actor_loss
                denom = denom + m.sum()
                Tc += 1
            if ((t + 1) % tbptt == 0) or (t == T - 1):
                truncated = truncated / (denom * Tc)
                actor_opt.zero_grad()
                truncated.backward()
                actor_gradient.append(om.norm_d([p.grad for p in actor.parameters()], 2))
                if record_grads:
                    chunk_grads.append(actor.flat_grads().clone())
                if clip_gradients > 0:
                    torch.nn.utils.clip_grad_norm_(actor.parameters(), max_norm=clip_gradients)
                actor_opt.step()
                truncated = None
                h = h.detach()
            if critic_on_obs:                      # ippo_lstm_multienvs.py:623 (Critic.forward ends with .squeeze(), :201)
                values = critic(obs[:, t]).squeeze()
            else:
                values = critic(states[:, t]).expand(-1, N)
            critic_loss = critic_loss + F.mse_loss(values[m], ret[:, t][m]) * m.sum()
            kl_divergence = kl_divergence + ((ratio - 1) - log_ratio)[m].mean(dim=-1).sum()
            clipped_ratio = clipped_ratio + ((ratio - 1.0).abs() > clip)[m].float().mean(dim=-1).sum()
        n = mask.sum()
        total_actor_loss = total_actor_loss / n
        critic_loss = critic_loss / n
        entropies = entropies / n
        kl_divergence = kl_divergence / n
        clipped_ratio = clipped_ratio / n
        critic_opt.zero_grad()
        critic_loss.backward()
        critic_gradient = om.norm_d([p.grad for p in critic.parameters()], 2)
        if record_grads:
            grads.append((chunk_grads, critic.flat_grads().clone()))
        if clip_gradients > 0:
            torch.nn.utils.clip_grad_norm_(critic.parameters(), max_norm=clip_gradients)
        critic_opt.step()
        stats["actor_loss"].append(total_actor_loss.item())
        stats["critic_loss"].append(critic_loss.item())
        stats["entropy"].append(entropies.item())
        stats["kl"].append(kl_divergence.item())
        stats["clipfrac"].append(float(clipped_ratio))
        stats["actor_grad_norm"].append(float(np.mean([float(g) for g in actor_gradient])))
        stats["critic_grad_norm"].append(float(critic_gradient))
        stats.setdefault("actor_chunk_grad_norms", []).append([float(g) for g in actor_gradient])
    if record_grads:
        stats["grads"] = grads
    return stats


def synthetic_old_logp(actor: GRUActor, batch, seed=1, sigma=0.05):
    """Old log-probs consistent with the recurrent policy: current log-prob (h unrolled from 0) + N(0, sigma^2)."""
    obs, actions, _, _, _, avail, _, _ = batch
    B, T, N, _ = obs.shape
    out = torch.zeros(B, T, N)
    h = None
    with torch.no_grad():
        for t in range(T):
            z, h = actor.logits(obs[:, t].reshape(B * N, -1), h, avail[:, t].reshape(B * N, -1))
            out[:, t] = Categorical(logits=z.reshape(B, N, -1)).log_prob(actions[:, t])
    g = torch.Generator().manual_seed(seed)
    return out + sigma * torch.randn(B, T, N, generator=g)
