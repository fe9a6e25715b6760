"""CPU test double of ``cleanmarl_b200.engine.Engine`` built on the oracle (TEST INFRASTRUCTURE).

It exists so that the host-side logic of ``cleanmarl_b200.mappo.MAPPO`` -- env sharding across ranks, the
"unnormalised sums + one all-reduce per epoch + identical Adam on every rank" protocol, step counters --
can run under ``gloo`` with world_size 2 on a machine without a GPU.  It lives in tests/ and is injected
through ``MAPPO(engine_factory=...)``; nothing in the product imports it.  Same tensors, same device
layout ([T][.][B], env-minor) and same call signatures as the real engine.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import mappo as om
from oracle import mappo_lstm as ol
from oracle import spread as osp


class OracleEngine:
    tensor_cores = False

    def __init__(self, shapes, device=None):
        self.shapes = shapes
        self.device = torch.device("cpu")
        s = shapes
        cin = s.obs_dim if s.critic_on_obs else s.state_dim
        self.recurrent = bool(getattr(s, "actor_recurrent", False))
        self.n_actor = s.obs_dim * s.actor_hidden + s.actor_hidden + s.actor_hidden ** 2 + s.actor_hidden + \
            s.n_actions * s.actor_hidden + s.n_actions
        if self.recurrent:
            Hh = s.actor_hidden
            self.n_actor = s.obs_dim * Hh + Hh + 2 * 3 * Hh * Hh + 2 * 3 * Hh + s.n_actions * Hh + s.n_actions
        self.n_critic = cin * s.critic_hidden + s.critic_hidden + s.critic_hidden ** 2 + s.critic_hidden + s.critic_hidden + 1
        self.n_params = self.n_actor + self.n_critic
        self.n_heads = s.n_agents if s.critic_on_obs else 1
        self.launches = 0
        self.weight_decay = (0.0, 0.0)

    def set_launch_chaining(self, on):
        pass                                        # a launch-scheduling hint of the CUDA library: no arithmetic

    def set_weight_decay(self, actor_wd, critic_wd):
        self.weight_decay = (float(actor_wd), float(critic_wd))

    # -- helpers -------------------------------------------------------------------------------
    def empty(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype)

    def alloc_rollout(self, with_obs=False):
        s = self.shapes
        T, B, N = s.n_steps, s.n_envs, s.n_agents
        return {"state": self.empty(T, s.state_dim, B), "actions": self.empty(T, N, B, dtype=torch.int32),
                "logp": self.empty(T, N, B), "reward": self.empty(T, B), "ep_return": self.empty(B, dtype=torch.float64),
                "values": self.empty(T, self.n_heads, B), "returns": self.empty(T, self.n_heads, B),
                "adv": self.empty(T, self.n_heads, B), "obs": None}

    def _nets(self, flat_actor=None, flat_critic=None):
        s = self.shapes
        cin = s.obs_dim if s.critic_on_obs else s.state_dim
        actor = (ol.GRUActor(s.obs_dim, s.actor_hidden, s.n_actions) if self.recurrent
                 else om.MLP(s.obs_dim, s.actor_hidden, 1, s.n_actions))
        critic = om.MLP(cin, s.critic_hidden, 1, 1)
        if flat_actor is not None:
            actor.load_flat(flat_actor)
        if flat_critic is not None:
            critic.load_flat(flat_critic)
        return actor, critic

    def _obs(self, state):                          # [T][S][B] -> [B,T,N,O]
        from cleanmarl_b200.engine import obs_from_state
        return obs_from_state(state, self.shapes.n_agents, self.shapes.obs_dim > 18).permute(3, 0, 1, 2).contiguous()

    # -- entries -------------------------------------------------------------------------------
    def env_reset(self, env, seed, episode):
        rng = np.random.default_rng([seed & 0xFFFFFFFF, episode])
        B = env.shape[1]
        env[0:6] = torch.from_numpy(rng.uniform(-1, 1, (6, B)))
        env[6:12] = 0
        env[12:18] = torch.from_numpy(rng.uniform(-1, 1, (6, B)))

    @staticmethod
    def _unpack(env):
        e = env.numpy()
        B = e.shape[1]
        return e[0:6].T.reshape(B, 3, 2).copy(), e[6:12].T.reshape(B, 3, 2).copy(), e[12:18].T.reshape(B, 3, 2).copy()

    def env_observe(self, env, state_out):
        pos, vel, lm = self._unpack(env)
        state_out.copy_(torch.from_numpy(osp.observe_batched(pos, vel, lm).reshape(env.shape[1], 54).T.copy()))

    def env_step(self, env, actions, state_out=None, reward_out=None):
        pos, vel, lm = self._unpack(env)
        B = env.shape[1]
        pos, vel, rew = osp.step_batched(pos, vel, lm, actions.t().numpy())
        env[0:6] = torch.from_numpy(pos.reshape(B, 6).T.copy()); env[6:12] = torch.from_numpy(vel.reshape(B, 6).T.copy())
        if reward_out is not None:
            reward_out.copy_(torch.from_numpy(rew[:, 0].astype(np.float32)))
        if state_out is not None:
            state_out.copy_(torch.from_numpy(osp.observe_batched(pos, vel, lm).reshape(B, 54).T.copy()))

    def rollout(self, actor_params, env, state, actions, logp, reward, *, noise=None, obs=None, ep_return=None,
                seed=0, episode=0):
        assert noise is not None, "the CPU test double needs explicit race noise"
        s = self.shapes
        B, T = s.n_envs, s.n_steps
        actor, _ = self._nets(actor_params)
        e = env.numpy()
        pos = e[0:6].T.reshape(B, 3, 2).copy(); vel = e[6:12].T.reshape(B, 3, 2).copy(); lm = e[12:18].T.reshape(B, 3, 2).copy()
        ids = np.broadcast_to(np.eye(3), (B, 3, 3))
        ret = np.zeros(B)
        h = None
        for t in range(T):
            raw = osp.observe_batched(pos, vel, lm)                                   # [B,3,18] f32
            o = np.concatenate([raw, ids], -1) if s.obs_dim > 18 else raw
            with torch.no_grad():
                if self.recurrent:
                    a, lp, h, _ = ol.rollout_act(actor, torch.from_numpy(o).float(), h,
                                                 torch.ones(B, 3, s.n_actions, dtype=torch.bool), noise[t].permute(2, 0, 1))
                else:
                    z = om.actor_logits(actor, torch.from_numpy(o).float())
                    a, lp = om.race_sample(z, noise[t].permute(2, 0, 1))               # [B,N,A]
            state[t] = torch.from_numpy(raw.reshape(B, 54).T.copy())
            actions[t] = a.t().to(torch.int32)
            logp[t] = lp.t()
            pos, vel, rew = osp.step_batched(pos, vel, lm, a.numpy())
            reward[t] = torch.from_numpy(rew[:, 0].astype(np.float32))
            ret += rew[:, 0]
        env[0:6] = torch.from_numpy(pos.reshape(B, 6).T.copy()); env[6:12] = torch.from_numpy(vel.reshape(B, 6).T.copy())
        if ep_return is not None:
            ep_return.copy_(torch.from_numpy(ret))
        self.launches += 1

    def critic_values(self, critic_params, values, *, state=None, obs=None):
        _, critic = self._nets(None, critic_params)
        with torch.no_grad():
            if self.shapes.critic_on_obs:
                v = critic(self._obs(state)).squeeze(-1)                               # [B,T,N]
                values.copy_(v.permute(1, 2, 0))
            else:
                v = critic(state.permute(2, 0, 1)).squeeze(-1)                          # [B,T]
                values.copy_(v.t().unsqueeze(1))
        self.launches += 1

    def td_lambda(self, values, reward, returns, adv, gamma, lam, *, mask=None):
        B = reward.shape[1]
        m = torch.ones(B, reward.shape[0], dtype=torch.bool) if mask is None else mask.t().bool()
        r, a = om.td_lambda_scan(values.permute(2, 0, 1).contiguous(), reward.t().contiguous(), m, gamma, lam)
        returns.copy_(r.permute(1, 2, 0)); adv.copy_(a.permute(1, 2, 0))
        self.launches += 1

    def normalize(self, x, n_heads, mode, phase, stats, *, mask=None):
        T, V, B = (x.shape[0], 1, x.shape[1]) if x.dim() == 2 else x.shape
        xv = x.reshape(T, V, B)
        m = torch.ones(T, B, dtype=torch.bool) if mask is None else mask.bool()
        if phase == 0:
            hm = xv.mean(dim=1)[m].double()
            stats[0], stats[1], stats[2], stats[3] = hm.sum(), (hm * hm).sum(), float(hm.numel()), 0.0
            return
        n = stats[2]; mean = stats[0] / n
        var = torch.clamp((stats[1] - stats[0] * mean) / (n - 1.0), min=0.0)
        mu, sd = mean.float(), torch.sqrt(var).float()
        if mode == 0:
            sel = m.unsqueeze(1).expand(T, V, B)
            xv[sel] = (xv[sel] - mu) / (sd + 1e-6)
        else:
            xv.copy_((xv - mu) / sd)

    def ppo_epoch_grads(self, params, grads, *, state=None, obs=None, actions, logp_old, adv, returns, mask=None,
                        avail=None, clip=0.2, ent_coef=0.001, value_clip=-1.0, values_old=None, env_begin=0,
                        env_count=None):
        s = self.shapes
        actor, critic = self._nets(params[:self.n_actor], params[self.n_actor:])
        T, B, N = s.n_steps, s.n_envs, s.n_agents
        o = self._obs(state)
        m = torch.ones(B, T, dtype=torch.bool) if mask is None else mask.t().bool()
        av = torch.ones(B, T, N, s.n_actions, dtype=torch.bool) if avail is None else avail.permute(3, 0, 1, 2).bool()
        exp = lambda x: x.permute(2, 0, 1).expand(B, T, N) if x.shape[1] == 1 else x.permute(2, 0, 1)
        sl = slice(env_begin, B if env_count is None else env_begin + env_count)
        kw = {} if value_clip <= 0 else dict(value_clip=value_clip, values_old=exp(values_old)[sl])
        out = om.ppo_epoch_flat(actor, critic, o[sl], actions.permute(2, 0, 1).long()[sl], logp_old.permute(2, 0, 1)[sl],
                                (o if s.critic_on_obs else state.permute(2, 0, 1))[sl], av[sl], m[sl], exp(adv)[sl],
                                exp(returns)[sl], clip, ent_coef, **kw)
        m = m[sl]
        out.actor_loss.backward(); out.critic_loss.backward()
        n = float(m.sum())
        grads[:self.n_actor] = actor.flat_grads() * n
        grads[self.n_actor:self.n_params] = critic.flat_grads() * n
        st = torch.tensor([out.actor_loss.item(), out.critic_loss.item(), out.entropy.item(), out.kl.item(),
                           float(out.clipfrac), 1.0, 0.0, 0.0]) * n
        st[6:] = 0
        grads[self.n_params:] = st
        self.launches += 3

    def clip_adam_step(self, params, grads, exp_avg, exp_avg_sq, *, step=1, step_dev=None, lr_actor=8e-4,
                       lr_critic=8e-4, beta1=0.9, beta2=0.999, eps=1e-8, max_norm=-1.0, stats_out=None):
        P, na = self.n_params, self.n_actor
        count = grads[P + 5]
        g = grads[:P] / count
        k = int(step_dev.item()) + 1 if step_dev is not None else step
        actor, critic = self._nets()
        norms = []
        for net, lo in ((actor, 0), (critic, na)):
            off, sq = lo, []
            for p in net.parameters():
                sq.append(torch.linalg.vector_norm(g[off:off + p.numel()])); off += p.numel()
            norms.append(torch.linalg.vector_norm(torch.stack(sq)))
        for (lo, hi, lr, nrm) in ((0, na, lr_actor, norms[0]), (na, P, lr_critic, norms[1])):
            gi = g[lo:hi]
            if max_norm > 0:
                gi = gi * torch.clamp(max_norm / (nrm + 1e-6), max=1.0)
            m, v = exp_avg[lo:hi], exp_avg_sq[lo:hi]
            m.lerp_(gi, 1 - beta1)
            v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
            bc1, bc2 = 1 - beta1 ** k, 1 - beta2 ** k
            denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
            wd = self.weight_decay[0 if lo == 0 else 1]
            if wd:
                params[lo:hi].mul_(1 - lr * wd)
            params[lo:hi].addcdiv_(m, denom, value=-(lr / bc1))
        if stats_out is not None:
            stats_out[:5] = grads[P:P + 5] / count
            stats_out[5], stats_out[6], stats_out[7] = norms[0], norms[1], count
        if step_dev is not None:
            step_dev.fill_(k)
        self.launches += 1

    # -- recurrent-actor entries (mappo_lstm_multienvs.py) ---------------------------------------------------
    def alloc_h_seq(self):
        s = self.shapes
        return torch.zeros(s.n_steps + 1, s.n_agents, s.actor_hidden, s.n_envs)

    def tbptt_chunk_grads(self, actor_params, grads, h_seq, t0, t1, *, state=None, obs=None, actions, logp_old, adv,
                          mask=None, avail=None, clip=0.2, ent_coef=0.001, stash=None):
        """Unnormalised sums of one truncated-BPTT chunk (autograd through the oracle's GRU actor), hidden state
        carried through ``h_seq`` exactly like the device kernel."""
        from torch.distributions.categorical import Categorical
        s = self.shapes
        T, B, N = s.n_steps, s.n_envs, s.n_agents
        actor, _ = self._nets(actor_params)
        o = self._obs(state)
        m = torch.ones(B, T, dtype=torch.bool) if mask is None else mask.t().bool()
        A = adv.permute(2, 0, 1).expand(B, T, N) if adv.shape[1] == 1 else adv.permute(2, 0, 1)
        acts = actions.permute(2, 0, 1).long()
        old = logp_old.permute(2, 0, 1)
        h = None if t0 == 0 else h_seq[t0].permute(2, 0, 1).reshape(B * N, -1).clone()
        loss = ent_s = kl_s = clip_s = 0.0
        for t in range(t0, t1):
            z, h = actor.logits(o[:, t].reshape(B * N, -1), h, None)
            with torch.no_grad():
                h_seq[t + 1] = h.reshape(B, N, -1).permute(1, 2, 0)
            dist = Categorical(logits=z.reshape(B, N, -1))
            lr = dist.log_prob(acts[:, t]) - old[:, t]
            ratio = torch.exp(lr)
            mt = m[:, t]
            pg = torch.min(A[:, t] * ratio, A[:, t] * torch.clamp(ratio, 1 - clip, 1 + clip))[mt].mean(dim=-1).sum()
            ent = dist.entropy()[mt].mean(dim=-1).sum()
            loss = loss + (-pg - ent_coef * ent)
            ent_s = ent_s + ent.detach()
            kl_s = kl_s + ((ratio - 1) - lr)[mt].mean(dim=-1).sum().detach()
            clip_s = clip_s + ((ratio - 1.0).abs() > clip)[mt].float().mean(dim=-1).sum()
        actor.zero_grad()
        loss.backward()
        na = self.n_actor
        grads[:na] = actor.flat_grads()
        grads[na:] = torch.tensor([loss.item(), 0.0, float(ent_s), float(kl_s), float(clip_s), float(m[:, t0:t1].sum()), 0.0, 0.0])
        self.launches += 2

    def critic_epoch_grads(self, critic_params, grads, *, state=None, obs=None, returns, mask=None):
        s = self.shapes
        T, B, N = s.n_steps, s.n_envs, s.n_agents
        _, critic = self._nets(None, critic_params)
        m = torch.ones(B, T, dtype=torch.bool) if mask is None else mask.t().bool()
        R = returns.permute(2, 0, 1)
        if s.critic_on_obs:
            v = critic(self._obs(state)).squeeze(-1)
        else:
            v = critic(state.permute(2, 0, 1))
        loss = (((v - R) ** 2).mean(dim=-1) * m.float()).sum()
        critic.zero_grad()
        loss.backward()
        nc = self.n_critic
        grads[:nc] = critic.flat_grads()
        grads[nc:] = torch.tensor([0.0, loss.item(), 0.0, 0.0, 0.0, float(m.sum()), 0.0, 0.0])
        self.launches += 2

    def adam_step_net(self, net, params, grads, exp_avg, exp_avg_sq, *, step=1, step_dev=None, lr=8e-4, beta1=0.9,
                      beta2=0.999, eps=1e-8, max_norm=-1.0, extra_div=1.0, stats_out=None):
        P = params.numel()
        count = grads[P + 5]
        g = grads[:P] / (count * extra_div)
        k = int(step_dev.item()) + 1 if step_dev is not None else step
        actor, critic = self._nets()
        module = actor if net == 0 else critic
        off, sq = 0, []
        for p in module.parameters():
            sq.append(torch.linalg.vector_norm(g[off:off + p.numel()])); off += p.numel()
        nrm = torch.linalg.vector_norm(torch.stack(sq))
        if max_norm > 0:
            g = g * torch.clamp(max_norm / (nrm + 1e-6), max=1.0)
        exp_avg.lerp_(g, 1 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** k, 1 - beta2 ** k
        denom = (exp_avg_sq.sqrt() / (bc2 ** 0.5)).add_(eps)
        if self.weight_decay[net]:
            params.mul_(1 - lr * self.weight_decay[net])
        params.addcdiv_(exp_avg, denom, value=-(lr / bc1))
        if stats_out is not None:
            stats_out[:5] = grads[P:P + 5]
            stats_out[5], stats_out[6], stats_out[7] = nrm, count, 0.0
        if step_dev is not None:
            step_dev.fill_(k)
        self.launches += 1
