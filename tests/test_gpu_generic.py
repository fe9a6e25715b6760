"""GPU parity tests of the layered ("generic") kernels (csrc/generic.cu): the shapes the reference's CLI accepts beyond its
defaults -- any ``*_num_layers`` / hidden width (MME:160-171, 186-196) -- and simple_spread_v3(N) with N != 3 agents,
through the same C-ABI entry points and against the same oracle as the fused path.  Run on the B200 box: ``pytest -m gpu``.
"""
import numpy as np
import pytest
import torch

from oracle import mappo as om
from oracle import spread as osp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cm():
    import cleanmarl_b200 as cm
    from cleanmarl_b200 import _lib
    _lib.load()
    return cm


def start_state(B, N, seed, crowded=True):
    rng = np.random.default_rng(seed)
    pos0 = rng.uniform(-1, 1, (B, N, 2))
    if crowded and N > 1:                                    # collisions: pull agent 1 close to agent 0 in a third of the envs
        close = rng.random(B) < 0.33
        pos0[close, 1] = pos0[close, 0] + rng.uniform(-0.2, 0.2, (int(close.sum()), 2))
    lm = rng.uniform(-1, 1, (B, N, 2))
    env = np.zeros((6 * N, B))
    env[0:2 * N] = pos0.reshape(B, 2 * N).T
    env[4 * N:6 * N] = lm.reshape(B, 2 * N).T
    return pos0, lm, env


# (n_agents, actor_layers, actor_hidden, critic_layers, critic_hidden)
SHAPES = [(3, 2, 32, 1, 128), (2, 1, 32, 1, 64), (5, 1, 48, 2, 96), (4, 3, 64, 1, 32), (1, 1, 32, 1, 64)]


@pytest.mark.parametrize("N,la,ha,lc,hc", SHAPES)
def test_generic_rollout_vs_oracle(cm, N, la, ha, lc, hc):
    """Device rollout of simple_spread_v3(N) with a layered actor vs the float64 oracle env + oracle actor, input driven:
    observations <= 1e-6 (> 99.9 % bit-exact), team reward <= 1e-6, episode return <= 1e-9, fp64 end state <= 1e-12; the
    oracle actor on the device observations with the same noise reproduces the device actions (ties excepted)."""
    from cleanmarl_b200 import engine as E
    B, Tn = 200, 25
    shapes = cm.Shapes.spread(B, N, True, actor_hidden=ha, actor_layers=la, critic_hidden=hc, critic_layers=lc)
    R = osp.raw_obs_dim(N)
    assert shapes.obs_dim == R + N and shapes.state_dim == N * R and shapes.env_rows == 6 * N
    eng = cm.Engine(shapes, device=0)
    dev = eng.device
    actor, critic = om.build_networks(1, obs_dim=shapes.obs_dim, state_dim=shapes.state_dim, actor_hidden=ha, actor_layers=la,
                                      critic_hidden=hc, critic_layers=lc)
    assert eng.n_actor == actor.flat_params().numel() and eng.n_critic == critic.flat_params().numel()
    pos0, lm, env = start_state(B, N, seed=N)
    env_d = torch.from_numpy(env).to(dev)
    q = om.draw_race_noise((Tn, N, 5, B), generator=torch.Generator().manual_seed(B)).to(dev)
    buf = eng.alloc_rollout(with_obs=True)
    eng.rollout(actor.flat_params().to(dev), env_d, buf["state"], buf["actions"], buf["logp"], buf["reward"], noise=q,
                obs=buf["obs"], ep_return=buf["ep_return"])
    torch.cuda.synchronize()
    acts = buf["actions"].cpu().numpy()                       # [T][N][B]
    ref = osp.rollout_batched(pos0, lm, np.transpose(acts, (0, 2, 1)))
    raw_d = buf["state"].cpu().numpy().reshape(Tn, N, R, B).transpose(0, 3, 1, 2)      # [T,B,N,R]
    assert np.abs(raw_d - ref["raw_obs"]).max() < 1e-6
    assert (raw_d == ref["raw_obs"]).mean() > 0.999
    assert np.abs(buf["reward"].cpu().numpy() - ref["reward"].astype(np.float32)).max() < 1e-6
    assert np.abs(buf["ep_return"].cpu().numpy() - ref["reward"].sum(0)).max() < 1e-9
    assert np.abs(env_d.cpu().numpy()[0:2 * N].T.reshape(B, N, 2) - ref["final_pos"]).max() < 1e-12
    assert torch.equal(buf["obs"], E.obs_from_state(buf["state"], N, True))
    obs_ref = buf["obs"].permute(0, 3, 1, 2).cpu()            # [T,B,N,O]
    with torch.no_grad():
        a_ref, lp_ref = om.race_sample(om.actor_logits(actor, obs_ref), q.permute(0, 3, 1, 2).cpu())
    a_dev = torch.from_numpy(np.transpose(acts, (0, 2, 1))).long()
    agree = a_ref == a_dev
    assert agree.float().mean() > 0.999
    assert (buf["logp"].permute(0, 2, 1).cpu()[agree] - lp_ref[agree]).abs().max() < 1e-5
    # the env duck-type one call at a time gives the same trajectory as the rollout kernel
    env2 = torch.from_numpy(env).to(dev)
    st, rw = eng.empty(shapes.state_dim, B), eng.empty(B)
    eng.env_observe(env2, st)
    assert torch.equal(st, buf["state"][0])
    eng.env_step(env2, buf["actions"][0].contiguous(), st, rw)
    assert torch.equal(st, buf["state"][1]) and torch.equal(rw, buf["reward"][0])
    # device reset: agents then landmarks ~ U(-1, 1), velocities zero; a function of (seed, episode)
    e1, e2 = eng.empty(6 * N, B, dtype=torch.float64), eng.empty(6 * N, B, dtype=torch.float64)
    eng.env_reset(e1, seed=5, episode=2); eng.env_reset(e2, seed=5, episode=2)
    assert torch.equal(e1, e2) and (e1[2 * N:4 * N] == 0).all() and float(e1.abs().max()) <= 1.0
    eng.env_reset(e2, seed=5, episode=3)
    assert not torch.equal(e1, e2)


@pytest.mark.parametrize("ippo", [False, True], ids=["mappo", "ippo"])
@pytest.mark.parametrize("N,la,ha,lc,hc", SHAPES[:4])
def test_generic_whole_iteration_vs_oracle(cm, N, la, ha, lc, hc, ippo):
    """One trainer iteration (rollout, critic forward, TD(lambda), 3 epochs with gradient clipping, Adam) on the layered
    kernels against the oracle on the same batch: advantages <= 1e-5, per-epoch losses <= 2e-5 relative, parameters as in
    ``test_tiny_and_ragged_env_counts``; graph replay == eager launches bit for bit."""
    from cleanmarl_b200.mappo import MAPPO, Args
    from cleanmarl_b200 import engine as E
    B = 70
    args = Args(batch_size=B, seed=6, n_agents=N, actor_num_layers=la, actor_hidden_dim=ha, critic_num_layers=lc,
                critic_hidden_dim=hc, clip_gradients=0.5)
    tr = MAPPO(args, ippo=ippo, use_graph=False)
    s = tr.engine.shapes
    actor, critic = om.build_networks(6, obs_dim=s.obs_dim, state_dim=s.obs_dim if ippo else s.state_dim, actor_hidden=ha,
                                      actor_layers=la, critic_hidden=hc, critic_layers=lc)
    assert torch.equal(tr.net.flat.cpu(), torch.cat([actor.flat_params(), critic.flat_params()]))      # reference init order
    _, _, env = start_state(B, N, seed=3)
    noise = torch.empty(25, N, 5, B).exponential_(1, generator=torch.Generator().manual_seed(B))
    tr.iteration(torch.from_numpy(env).cuda(), noise.cuda())
    torch.cuda.synchronize()
    batch = tuple(t.cpu() for t in tr.get_batch())
    assert batch[0].shape == (B, 25, N, s.obs_dim) and batch[4].shape == (B, 25, s.state_dim)
    ret, adv = om.td_lambda_batched(critic, batch[0] if ippo else batch[4], batch[3], batch[7], 0.99, 0.95, N)
    assert (E.heads_to_reference(tr.buf["adv"], N).cpu() - adv).abs().max() < 1e-5
    aopt, copt = om.make_optimizers(actor, critic)
    st = om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, clip_gradients=0.5,
                       critic_on_obs=ippo, flat=True)
    final = torch.cat([actor.flat_params(), critic.flat_params()])
    dp = (tr.net.flat.cpu() - final).abs()
    assert (dp < 3e-6).float().mean() > 0.99 and dp.max() < 2e-4
    es = tr.epoch_stats.cpu()
    for ep in range(3):
        for k, key in ((0, "actor_loss"), (1, "critic_loss"), (2, "entropy"), (5, "actor_grad_norm"), (6, "critic_grad_norm")):
            assert abs(es[ep, k].item() - st[key][ep]) <= 2e-5 * abs(st[key][ep]) + 1e-7, (ep, key)
        assert es[ep, 7].item() == B * 25
    outs = []
    for graph in (False, True):
        t2 = MAPPO(args, ippo=ippo, use_graph=graph)
        for _ in range(3):
            t2.iteration()
        torch.cuda.synchronize()
        outs.append((t2.net.flat.clone(), t2.epoch_stats.clone(), t2.buf["actions"].clone()))
    assert all(torch.equal(a, b) for a, b in zip(*outs))


@pytest.mark.parametrize("N,la,ha,lc,hc", [(3, 2, 128, 2, 128), (6, 1, 32, 1, 64)])
def test_generic_epoch_gradients_synthetic(cm, N, la, ha, lc, hc):
    """cmarl_ppo_epoch_grads(_ex) on the layered kernels: seeded synthetic batch with ragged masks and random avail, the
    clipped value loss and an env-block range, per-tensor gradients <= 2e-5 of the tensor's max vs the oracle's autograd
    (relu-kink samples masked, see tests/test_gpu_parity.py), deterministic run to run; several time-step chunks."""
    from cleanmarl_b200 import engine as E
    from test_gpu_parity import _relu_kink_samples, ragged_mask
    B, Tn = 300, 25
    R = osp.raw_obs_dim(N)
    shapes = cm.Shapes.spread(B, N, True, actor_hidden=ha, actor_layers=la, critic_hidden=hc, critic_layers=lc)
    actor, critic = om.build_networks(5, obs_dim=R + N, state_dim=N * R, actor_hidden=ha, actor_layers=la, critic_hidden=hc,
                                      critic_layers=lc)
    gen = torch.Generator().manual_seed(N)
    raw = torch.rand(B, Tn, N, R, generator=gen) * 4 - 2
    obs = torch.cat([raw, torch.eye(N).expand(B, Tn, N, N)], dim=-1).contiguous()
    states = raw.reshape(B, Tn, N * R).contiguous()
    actions = torch.randint(0, 5, (B, Tn, N), generator=gen)
    avail = torch.rand(B, Tn, N, 5, generator=gen) > 0.2
    avail.scatter_(-1, actions.unsqueeze(-1), True)
    with torch.no_grad():
        logp = torch.distributions.Categorical(logits=om.actor_logits(actor, obs, avail)).log_prob(actions) \
            + 0.1 * torch.randn(B, Tn, N, generator=gen)
    mask = ragged_mask(B, Tn, gen) & ~_relu_kink_samples(actor, critic, obs, states)
    adv = (torch.randn(B, Tn, 1, generator=gen) * 3).expand(B, Tn, N).contiguous()
    ret = (torch.randn(B, Tn, 1, generator=gen) * 5).expand(B, Tn, N).contiguous()
    v_old = (ret[..., :1] + torch.randn(B, Tn, 1, generator=gen) * 0.6).expand(B, Tn, N).contiguous()
    batch = (obs, actions, logp, torch.zeros(B, Tn), states, avail, torch.zeros(B, Tn), mask)
    eng = cm.Engine(shapes, device=0)
    dev = eng.device
    d = E.to_device_layout(batch, dev, with_obs=False)
    params = torch.cat([actor.flat_params(), critic.flat_params()]).to(dev)
    kw = dict(state=d["state"], actions=d["actions"], logp_old=d["logp"], adv=E.heads_to_device(adv, 1, dev),
              returns=E.heads_to_device(ret, 1, dev), mask=d["mask"], avail=d["avail"], clip=0.2, ent_coef=0.001)
    g1, g2 = eng.empty(eng.n_params + 8), eng.empty(eng.n_params + 8)
    for lo, hi, vc in ((0, B, -1.0), (0, B, 0.2), (100, 231, 0.2)):
        sl = slice(lo, hi)
        ext = dict(value_clip=vc, values_old=E.heads_to_device(v_old, 1, dev), env_begin=lo, env_count=hi - lo)
        eng.ppo_epoch_grads(params, g1, **kw, **ext)
        eng.ppo_epoch_grads(params, g2, **kw, **ext)
        assert torch.equal(g1, g2)
        actor.zero_grad(); critic.zero_grad()
        okw = {} if vc <= 0 else dict(value_clip=vc, values_old=v_old[sl])
        out = om.ppo_epoch_flat(actor, critic, obs[sl], actions[sl], logp[sl], states[sl], avail[sl], mask[sl], adv[sl], ret[sl],
                                0.2, 0.001, **okw)
        out.actor_loss.backward(); out.critic_loss.backward()
        g = g1.cpu()
        n = g[eng.n_params + 5].item()
        assert n == float(mask[sl].sum())
        ref = torch.cat([actor.flat_grads(), critic.flat_grads()])
        off = 0
        for net in (actor, critic):
            for p in net.parameters():
                k = p.numel()
                a, b = g[off:off + k] / n, ref[off:off + k]
                scale = max(b.abs().max().item(), 1e-6)
                assert (a - b).abs().max().item() <= 2e-5 * scale + 1e-9, (lo, hi, vc, tuple(p.shape))
                off += k
        st = g[eng.n_params:] / n
        assert abs(st[0].item() - out.actor_loss.item()) < 1e-5 * abs(out.actor_loss.item()) + 1e-7
        assert abs(st[1].item() - out.critic_loss.item()) < 1e-5 * abs(out.critic_loss.item())
        assert abs(st[2].item() - out.entropy.item()) < 1e-5 * abs(out.entropy.item())
        assert abs(st[4].item() - float(out.clipfrac)) < 1e-6


def test_generic_actor_act_and_adam(cm):
    """K2 alone and K8 on a layered shape: logits <= 2e-6, actions of the oracle sampler on the device logits; three Adam
    steps with gradient clipping vs torch.optim.Adam <= 1e-7 (incl. the device-resident step counter)."""
    N, B = 4, 500
    R = osp.raw_obs_dim(N)
    shapes = cm.Shapes.spread(B, N, True, n_steps=1, actor_hidden=96, actor_layers=2, critic_hidden=128, critic_layers=2)
    eng = cm.Engine(shapes, device=0)
    dev = eng.device
    actor, critic = om.build_networks(3, obs_dim=R + N, state_dim=N * R, actor_hidden=96, actor_layers=2, critic_hidden=128,
                                      critic_layers=2)
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(B, N, R + N, generator=gen) * 2 - 1
    avail = torch.rand(B, N, 5, generator=gen) > 0.2
    q = torch.empty(B, N, 5).exponential_(1, generator=gen)
    actions, logp, logits = eng.empty(N, B, dtype=torch.int32), eng.empty(N, B), eng.empty(N, 5, B)
    eng.actor_act(actor.flat_params().to(dev), x.permute(1, 2, 0).contiguous().to(dev), q.permute(1, 2, 0).contiguous().to(dev),
                  actions, logp, avail=avail.permute(1, 2, 0).contiguous().to(torch.uint8).to(dev), logits=logits)
    with torch.no_grad():
        ref_logits = om.actor_logits(actor, x, avail)
    lg = logits.permute(2, 0, 1).cpu()
    assert (lg - ref_logits)[avail].abs().max() < 2e-6
    a2, lp2 = om.race_sample(lg, q)
    assert (a2 == actions.permute(1, 0).cpu().long()).float().mean() > 0.9995
    # Adam
    aopt, copt = om.make_optimizers(actor, critic)
    params = torch.cat([actor.flat_params(), critic.flat_params()]).to(dev)
    m, v = torch.zeros_like(params), torch.zeros_like(params)
    p2, m2, v2 = params.clone(), m.clone(), v.clone()
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    stats = eng.empty(8)
    for step in range(1, 4):
        ga = torch.randn(eng.n_actor, generator=gen) * 0.3
        gc = torch.randn(eng.n_critic, generator=gen) * 2.0
        n = 1600.0
        grads = torch.cat([ga * n, gc * n, torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, n, 0.0, 0.0])]).to(dev)
        for net, gflat in ((actor, ga), (critic, gc)):
            off = 0
            for p in net.parameters():
                p.grad = gflat[off:off + p.numel()].reshape(p.shape).clone()
                off += p.numel()
        na, nc = om.norm_d([p.grad for p in actor.parameters()]), om.norm_d([p.grad for p in critic.parameters()])
        torch.nn.utils.clip_grad_norm_(actor.parameters(), max_norm=0.5)
        torch.nn.utils.clip_grad_norm_(critic.parameters(), max_norm=0.5)
        aopt.step(); copt.step()
        eng.clip_adam_step(params, grads, m, v, step=step, max_norm=0.5, stats_out=stats)
        eng.clip_adam_step(p2, grads, m2, v2, step_dev=cnt, max_norm=0.5)
        ref = torch.cat([actor.flat_params(), critic.flat_params()])
        assert (params.cpu() - ref).abs().max().item() < 1e-7
        s = stats.cpu()
        assert abs(s[5].item() - float(na)) < 1e-5 * float(na) and abs(s[6].item() - float(nc)) < 1e-5 * float(nc)
        assert s[7].item() == n
    assert int(cnt.item()) == 3 and torch.equal(params, p2)
