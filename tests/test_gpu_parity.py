"""GPU parity tests: every kernel of libcmarl_b200.so against the CPU oracle (oracle/), through the
C ABI.  Run on the B200 box: ``pytest -m gpu``.  Tolerances are stated per test.
"""
import numpy as np
import pytest
import torch

from oracle import mappo as om
from oracle import spread as osp

pytestmark = pytest.mark.gpu


def T(x):
    return torch.from_numpy(np.asarray(x))


@pytest.fixture(scope="module")
def cm():
    import cleanmarl_b200 as cm
    from cleanmarl_b200 import _lib
    _lib.load()          # raises if the CUDA extension is missing: no fallback
    return cm


TC = [False, True]      # fp32 FFMA chain / tcgen05 3xTF32 chain: the same tolerances hold for both


def make_engine(cm, B, T_=25, tc=None, **kw):
    return cm.Engine(cm.Shapes(n_envs=B, n_steps=T_, **kw), device=0, tensor_cores=tc)


def golden_setup(cm, g, ippo, tc):
    """Networks (the reference's init for the fixture's seed / widths / layer counts) and the engine of a G8 fixture."""
    kw = dict(actor_hidden=int(g["actor_hidden_dim"]), critic_hidden=int(g["critic_hidden_dim"]))
    if "actor_num_layers" in g:
        kw.update(actor_layers=int(g["actor_num_layers"]), critic_layers=int(g["critic_num_layers"]))
    actor, critic = om.build_networks(int(g["seed"]), state_dim=21 if ippo else 54, **kw)
    eng = make_engine(cm, int(g["B"]), tc=tc, critic_on_obs=ippo, **kw)
    return actor, critic, eng


def flat_params(actor, critic, device):
    return torch.cat([actor.flat_params(), critic.flat_params()]).to(device).contiguous()


def ragged_mask(B, Tn, gen, frac=0.4):
    lengths = torch.where(torch.rand(B, generator=gen) < frac,
                          torch.randint(1, Tn + 1, (B,), generator=gen), torch.full((B,), Tn))
    return (torch.arange(Tn)[None, :] < lengths[:, None])


# ----------------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("B,V", [(1000, 1), (333, 3), (4096, 1)])
def test_td_lambda_scan_bit_exact(cm, B, V):
    """K5 vs oracle.td_lambda_scan on identical values: bit-exact (same fp32 op order, MME:496-504)."""
    from cleanmarl_b200 import engine as E
    Tn = 25
    g = torch.Generator().manual_seed(B + V)
    values = torch.randn(B, Tn, V, generator=g) * 5
    reward = -torch.rand(B, Tn, generator=g) * 4
    mask = ragged_mask(B, Tn, g)
    ret, adv = om.td_lambda_scan(values, reward, mask, 0.99, 0.95)
    eng = make_engine(cm, B, critic_on_obs=(V == 3), critic_hidden=32 if V == 3 else 64)
    dev = eng.device
    v_d = values.permute(1, 2, 0).contiguous().to(dev)
    r_d = reward.permute(1, 0).contiguous().to(dev)
    m_d = mask.permute(1, 0).contiguous().to(torch.uint8).to(dev)
    ret_d, adv_d = torch.empty_like(v_d), torch.empty_like(v_d)
    eng.td_lambda(v_d, r_d, ret_d, adv_d, 0.99, 0.95, mask=m_d)
    assert torch.equal(ret_d.permute(2, 0, 1).cpu(), ret)
    assert torch.equal(adv_d.permute(2, 0, 1).cpu(), adv)
    # no mask == all ones
    ret1, adv1 = om.td_lambda_scan(values, reward, torch.ones_like(mask), 0.99, 0.95)
    eng.td_lambda(v_d, r_d, ret_d, adv_d, 0.99, 0.95)
    assert torch.equal(ret_d.permute(2, 0, 1).cpu(), ret1) and torch.equal(adv_d.permute(2, 0, 1).cpu(), adv1)


# ----------------------------------------------------------------------------------------- K4 (+K5) on the golden run
@pytest.mark.parametrize("tc", TC, ids=["ffma", "tc"])
@pytest.mark.parametrize("name,ippo", [("g8_mappo", False), ("g8_ippo", True), ("g8_mappo_deep", False)])
def test_critic_td_lambda_vs_reference_run(cm, golden, name, ippo, tc):
    """K4+K5 on the batch of a real reference iteration: returns/advantages within 1e-5 (north star).  g8_mappo_deep
    (--actor_num_layers 2 --critic_hidden_dim 128) runs the layered kernels (csrc/generic.cu)."""
    from cleanmarl_b200 import engine as E
    g = golden(name)
    actor, critic, eng = golden_setup(cm, g, ippo, tc)
    B = int(g["B"])
    dev = eng.device
    batch = tuple(T(g[k]) for k in ("obs", "actions", "log_probs", "reward", "states", "avail", "done", "mask"))
    d = E.to_device_layout(batch, dev)
    values = eng.empty(25, eng.n_heads, B)
    cp = critic.flat_params().to(dev)
    for use_obs in ((True, False) if ippo else (False,)):
        eng.critic_values(cp, values, state=d["state"], obs=d["obs"] if use_obs else None)
        with torch.no_grad():
            v_ref = critic(batch[0] if ippo else batch[4]).squeeze(-1)      # [B,T] or [B,T,N]
        v_ref = v_ref.unsqueeze(-1) if not ippo else v_ref
        assert (values.permute(2, 0, 1).cpu() - v_ref).abs().max() < 2e-6
        ret, adv = torch.empty_like(values), torch.empty_like(values)
        eng.td_lambda(values, d["reward"], ret, adv, float(g["gamma"]), float(g["td_lambda"]), mask=d["mask"])
        ret_r = E.heads_to_reference(ret, 3).cpu()
        adv_r = E.heads_to_reference(adv, 3).cpu()
        assert (ret_r - T(g["return_lambda"])).abs().max() < 1e-5
        assert (adv_r - T(g["advantages"])).abs().max() < 1e-5


# ----------------------------------------------------------------------------------------- K7
def _oracle_epoch(actor, critic, batch, adv, ret, ippo, clip=0.2, ent=0.001, flat=True):
    obs, actions, logp, reward, states, avail, done, mask = batch
    actor.zero_grad(); critic.zero_grad()
    fn = om.ppo_epoch_flat if flat else om.ppo_epoch_loop
    out = fn(actor, critic, obs, actions, logp, obs if ippo else states, avail, mask, adv, ret, clip, ent)
    out.actor_loss.backward(); out.critic_loss.backward()
    return out, actor.flat_grads(), critic.flat_grads()


def _check_epoch(eng, E, actor, critic, batch, adv, ret, ippo, use_obs, use_mask=True, use_avail=True, flat=True):
    dev = eng.device
    d = E.to_device_layout(batch, dev)
    params = flat_params(actor, critic, dev)
    grads = eng.empty(eng.n_params + 8)
    eng.ppo_epoch_grads(params, grads, state=d["state"], obs=d["obs"] if use_obs else None, actions=d["actions"],
                        logp_old=d["logp"], adv=E.heads_to_device(adv, eng.n_heads, dev),
                        returns=E.heads_to_device(ret, eng.n_heads, dev), mask=d["mask"] if use_mask else None,
                        avail=d["avail"] if use_avail else None, clip=0.2, ent_coef=0.001)
    out, ga, gc = _oracle_epoch(actor, critic, batch, adv, ret, ippo, flat=flat)
    gcpu = grads.cpu()
    n = gcpu[eng.n_params + 5].item()
    assert n == float(batch[7].sum())
    g_a = gcpu[:eng.n_actor] / n
    g_c = gcpu[eng.n_actor:eng.n_params] / n
    st = gcpu[eng.n_params:] / n
    # per-tensor gradient check: max abs error <= 2e-5 of the tensor's max magnitude (fp32 reassociation)
    for name, gd, gr, net in (("actor", g_a, ga, actor), ("critic", g_c, gc, critic)):
        off = 0
        for p in net.parameters():
            k = p.numel()
            a, b = gd[off:off + k], gr[off:off + k]
            scale = max(b.abs().max().item(), 1e-6)
            assert (a - b).abs().max().item() <= 2e-5 * scale + 1e-9, (name, tuple(p.shape), (a - b).abs().max().item(), scale)
            off += k
    rel = lambda x, y: abs(x - y) / max(abs(y), 1e-6)
    assert rel(st[0].item(), out.actor_loss.item()) < 1e-5
    assert rel(st[1].item(), out.critic_loss.item()) < 1e-5
    assert rel(st[2].item(), out.entropy.item()) < 1e-5
    assert abs(st[3].item() - out.kl.item()) < 1e-6 + 1e-4 * abs(out.kl.item())
    assert abs(st[4].item() - float(out.clipfrac)) < 1e-6
    return grads


@pytest.mark.parametrize("tc", TC, ids=["ffma", "tc"])
@pytest.mark.parametrize("name,ippo", [("g8_mappo", False), ("g8_ippo", True), ("g8_mappo_deep", False)])
def test_ppo_epoch_grads_vs_reference_loop(cm, golden, name, ippo, tc):
    """K7 on the batch of a real reference iteration vs autograd through the reference's own loop form."""
    from cleanmarl_b200 import engine as E
    g = golden(name)
    actor, critic, eng = golden_setup(cm, g, ippo, tc)
    batch = tuple(T(g[k]) for k in ("obs", "actions", "log_probs", "reward", "states", "avail", "done", "mask"))
    for use_obs in (False, True):
        _check_epoch(eng, E, actor, critic, batch, T(g["advantages"]), T(g["return_lambda"]), ippo, use_obs, flat=False)


@pytest.mark.parametrize("tc", TC, ids=["ffma", "tc"])
@pytest.mark.parametrize("B,ippo,hid", [(300, False, (32, 64)), (1024, False, (32, 64)), (515, True, (32, 32)),
                                        (256, False, (64, 32)), (260, True, (64, 64))])
def test_ppo_epoch_grads_synthetic(cm, B, ippo, hid, tc):
    """K7 on seeded synthetic batches: full TMA tiles + ragged tail tile, ragged masks, random avail,
    obs rebuilt from state vs explicit obs, both hidden widths."""
    from cleanmarl_b200 import engine as E
    ha, hc = hid
    actor, critic = om.build_networks(5, state_dim=21 if ippo else 54, actor_hidden=ha, critic_hidden=hc)
    batch = list(om.synthetic_batch(B, seed=B, actor=actor))
    gen = torch.Generator().manual_seed(B)
    batch[7] = ragged_mask(B, 25, gen)
    avail = torch.rand(B, 25, 3, 5, generator=gen) > 0.2
    avail.scatter_(-1, batch[1].unsqueeze(-1), True)            # the taken action is always available
    batch[5] = avail
    with torch.no_grad():
        batch[2] = torch.distributions.Categorical(logits=om.actor_logits(actor, batch[0], avail)).log_prob(batch[1]) \
            + 0.1 * torch.randn(B, 25, 3, generator=gen)
    V = 3 if ippo else 1
    adv = torch.randn(B, 25, V, generator=gen).expand(B, 25, 3).contiguous() * 3
    ret = torch.randn(B, 25, V, generator=gen).expand(B, 25, 3).contiguous() * 5
    eng = make_engine(cm, B, tc=tc, critic_on_obs=ippo, actor_hidden=ha, critic_hidden=hc)
    for use_obs in (False, True):
        _check_epoch(eng, E, actor, critic, tuple(batch), adv, ret, ippo, use_obs)
    g1 = _check_epoch(eng, E, actor, critic, tuple(batch), adv, ret, ippo, False).clone()
    g2 = _check_epoch(eng, E, actor, critic, tuple(batch), adv, ret, ippo, False)
    assert torch.equal(g1, g2)                                   # deterministic (fixed-order reductions)


# ----------------------------------------------------------------------------------------- K8
@pytest.mark.parametrize("max_norm", [-1.0, 0.5])
def test_clip_adam_matches_torch_adam(cm, max_norm):
    """K8 vs torch.optim.Adam (+clip_grad_norm_) for 3 steps: parameters within 1e-7 abs."""
    actor, critic = om.build_networks(3)
    eng = make_engine(cm, 64)
    dev = eng.device
    aopt, copt = om.make_optimizers(actor, critic)
    params = flat_params(actor, critic, dev)
    m, v = torch.zeros_like(params), torch.zeros_like(params)
    gen = torch.Generator().manual_seed(0)
    stats = eng.empty(8)
    for step in range(1, 4):
        ga = torch.randn(eng.n_actor, generator=gen) * 0.3
        gc = torch.randn(eng.n_critic, generator=gen) * 2.0
        n = 1600.0
        grads = torch.cat([ga * n, gc * n, torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, n, 0.0, 0.0]) * 1.0]).to(dev)
        for net, gflat in ((actor, ga), (critic, gc)):
            off = 0
            for p in net.parameters():
                p.grad = gflat[off:off + p.numel()].reshape(p.shape).clone()
                off += p.numel()
        na, nc = om.norm_d([p.grad for p in actor.parameters()]), om.norm_d([p.grad for p in critic.parameters()])
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_(actor.parameters(), max_norm=max_norm)
            torch.nn.utils.clip_grad_norm_(critic.parameters(), max_norm=max_norm)
        aopt.step(); copt.step()
        eng.clip_adam_step(params, grads, m, v, step=step, max_norm=max_norm, stats_out=stats)
        ref = torch.cat([actor.flat_params(), critic.flat_params()])
        assert (params.cpu() - ref).abs().max().item() < 1e-7
        s = stats.cpu()
        assert abs(s[5].item() - float(na)) < 1e-5 * float(na) and abs(s[6].item() - float(nc)) < 1e-5 * float(nc)
        assert abs(s[0].item() - 1.0 / n) < 1e-9 and s[7].item() == n
    # device-resident step counter (CUDA-graph friendly) gives the same result as the host step
    p2 = flat_params(*om.build_networks(3), dev)
    m2, v2 = torch.zeros_like(p2), torch.zeros_like(p2)
    p3, m3, v3 = p2.clone(), m2.clone(), v2.clone()
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    for step in range(1, 3):
        eng.clip_adam_step(p2, grads, m2, v2, step=step, max_norm=max_norm)
        eng.clip_adam_step(p3, grads, m3, v3, step_dev=cnt, max_norm=max_norm)
    assert int(cnt.item()) == 2 and torch.equal(p2, p3)


# ----------------------------------------------------------------------------------------- K2
def test_actor_act_vs_reference_sample(cm, golden):
    """K2 on the inputs of a real ``Actor.act`` call (g3): logits within 2e-6, log-probs within 1e-5,
    action indices bit-exact wherever the race is not a numerical tie (|margin| > 1e-5 relative)."""
    g = golden("g3_sample")
    B = g["x"].shape[0]
    eng = make_engine(cm, B, T_=1)
    dev = eng.device
    obs = T(g["x"]).permute(1, 2, 0).contiguous().to(dev)                 # [N][O][B]
    avail = T(g["avail"]).permute(1, 2, 0).contiguous().to(torch.uint8).to(dev)
    q = T(g["q"]).permute(1, 2, 0).contiguous().to(dev)
    actions = eng.empty(3, B, dtype=torch.int32)
    logp, logits = eng.empty(3, B), eng.empty(3, 5, B)
    eng.actor_act(T(g["params"]).to(dev), obs, q, actions, logp, avail=avail, logits=logits)
    lg = logits.permute(2, 0, 1).cpu()
    ref_logits = T(g["logits"])
    assert (lg - ref_logits).abs().max() < 2e-6 * max(1.0, ref_logits[ref_logits > -1e8].abs().max().item())
    act = actions.permute(1, 0).cpu().long()
    ref_act = T(g["actions"])
    # race margin of the reference decision
    probs = torch.softmax(ref_logits, dim=-1)
    r = probs / T(g["q"])
    top2 = r.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]) / top2[..., 0]
    decisive = margin > 1e-5
    assert decisive.float().mean() > 0.999
    assert torch.equal(act[decisive], ref_act[decisive])
    same = act == ref_act
    assert (logp.permute(1, 0).cpu()[same] - T(g["logp"])[same]).abs().max() < 1e-5
    # oracle sampler on the DEVICE logits must reproduce the device actions bit-exactly
    a2, _ = om.race_sample(lg, T(g["q"]))
    assert (a2 == act).float().mean() > 0.9995


# ----------------------------------------------------------------------------------------- K1+K2+K3
@pytest.mark.parametrize("hid", [32, 64])
@pytest.mark.parametrize("path", ["tc", "ffma"])
@pytest.mark.parametrize("B", [96, 1000])
def test_rollout_vs_oracle(cm, B, path, hid, monkeypatch):
    """Device rollout (layer 2 on the tensor cores: rollout_tc_kernel, the default; CUDA-core rollout_kernel) vs the
    float64 oracle env + oracle actor, input driven (start positions and race
    noise supplied).  Physics: the oracle env replays the DEVICE actions open loop -> observations within
    1e-6 and team reward within 1e-6 of the oracle at every step.  Policy: the oracle actor on the device
    observations with the same noise reproduces the device actions (ties excepted) and log-probs."""
    from cleanmarl_b200 import engine as E
    monkeypatch.setenv("CMARL_ROLLOUT", path)
    Tn = 25
    actor, critic = om.build_networks(1, actor_hidden=hid)
    eng = make_engine(cm, B, actor_hidden=hid)
    dev = eng.device
    rng = np.random.default_rng(B)
    pos0 = rng.uniform(-1, 1, (B, 3, 2))
    # make collisions common: pull agent 1 close to agent 0 in a third of the envs
    close = rng.random(B) < 0.33
    pos0[close, 1] = pos0[close, 0] + rng.uniform(-0.2, 0.2, (int(close.sum()), 2))
    lm = rng.uniform(-1, 1, (B, 3, 2))
    env = np.zeros((18, B))
    env[0:6] = pos0.reshape(B, 6).T
    env[12:18] = lm.reshape(B, 6).T
    env_d = torch.from_numpy(env).to(dev)
    q = om.draw_race_noise((Tn, 3, 5, B), generator=torch.Generator().manual_seed(B)).to(dev)
    buf = eng.alloc_rollout(with_obs=True)
    eng.rollout(actor.flat_params().to(dev), env_d, buf["state"], buf["actions"], buf["logp"], buf["reward"],
                noise=q, obs=buf["obs"], ep_return=buf["ep_return"])
    torch.cuda.synchronize()
    acts = buf["actions"].cpu().numpy()                       # [T][N][B]
    ref = osp.rollout_batched(pos0, lm, np.transpose(acts, (0, 2, 1)))
    raw_d = buf["state"].cpu().numpy().reshape(Tn, 3, 18, B).transpose(0, 3, 1, 2)     # [T,B,3,18]
    err = np.abs(raw_d - ref["raw_obs"]).max()
    assert err < 1e-6, err
    bit_exact = (raw_d == ref["raw_obs"]).mean()
    assert bit_exact > 0.999, bit_exact
    rew_d = buf["reward"].cpu().numpy()
    assert np.abs(rew_d - ref["reward"].astype(np.float32)).max() < 1e-6
    assert np.abs(buf["ep_return"].cpu().numpy() - ref["reward"].sum(0)).max() < 1e-9
    assert np.abs(env_d.cpu().numpy()[0:6].T.reshape(B, 3, 2) - ref["final_pos"]).max() < 1e-12
    # obs output = raw + one-hot ids, and equals what K7 rebuilds from state
    assert torch.equal(buf["obs"], E.obs_from_state(buf["state"]))
    # policy parity on the device observations
    obs_ref = buf["obs"].permute(0, 3, 1, 2).cpu()            # [T,B,N,O]
    with torch.no_grad():
        logits = om.actor_logits(actor, obs_ref)
        a_ref, lp_ref = om.race_sample(logits, q.permute(0, 3, 1, 2).cpu())
    a_dev = torch.from_numpy(np.transpose(acts, (0, 2, 1))).long()
    agree = (a_ref == a_dev)
    assert agree.float().mean() > 0.9995, agree.float().mean()
    lp_dev = buf["logp"].permute(0, 2, 1).cpu()
    assert (lp_dev[agree] - lp_ref[agree]).abs().max() < 1e-5


def test_rollout_device_rng_and_reset(cm):
    """Device RNG path: reset draws U(-1,1) positions, zero velocities; actions follow the policy
    distribution (chi-square-ish frequency check); same (seed, episode) -> identical rollout."""
    B = 4096
    actor, _ = om.build_networks(1)
    eng = make_engine(cm, B)
    dev = eng.device
    env = eng.empty(18, B, dtype=torch.float64)
    outs = []
    for rep in range(2):
        eng.env_reset(env, seed=7, episode=3)
        e0 = env.clone()
        buf = eng.alloc_rollout()
        eng.rollout(actor.flat_params().to(dev), env, buf["state"], buf["actions"], buf["logp"], buf["reward"],
                    seed=7, episode=3)
        outs.append((e0, buf))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1]["actions"], outs[1][1]["actions"])
    e0 = outs[0][0].cpu().numpy()
    assert (np.abs(e0[0:6]) <= 1).all() and (np.abs(e0[12:18]) <= 1).all() and (e0[6:12] == 0).all()
    assert abs(e0[0:6].mean()) < 0.02 and abs(e0[0:6].var() - 1 / 3) < 0.02
    eng.env_reset(env, seed=7, episode=4)
    assert not torch.equal(env, outs[0][0])
    # empirical action frequencies at t=0 vs the policy's probabilities
    buf = outs[0][1]
    from cleanmarl_b200 import engine as E
    obs0 = E.obs_from_state(buf["state"])[0].permute(2, 0, 1).cpu()           # [B,N,O]
    with torch.no_grad():
        p = torch.softmax(om.actor_logits(actor, obs0), -1).mean(dim=(0, 1))
    freq = torch.bincount(buf["actions"][0].flatten().cpu().long(), minlength=5).float() / (3 * B)
    assert (freq - p).abs().max() < 0.02


# ----------------------------------------------------------------------------------------- whole iteration
@pytest.mark.parametrize("tc", TC, ids=["ffma", "tc"])
@pytest.mark.parametrize("name,ippo", [("g8_mappo", False), ("g8_mappo_flags", False), ("g8_ippo", True),
                                       ("g8_mappo_deep", False)])
def test_whole_update_vs_reference_run(cm, golden, name, ippo, tc):
    """K4+K5(+K6)+3x(K7+K8) from the reference's initial parameters on the reference's batch: per-epoch
    statistics and the final parameters follow the unmodified reference run (g8_mappo_deep: the reference run with
    --actor_num_layers 2 --critic_hidden_dim 128 --clip_gradients 0.5, on the layered kernels)."""
    from cleanmarl_b200 import engine as E
    g = golden(name)
    B = int(g["B"])
    actor, critic, eng = golden_setup(cm, g, ippo, tc)
    dev = eng.device
    batch = tuple(T(g[k]) for k in ("obs", "actions", "log_probs", "reward", "states", "avail", "done", "mask"))
    d = E.to_device_layout(batch, dev)
    params = flat_params(actor, critic, dev)
    values = eng.empty(25, eng.n_heads, B)
    ret, adv = torch.empty_like(values), torch.empty_like(values)
    stats64 = eng.empty(4, dtype=torch.float64)
    # (the golden reward is already normalised when the flag is set: RolloutBuffer.get_batch does it)
    eng.critic_values(params[eng.n_actor:], values, state=d["state"], obs=None)
    eng.td_lambda(values, d["reward"], ret, adv, float(g["gamma"]), float(g["td_lambda"]), mask=d["mask"])
    if bool(g["normalize_advantage"]):
        eng.normalize(adv, eng.n_heads, 1, 0, stats64, mask=d["mask"]); eng.normalize(adv, eng.n_heads, 1, 1, stats64, mask=d["mask"])
    if bool(g["normalize_return"]):
        eng.normalize(ret, eng.n_heads, 1, 0, stats64, mask=d["mask"]); eng.normalize(ret, eng.n_heads, 1, 1, stats64, mask=d["mask"])
    m_ = d["mask"].bool().permute(1, 0).cpu()
    tol = 1e-5 if not bool(g["normalize_advantage"]) else 2e-6
    assert (E.heads_to_reference(adv, 3).cpu() - T(g["advantages"]))[m_].abs().max() < tol
    assert (E.heads_to_reference(ret, 3).cpu() - T(g["return_lambda"]))[m_].abs().max() < tol
    m, v = torch.zeros_like(params), torch.zeros_like(params)
    grads, stats = eng.empty(eng.n_params + 8), eng.empty(8)
    for ep in range(int(g["epochs"])):
        eng.ppo_epoch_grads(params, grads, state=d["state"], actions=d["actions"], logp_old=d["logp"], adv=adv,
                            returns=ret, mask=d["mask"], avail=d["avail"], clip=float(g["ppo_clip"]),
                            ent_coef=float(g["entropy_coef"]))
        eng.clip_adam_step(params, grads, m, v, step=ep + 1, lr_actor=float(g["lr_actor"]),
                           lr_critic=float(g["lr_critic"]), max_norm=float(g["clip_gradients"]), stats_out=stats)
        s = stats.cpu().numpy()
        ref = [g["actor_losses"][ep], g["critic_losses"][ep], g["entropies"][ep], g["kls"][ep], g["clipfracs"][ep],
               g["actor_grad_norms"][ep], g["critic_grad_norms"][ep]]
        for k in (0, 1, 2, 5, 6):
            assert abs(s[k] - ref[k]) <= 2e-5 * abs(ref[k]) + 1e-7, (ep, k, s[k], ref[k])
        assert abs(s[3] - ref[3]) < 1e-6 + 1e-3 * abs(ref[3])
        assert abs(s[4] - ref[4]) < 1e-6
    final = torch.cat([T(g["actor_final"]), T(g["critic_final"])])
    # 3 Adam steps of lr 8e-4 move each parameter by <= 2.4e-3; agreement to 1e-6 abs (the deep fixture's 150 samples leave
    # a few gradients near Adam's eps, where a 1e-9 difference is a visible fraction of lr: 99.5 % within 1e-6, all within 5e-6)
    dp = (params.cpu() - final).abs()
    if name == "g8_mappo_deep":
        assert (dp < 1e-6).float().mean() > 0.995 and dp.max().item() < 5e-6
    else:
        assert dp.max().item() < 1e-6


def test_reward_normalisation(cm):
    """K6 mode 0 vs RolloutBuffer.get_batch's normalize_reward (MME:143-146)."""
    B, Tn = 777, 25
    gen = torch.Generator().manual_seed(2)
    reward = -torch.rand(B, Tn, generator=gen) * 4
    mask = ragged_mask(B, Tn, gen)
    ref = om.normalize_reward_(reward, mask)
    eng = make_engine(cm, B)
    dev = eng.device
    r = reward.permute(1, 0).contiguous().to(dev)
    st = eng.empty(4, dtype=torch.float64)
    m = mask.permute(1, 0).contiguous().to(torch.uint8).to(dev)
    eng.normalize(r, 1, 0, 0, st, mask=m)
    eng.normalize(r, 1, 0, 1, st, mask=m)
    out = r.permute(1, 0).cpu()
    assert (out - ref)[mask].abs().max() < 2e-6
    assert torch.equal(out[~mask], reward[~mask])


# ----------------------------------------------------------------------------------------- multi-GPU (NCCL)
@pytest.mark.parametrize("comm", ["p2p", "nccl"])
@pytest.mark.parametrize("flags", ["plain", "flags", "recurrent"])
def test_two_gpus_nccl_equal_one(cm, tmp_path, flags, comm):
    """Envs sharded over 2 GPUs (one process per GPU; gradient sums exchanged inside the Adam kernel over peer memory
    (p2p, the default) or by one NCCL all-reduce of 9 678 floats per epoch) == 1 GPU on all envs: replicas bit-identical
    to each other, parameters within fp32 reassociation of the single-GPU run."""
    import os
    import socket
    import subprocess
    import sys
    from pathlib import Path
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import mgpu_worker
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    B = 1024
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(Path(__file__).resolve().parent / "mgpu_worker.py"), str(tmp_path), str(B), flags]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, CMARL_COMM=comm))
    assert r.returncode == 0, r.stderr[-2000:]
    two = torch.load(tmp_path / "mgpu.pt")
    assert two["comm"] == comm
    kw = {"flags": {"normalize_advantage": True, "clip_gradients": 0.5}, "recurrent": {"recurrent": True}}.get(flags, {})
    one = mgpu_worker.run(B, 0, 1, 0, **kw)
    assert two["step"] == one.step
    dp = (two["params"] - one.net.flat.cpu()).abs()
    if flags == "recurrent":       # 24 actor Adam steps: near-eps gradients amplify reassociation differences
        assert (dp < 3e-6).float().mean() > 0.995 and dp.max() < 1e-4
    elif flags == "flags":
        # Normalised advantages: a handful of parameters (16 of 9 670, profiles/tools/mgpu_diag.py) have a gradient at the
        # round-off level of their tensor (|g| ~ 1e-9 max|g|), so Adam's g / sqrt(v) moves them by ~lr per step in a direction
        # that rounding decides -- in the reference as well.  One and two GPUs group the tiles of the chain kernels'
        # persistent TMEM accumulators differently (4 tiles per flush); with CMARL_TC_FLUSH=1 the two runs agree to 3e-8.
        assert (dp < 2e-6).float().mean() > 0.995 and dp.max() < 1e-4
    else:
        assert dp.max() < 2e-6
    ref = one.epoch_stats.cpu()
    assert (two["stats"] - ref).abs().max() < 1e-4 * ref.abs().max()


# ----------------------------------------------------------------------------------------- CUDA-graph replay
@pytest.mark.parametrize("flags", [{}, {"normalize_reward": True, "normalize_advantage": True, "normalize_return": True,
                                        "clip_gradients": 0.5}], ids=["plain", "normalised+clip"])
@pytest.mark.parametrize("recurrent", [False, True], ids=["mlp", "recurrent"])
def test_graph_replay_equals_eager_launches(cm, recurrent, flags):
    """The trainer's default mode (one captured CUDA graph per iteration, Philox keyed by a device episode counter)
    produces bit-identical parameters, statistics and rollouts to eager launches with the by-value episode argument."""
    from cleanmarl_b200.mappo import MAPPO, Args, ArgsRecurrent
    cls = ArgsRecurrent if recurrent else Args
    out = []
    for graph in (False, True):
        tr = MAPPO(cls(batch_size=512, seed=4, **flags), use_graph=graph)
        for _ in range(4):
            tr.iteration()
        torch.cuda.synchronize()
        assert tr.use_graph == graph and (len(tr._graphs) == 1) == graph
        out.append((tr.net.flat.clone(), tr.epoch_stats.clone(), tr.buf["actions"].clone(), tr.buf["ep_return"].clone(),
                    tr.step, tr.training_step, tr.episode, tr.num_episodes))
    for a, b in zip(*out):
        assert torch.equal(a, b) if isinstance(a, torch.Tensor) else a == b
    # host start states through the second captured graph (reset = False)
    tr = MAPPO(cls(batch_size=512, seed=4), use_graph=True)
    tr2 = MAPPO(cls(batch_size=512, seed=4), use_graph=False)
    env = torch.empty(18, 512, dtype=torch.float64).uniform_(-1, 1)
    env[6:12] = 0
    env = env.cuda()
    for _ in range(3):
        tr.iteration(env_init=env)
        tr2.iteration(env_init=env)
    torch.cuda.synchronize()
    assert torch.equal(tr.net.flat, tr2.net.flat) and torch.equal(tr.buf["actions"], tr2.buf["actions"])


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("kind", ["mlp", "ippo", "recurrent", "mlp_flags"])
def test_launch_chaining_changes_nothing(cm, kind, graph):
    """Launch chaining (cmarl_ctx_set_launch_chaining: every kernel of an iteration a programmatic dependent of the one in
    front of it) only moves launch latency and prologues under the predecessor: parameters, statistics and rollouts are
    bit-identical to normally serialised launches, iteration after iteration (a kernel reading its predecessor's output
    too early would show up here)."""
    from cleanmarl_b200.mappo import MAPPO, Args, ArgsRecurrent
    cls = ArgsRecurrent if kind == "recurrent" else Args
    flags = dict(normalize_advantage=True, normalize_reward=True, clip_gradients=0.5) if kind == "mlp_flags" else {}
    out = []
    for chain in (False, True):
        tr = MAPPO(cls(batch_size=4096 if kind == "mlp" else 640, seed=7, **flags), ippo=kind == "ippo", use_graph=graph)
        tr.chain = chain
        for _ in range(6):
            tr.iteration()
        torch.cuda.synchronize()
        out.append((tr.net.flat.clone(), tr.exp_avg_sq.clone(), tr.epoch_stats.clone(), tr.buf["actions"].clone(),
                    tr.buf["returns"].clone(), tr.buf["ep_return"].clone()))
    for a, b in zip(*out):
        assert torch.equal(a, b)


# ----------------------------------------------------------------------------------------- full size / tiny size
def test_full_size_iteration_properties(cm):
    """BASELINE's largest per-GPU size (65 536 envs, 1.6 M env-steps per rollout) through size-independent properties:
    buffers finite and in range, physics consistent step to step, TD(lambda) recursion re-evaluated with separately
    rounded torch ops on the device bit for bit, A == R - V bit for bit, valid count exact, parameters move."""
    from cleanmarl_b200.mappo import MAPPO, Args
    B, Tn = 65536, 25
    tr = MAPPO(Args(batch_size=B, seed=2), use_graph=False)
    p0 = tr.net.flat.clone()
    tr.iteration()
    torch.cuda.synchronize()
    buf = tr.buf
    st, acts, logp, rew = buf["state"], buf["actions"], buf["logp"], buf["reward"]
    assert torch.isfinite(st).all() and torch.isfinite(logp).all() and torch.isfinite(rew).all()
    assert int(acts.min()) >= 0 and int(acts.max()) <= 4 and float(logp.max()) <= 0.0
    assert float(rew.max()) <= 0.0                                     # -distances - collisions
    raw = st.reshape(Tn, 3, 18, B)
    assert (raw[:, :, 14:18] == 0).all()                               # silent agents: 4 zero communication slots
    # p_pos(t+1) = p_pos(t) + p_vel(t) * dt (positions first, fp64 inside the kernel; fp32 observations here)
    pred = raw[:-1, :, 2:4].double() + raw[:-1, :, 0:2].double() * 0.1
    assert (raw[1:, :, 2:4].double() - pred).abs().max() < 1e-6
    # every agent sees the same landmarks: landmark_rel + own pos is agent-independent
    lm = raw[:, :, 4:10].reshape(Tn, 3, 3, 2, B) + raw[:, :, None, 2:4]
    assert (lm[:, 0] - lm[:, 1]).abs().max() < 1e-6 and (lm[:, 0] - lm[:, 2]).abs().max() < 1e-6
    # action frequencies of the freshly initialised policy are near uniform
    freq = torch.bincount(acts.flatten().long(), minlength=5).float() / acts.numel()
    assert (freq - 0.2).abs().max() < 0.05
    # TD(lambda): R_t = r_t + g (l R_{t+1} + (1 - l) V_{t+1}), bootstrap 0 at the end; separately rounded fp32 ops
    V, R, A = buf["values"][:, 0], buf["returns"][:, 0], buf["adv"][:, 0]
    g = torch.tensor(0.99, device=V.device); l = torch.tensor(0.95, device=V.device)
    oml = torch.tensor(1 - 0.95, dtype=torch.float32, device=V.device)
    last = torch.zeros(B, device=V.device)
    for t in reversed(range(Tn)):
        nv = V[t + 1] if t + 1 < Tn else torch.zeros(B, device=V.device)
        inner = l * last
        inner = inner + oml * nv
        rt = rew[t] + g * inner
        assert torch.equal(rt, R[t]), t
        last = rt
    assert torch.equal(A, R - V)
    s = tr.epoch_stats.cpu()
    assert (s[:, 7] == B * Tn).all() and torch.isfinite(s).all()
    assert not torch.equal(tr.net.flat, p0) and torch.isfinite(tr.net.flat).all()
    assert abs(float(buf["ep_return"].mean()) - float(rew.double().sum(0).mean())) < 1e-5


@pytest.mark.parametrize("B", [1, 5, 33])
@pytest.mark.parametrize("recurrent", [False, True], ids=["mlp", "recurrent"])
def test_tiny_and_ragged_env_counts(cm, B, recurrent):
    """One env, a handful, one more than a warp: every kernel's ragged-tile path against the oracle (whole iteration)."""
    from cleanmarl_b200.mappo import MAPPO, Args, ArgsRecurrent
    from cleanmarl_b200 import engine as E
    from oracle import mappo_lstm as ol
    tr = MAPPO((ArgsRecurrent if recurrent else Args)(batch_size=B, seed=6), use_graph=False)
    actor, critic = (ol if recurrent else om).build_networks(6)
    g = torch.Generator().manual_seed(B)
    env = torch.zeros(18, B, dtype=torch.float64)
    env[0:6] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    env[12:18] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    noise = torch.empty(25, 3, 5, B).exponential_(1, generator=g)
    tr.iteration(env.cuda(), noise.cuda())
    torch.cuda.synchronize()
    batch = tuple(t.cpu() for t in tr.get_batch())
    ret, adv = om.td_lambda_batched(critic, batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    assert (E.heads_to_reference(tr.buf["adv"], 3).cpu() - adv).abs().max() < 1e-5
    aopt, copt = om.make_optimizers(actor, critic)
    if recurrent:
        ol.ppo_update_tbptt(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, tbptt=10)
    else:
        om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001)
    final = torch.cat([actor.flat_params(), critic.flat_params()])
    dp = (tr.net.flat.cpu() - final).abs()
    assert (dp < 3e-6).float().mean() > 0.99 and dp.max() < 2e-4       # near-eps Adam gradients: see test_gpu_recurrent


# ----------------------------------------------------------------------------------------- evaluation loop (MME:614-644)
def test_evaluate_matches_oracle_closed_loop(cm):
    """``evaluate()`` (sampled policy on ``num_eval_ep`` parallel device envs, MME:614-644) against the oracle rolling the
    SAME start states and race noise closed loop (oracle env + oracle actor + exponential race): per-env episode returns
    equal to 1e-9 wherever every race of the episode was decided identically (ties excepted, > 90 % of the envs), and
    (mean, population std, length) as np.mean / np.std / np.mean of MME:642-644.  The eval context is built once."""
    from cleanmarl_b200.mappo import MAPPO, Args, evaluate
    n, Tn = 64, 25
    tr = MAPPO(Args(batch_size=256, seed=9), use_graph=False)
    tr.iteration()                                                   # one update: the policy is no longer the initial one
    torch.cuda.synchronize()
    actor, _ = om.build_networks(9)
    actor.load_flat(tr.net.actor.cpu())
    g = torch.Generator().manual_seed(21)
    env = torch.zeros(18, n, dtype=torch.float64)
    env[0:6] = torch.rand(6, n, generator=g, dtype=torch.float64) * 2 - 1
    env[12:18] = torch.rand(6, n, generator=g, dtype=torch.float64) * 2 - 1
    noise = torch.empty(Tn, 3, 5, n).exponential_(1, generator=g)
    mean, std, length = evaluate(tr, n, seed=0, env_init=env.cuda(), noise=noise.cuda())
    eng_eval, buf, _ = tr._eval_ctx[n]
    acts_dev = buf["actions"].permute(0, 2, 1).cpu().long()          # [T,B,N]
    ret_dev = buf["ep_return"].cpu().numpy()
    # oracle, closed loop
    e = env.numpy()
    pos = e[0:6].T.reshape(n, 3, 2).copy(); vel = np.zeros_like(pos); lm = e[12:18].T.reshape(n, 3, 2).copy()
    ids = np.broadcast_to(np.eye(3), (n, 3, 3))
    total = np.zeros(n)
    same = np.ones(n, dtype=bool)
    for t in range(Tn):
        raw = osp.observe_batched(pos, vel, lm)
        obs = torch.from_numpy(np.concatenate([raw, ids], axis=-1)).float()
        with torch.no_grad():
            a, _ = om.race_sample(om.actor_logits(actor, obs), noise[t].permute(2, 0, 1))
        same &= (a == acts_dev[t]).all(dim=1).numpy()
        pos, vel, rew = osp.step_batched(pos, vel, lm, acts_dev[t].numpy())     # follow the device where a tie split them
        total += rew[:, 0]
    assert same.mean() > 0.9, same.mean()
    assert np.abs(ret_dev - total).max() < 1e-9                      # physics + reward along the device's own actions
    assert abs(mean - float(np.mean(ret_dev))) < 1e-12 and abs(std - float(np.std(ret_dev))) < 1e-9 and length == 25.0
    if same.all():
        # (the oracle's own closed loop took exactly these actions: its mean / std are the ones above)
        assert abs(mean - float(np.mean(total))) < 1e-9
    # one context per evaluation size, reused; device-drawn evaluations are a function of the seed
    r1 = evaluate(tr, n, seed=5)
    assert tr._eval_ctx[n][0] is eng_eval and len(tr._eval_ctx) == 1
    assert evaluate(tr, n, seed=5) == r1 and evaluate(tr, n, seed=6) != r1
    assert -60.0 < r1[0] < -5.0 and r1[1] > 0.0


def test_trainer_wide_actor_and_critic(cm):
    """--actor_hidden_dim 64 with the default 64-wide MAPPO critic: 13 638 parameters, more than one clip_adam_kernel
    thread row of 12 288 (the kernel holds 16 384) -- the whole iteration against the oracle, as the tiny-size test."""
    from cleanmarl_b200.mappo import MAPPO, Args
    from cleanmarl_b200 import engine as E
    B = 130
    tr = MAPPO(Args(batch_size=B, seed=6, actor_hidden_dim=64, critic_hidden_dim=64, clip_gradients=0.5), use_graph=False)
    assert tr.engine.n_params == 13638
    actor, critic = om.build_networks(6, actor_hidden=64, critic_hidden=64)
    assert torch.equal(tr.net.flat.cpu(), torch.cat([actor.flat_params(), critic.flat_params()]))
    g = torch.Generator().manual_seed(B)
    env = torch.zeros(18, B, dtype=torch.float64)
    env[0:6] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    env[12:18] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    noise = torch.empty(25, 3, 5, B).exponential_(1, generator=g)
    tr.iteration(env.cuda(), noise.cuda())
    torch.cuda.synchronize()
    batch = tuple(t.cpu() for t in tr.get_batch())
    ret, adv = om.td_lambda_batched(critic, batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    assert (E.heads_to_reference(tr.buf["adv"], 3).cpu() - adv).abs().max() < 1e-5
    aopt, copt = om.make_optimizers(actor, critic)
    st = om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, clip_gradients=0.5)
    final = torch.cat([actor.flat_params(), critic.flat_params()])
    dp = (tr.net.flat.cpu() - final).abs()
    assert (dp < 3e-6).float().mean() > 0.99 and dp.max() < 2e-4
    s = tr.epoch_stats.cpu()
    for ep in range(3):
        assert abs(s[ep, 0].item() - st["actor_loss"][ep]) < 2e-5 * abs(st["actor_loss"][ep]) + 1e-7
        assert abs(s[ep, 1].item() - st["critic_loss"][ep]) < 2e-5 * abs(st["critic_loss"][ep]) + 1e-7


# ----------------------------------------------------------------------------------------- BASELINE configs[1] / [2] size
def _relu_kink_samples(actor, critic, obs, critic_in, tol=2e-6):
    """(b, t) pairs in which some hidden unit's pre-activation lies within float rounding of zero (|pre| < tol * (|x||W| + |b|),
    evaluated in fp64).  relu' is discontinuous there: two correct fp32 evaluations may disagree on the unit's mask, and ONE
    flipped (sample, unit) changes a gradient that is a sum over n samples by ~1/sqrt(n) of its size (1.8e-3 at n = 307 200
    -- measured: profiles/grad_accuracy_r2.md), far above any reassociation tolerance.  At 10^7 pre-activations per
    network a handful of such samples always exists, so the full-size test masks them out (both sides) instead of widening
    the tolerance."""
    def kinks(net, x):
        x = x.double()
        amb = torch.zeros(x.shape[:-1], dtype=torch.bool)
        for lin in list(net.linears)[:-1]:
            W, b = lin.weight.double(), lin.bias.double()
            pre = x @ W.T + b
            amb |= (pre.abs() < tol * (x.abs() @ W.abs().T + b.abs())).any(-1)
            x = torch.relu(pre)
        return amb
    with torch.no_grad():
        a = kinks(actor, obs).any(-1)
        c = kinks(critic, critic_in)
    return a | (c.any(-1) if c.dim() == 3 else c)


@pytest.mark.parametrize("ippo", [False, True], ids=["mappo", "ippo"])
def test_whole_update_at_baseline_size(cm, ippo):
    """The whole update at BASELINE.json's own size (num_envs = 4096: 800 critic / 2 400 actor tiles over the persistent
    148 / 296-CTA grids, several rounds per CTA) on the SURVEY 8(d) synthetic batch against the oracle
    (``td_lambda_batched`` + ``ppo_update(flat=True)``): advantages <= 1e-5, first-epoch gradients per tensor
    <= 2e-5 of the tensor's max (samples sitting on a relu kink masked out, see ``_relu_kink_samples``), per-epoch losses
    <= 2e-5 relative, gradient norms <= 2e-4 (later epochs may flip a kink), parameters <= 1e-6 after the 3 epochs."""
    from cleanmarl_b200 import engine as E
    B, Tn = 4096, 25
    actor, critic = (om.build_networks(1, state_dim=21, critic_hidden=32) if ippo else om.build_networks(1))
    batch = list(om.synthetic_batch(B, seed=1, actor=actor))
    eng = make_engine(cm, B, tc=True, critic_on_obs=ippo, critic_hidden=32 if ippo else 64)
    dev = eng.device
    params = flat_params(actor, critic, dev)
    d = E.to_device_layout(tuple(batch), dev, with_obs=False)
    values = eng.empty(Tn, eng.n_heads, B)
    ret, adv = torch.empty_like(values), torch.empty_like(values)
    eng.critic_values(params[eng.n_actor:], values, state=d["state"])
    eng.td_lambda(values, d["reward"], ret, adv, 0.99, 0.95)
    ret_o, adv_o = om.td_lambda_batched(critic, batch[0] if ippo else batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    assert (E.heads_to_reference(adv, 3).cpu() - adv_o).abs().max() < 1e-5
    assert (E.heads_to_reference(ret, 3).cpu() - ret_o).abs().max() < 1e-5
    kink = _relu_kink_samples(actor, critic, batch[0], batch[0] if ippo else batch[4])
    assert 0 < int(kink.sum()) < 0.01 * B * Tn, int(kink.sum())
    batch[7] = batch[7] & ~kink
    d = E.to_device_layout(tuple(batch), dev, with_obs=False)
    aopt, copt = om.make_optimizers(actor, critic)
    st = om.ppo_update(actor, critic, aopt, copt, tuple(batch), adv_o, ret_o, epochs=3, clip=0.2, ent_coef=0.001,
                       critic_on_obs=ippo, flat=True, record_grads=True)
    adv_d, ret_d = E.heads_to_device(adv_o, eng.n_heads, dev), E.heads_to_device(ret_o, eng.n_heads, dev)
    m, v = torch.zeros_like(params), torch.zeros_like(params)
    grads, stats = eng.empty(eng.n_params + 8), eng.empty(8)
    n = float(batch[7].sum())
    for ep in range(3):
        eng.ppo_epoch_grads(params, grads, state=d["state"], actions=d["actions"], logp_old=d["logp"], adv=adv_d,
                            returns=ret_d, mask=d["mask"], clip=0.2, ent_coef=0.001)
        if ep == 0:
            gcpu = grads.cpu()
            assert gcpu[eng.n_params + 5].item() == n
            ga, gc = st["grads"][0]
            ref = torch.cat([ga, gc])
            off = 0
            for net in (actor, critic):
                for p in net.parameters():
                    k = p.numel()
                    a, b = gcpu[off:off + k] / n, ref[off:off + k]
                    scale = max(b.abs().max().item(), 1e-6)
                    assert (a - b).abs().max().item() <= 2e-5 * scale + 1e-9, (tuple(p.shape), (a - b).abs().max().item(), scale)
                    off += k
        eng.clip_adam_step(params, grads, m, v, step=ep + 1, stats_out=stats)
        s = stats.cpu().numpy()
        ref = [st["actor_loss"][ep], st["critic_loss"][ep], st["entropy"][ep], st["kl"][ep], st["clipfrac"][ep],
               st["actor_grad_norm"][ep], st["critic_grad_norm"][ep]]
        for k in (0, 1, 2):
            assert abs(s[k] - ref[k]) <= 2e-5 * abs(ref[k]) + 1e-7, (ep, k, s[k], ref[k])
        for k in (5, 6):
            assert abs(s[k] - ref[k]) <= (2e-5 if ep == 0 else 2e-4) * abs(ref[k]) + 1e-7, (ep, k, s[k], ref[k])
        assert abs(s[3] - ref[3]) < 1e-6 + 1e-3 * abs(ref[3]) and abs(s[4] - ref[4]) < 1e-6
    final = torch.cat([actor.flat_params(), critic.flat_params()])
    dp = (params.cpu() - final).abs()
    # (a gradient within ~1e-7 of Adam's eps = 1e-8 turns a 1e-9 reassociation difference into a visible fraction of lr)
    assert (dp < 1e-6).float().mean() > 0.995 and dp.max().item() < 2e-5, (dp.max().item(), (dp < 1e-6).float().mean().item())


# ----------------------------------------------------------------------------------------- beyond-reference options (north_star)
@pytest.mark.parametrize("tc", TC, ids=["ffma", "tc"])
@pytest.mark.parametrize("ippo", [False, True], ids=["mappo", "ippo"])
def test_value_clip_and_minibatch_gradients(cm, ippo, tc):
    """cmarl_ppo_epoch_grads_ex: the clipped value loss and the env-block (minibatch) range against the oracle's
    autograd on the same block (ragged masks, random avail), blocks that start / end inside a 128-sample tile; and the
    plain entry == the _ex entry with the options off, bit for bit."""
    from cleanmarl_b200 import engine as E
    B, Tn = 700, 25
    actor, critic = om.build_networks(5, state_dim=21 if ippo else 54, critic_hidden=32 if ippo else 64)
    batch = list(om.synthetic_batch(B, seed=3, actor=actor))
    gen = torch.Generator().manual_seed(3)
    batch[7] = ragged_mask(B, Tn, gen) & ~_relu_kink_samples(actor, critic, batch[0], batch[0] if ippo else batch[4])
    V = 3 if ippo else 1
    adv = (torch.randn(B, Tn, V, generator=gen) * 3).expand(B, Tn, 3).contiguous()
    ret = (torch.randn(B, Tn, V, generator=gen) * 5).expand(B, Tn, 3).contiguous()
    v_old = (ret[..., :V] + torch.randn(B, Tn, V, generator=gen) * 0.6).expand(B, Tn, 3).contiguous()   # |V - V_old| both sides of 0.2
    eng = make_engine(cm, B, tc=tc, critic_on_obs=ippo, critic_hidden=32 if ippo else 64)
    dev = eng.device
    d = E.to_device_layout(tuple(batch), dev, with_obs=False)
    params = flat_params(actor, critic, dev)
    kw = dict(state=d["state"], actions=d["actions"], logp_old=d["logp"], adv=E.heads_to_device(adv, eng.n_heads, dev),
              returns=E.heads_to_device(ret, eng.n_heads, dev), mask=d["mask"], clip=0.2, ent_coef=0.001)
    vold_d = E.heads_to_device(v_old, eng.n_heads, dev)
    g0, g1 = eng.empty(eng.n_params + 8), eng.empty(eng.n_params + 8)
    eng.ppo_epoch_grads(params, g0, **kw)
    eng.ppo_epoch_grads(params, g1, **kw, value_clip=-1.0, values_old=vold_d, env_begin=0, env_count=B)
    assert torch.equal(g0, g1)
    obs, actions, logp, reward, states, avail, done, mask = batch
    for lo, hi, vc in ((0, B, 0.2), (0, 233, 0.2), (233, 466, -1.0), (466, 700, 0.2), (130, 131, 0.2)):
        sl = slice(lo, hi)
        eng.ppo_epoch_grads(params, g1, **kw, value_clip=vc, values_old=vold_d, env_begin=lo, env_count=hi - lo)
        actor.zero_grad(); critic.zero_grad()
        ext = {} if vc <= 0 else dict(value_clip=vc, values_old=v_old[sl])
        out = om.ppo_epoch_flat(actor, critic, obs[sl], actions[sl], logp[sl], (obs if ippo else states)[sl], avail[sl],
                                mask[sl], adv[sl], ret[sl], 0.2, 0.001, **ext)
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
        assert abs(g[eng.n_params + 1].item() / n - out.critic_loss.item()) < 1e-5 * abs(out.critic_loss.item())
        assert abs(g[eng.n_params + 0].item() / n - out.actor_loss.item()) < 1e-5 * abs(out.actor_loss.item()) + 1e-7


def test_trainer_with_value_clip_and_minibatches(cm):
    """--value_clip 0.2 --num_minibatches 4 through the trainer (graph replay == eager, 12 optimizer steps per iteration)
    against the oracle's update with the same options."""
    from cleanmarl_b200.mappo import MAPPO, Args
    from cleanmarl_b200 import engine as E
    B = 520
    g = torch.Generator().manual_seed(B)
    env = torch.zeros(18, B, dtype=torch.float64)
    env[0:6] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    env[12:18] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    noise = torch.empty(25, 3, 5, B).exponential_(1, generator=g)
    tr = MAPPO(Args(batch_size=B, seed=6, value_clip=0.2, num_minibatches=4), use_graph=False)
    tr.iteration(env.cuda(), noise.cuda())
    torch.cuda.synchronize()
    assert tr.training_step == 12
    actor, critic = om.build_networks(6)
    batch = tuple(t.cpu() for t in tr.get_batch())
    ret, adv = om.td_lambda_batched(critic, batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    with torch.no_grad():
        v_old = critic(batch[4]).expand(B, 25, 3).contiguous()
    aopt, copt = om.make_optimizers(actor, critic)
    st = om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, flat=True,
                       value_clip=0.2, values_old=v_old, num_minibatches=4)
    dp = (tr.net.flat.cpu() - torch.cat([actor.flat_params(), critic.flat_params()])).abs()
    assert (dp < 3e-6).float().mean() > 0.99 and dp.max() < 2e-4
    s = tr.epoch_stats.cpu()
    for ep in range(3):
        assert abs(s[ep, 1].item() - np.mean(st["critic_loss"][4 * ep:4 * ep + 4])) < 2e-5 * abs(s[ep, 1].item())
    # graph replay of the 12-step iteration == eager
    outs = []
    for graph in (False, True):
        t2 = MAPPO(Args(batch_size=B, seed=6, value_clip=0.2, num_minibatches=4), use_graph=graph)
        for _ in range(3):
            t2.iteration()
        torch.cuda.synchronize()
        outs.append((t2.net.flat.clone(), t2.epoch_stats.clone(), t2.training_step))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2] == 36


def test_fused_reduce_adam_equals_separate_launches(cm):
    """cmarl_reduce_clip_adam_step (partial reduction + Adam in one launch) == cmarl_ppo_epoch_grads followed by
    cmarl_clip_adam_step, bit for bit: reduced gradient sums, parameters, moments, statistics, device step counter --
    MAPPO and IPPO, with gradient clipping, over three epochs; and the trainer with it on == off."""
    from cleanmarl_b200 import engine as E
    from cleanmarl_b200.mappo import MAPPO, Args
    for ippo in (False, True):
        B = 1100
        actor, critic = om.build_networks(4, state_dim=21 if ippo else 54, critic_hidden=32 if ippo else 64)
        batch = om.synthetic_batch(B, seed=4, actor=actor)
        eng = make_engine(cm, B, tc=True, critic_on_obs=ippo, critic_hidden=32 if ippo else 64)   # (entry points called directly)
        dev = eng.device
        d = E.to_device_layout(batch, dev, with_obs=False)
        gen = torch.Generator().manual_seed(1)
        V = eng.n_heads
        adv = (torch.randn(25, V, B, generator=gen) * 3).to(dev)
        ret = (torch.randn(25, V, B, generator=gen) * 5).to(dev)
        kw = dict(state=d["state"], actions=d["actions"], logp_old=d["logp"], adv=adv, returns=ret, clip=0.2, ent_coef=0.001)
        outs = []
        for fused in (False, True):
            params = flat_params(actor, critic, dev)
            m, v = torch.zeros_like(params), torch.zeros_like(params)
            grads, stats = eng.empty(eng.n_params + 8), eng.empty(3, 8)
            cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            keep = []
            for ep in range(3):
                if fused:
                    eng.ppo_epoch_grads(params, None, **kw)
                    eng.reduce_clip_adam_step(params, grads, m, v, step_dev=cnt, max_norm=0.5, stats_out=stats[ep])
                else:
                    eng.ppo_epoch_grads(params, grads, **kw)
                    eng.clip_adam_step(params, grads, m, v, step_dev=cnt, max_norm=0.5, stats_out=stats[ep])
                keep.append(grads.clone())
            outs.append((params, m, v, stats, cnt, *keep))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
    import os
    res = []
    for flag in ("0", "1"):
        os.environ["CMARL_FUSED_UPDATE"] = flag
        try:
            tr = MAPPO(Args(batch_size=700, seed=8, clip_gradients=0.5))
            assert tr.engine.fused_update == (flag == "1")
            for _ in range(4):
                tr.iteration()
            torch.cuda.synchronize()
            res.append((tr.net.flat.clone(), tr.epoch_stats.clone(), tr.launches_per_iteration))
        finally:
            os.environ.pop("CMARL_FUSED_UPDATE", None)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert res[0][2] == 17 and res[1][2] == 14                      # launches per iteration: three fewer
