"""GPU parity tests of the recurrent-actor path (BASELINE config 4, cleanmarl/mappo_lstm_multienvs.py) against the
CPU oracle (oracle/mappo_lstm.py, pinned bit-exact to the unmodified reference) and the golden reference runs,
through the C ABI.  Run on the B200 box: ``pytest -m gpu``.  Tolerances are stated per test.
"""
import numpy as np
import pytest
import torch

from oracle import mappo as om
from oracle import mappo_lstm as ol
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


def make_engine(cm, B, **kw):
    return cm.Engine(cm.Shapes(n_envs=B, actor_recurrent=True, **kw), device=0)


def ragged_mask(B, Tn, gen, frac=0.4):
    lengths = torch.where(torch.rand(B, generator=gen) < frac,
                          torch.randint(1, Tn + 1, (B,), generator=gen), torch.full((B,), Tn))
    return (torch.arange(Tn)[None, :] < lengths[:, None])


# ----------------------------------------------------------------------------------------- K2 (recurrent)
def test_recurrent_act_vs_reference_fixture(cm, golden):
    """Five chained Actor.act(x, h, avail) calls of the real reference (g3_recurrent): hidden state and logits within
    2e-6, sampled action bit-exact wherever the race is decided by more than 1e-5, log-prob within 2e-6."""
    g = golden("g3_recurrent")
    M = g["x"].shape[1]                      # 768 rows = 256 envs x 3 "agents" for the device layout
    B = M // 3
    eng = make_engine(cm, B)
    dev = eng.device
    params = T(g["params"]).to(dev)
    h = None
    for t in range(5):
        # fixture rows are (env, agent)-major like the reference's reshape(B * N, -1)
        x = T(g["x"][t]).reshape(B, 3, 21).permute(1, 2, 0).contiguous().to(dev)            # [N][O][B]
        av = T(g["avail"][t]).reshape(B, 3, 5).permute(1, 2, 0).contiguous().to(torch.uint8).to(dev)
        q = T(g["q"][t]).reshape(B, 3, 5).permute(1, 2, 0).contiguous().to(dev)
        actions = eng.empty(3, B, dtype=torch.int32)
        logp, logits, h_out = eng.empty(3, B), eng.empty(3, 5, B), eng.empty(3, 32, B)
        eng.actor_act_recurrent(params, x, q, actions, logp, h_out, h_in=h, avail=av, logits=logits)
        h_ref = T(g[f"h{t}"]).reshape(B, 3, 32)
        z_ref = T(g[f"logits{t}"]).reshape(B, 3, 5)
        assert (h_out.permute(2, 0, 1).cpu() - h_ref).abs().max() < 2e-6
        zd = logits.permute(2, 0, 1).cpu()
        live = z_ref > -1e8
        assert (zd[live] - z_ref[live]).abs().max() < 2e-6 and (zd[~live] == -1e9).all()
        # race margin of the reference: best vs second best of probs / q
        p = torch.softmax(z_ref, -1) / T(g["q"][t]).reshape(B, 3, 5)
        top2 = p.topk(2, dim=-1).values
        decided = (top2[..., 0] - top2[..., 1]) > 1e-5 * top2[..., 0]
        a_ref = T(g[f"actions{t}"]).reshape(B, 3)
        a_dev = actions.t().cpu().long()
        assert decided.float().mean() > 0.999
        assert torch.equal(a_dev[decided], a_ref[decided])
        same = a_dev == a_ref
        assert (logp.t().cpu()[same] - T(g[f"logp{t}"]).reshape(B, 3)[same]).abs().max() < 2e-6
        h = h_out


# ----------------------------------------------------------------------------------------- K1+K2+K3 (recurrent rollout)
def test_recurrent_rollout_matches_oracle(cm):
    """Device rollout with the GRU actor vs the oracle stepping (oracle.spread + oracle.mappo_lstm.rollout_act) on
    the same start states and race noise: actions equal for every env whose races were all decided by > 1e-5 (the hidden
    state makes an env's later steps depend on its earlier actions), observations within 1e-6, log-probs within 5e-6."""
    B, Tn = 512, 25
    eng = make_engine(cm, B)
    dev = eng.device
    actor, critic = ol.build_networks(2)
    pa = actor.flat_params().to(dev)
    env = eng.empty(18, B, dtype=torch.float64)
    eng.env_reset(env, seed=5, episode=0)
    env0 = env.cpu().numpy().copy()
    q = om.draw_race_noise((Tn, 3, 5, B), generator=torch.Generator().manual_seed(3))
    buf = eng.alloc_rollout()
    eng.rollout(pa, env, buf["state"], buf["actions"], buf["logp"], buf["reward"], noise=q.to(dev),
                ep_return=buf["ep_return"])
    torch.cuda.synchronize()
    # oracle: step the CPU env with the oracle's own sampled actions
    pos = env0[0:6].T.reshape(B, 3, 2).copy(); vel = env0[6:12].T.reshape(B, 3, 2).copy()
    lm = env0[12:18].T.reshape(B, 3, 2).copy()
    ids = np.broadcast_to(np.eye(3), (B, 3, 3))
    avail = torch.ones(B, 3, 5, dtype=torch.bool)
    h = None
    ok = torch.ones(B, dtype=torch.bool)
    acts_d = buf["actions"].cpu().long()          # [T][N][B]
    logp_d = buf["logp"].cpu()
    state_d = buf["state"].cpu()
    rew_d = buf["reward"].cpu()
    for t in range(Tn):
        raw = osp.observe_batched(pos, vel, lm)
        o = torch.from_numpy(np.concatenate([raw, ids], -1)).float()
        a, lp, h, z = ol.rollout_act(actor, o, h, avail, q[t].permute(2, 0, 1))
        p = torch.softmax(z, -1) / q[t].permute(2, 0, 1)
        top2 = p.topk(2, dim=-1).values
        ok &= ((top2[..., 0] - top2[..., 1]) > 1e-5 * top2[..., 0]).all(dim=-1)
        sd = state_d[t].reshape(3, 18, B).permute(2, 0, 1)
        assert (sd[ok] - torch.from_numpy(raw)[ok]).abs().max() < 1e-6, t
        assert torch.equal(acts_d[t].t()[ok], a[ok]), t
        assert (logp_d[t].t()[ok] - lp[ok]).abs().max() < 5e-6, t
        pos, vel, rew = osp.step_batched(pos, vel, lm, a.numpy())
        assert (rew_d[t][ok] - torch.from_numpy(rew[:, 0]).float()[ok]).abs().max() < 1e-5, t
    assert ok.float().mean() > 0.97


# ----------------------------------------------------------------------------------------- K7a / K7b / K8 (TBPTT)
def _device_update(eng, E, actor, critic, batch, adv, ret, *, epochs, tbptt, clip_gradients, lr_a, lr_c, use_mask=True,
                   use_avail=True, use_obs=False, weight_decay=0.0, stash=True):
    from cleanmarl_b200.mappo import tbptt_chunks
    dev = eng.device
    eng.set_weight_decay(weight_decay, weight_decay)
    d = E.to_device_layout(batch, dev)
    na = eng.n_actor
    flat = torch.cat([actor.flat_params(), critic.flat_params()]).to(dev).contiguous()
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    ga, gc = eng.empty(na + 8), eng.empty(eng.n_critic + 8)
    h_seq = eng.alloc_h_seq()
    gate_stash = eng.alloc_gate_stash() if stash else None
    step_a = torch.zeros(1, dtype=torch.int32, device=dev)
    step_c = torch.zeros(1, dtype=torch.int32, device=dev)
    adv_d = E.heads_to_device(adv, eng.n_heads, dev)
    ret_d = E.heads_to_device(ret, eng.n_heads, dev)
    chunks = tbptt_chunks(25, tbptt)
    out = {"chunk_grads": [], "critic_grads": [], "chunk_stats": [], "critic_stats": []}
    for ep in range(epochs):
        cg, cs = [], []
        for (t0, t1) in chunks:
            eng.tbptt_chunk_grads(flat[:na], ga, h_seq, t0, t1, state=d["state"], obs=d["obs"] if use_obs else None,
                                  actions=d["actions"], logp_old=d["logp"], adv=adv_d,
                                  mask=d["mask"] if use_mask else None, avail=d["avail"] if use_avail else None,
                                  clip=0.2, ent_coef=0.001, stash=gate_stash)
            st = eng.empty(8)
            cg.append(ga.cpu().clone())
            eng.adam_step_net(0, flat[:na], ga, m[:na], v[:na], step_dev=step_a, lr=lr_a, max_norm=clip_gradients,
                              extra_div=t1 - t0, stats_out=st)
            cs.append(st.cpu())
        eng.critic_epoch_grads(flat[na:], gc, state=d["state"], returns=ret_d, mask=d["mask"] if use_mask else None)
        st = eng.empty(8)
        out["critic_grads"].append(gc.cpu().clone())
        eng.adam_step_net(1, flat[na:], gc, m[na:], v[na:], step_dev=step_c, lr=lr_c, max_norm=clip_gradients, stats_out=st)
        out["chunk_grads"].append(cg); out["chunk_stats"].append(torch.stack(cs)); out["critic_stats"].append(st.cpu())
    torch.cuda.synchronize()
    out["params"] = flat.cpu()
    out["h_seq"] = h_seq.cpu()
    return out


@pytest.mark.parametrize("name", ["g8_mappo_lstm", "g8_mappo_lstm_flags", "g8_ippo_lstm"])
def test_tbptt_update_vs_reference_run(cm, golden, name):
    """The whole recurrent update on the batch of a real reference iteration (B = 6): per-epoch losses / statistics
    within 2e-5 relative, mean chunk gradient norm within 1e-4 relative, parameters after the reference's 3 epochs
    (9-12 actor Adam steps, 3 critic steps) within 3e-6."""
    from cleanmarl_b200 import engine as E
    g = golden(name)
    ippo = name == "g8_ippo_lstm"        # ippo_lstm_multienvs.py: critic on obs (V = 3 heads), AdamW (wd 0.01), tbptt 5
    B = int(g["B"])
    if ippo:
        actor, critic = ol.build_networks(int(g["seed"]), state_dim=21, critic_hidden=32)
        eng = make_engine(cm, B, critic_on_obs=True, critic_hidden=32)
    else:
        actor, critic = ol.build_networks(int(g["seed"]))
        eng = make_engine(cm, B)
    batch = tuple(T(g[k]) for k in ("obs", "actions", "log_probs", "reward", "states", "avail", "done", "mask"))
    adv, ret = T(g["advantages"]), T(g["return_lambda"])
    out = _device_update(eng, E, actor, critic, batch, adv, ret, epochs=int(g["epochs"]), tbptt=int(g["tbptt"]),
                         clip_gradients=float(g["clip_gradients"]), lr_a=float(g["lr_actor"]), lr_c=float(g["lr_critic"]),
                         weight_decay=0.01 if ippo else 0.0)
    na = eng.n_actor
    for ep in range(int(g["epochs"])):
        cs, ks = out["chunk_stats"][ep], out["critic_stats"][ep]
        n = cs[:, 6].sum()
        assert n == B * 25
        rel = lambda a, b: abs(float(a) - float(b)) / (abs(float(b)) + 1e-12)
        assert rel(cs[:, 0].sum() / n, g["actor_losses"][ep]) < 2e-5 or abs(float(cs[:, 0].sum() / n) - float(g["actor_losses"][ep])) < 2e-7
        assert rel(ks[1] / ks[6], g["critic_losses"][ep]) < 2e-5
        assert rel(cs[:, 2].sum() / n, g["entropies"][ep]) < 2e-5
        assert abs(float(cs[:, 3].sum() / n) - float(g["kls"][ep])) < 1e-6
        assert abs(float(cs[:, 4].sum() / n) - float(g["clipfracs"][ep])) < 1e-6
        assert rel(cs[:, 5].mean(), g["actor_grad_norms"][ep]) < 1e-4
        assert rel(ks[5], g["critic_grad_norms"][ep]) < 1e-4
    assert (out["params"][:na] - T(g["actor_final"])).abs().max() < 3e-6
    assert (out["params"][na:] - T(g["critic_final"])).abs().max() < 3e-6


@pytest.mark.parametrize("B,use_obs,tbptt", [(200, False, 10), (77, True, 7), (1024, False, 25), (130, False, 12)])   # 12: the last chunk is ONE step
def test_tbptt_chunk_gradients_vs_oracle(cm, B, use_obs, tbptt):
    """Per-chunk actor gradients (ragged masks, masked-out actions, ragged last tile, explicit obs / obs rebuilt from
    state) against autograd through the oracle's truncated-BPTT loop: each chunk's flat gradient within 2e-5 of its own
    max-norm; chunk-end hidden states within 2e-6; parameters after one epoch within 2e-6."""
    from cleanmarl_b200 import engine as E
    from cleanmarl_b200.mappo import tbptt_chunks
    gen = torch.Generator().manual_seed(B)
    actor, critic = ol.build_networks(B)
    batch = list(om.synthetic_batch(B, seed=B + 1))
    batch[7] = ragged_mask(B, 25, gen)
    avail = torch.ones(B, 25, 3, 5, dtype=torch.bool)
    avail[torch.rand(B, 25, 3, 5, generator=gen) < 0.1] = False
    avail[..., 0] = True
    batch[5] = avail
    # actions must be available ones; old log-probs near the current policy so both clip sides are exercised
    probs = avail.float() / avail.float().sum(-1, keepdim=True)
    batch[1] = torch.multinomial(probs.reshape(-1, 5), 1, generator=gen).reshape(B, 25, 3)
    batch[2] = ol.synthetic_old_logp(actor, batch, seed=3)
    batch = tuple(batch)
    adv = torch.randn(B, 25, 1, generator=gen).expand(B, 25, 3).contiguous()
    ret = torch.randn(B, 25, 1, generator=gen).expand(B, 25, 3).contiguous()
    eng = make_engine(cm, B)
    out = _device_update(eng, E, actor, critic, batch, adv, ret, epochs=1, tbptt=tbptt, clip_gradients=-1.0,
                         lr_a=8e-4, lr_c=8e-4, use_obs=use_obs)
    # `out` ran the tcgen05 kernels (tc_gru.cu: a gate stash is how its two kernels meet).  The fp32 FFMA kernel (gru.cu)
    # with and without a stash (recompute variant) gives the same gradients and parameters bit for bit, and the tcgen05
    # pair agrees with it to the tolerance both are held to against the oracle below.
    import os
    os.environ["CMARL_TBPTT"] = "ffma"
    try:
        out1 = _device_update(eng, E, actor, critic, batch, adv, ret, epochs=1, tbptt=tbptt, clip_gradients=-1.0,
                              lr_a=8e-4, lr_c=8e-4, use_obs=use_obs)
    finally:
        os.environ.pop("CMARL_TBPTT", None)
    out2 = _device_update(eng, E, actor, critic, batch, adv, ret, epochs=1, tbptt=tbptt, clip_gradients=-1.0,
                          lr_a=8e-4, lr_c=8e-4, use_obs=use_obs, stash=False)
    assert torch.equal(out1["params"], out2["params"])
    assert all(torch.equal(x, y) for x, y in zip(out1["chunk_grads"][0], out2["chunk_grads"][0]))
    assert (out["h_seq"] - out1["h_seq"]).abs().max() < 2e-6
    for x, y in zip(out["chunk_grads"][0], out1["chunk_grads"][0]):
        assert (x[:eng.n_actor] - y[:eng.n_actor]).abs().max() <= 2e-5 * y[:eng.n_actor].abs().max()
    aopt, copt = om.make_optimizers(actor, critic)
    st = ol.ppo_update_tbptt(actor, critic, aopt, copt, batch, adv, ret, epochs=1, clip=0.2, ent_coef=0.001,
                             tbptt=tbptt, record_grads=True)
    chunk_grads, critic_grad = st["grads"][0]
    na = eng.n_actor
    chunks = tbptt_chunks(25, tbptt)
    mask = batch[7]
    for ci, (t0, t1) in enumerate(chunks):
        n_valid = float(mask[:, t0:t1].sum())
        dev_g = out["chunk_grads"][0][ci]
        assert dev_g[na + 5] == n_valid
        gd = dev_g[:na] / (n_valid * (t1 - t0))
        gr = chunk_grads[ci]
        assert (gd - gr).abs().max() <= 2e-5 * gr.abs().max(), (ci, float((gd - gr).abs().max()), float(gr.abs().max()))
    gc = out["critic_grads"][0]
    assert (gc[:eng.n_critic] / gc[eng.n_critic + 5] - critic_grad).abs().max() <= 2e-5 * critic_grad.abs().max()
    # Adam's first step is lr * g / (|g| + eps): where |g| is not far above eps = 1e-8 (dead units), a 1e-9 difference
    # in g moves the parameter by a visible fraction of lr; everywhere else the parameters agree to 2e-6
    dp_a = (out["params"][:na] - actor.flat_params()).abs()
    dp_c = (out["params"][na:] - critic.flat_params()).abs()
    solid_a = chunk_grads[-1].abs() > 1e-5 * chunk_grads[-1].abs().max()
    solid_c = critic_grad.abs() > 1e-5 * critic_grad.abs().max()
    assert dp_a[solid_a].max() < 2e-6 and dp_c[solid_c].max() < 2e-6
    assert dp_a.max() < 8e-4 * len(chunks) and dp_c.max() < 8e-4


def test_trainer_recurrent_iteration_runs_and_matches_oracle(cm):
    """MAPPO(ArgsRecurrent).iteration() end to end on the device (rollout with the GRU actor, critic, TD(lambda),
    truncated-BPTT epochs) == the oracle's update on the batch the device collected: > 99.5 % of the parameters within
    3e-6, all within 1e-4; logged scalars within 2e-5 relative."""
    from cleanmarl_b200.mappo import MAPPO, ArgsRecurrent
    B = 256
    tr = MAPPO(ArgsRecurrent(batch_size=B, seed=9))
    assert tr.engine.n_actor == 7205
    actor, critic = ol.build_networks(9)
    assert torch.equal(torch.cat([actor.flat_params(), critic.flat_params()]), tr.net.flat.cpu())
    tr.iteration()
    torch.cuda.synchronize()
    batch = tuple(t.cpu() for t in tr.get_batch())
    ret, adv = om.td_lambda_batched(critic, batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    from cleanmarl_b200 import engine as E
    assert (E.heads_to_reference(tr.buf["adv"], 3).cpu() - adv).abs().max() < 1e-5
    aopt, copt = om.make_optimizers(actor, critic)
    st = ol.ppo_update_tbptt(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, tbptt=10)
    final = torch.cat([actor.flat_params(), critic.flat_params()])
    # 12 Adam steps: parameters whose gradient sits near Adam's eps amplify 1e-9 gradient differences (see above)
    dp = (tr.net.flat.cpu() - final).abs()
    assert (dp < 3e-6).float().mean() > 0.995 and dp.max() < 1e-4
    sc = tr.train_scalars()
    assert abs(sc["actor_loss"] - np.mean(st["actor_loss"])) < 2e-5 * abs(np.mean(st["actor_loss"])) + 1e-6
    assert abs(sc["critic_loss"] - np.mean(st["critic_loss"])) < 2e-5 * abs(np.mean(st["critic_loss"]))
    assert abs(sc["actor_gradients"] - np.mean(st["actor_grad_norm"])) < 1e-4 * np.mean(st["actor_grad_norm"])
    assert tr.step == B * 25 and tr.training_step == 3


def test_tcgen05_chunk_kernels_at_baseline_size(cm):
    """BASELINE configs[3] size (8 192 envs: 192 tiles of 128 envs over 296 forward / 148 backward CTAs, i.e. CTAs that
    walk several tiles and add a second tile's sums into their partial row), one epoch of three chunks:
    * forward: hidden states and gate stash of tc_gru_fwd_kernel within 2e-6 of the fp32 FFMA kernel's (csrc/gru.cu, itself
      held to the oracle at the sizes the oracle finishes quickly);
    * backward: tc_gru_bwd_kernel against the FFMA backward pass ON THE SAME forward results (CMARL_TBPTT=tcfwd), so both
      see identical relu' masks: every chunk's gradient sums within 2e-5 of the chunk gradient's max (per tensor: 2e-5 of
      the tensor's max), statistics within 1e-6 relative."""
    import os
    from cleanmarl_b200 import engine as E
    from cleanmarl_b200.mappo import tbptt_chunks
    B = 8192
    gen = torch.Generator().manual_seed(5)
    actor, critic = ol.build_networks(5)
    batch = list(om.synthetic_batch(B, seed=6))
    batch[7] = ragged_mask(B, 25, gen)
    batch[2] = ol.synthetic_old_logp(actor, batch, seed=3)
    adv = torch.randn(B, 25, 1, generator=gen).expand(B, 25, 3).contiguous()
    eng = make_engine(cm, B)
    dev = eng.device
    d = E.to_device_layout(tuple(batch), dev)
    adv_d = E.heads_to_device(adv, eng.n_heads, dev)
    na = eng.n_actor
    flat = torch.cat([actor.flat_params(), critic.flat_params()]).to(dev).contiguous()
    chunks = tbptt_chunks(25, 10)
    res = {}
    for mode in ("ffma", "tcfwd", "tc"):
        os.environ["CMARL_TBPTT"] = mode
        try:
            h_seq, stash, ga = eng.alloc_h_seq(), eng.alloc_gate_stash(), eng.empty(na + 8)
            stash.zero_()
            grads = []
            for (t0, t1) in chunks:
                eng.tbptt_chunk_grads(flat[:na], ga, h_seq, t0, t1, state=d["state"], actions=d["actions"],
                                      logp_old=d["logp"], adv=adv_d, mask=d["mask"], clip=0.2, ent_coef=0.001, stash=stash)
                grads.append(ga.clone())
            torch.cuda.synchronize()
            res[mode] = (h_seq.clone(), stash.clone(), grads)
        finally:
            os.environ.pop("CMARL_TBPTT", None)
    assert (res["tcfwd"][0] - res["ffma"][0]).abs().max() < 2e-6
    assert (res["tcfwd"][1] - res["ffma"][1]).abs().max() < 2e-6
    assert torch.equal(res["tc"][0], res["tcfwd"][0]) and torch.equal(res["tc"][1], res["tcfwd"][1])
    sizes = [32 * 21, 32, 3072, 3072, 96, 96, 160, 5]
    for g_tc, g_ref in zip(res["tc"][2], res["tcfwd"][2]):
        assert (g_tc[:na] - g_ref[:na]).abs().max() <= 2e-5 * g_ref[:na].abs().max()
        o = 0
        for n in sizes:
            assert (g_tc[o:o + n] - g_ref[o:o + n]).abs().max() <= 2e-5 * g_ref[o:o + n].abs().max(), (o, n)
            o += n
        assert (g_tc[na:] - g_ref[na:]).abs().max() <= 1e-6 * g_ref[na:].abs().max()
