"""Pin the oracle (oracle/mappo.py) against fixtures produced by the unmodified reference.

CPU only.  Fixtures: tests/golden/*.npz (see tests/golden/gen_golden.py).
"""
import numpy as np
import pytest
import torch

from oracle import mappo as om


def T(x):
    return torch.from_numpy(np.asarray(x))


def test_g1_param_init_bit_exact(golden):
    g = golden("g1_params")
    for seed in (1, 7):
        a, c = om.build_networks(seed)
        assert np.array_equal(a.flat_params().numpy(), g[f"mappo_actor_s{seed}"])
        assert np.array_equal(c.flat_params().numpy(), g[f"mappo_critic_s{seed}"])
        a, c = om.build_networks(seed, state_dim=21, critic_hidden=32)
        assert np.array_equal(a.flat_params().numpy(), g[f"ippo_actor_s{seed}"])
        assert np.array_equal(c.flat_params().numpy(), g[f"ippo_critic_s{seed}"])
    a, c = om.build_networks(3, actor_hidden=64, actor_layers=2, critic_hidden=128, critic_layers=2)
    assert np.array_equal(a.flat_params().numpy(), g["mappo_actor_wide"])
    assert np.array_equal(c.flat_params().numpy(), g["mappo_critic_wide"])
    assert a.flat_params().numel() == 64 * 21 + 64 + 2 * (64 * 64 + 64) + 5 * 64 + 5


def test_g3_categorical_sample_is_exponential_race(golden):
    g = golden("g3_sample")
    actor = om.MLP(21, 32, 1, 5)
    actor.load_flat(T(g["params"]))
    with torch.no_grad():
        logits = om.actor_logits(actor, T(g["x"]), T(g["avail"]))
        assert np.array_equal(logits.numpy(), g["logits"])
        actions, logp = om.race_sample(logits, T(g["q"]))
    assert np.array_equal(actions.numpy(), g["actions"])          # bit-exact indices
    assert np.array_equal(logp.numpy(), g["logp"])
    # masked actions are never drawn
    assert not ((~g["avail"]) & (np.eye(5, dtype=bool)[g["actions"]])).any()


def test_log_domain_race_is_the_same_decision(golden):
    """rollout_tc_kernel decides the exponential race as argmax_a (z_a - log q_a) instead of the reference's
    argmax_a softmax(z)_a / q_a (MME:172-176 -> torch.multinomial): the same decision on the reference's own G3 draws
    (the unmodified Actor.act under the reference's seed) and on 10^6 random races, up to races closer than a few ulp;
    log_prob = z_a - logsumexp(z) equals Categorical.log_prob to fp32 round-off."""
    g = golden("g3_sample")
    logits, q = T(g["logits"]), T(g["q"])
    a_log = torch.argmax(logits - torch.log(q), dim=-1)
    assert np.array_equal(a_log.numpy(), g["actions"])
    gen = torch.Generator().manual_seed(5)
    z = torch.randn(1_000_000, 5, generator=gen) * 2.0
    qq = om.draw_race_noise((1_000_000, 5), generator=gen)
    a_ref, lp_ref = om.race_sample(z, qq)
    a_new = torch.argmax(z - torch.log(qq), dim=-1)
    agree = a_new == a_ref
    assert agree.float().mean() > 0.99999
    # the disagreements are races decided in the last bits: the runner-up is within 1e-5 relative of the winner
    r = torch.softmax(z[~agree], -1) / qq[~agree]
    top2 = r.topk(2, dim=-1).values
    assert ((top2[:, 0] - top2[:, 1]) <= 1e-5 * top2[:, 0]).all()
    lp_new = z.gather(-1, a_new[:, None])[:, 0] - torch.logsumexp(z, -1)
    assert (lp_new[agree] - lp_ref[agree]).abs().max() < 2e-6


def _episodes(g, tag, n):
    eps = []
    for i in range(n):
        eps.append({k: list(g[f"{tag}_ep{i}_{k}"]) for k in
                    ("obs", "actions", "log_prob", "reward", "states", "done", "avail_actions")})
    return eps


@pytest.mark.parametrize("tag,norm", [("plain", False), ("normr", True)])
def test_g4_buffer_collate_bit_exact(golden, tag, norm):
    g = golden("g4_buffer")
    batch = om.collate(_episodes(g, tag, len(g["lengths"])), 3, 21, 54, 5, normalize_reward=norm)
    names = ("obs", "actions", "log_probs", "reward", "states", "avail", "done", "mask")
    for n, t in zip(names, batch):
        ref = g[f"{tag}_{n}"]
        assert t.numpy().dtype == ref.dtype, n
        assert np.array_equal(t.numpy(), ref), n


def _batch(g):
    return (T(g["obs"]), T(g["actions"]), T(g["log_probs"]), T(g["reward"]), T(g["states"]),
            T(g["avail"]), T(g["done"]), T(g["mask"]))


def golden_networks(g, ippo=False):
    """The reference's networks for a G8 fixture: seed, widths and (fixtures made with non-default shapes) layer counts."""
    kw = dict(actor_hidden=int(g["actor_hidden_dim"]), critic_hidden=int(g["critic_hidden_dim"]))
    if "actor_num_layers" in g:
        kw.update(actor_layers=int(g["actor_num_layers"]), critic_layers=int(g["critic_num_layers"]))
    if ippo:
        kw.update(state_dim=21)
    return om.build_networks(int(g["seed"]), **kw)


@pytest.mark.parametrize("name,ippo", [("g8_mappo", False), ("g8_mappo_flags", False), ("g8_ippo", True), ("g8_mappo_deep", False)])
def test_g8_whole_iteration(golden, name, ippo):
    """TD(lambda) + normalisation + 3 PPO epochs + Adam reproduce the reference run bit for bit (g8_mappo_deep: the
    reference run with --actor_num_layers 2 --critic_hidden_dim 128 --clip_gradients 0.5)."""
    g = golden(name)
    actor, critic = golden_networks(g, ippo)
    batch = _batch(g)
    obs, actions, logp, reward, states, avail, done, mask = batch
    critic_in = obs if ippo else states
    gamma, lam = float(g["gamma"]), float(g["td_lambda"])
    ret, adv = om.td_lambda_loop(critic, critic_in, reward, mask, gamma, lam, 3)
    # the batched form (used at large sizes) stays within the stated tolerance of the loop form
    ret_b, adv_b = om.td_lambda_batched(critic, critic_in, reward, mask, gamma, lam, 3)
    assert (ret_b - ret).abs().max() < 1e-5 and (adv_b - adv).abs().max() < 1e-5
    if bool(g["normalize_advantage"]):
        adv = om.normalize_masked(adv, mask)
    if bool(g["normalize_return"]):
        ret = om.normalize_masked(ret, mask)
    assert np.array_equal(ret.numpy(), g["return_lambda"])
    assert np.array_equal(adv.numpy(), g["advantages"])
    if not ippo:
        # MAPPO: the scalar advantage is broadcast to the agents (MME:484-485, 496, 502)
        assert (adv[..., 0:1] == adv).all()

    aopt, copt = om.make_optimizers(actor, critic, float(g["lr_actor"]), float(g["lr_critic"]))
    stats = om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=int(g["epochs"]),
                          clip=float(g["ppo_clip"]), ent_coef=float(g["entropy_coef"]),
                          clip_gradients=float(g["clip_gradients"]), critic_on_obs=ippo)
    np.testing.assert_array_equal(np.array(stats["actor_loss"]), g["actor_losses"])
    np.testing.assert_array_equal(np.array(stats["critic_loss"]), g["critic_losses"])
    np.testing.assert_array_equal(np.array(stats["entropy"]), g["entropies"])
    np.testing.assert_array_equal(np.array(stats["kl"]), g["kls"])
    np.testing.assert_array_equal(np.array(stats["clipfrac"]), g["clipfracs"])
    np.testing.assert_array_equal(np.array(stats["actor_grad_norm"]), g["actor_grad_norms"])
    np.testing.assert_array_equal(np.array(stats["critic_grad_norm"]), g["critic_grad_norms"])
    assert np.array_equal(actor.flat_params().numpy(), g["actor_final"])
    assert np.array_equal(critic.flat_params().numpy(), g["critic_final"])


def test_flat_epoch_matches_loop_epoch(golden):
    g = golden("g8_mappo")
    actor, critic = om.build_networks(int(g["seed"]))
    obs, actions, logp, reward, states, avail, done, mask = _batch(g)
    adv, ret = T(g["advantages"]), T(g["return_lambda"])
    outs = []
    for fn in (om.ppo_epoch_loop, om.ppo_epoch_flat):
        actor.zero_grad(); critic.zero_grad()
        o = fn(actor, critic, obs, actions, logp, states, avail, mask, adv, ret, 0.2, 0.001)
        o.actor_loss.backward(); o.critic_loss.backward()
        outs.append((o, actor.flat_grads().clone(), critic.flat_grads().clone()))
    (a, ga, gc), (b, gb, gd) = outs
    assert abs(a.actor_loss.item() - b.actor_loss.item()) < 1e-5 * abs(a.actor_loss.item())
    assert abs(a.critic_loss.item() - b.critic_loss.item()) < 1e-5 * abs(a.critic_loss.item())
    assert (ga - gb).abs().max() <= 1e-5 * ga.abs().max()
    assert (gc - gd).abs().max() <= 1e-5 * gc.abs().max()


def test_scan_matches_loop_bit_exact_given_values(golden):
    """td_lambda_scan is the bit-exact target of the CUDA scan: same fp32 op order as MME:496-504."""
    g = golden("g8_mappo")
    actor, critic = om.build_networks(int(g["seed"]))
    obs, actions, logp, reward, states, avail, done, mask = _batch(g)
    B, Tn = reward.shape
    # batch-of-1 critic values, exactly what the loop sees
    with torch.no_grad():
        v = torch.stack([torch.stack([critic(states[b, t]) for t in range(Tn)]) for b in range(B)])
    ret, adv = om.td_lambda_scan(v.expand(B, Tn, 3), reward, mask, float(g["gamma"]), float(g["td_lambda"]))
    assert np.array_equal(ret.numpy(), g["return_lambda"])
    assert np.array_equal(adv.numpy(), g["advantages"])


# ------------------------------------------------------------------ recurrent actor (mappo_lstm_multienvs.py)
from oracle import mappo_lstm as ol  # noqa: E402


def test_g1_recurrent_param_init_bit_exact(golden):
    g = golden("g1_params")
    for seed in (1, 7):
        a, c = ol.build_networks(seed)
        assert np.array_equal(a.flat_params().numpy(), g[f"lstm_actor_s{seed}"])
        assert np.array_equal(c.flat_params().numpy(), g[f"lstm_critic_s{seed}"])
    assert a.flat_params().numel() == 7205
    assert [n for n, _ in a.named_parameters()] == list(g["lstm_actor_names"])


def test_g3_recurrent_act_carries_hidden_state(golden):
    g = golden("g3_recurrent")
    actor = ol.GRUActor(21, 32, 5)
    actor.load_flat(T(g["params"]))
    h = None
    for t in range(5):
        x, avail, q = T(g["x"][t]), T(g["avail"][t]), T(g["q"][t])
        a, lp, h, z = ol.rollout_act(actor, x[:, None, :], h, avail[:, None, :], q[:, None, :])
        assert np.array_equal(z[:, 0].numpy(), g[f"logits{t}"])
        assert np.array_equal(h.numpy(), g[f"h{t}"])
        assert np.array_equal(a[:, 0].numpy(), g[f"actions{t}"])
        assert np.array_equal(lp[:, 0].numpy(), g[f"logp{t}"])


def test_tbptt_chunks():
    assert ol.tbptt_chunks(25, 10) == [(0, 10), (10, 20), (20, 25)]
    assert ol.tbptt_chunks(25, 7) == [(0, 7), (7, 14), (14, 21), (21, 25)]
    assert ol.tbptt_chunks(25, 25) == [(0, 25)]
    assert ol.tbptt_chunks(25, 40) == [(0, 25)]


@pytest.mark.parametrize("name", ["g8_mappo_lstm", "g8_mappo_lstm_flags", "g8_ippo_lstm"])
def test_g8_recurrent_whole_iteration(golden, name):
    """TD(lambda) + truncated-BPTT epochs (actor Adam step per chunk) reproduce the reference run bit for bit
    (``g8_ippo_lstm``: ippo_lstm_multienvs.py -- critic on the observations, AdamW, tbptt 5)."""
    g = golden(name)
    ippo = name == "g8_ippo_lstm"
    if ippo:
        actor, critic = ol.build_networks(int(g["seed"]), state_dim=21, critic_hidden=int(g["critic_hidden_dim"]))
        assert str(g["optimizer"]) == "AdamW" and int(g["tbptt"]) == 5
    else:
        actor, critic = ol.build_networks(int(g["seed"]))
    batch = _batch(g)
    obs, actions, logp, reward, states, avail, done, mask = batch
    ret, adv = om.td_lambda_loop(critic, obs if ippo else states, reward, mask, float(g["gamma"]), float(g["td_lambda"]), 3)
    if bool(g["normalize_advantage"]):
        adv = om.normalize_masked(adv, mask)
    if bool(g["normalize_return"]):
        ret = om.normalize_masked(ret, mask)
    assert np.array_equal(ret.numpy(), g["return_lambda"])
    assert np.array_equal(adv.numpy(), g["advantages"])
    aopt, copt = om.make_optimizers(actor, critic, float(g["lr_actor"]), float(g["lr_critic"]),
                                    name=str(g["optimizer"]) if "optimizer" in g.files else "Adam")
    stats = ol.ppo_update_tbptt(actor, critic, aopt, copt, batch, adv, ret, epochs=int(g["epochs"]),
                                clip=float(g["ppo_clip"]), ent_coef=float(g["entropy_coef"]),
                                tbptt=int(g["tbptt"]), clip_gradients=float(g["clip_gradients"]), critic_on_obs=ippo)
    np.testing.assert_array_equal(np.array(stats["actor_loss"]), g["actor_losses"])
    np.testing.assert_array_equal(np.array(stats["critic_loss"]), g["critic_losses"])
    np.testing.assert_array_equal(np.array(stats["entropy"]), g["entropies"])
    np.testing.assert_array_equal(np.array(stats["kl"]), g["kls"])
    np.testing.assert_array_equal(np.array(stats["clipfrac"]), g["clipfracs"])
    np.testing.assert_allclose(np.array(stats["actor_grad_norm"]), g["actor_grad_norms"], rtol=1e-7)
    np.testing.assert_array_equal(np.array(stats["critic_grad_norm"]), g["critic_grad_norms"])
    assert np.array_equal(actor.flat_params().numpy(), g["actor_final"])
    assert np.array_equal(critic.flat_params().numpy(), g["critic_final"])
