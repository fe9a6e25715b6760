"""Generate the golden fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/gen_golden.py

Needs ``/root/reference`` (or ``$CMARL_REFERENCE_DIR``); the GPU box never runs this, it only
reads the committed ``*.npz``.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so every fixture is produced by executing the reference's real code:

g0_args.json       the ``Args`` dataclasses of mappo_multienvs.py / ippo_multienvs.py (names, types, defaults).
g1_params.npz      ``Actor``/``Critic`` of MME:160-200 built after ``torch.manual_seed`` in the
                   order of MME:329-339 (MAPPO and IPPO shapes).
g3_sample.npz      ``Actor.act`` (MME:172-176) under a known generator state + the exponential
                   noise that state produces (pins "Categorical.sample == argmax(p/q)").
g4_buffer.npz      ``RolloutBuffer.add/get_batch`` (MME:103-157) on ragged synthetic episodes,
                   with and without ``normalize_reward``.
g8_mappo*.npz      one whole iteration of ``mappo_multienvs.py`` executed with ``runpy`` on the
                   numpy stand-in env (rollout with real worker processes, collate, TD(lambda)
                   loop MME:484-512, 3 PPO epochs MME:521-603): the batch, returns, advantages,
                   per-epoch statistics and the updated parameters.  ``_flags`` = all
                   normalize_* flags on + gradient clipping.
g8_mappo_deep.npz  the same with ``--actor_num_layers 2 --critic_hidden_dim 128 --clip_gradients 0.5`` (``--only-deep``).
g8_ippo.npz        the same for ``ippo_multienvs.py``.
g1_params.npz      also holds the recurrent ``Actor`` (fc1 + GRUCell + fc2) / ``Critic`` of
                   ``mappo_lstm_multienvs.py:162-200, 327-338`` (keys ``lstm_*``).
g3_recurrent.npz   five consecutive ``Actor.act(x, h, avail)`` calls of the recurrent actor
                   (``mappo_lstm_multienvs.py:170-184``) with the hidden state carried, under a known
                   generator state + the exponential noise that state produces.
g8_ippo_lstm.npz   the same for ``ippo_lstm_multienvs.py`` (decentralised critic on obs, AdamW, tbptt 5).
g8_mappo_lstm*.npz one whole iteration of ``mappo_lstm_multienvs.py`` (rollout with the GRU actor, TD(lambda),
                   3 epochs of truncated BPTT with an actor Adam step per chunk, ``mappo_lstm_multienvs.py:551-664``);
                   ``_flags`` = tbptt 7, advantage normalisation, gradient clipping.
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent.parent
# CMARL_GOLDEN_OUT=<dir>: regenerate into another directory (tests/golden/compare_golden.py checks the result against the committed files)
HERE = Path(__import__("os").environ.get("CMARL_GOLDEN_OUT") or Path(__file__).resolve().parent)
sys.path.insert(0, str(REPO))

from oracle import ref_loader  # noqa: E402


def flat(module):
    return torch.cat([p.detach().reshape(-1) for p in module.parameters()]).numpy()


def g0(ref, ref_ippo, ref_lstm=None, ref_ippo_lstm=None):
    """The reference's CLI dataclasses (MME:18-79, ippo_multienvs.py:18-79): field names, types, defaults."""
    import dataclasses
    import json
    out = {}
    mods = [("mappo_multienvs", ref), ("ippo_multienvs", ref_ippo)]
    if ref_lstm is not None:
        mods.append(("mappo_lstm_multienvs", ref_lstm))
    if ref_ippo_lstm is not None:
        mods.append(("ippo_lstm_multienvs", ref_ippo_lstm))
        for single in ("mappo", "ippo", "mappo_lstm", "ippo_lstm"):          # the single-env scripts (same update code)
            mods.append((single, ref_loader.load_module(single + ".py")))
    for name, mod in mods:
        out[name] = [{"name": f.name, "type": getattr(f.type, "__name__", str(f.type)), "default": f.default}
                     for f in dataclasses.fields(mod.Args)]
    (HERE / "g0_args.json").write_text(json.dumps(out, indent=1))


def g1(ref, ref_ippo):
    out = {}
    for seed in (1, 7):
        torch.manual_seed(seed)
        a = ref.Actor(21, 32, 1, 5)
        c = ref.Critic(54, 64, 1)
        out[f"mappo_actor_s{seed}"] = flat(a)
        out[f"mappo_critic_s{seed}"] = flat(c)
        torch.manual_seed(seed)
        a = ref_ippo.Actor(21, 32, 1, 5)
        c = ref_ippo.Critic(21, 32, 1)
        out[f"ippo_actor_s{seed}"] = flat(a)
        out[f"ippo_critic_s{seed}"] = flat(c)
    torch.manual_seed(3)
    a = ref.Actor(21, 64, 2, 5)
    c = ref.Critic(54, 128, 2)
    out["mappo_actor_wide"] = flat(a)
    out["mappo_critic_wide"] = flat(c)
    np.savez_compressed(HERE / "g1_params.npz", **out)


def g1_lstm(ref_lstm):
    out = dict(np.load(HERE / "g1_params.npz"))
    for seed in (1, 7):
        torch.manual_seed(seed)
        a = ref_lstm.Actor(21, 32, 5)
        c = ref_lstm.Critic(54, 64, 1)
        out[f"lstm_actor_s{seed}"] = flat(a)
        out[f"lstm_critic_s{seed}"] = flat(c)
    out["lstm_actor_names"] = np.array([n for n, _ in a.named_parameters()])
    np.savez_compressed(HERE / "g1_params.npz", **out)


def g3_recurrent(ref_lstm):
    torch.manual_seed(13)
    actor = ref_lstm.Actor(21, 32, 5)
    M, steps = 768, 5
    x = torch.randn(steps, M, 21)
    avail = torch.ones(steps, M, 5, dtype=torch.bool)
    avail[:, ::5, 2] = False
    out = {"params": flat(actor), "x": x.numpy(), "avail": avail.numpy()}
    with torch.no_grad():
        torch.manual_seed(77)
        h = None
        for t in range(steps):
            actions, logp, h = actor.act(x[t], h=h, avail_action=avail[t])
            out[f"actions{t}"] = actions.numpy(); out[f"logp{t}"] = logp.numpy(); out[f"h{t}"] = h.numpy()
        torch.manual_seed(77)
        q = torch.stack([torch.empty(M, 5).exponential_(1) for _ in range(steps)])
        h = None
        for t in range(steps):
            z, h = actor.logits(x[t], h, avail[t])
            out[f"logits{t}"] = z.numpy()
    out["q"] = q.numpy()
    np.savez_compressed(HERE / "g3_recurrent.npz", **out)


def g3(ref):
    torch.manual_seed(11)
    actor = ref.Actor(21, 32, 1, 5)
    x = torch.randn(2048, 3, 21)
    avail = torch.ones(2048, 3, 5, dtype=torch.bool)
    avail[::7, 1, 3] = False
    with torch.no_grad():
        logits = actor.logits(x, avail)
        torch.manual_seed(99)
        actions, logp = actor.act(x, avail)
        torch.manual_seed(99)
        q = torch.empty(2048 * 3, 5).exponential_(1)
    np.savez_compressed(HERE / "g3_sample.npz", params=flat(actor), x=x.numpy(), avail=avail.numpy(),
                        logits=logits.numpy(), q=q.numpy().reshape(2048, 3, 5),
                        actions=actions.numpy(), logp=logp.numpy())


def g4(ref):
    rng = np.random.default_rng(5)
    lengths = [25, 7, 25, 1, 13]
    out = {"lengths": np.array(lengths)}
    for tag, norm in (("plain", False), ("normr", True)):
        rb = ref.RolloutBuffer(len(lengths), 3, 21, 54, 5, normalize_reward=norm)
        eps_np = []
        for L in lengths:
            ep = {
                "obs": [rng.standard_normal((3, 21)) for _ in range(L)],
                "actions": [torch.from_numpy(rng.integers(0, 5, 3)) for _ in range(L)],
                "log_prob": [torch.from_numpy(-rng.random(3).astype(np.float32)) for _ in range(L)],
                "reward": [float(-rng.random() * 4) for _ in range(L)],
                "states": [rng.standard_normal(54).astype(np.float32) for _ in range(L)],
                "done": [False] * L,
                "avail_actions": [rng.integers(0, 2, (3, 5)) for _ in range(L)],
            }
            eps_np.append({k: np.stack([np.asarray(v) for v in vals]) for k, vals in ep.items()})
            rb.add(ep)
        batch = rb.get_batch()
        names = ("obs", "actions", "log_probs", "reward", "states", "avail", "done", "mask")
        for n, t in zip(names, batch):
            out[f"{tag}_{n}"] = t.numpy()
        for i, ep in enumerate(eps_np):
            for k, v in ep.items():
                out[f"{tag}_ep{i}_{k}"] = v
    np.savez_compressed(HERE / "g4_buffer.npz", **out)


def g8(script, tag, extra, B=6, seed=1):
    argv = ["--env_type", "pz", "--env_name", "simple_spread_v3", "--batch_size", str(B),
            "--total_timesteps", str(B * 25), "--eval_steps", "1000000000", "--seed", str(seed)] + extra
    with tempfile.TemporaryDirectory() as tmp:
        g = ref_loader.run_script(argv, script=script, cwd=tmp)
    args = g["args"]
    out = {
        "seed": np.array(seed), "B": np.array(B),
        "gamma": np.array(args.gamma), "td_lambda": np.array(args.td_lambda),
        "ppo_clip": np.array(args.ppo_clip), "entropy_coef": np.array(args.entropy_coef),
        "clip_gradients": np.array(args.clip_gradients), "epochs": np.array(args.epochs),
        "lr_actor": np.array(args.learning_rate_actor), "lr_critic": np.array(args.learning_rate_critic),
        "normalize_reward": np.array(args.normalize_reward),
        "normalize_advantage": np.array(args.normalize_advantage),
        "normalize_return": np.array(args.normalize_return),
        "actor_hidden_dim": np.array(args.actor_hidden_dim), "critic_hidden_dim": np.array(args.critic_hidden_dim),
        "actor_num_layers": np.array(args.actor_num_layers), "critic_num_layers": np.array(args.critic_num_layers),
        "obs": g["b_obs"].numpy(), "actions": g["b_actions"].numpy(), "log_probs": g["b_log_probs"].numpy(),
        "reward": g["b_reward"].numpy(), "states": g["b_states"].numpy(),
        "avail": g["b_avail_actions"].numpy(), "done": g["b_done"].numpy(), "mask": g["b_mask"].numpy(),
        "return_lambda": g["return_lambda"].numpy(), "advantages": g["advantages"].numpy(),
        "actor_final": flat(g["actor"]), "critic_final": flat(g["critic"]),
        "actor_losses": np.array(g["actor_losses"]), "critic_losses": np.array(g["critic_losses"]),
        "entropies": np.array(g["entropies_bonuses"]), "kls": np.array(g["kl_divergences"]),
        "clipfracs": np.array([float(x) for x in g["clipped_ratios"]]),
        "actor_grad_norms": np.array([float(x) for x in g["actor_gradients"]]),
        "critic_grad_norms": np.array([float(x) for x in g["critic_gradients"]]),
        "step": np.array(g["step"]), "training_step": np.array(g["training_step"]),
    }
    if hasattr(args, "tbptt"):
        out["tbptt"] = np.array(args.tbptt)
    out["optimizer"] = np.array(args.optimizer)
    np.savez_compressed(HERE / f"g8_{tag}.npz", **out)
    print(tag, "step", g["step"], "actor_losses", g["actor_losses"])


def main():
    if ref_loader.reference_dir() is None:
        raise SystemExit("reference sources not found")
    ref = ref_loader.load_module("mappo_multienvs.py")
    ref_ippo = ref_loader.load_module("ippo_multienvs.py")
    g0(ref, ref_ippo, ref_loader.load_module("mappo_lstm_multienvs.py"), ref_loader.load_module("ippo_lstm_multienvs.py"))
    if "--only-args" in sys.argv:
        return
    if "--only-lstm" in sys.argv:
        ref_lstm = ref_loader.load_module("mappo_lstm_multienvs.py")
        g1_lstm(ref_lstm)
        g3_recurrent(ref_lstm)
        g8("mappo_lstm_multienvs.py", "mappo_lstm", [], seed=4)
        g8("mappo_lstm_multienvs.py", "mappo_lstm_flags",
           ["--tbptt", "7", "--normalize_advantage", "--clip_gradients", "0.5"], seed=5)
        g8("ippo_lstm_multienvs.py", "ippo_lstm", [], seed=6)
        return
    if "--only-deep" in sys.argv:
        # shapes beyond the defaults (MME:160-171, 186-196): two hidden->hidden blocks in the actor, a 128-wide critic
        g8("mappo_multienvs.py", "mappo_deep", ["--actor_num_layers", "2", "--critic_hidden_dim", "128", "--clip_gradients", "0.5"], seed=7)
        return
    if "--only-ippo-lstm" in sys.argv:
        g8("ippo_lstm_multienvs.py", "ippo_lstm", [], seed=6)
        return
    g1(ref, ref_ippo)
    g3(ref)
    g4(ref)
    g8("mappo_multienvs.py", "mappo", [])
    g8("mappo_multienvs.py", "mappo_flags",
       ["--normalize_reward", "--normalize_advantage", "--normalize_return", "--clip_gradients", "0.5"], seed=2)
    g8("ippo_multienvs.py", "ippo", [], seed=3)
    ref_lstm = ref_loader.load_module("mappo_lstm_multienvs.py")
    g1_lstm(ref_lstm)
    g3_recurrent(ref_lstm)
    g8("mappo_lstm_multienvs.py", "mappo_lstm", [], seed=4)
    g8("mappo_lstm_multienvs.py", "mappo_lstm_flags",
       ["--tbptt", "7", "--normalize_advantage", "--clip_gradients", "0.5"], seed=5)
    g8("ippo_lstm_multienvs.py", "ippo_lstm", [], seed=6)


if __name__ == "__main__":
    main()
