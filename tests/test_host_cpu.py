"""CPU tests (no GPU): the C-ABI library loads and exports every symbol of include/cmarl_b200.h, the host mirror of
the reference's CLI / objects behaves like the reference's, missing-GPU paths fail loudly, and the multi-GPU host
logic (env sharding, one all-reduce of unnormalised sums per epoch, identical Adam on every rank) is exercised
under ``gloo`` with world_size 2 on a CPU test double of the engine (tests/fake_engine.py)."""
import ctypes as C
import dataclasses
import json
import os
import re
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "tests"))


# ------------------------------------------------------------------------------------------ C ABI
def header_symbols():
    text = (REPO / "include" / "cmarl_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmarl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from cleanmarl_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 26
    lib = C.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cmarl_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "cleanmarl_b200/_lib.py must bind exactly the header's entry points"
    _lib.load()
    assert _lib.load().cmarl_version() == _lib.VERSION == 104


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from cleanmarl_b200 import _lib
    import cleanmarl_b200 as cm
    lib = _lib.load()
    cfg = _lib.Config(0, 64, 25, 3, 21, 54, 5, 32, 1, 64, 1, 0, 0, 0)
    h = C.c_void_p()
    rc = lib.cmarl_ctx_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and not h.value and lib.cmarl_last_error()          # no CPU fallback inside the library
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cm.Engine(cm.Shapes(n_envs=64))


def test_ctx_rejects_unsupported_configurations():
    from cleanmarl_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    # (layer counts / widths / agent counts beyond the fused kernels' defaults are served by the layered kernels; what stays
    # invalid: inconsistent dimensions, out-of-range values, the recurrent actor on non-default shapes)
    for bad in (dict(n_agents=4), dict(n_agents=9, state_dim=9 * 54, obs_dim=63), dict(n_actions=6), dict(actor_layers=0),
                dict(critic_layers=7), dict(actor_hidden=300), dict(critic_hidden=0), dict(obs_dim=20), dict(state_dim=50),
                dict(n_envs=0), dict(actor_recurrent=2), dict(actor_recurrent=1, actor_hidden=64),
                dict(actor_recurrent=1, critic_layers=2), dict(n_landmarks=2)):
        base = dict(device=0, n_envs=64, n_steps=25, n_agents=3, obs_dim=21, state_dim=54, n_actions=5, actor_hidden=32,
                    actor_layers=1, critic_hidden=64, critic_layers=1, critic_on_obs=0, actor_recurrent=0, n_landmarks=0)
        base.update(bad)
        cfg = _lib.Config(*base.values())
        assert lib.cmarl_ctx_create(C.byref(cfg), C.byref(h)) < 0, bad      # argument error, before any CUDA call
        assert lib.cmarl_last_error()


# ------------------------------------------------------------------------------------------ CLI mirror
def test_args_mirror_the_reference_dataclass():
    """Same fields, order, types and defaults as MME:18-79 / ippo_multienvs.py (fixture made from the reference by
    tests/golden/gen_golden.py); documented deviations: env_type/env_name default to the one implemented env, device
    defaults to cuda."""
    from cleanmarl_b200.mappo import Args
    from cleanmarl_b200.ippo_multienvs import Args as IppoArgs
    from cleanmarl_b200.mappo_lstm_multienvs import Args as LstmArgs
    from cleanmarl_b200.ippo_lstm_multienvs import Args as IppoLstmArgs
    ref = json.loads((REPO / "tests" / "golden" / "g0_args.json").read_text())
    deviations = {"env_type": "pz", "env_name": "simple_spread_v3", "device": "cuda"}
    import importlib
    singles = [(n, importlib.import_module(f"cleanmarl_b200.single.{n}").Args) for n in ("mappo", "ippo", "mappo_lstm", "ippo_lstm")]
    for name, cls in [("mappo_multienvs", Args), ("ippo_multienvs", IppoArgs), ("mappo_lstm_multienvs", LstmArgs),
                      ("ippo_lstm_multienvs", IppoLstmArgs)] + singles:
        from cleanmarl_b200.mappo import EXTENSION_FIELDS
        ours = {f.name: f for f in dataclasses.fields(cls) if f.name not in EXTENSION_FIELDS}
        assert all(f.default in (-1, 1, 3) for f in dataclasses.fields(cls) if f.name in EXTENSION_FIELDS)   # defaults = the reference (off; 3 agents)
        theirs = {f["name"]: f for f in ref[name]}
        assert sorted(ours) == sorted(theirs)            # (ippo_multienvs.py lists ppo_clip/entropy_coef before epochs;
        if name == "mappo_multienvs":                    #  flag order is irrelevant to the keyword CLI)
            assert list(ours) == list(theirs)
        for k, f in ours.items():
            assert getattr(f.type, "__name__", str(f.type)) == theirs[k]["type"], k
            assert f.default == deviations.get(k, theirs[k]["default"]), k


def test_cli_parses_like_tyro_reference_and_rejects_what_is_not_built():
    import tyro
    from cleanmarl_b200.mappo import Args, validate_args
    a = tyro.cli(Args, args=["--batch_size", "4096", "--no-agent-ids", "--td-lambda", "0.9", "--normalize_advantage"])
    assert a.batch_size == 4096 and a.agent_ids is False and a.td_lambda == 0.9 and a.normalize_advantage is True
    validate_args(a)
    for bad in (dict(env_type="smaclite"), dict(env_name="simple_tag_v3"), dict(device="cpu"), dict(optimizer="SGD"),
                dict(actor_num_layers=0), dict(critic_hidden_dim=512), dict(n_agents=9), dict(num_minibatches=0), dict(batch_size=0)):
        with pytest.raises(SystemExit):
            validate_args(dataclasses.replace(a, **bad))
    validate_args(dataclasses.replace(a, optimizer="AdamW"))
    validate_args(dataclasses.replace(a, actor_num_layers=2, critic_hidden_dim=128, n_agents=5))     # layered kernels
    from cleanmarl_b200.mappo import ArgsRecurrent
    r = tyro.cli(ArgsRecurrent, args=["--batch_size", "8192", "--tbptt", "5"])
    assert r.tbptt == 5 and r.num_eval_ep == 5
    validate_args(r)
    for bad in (dict(tbptt=0), dict(actor_hidden_dim=64), dict(critic_num_layers=2)):
        with pytest.raises(SystemExit):
            validate_args(dataclasses.replace(r, **bad))


def test_layout_round_trip_is_bit_exact():
    """Device layout ([T][.][B], env-minor) <-> the reference's get_batch 8-tuple (MME:148-157): pure transposes."""
    from cleanmarl_b200 import engine as E
    from oracle import mappo as om
    actor, _ = om.build_networks(1)
    batch = om.synthetic_batch(37, seed=5, actor=actor)
    d = E.to_device_layout(batch, torch.device("cpu"))
    back = E.to_reference_layout(d)
    for i, (x, y) in enumerate(zip(batch, back)):
        assert x.dtype == y.dtype and torch.equal(x, y), i
    d2 = {k: v for k, v in d.items() if k != "obs"}
    back2 = E.to_reference_layout(d2)                     # obs rebuilt from state + one-hot ids
    assert torch.equal(back2[0], batch[0])
    adv = torch.randn(37, 25, 1).expand(37, 25, 3).contiguous()
    assert torch.equal(E.heads_to_reference(E.heads_to_device(adv, 1, "cpu"), 3), adv)


# ------------------------------------------------------------------------------------------ host logic on the CPU double
def _trainer(batch_size, rank=0, world=1, recurrent=False, **kw):
    from cleanmarl_b200.mappo import MAPPO, Args, ArgsRecurrent
    from fake_engine import OracleEngine
    return MAPPO((ArgsRecurrent if recurrent else Args)(batch_size=batch_size, seed=3, **kw), rank=rank, world_size=world,
                 engine_factory=lambda shapes, dev: OracleEngine(shapes))


def _inputs(B, seed=11):
    g = torch.Generator().manual_seed(seed)
    env = torch.zeros(18, B, dtype=torch.float64)
    env[0:6] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    env[12:18] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    noise = torch.empty(25, 3, 5, B).exponential_(1, generator=g)
    return env, noise


def test_trainer_iteration_matches_oracle_update():
    """MAPPO.iteration (collect -> advantages -> epochs) on the CPU double == the oracle's ppo_update on the same batch."""
    from oracle import mappo as om
    B = 12
    env, noise = _inputs(B)
    tr = _trainer(B)
    p0 = tr.net.flat.clone()
    tr.iteration(env.clone(), noise)
    assert tr.step == B * 25 and tr.training_step == 3 and tr.num_episodes == B
    batch = tr.get_batch()
    actor, critic = om.build_networks(3)
    assert torch.equal(torch.cat([actor.flat_params(), critic.flat_params()]), p0)      # reference init (seed 3)
    ret, adv = om.td_lambda_batched(critic, batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    aopt, copt = om.make_optimizers(actor, critic)
    st = om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, flat=True)
    final = torch.cat([actor.flat_params(), critic.flat_params()])
    assert (tr.net.flat - final).abs().max() < 1e-6
    sc = tr.train_scalars()
    assert abs(sc["actor_loss"] - np.mean(st["actor_loss"])) < 1e-5 * abs(np.mean(st["actor_loss"])) + 1e-7


def test_trainer_extensions_value_clip_and_minibatches():
    """The two default-off options beyond the reference (--value_clip, --num_minibatches): the trainer's loop (one
    optimizer step per contiguous env block, clipped value loss against the rollout-time values) on the CPU double ==
    the oracle's ppo_update with the same options; and with the options at their defaults nothing changes."""
    from oracle import mappo as om
    B = 12
    env, noise = _inputs(B)
    tr = _trainer(B, value_clip=0.2, num_minibatches=3)
    tr.iteration(env.clone(), noise)
    assert tr.training_step == 9 and tr.step == B * 25
    batch = tr.get_batch()
    actor, critic = om.build_networks(3)
    ret, adv = om.td_lambda_batched(critic, batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    with torch.no_grad():
        v_old = critic(batch[4]).expand(B, 25, 3).contiguous()
    aopt, copt = om.make_optimizers(actor, critic)
    st = om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, flat=True,
                       value_clip=0.2, values_old=v_old, num_minibatches=3)
    assert len(st["actor_loss"]) == 9
    assert (tr.net.flat - torch.cat([actor.flat_params(), critic.flat_params()])).abs().max() < 1e-6
    sc = tr.train_scalars()
    assert abs(sc["critic_loss"] - np.mean(st["critic_loss"])) < 1e-5 * abs(np.mean(st["critic_loss"])) + 1e-7
    # the clipped loss really differs from the plain MSE on this batch
    plain = _trainer(B)
    plain.iteration(env.clone(), noise)
    assert (plain.net.flat - tr.net.flat).abs().max() > 1e-5
    off = _trainer(B, value_clip=-1, num_minibatches=1)
    off.iteration(env.clone(), noise)
    assert torch.equal(off.net.flat, plain.net.flat)


def test_recurrent_trainer_iteration_matches_oracle_update():
    """MAPPO(ArgsRecurrent).iteration on the CPU double (truncated-BPTT chunks, actor step per chunk, hidden state carried
    through h_seq, critic step per epoch, mappo_lstm_multienvs.py:551-664) == the oracle's ppo_update_tbptt."""
    from oracle import mappo as om
    from oracle import mappo_lstm as ol
    B = 10
    env, noise = _inputs(B)
    tr = _trainer(B, recurrent=True, tbptt=7)
    assert tr.chunks == [(0, 7), (7, 14), (14, 21), (21, 25)] and tr.engine.n_actor == 7205
    p0 = tr.net.flat.clone()
    tr.iteration(env.clone(), noise)
    batch = tr.get_batch()
    actor, critic = ol.build_networks(3)
    assert torch.equal(torch.cat([actor.flat_params(), critic.flat_params()]), p0)      # reference init (seed 3)
    ret, adv = om.td_lambda_batched(critic, batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    aopt, copt = om.make_optimizers(actor, critic)
    st = ol.ppo_update_tbptt(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, tbptt=7)
    final = torch.cat([actor.flat_params(), critic.flat_params()])
    assert (tr.net.flat - final).abs().max() < 2e-6
    sc = tr.train_scalars()
    assert abs(sc["actor_loss"] - np.mean(st["actor_loss"])) < 1e-5 * abs(np.mean(st["actor_loss"])) + 1e-7
    assert abs(sc["critic_loss"] - np.mean(st["critic_loss"])) < 1e-5 * abs(np.mean(st["critic_loss"]))
    assert abs(sc["actor_gradients"] - np.mean(st["actor_grad_norm"])) < 1e-5 * np.mean(st["actor_grad_norm"])
    assert tr.training_step == 3 and int(tr.adam_step_a) == 12 and int(tr.adam_step) == 3


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _rank_main(rank, world, port, B, out_dir, flags):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
    env, noise = _inputs(B)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    tr = _trainer(B, rank, world, **flags)
    assert tr.B == per
    for _ in range(2):
        tr.iteration(env[:, sl].clone(), noise[..., sl].contiguous())
    roll = tr.rollout_scalars()
    torch.save({"params": tr.net.flat, "stats": tr.epoch_stats, "step": tr.step, "ep_reward": roll["ep_reward"],
                "episodes": tr.num_episodes}, Path(out_dir) / f"rank{rank}.pt")
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("flags", [{}, {"normalize_advantage": True, "normalize_reward": True, "clip_gradients": 0.5},
                                   {"recurrent": True}],
                         ids=["plain", "normalised+clip", "recurrent"])
def test_two_ranks_gloo_equal_one_rank(tmp_path, flags):
    """Envs sharded over 2 ranks (gloo, CPU double) == 1 rank on all envs: same parameters (fp32 reassociation only),
    identical replicas, global step / episode counters and global normalisation statistics."""
    import torch.multiprocessing as mp
    B, world = 16, 2
    mp.spawn(_rank_main, args=(world, _free_port(), B, str(tmp_path), flags), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(world))
    assert torch.equal(r0["params"], r1["params"]), "replicas must stay bit-identical"
    assert torch.equal(r0["stats"], r1["stats"])
    env, noise = _inputs(B)
    tr = _trainer(B, **flags)
    for _ in range(2):
        tr.iteration(env.clone(), noise)
    assert r0["step"] == tr.step == 2 * B * 25 and r0["episodes"] == tr.num_episodes
    assert (r0["params"] - tr.net.flat).abs().max() < 2e-6
    assert (r0["stats"] - tr.epoch_stats).abs().max() < 1e-4 * tr.epoch_stats.abs().max()
    assert abs(r0["ep_reward"] - tr.rollout_scalars()["ep_reward"]) < 1e-9


# ------------------------------------------------------------------------------------------ the script's outer loop
def test_cli_main_loop_logs_like_the_reference(tmp_path, monkeypatch):
    """The drop-in script's ``while step < total_timesteps`` loop on the CPU double with a recording writer: tag names,
    x-axis (env steps, MME:435) and cadence of MME:454-468 (mean over ALL episodes pending since the last line, logged
    once more than ``log_every`` are pending), MME:605-612 (eight train scalars every iteration) and MME:614 (eval when
    ``training_step / epochs`` is a multiple of ``eval_steps``)."""
    from cleanmarl_b200 import mappo_multienvs as cli
    from cleanmarl_b200.mappo import MAPPO
    from fake_engine import OracleEngine
    monkeypatch.chdir(tmp_path)
    returns, evals, log = [], [], []

    class CpuTrainer(MAPPO):
        def __init__(self, args, device_index=0, rank=0, world_size=1, ippo=False):
            super().__init__(args, rank=rank, world_size=world_size, ippo=ippo,
                             engine_factory=lambda shapes, dev: OracleEngine(shapes))
            self._g = torch.Generator().manual_seed(5)

        def iteration(self):
            noise = torch.empty(25, 3, 5, self.B).exponential_(1, generator=self._g)
            super().iteration(None, noise)
            returns.append(self.buf["ep_return"].clone())

    class Recorder:
        def __init__(self, logdir):
            log.append(("dir", logdir))

        def add_text(self, tag, text):
            log.append(("text", tag, text))

        def add_scalar(self, tag, value, step):
            log.append(("scalar", tag, float(value), int(step)))

        def close(self):
            log.append(("close",))

    def fake_eval(trainer, n, seed):
        evals.append((trainer.training_step, n, seed))
        return -30.0, 2.0, 25.0

    tr = cli.main(["--batch_size", "4", "--total_timesteps", "600", "--log_every", "10", "--eval_steps", "2",
                   "--num_eval_ep", "3", "--seed", "3"], trainer_cls=CpuTrainer, evaluate_fn=fake_eval,
                  SummaryWriter=Recorder)
    assert tr.step == 600 and tr.training_step == 18 and tr.num_episodes == 24 and len(returns) == 6
    assert re.fullmatch(r"runs/MAPPO-multienvs-pz__simple_spread_v3__\d{4}-\d\d-\d\d_\d\d-\d\d-\d\d", log[0][1])   # MME:345-357
    assert log[1][:2] == ("text", "hyperparameters") and "|batch_size|4|" in log[1][2] and log[-1] == ("close",)
    scalars = [e[1:] for e in log if e[0] == "scalar"]
    by_step = {s: [(t, v) for t, v, st in scalars if st == s] for s in range(100, 700, 100)}
    train_tags = ["train/actor_loss", "train/critic_loss", "train/entropy", "train/kl_divergence", "train/clipped_ratios",
                  "train/actor_gradients", "train/critic_gradients", "train/num_updates"]
    for k, step in enumerate(range(100, 700, 100), start=1):
        tags = [t for t, _ in by_step[step]]
        assert sorted(t for t in tags if t.startswith("train/")) == sorted(train_tags)
        assert dict(by_step[step])["train/num_updates"] == 3 * k
        assert ("eval/ep_reward" in tags) == (k % 2 == 0)
        assert ("rollout/ep_reward" in tags) == (k % 3 == 0)         # 4, 8, 12 > 10 episodes pending
        if k % 2 == 0:
            assert [dict(by_step[step])[t] for t in ("eval/ep_reward", "eval/std_ep_reward", "eval/ep_length")] == [-30.0, 2.0, 25.0]
        if k % 3 == 0:
            d = dict(by_step[step])
            pending = torch.cat(returns[k - 3:k]).double()
            assert abs(d["rollout/ep_reward"] - float(pending.mean())) < 1e-9
            assert d["rollout/ep_length"] == 25.0 and d["rollout/num_episodes"] == 4 * k
    assert [e[0] for e in evals] == [6, 12, 18] and all(e[1] == 3 for e in evals)


@pytest.mark.parametrize("script,prefix,ippo,recurrent", [("mappo", "MAPPO", False, False), ("ippo", "IPPO", True, False),
                                                         ("mappo_lstm", "MAPPO-lstm", False, True), ("ippo_lstm", "IPPO-lstm", True, True)])
def test_single_env_scripts_run_the_shared_loop(tmp_path, monkeypatch, script, prefix, ippo, recurrent):
    """``cleanmarl_b200/single/*.py`` (drop-ins for the single-env ``mappo.py`` / ``ippo.py`` / ``*_lstm.py``): their own
    ``Args`` class, the reference's run-directory prefix (mappo.py:277-279), and the shared
    training loop -- run for ten iterations on the CPU double."""
    import importlib
    from cleanmarl_b200 import mappo_multienvs as cli
    from cleanmarl_b200.mappo import MAPPO
    from fake_engine import OracleEngine
    mod = importlib.import_module(f"cleanmarl_b200.single.{script}")
    monkeypatch.chdir(tmp_path)
    log, evals = [], []

    class CpuTrainer(MAPPO):
        def __init__(self, args, device_index=0, rank=0, world_size=1, ippo=False):
            super().__init__(args, rank=rank, world_size=world_size, ippo=ippo,
                             engine_factory=lambda shapes, dev: OracleEngine(shapes))
            self._g = torch.Generator().manual_seed(5)

        def iteration(self):
            super().iteration(None, torch.empty(25, 3, 5, self.B).exponential_(1, generator=self._g))

    class Recorder:
        def __init__(self, logdir): log.append(("dir", logdir))
        def add_text(self, tag, text): pass
        def add_scalar(self, tag, value, step): log.append((tag, float(value), int(step)))
        def close(self): pass

    def fake_eval(trainer, n, seed):
        evals.append(trainer.training_step)
        return -30.0, 2.0, 25.0

    src = (Path(mod.__file__).read_text())
    assert f'run_prefix="{prefix}"' in src
    tr = cli.main(["--batch_size", "2", "--total_timesteps", "500", "--epochs", "1", "--eval_steps", "10"], algo=prefix, ippo=ippo, args_cls=mod.Args,
                  run_prefix=prefix, trainer_cls=CpuTrainer, evaluate_fn=fake_eval, SummaryWriter=Recorder)
    assert tr.ippo == ippo and tr.recurrent == recurrent          # (the Args defaults are checked against the reference fixture above)
    assert re.fullmatch(rf"runs/{re.escape(prefix)}-pz__simple_spread_v3__\d{{4}}-\d\d-\d\d_\d\d-\d\d-\d\d", log[0][1])
    assert tr.step == 500 and tr.training_step == 10 and evals == [10]      # (training_step / epochs) % 10 == 0 at the 10th iteration
    assert sum(1 for e in log if e[0] == "train/actor_loss") == 10 and sum(1 for e in log if e[0] == "eval/ep_reward") == 1


def test_env_duck_type_on_the_cpu_double():
    """SpreadVecEnv (env/common_interface.py:5-23 with a leading env axis) on the engine double; the same body runs against
    the CUDA library in tests/test_gpu_vecenv.py."""
    from cleanmarl_b200.engine import Shapes
    from cleanmarl_b200.mappo import SpreadVecEnv
    from env_contract import check_env_duck_type
    from fake_engine import OracleEngine
    B = 300
    check_env_duck_type(lambda seed: SpreadVecEnv(OracleEngine(Shapes(n_envs=B)), agent_ids=True, seed=seed), B,
                        state_tol=0.0, obs_tol=0.0)


def test_fast_divmod_model_is_exact_below_2_pow_24():
    """csrc/tc_chain.cu:101-108 decodes tile indices as q = trunc(float(u) * (1.0f / d)) with one correction step either
    way; csrc/chain.cu:349 refuses launches with 2^24 tiles or more.  numpy float32 performs the same IEEE operations
    (exact int -> float below 2^24, round-to-nearest multiply and divide, truncating conversion), so the claim 'the
    estimate is within one of the quotient' can be checked here: multiples of d and their neighbours, the top of the range
    and random indices, for every divisor the kernels can meet (tiles per time step <= 2^19 at 64 Mi envs, agent groups <= 3)
    and a sweep of others."""
    rng = np.random.default_rng(0)
    top = (1 << 24) - 1
    divisors = np.unique(np.concatenate([np.arange(1, 4100), 2 ** np.arange(12, 24), 2 ** np.arange(12, 24) - 1,
                                         2 ** np.arange(12, 24) + 1, rng.integers(4100, 1 << 23, 2000)])).astype(np.int64)
    for d in divisors:
        inv = np.float32(1.0) / np.float32(d)
        k = rng.integers(0, top // d + 1, 64)
        u = np.unique(np.clip(np.concatenate([k * d, k * d - 1, k * d + 1, k * d + d - 1, [0, top, top - 1, top - d]]), 0, top))
        q = np.trunc(u.astype(np.float32) * inv).astype(np.int64)
        assert (np.abs(q - u // d) <= 1).all(), d                    # what the single correction step relies on
        r = u - q * d
        q = np.where(r < 0, q - 1, np.where(r >= d, q + 1, q))
        r = np.where(r < 0, r + d, np.where(r >= d, r - d, r))
        assert np.array_equal(q, u // d) and np.array_equal(r, u % d), d
    u = rng.integers(0, top + 1, 200000)
    for d in (1, 2, 3, 32, 512, 8192, 524288):
        inv = np.float32(1.0) / np.float32(d)
        q = np.trunc(u.astype(np.float32) * inv).astype(np.int64)
        assert (np.abs(q - u // d) <= 1).all(), d


def test_tf32_split_model():
    """csrc/tc_ptx.cuh:185-191: hi = bits(x) + 0x1000 & ~0x1FFF (round to nearest, ties away, on the sign-magnitude
    pattern), lo = the same on x - hi.  Checked on the same integer / fp32 operations in numpy: hi and lo are tf32 values
    (low 13 mantissa bits clear), x - hi is exact, and what the 3-term product scheme of csrc/tc_chain.cu drops,
    |x - hi - lo|, is at most 2^-23 |x| for normal x -- one fp32 ulp at the bottom of a binade (header of tc_chain.cu: 3xTF32
    measured 2.4e-7 against 1.8e-7 for a plain fp32 GEMM)."""
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(200000) * 10.0 ** rng.integers(-6, 7, 200000),
                        [0.0, -0.0, 1.0, -1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 + 2.0 ** -23, 2.0 - 2.0 ** -23,
                         -(2.0 - 2.0 ** -23), 3.0e38, 1.2e-38]]).astype(np.float32)

    def rna(v):
        return ((v.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)

    hi = rna(x)
    d = x - hi                                                        # fp32 subtraction, as on the device
    lo = rna(d)
    assert not (hi.view(np.uint32) & 0x1FFF).any() and not (lo.view(np.uint32) & 0x1FFF).any()
    assert np.array_equal(d.astype(np.float64), x.astype(np.float64) - hi.astype(np.float64))       # exact
    ax = np.abs(x.astype(np.float64))
    assert (np.abs(d.astype(np.float64)) <= ax * 2.0 ** -11).all()
    res = np.abs(x.astype(np.float64) - hi.astype(np.float64) - lo.astype(np.float64))
    normal = ax >= 2.0 ** -100                                        # gradual underflow: absolute, not relative, error
    assert (res[normal] <= ax[normal] * 2.0 ** -23).all()
    assert (res[normal] / ax[normal]).max() > 2.0 ** -23.5            # and the bound is attained: not 2^-24
    assert hi[-4] == np.float32(2.0) and hi[-3] == np.float32(-2.0)                                  # mantissa carry rounds up
    assert hi[-6] == np.float32(1.0 + 2.0 ** -10) and lo[-6] == np.float32(-(2.0 ** -11))            # tie: away from zero
    assert np.array_equal(np.signbit(hi[-10:-8]), [False, True]) and not hi[-10:-8].any()            # +-0 stay +-0


def test_adam_bias_correction_model():
    """csrc/exact.cu:303-311 forms beta^step by repeated squaring in fp64 where torch's ``_single_tensor_adam`` (MME:593-594)
    evaluates the python-float ``beta ** step``; the kernel then uses ``(float)(-(lr / bc1))`` and ``(float)sqrt(bc2)``.
    Same fp64 operations here: the float32 values the update sees are identical for the steps a run can reach (every step
    to 4096, then a sweep to 10^6)."""
    def by_squaring(beta, step):
        p, b, e = 1.0, beta, step
        while e:
            if e & 1:
                p *= b
            b *= b
            e >>= 1
        return p

    rng = np.random.default_rng(2)
    steps = np.unique(np.concatenate([np.arange(1, 4097), rng.integers(4097, 10 ** 6 + 1, 4000), [10 ** 6]]))
    for lr in (8e-4, 1e-3, 5e-4):
        for step in steps.tolist():
            bc1_t, bc2_t = 1 - 0.9 ** step, 1 - 0.999 ** step
            bc1_k, bc2_k = 1.0 - by_squaring(0.9, step), 1.0 - by_squaring(0.999, step)
            assert np.float32(-(lr / bc1_t)) == np.float32(-(lr / bc1_k)), step
            assert np.float32(np.sqrt(bc2_t)) == np.float32(np.sqrt(bc2_k)), step
