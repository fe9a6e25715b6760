"""bench.py -- throughput of the MAPPO multi-env training path (rollout + TD(lambda)/GAE + PPO epochs).

    python bench.py --gpus N --steps K --warmup W            # this repo (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path, host cores

Metric (BASELINE.json): agent-env-steps/s of one full training iteration; a "step" is one iteration =
B*T*N agent-env-steps.  Workload at N=1: BASELINE.json configs[1] (mappo_multienvs.py, simple_spread_v3,
3 agents, num_envs=4096); N>1 keeps 4096 envs per GPU (weak scaling, envs sharded, one gradient
all-reduce per PPO epoch).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

T_STEPS, N_AGENTS, N_ACT = 25, 3, 5
METRIC = "agent_env_steps_per_sec"
UNIT = "agent-env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=None,
                    help="default 4096 (BASELINE configs[1], [2]); 8192 for --algo mappo_lstm (configs[3])")
    ap.add_argument("--algo", default="mappo", choices=["mappo", "ippo", "mappo_lstm"])
    ap.add_argument("--ref-envs", type=int, default=32, help="envs in the reference arm's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the short IPPO / recurrent runs at N=1")
    ap.add_argument("--gae-envs", type=int, default=1 << 20, help="envs for the stand-alone GAE roofline probe")
    a = ap.parse_args()
    if a.envs_per_gpu is None:
        a.envs_per_gpu = 8192 if a.algo == "mappo_lstm" else 4096
    return a


SCRIPTS = {"mappo": "mappo_multienvs.py", "ippo": "ippo_multienvs.py", "mappo_lstm": "mappo_lstm_multienvs.py"}


def peaks():
    """(HBM GB/s, dense bf16 TFLOP/s burst, SM max MHz, source)"""
    p = REPO / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return (float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), float(d.get("sm_max_mhz", 1965.0)),
                "measured (MEASURED_PEAKS.json)")
    return 6650.0, 1590.0, 1965.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch measured by ncu for this round's kernels (committed under profiles/; ncu cannot run inside
    the timed bench).  {} when the file is missing."""
    for name in ("traffic_r2.json", "traffic_r1.json"):
        p = REPO / "profiles" / name
        try:
            d = json.loads(p.read_text())
            d["_file"] = f"profiles/{name}"
            return d
        except Exception:        # noqa: BLE001
            continue
    return {}


def workload_config(a, world, envs_per_gpu=None, algo=None, note=""):
    algo = algo or a.algo
    epg = envs_per_gpu or a.envs_per_gpu
    return {
        "workload": f"{SCRIPTS[algo]} simple_spread_v3, 3 agents, T=25, num_envs={epg * world} "
                    f"({epg}/GPU){note}, 3 PPO epochs, actor "
                    f"{'21-32-GRU(32)-5, truncated BPTT 10 (3 actor steps per epoch)' if algo == 'mappo_lstm' else '21-32-32-5'}"
                    f", critic {'21-32-32-1 per agent' if algo == 'ippo' else '54-64-64-1'}",
        "global_batch": epg * world,
        "envs_per_gpu": epg,
        "agent_env_steps_per_step": epg * world * T_STEPS * N_AGENTS,
        "parallelism": (f"dp{world} (envs sharded, " + ("1 all-reduce of 7213 floats per TBPTT chunk + 1 of 7753 per epoch)"
                        if algo == "mappo_lstm" else "1 all-reduce of 9678 floats per epoch)")) if world > 1 else "single GPU",
        "l2": "flushed between timed iterations (256 MiB write, outside the per-step event pairs)",
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + clock-event (throttle) reasons sampled through NVML every 50 ms during the timed region."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_sm = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception as e:           # noqa: BLE001
            print(f"clock sampling unavailable: {e}", file=sys.stderr)
            return
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}

        def loop():
            while not self._stop.is_set():
                try:
                    self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    try:
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except AttributeError:
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for name, bit in bits.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:        # noqa: BLE001
                    pass
                time.sleep(0.05)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        return {"sm_mhz": int(statistics.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_port_iteration(B, seed=1, algo="mappo"):
    """One full iteration of the reference's arithmetic on the host (oracle port, torch CPU): rollout on the
    vectorised numpy env, the reference's TD(lambda) loop, 3 PPO epochs + Adam.  Returns seconds."""
    import numpy as np
    import torch
    from oracle import mappo as om
    from oracle import mappo_lstm as ol
    from oracle import spread as osp
    lstm = algo == "mappo_lstm"
    ippo = algo == "ippo"
    if lstm:
        actor, critic = ol.build_networks(seed)
    elif ippo:
        actor, critic = om.build_networks(seed, state_dim=21, critic_hidden=32)
    else:
        actor, critic = om.build_networks(seed)
    aopt, copt = om.make_optimizers(actor, critic)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    pos = rng.uniform(-1, 1, (B, 3, 2)); vel = np.zeros_like(pos); lm = rng.uniform(-1, 1, (B, 3, 2))
    eps = {k: [] for k in ("obs", "actions", "log_prob", "reward", "states")}
    ids = np.broadcast_to(np.eye(3), (B, 3, 3))
    avail1 = torch.ones(B, 3, 5, dtype=torch.bool)
    h = None
    for t in range(T_STEPS):
        raw = osp.observe_batched(pos, vel, lm)
        obs = np.concatenate([raw, ids], axis=-1)
        with torch.no_grad():
            if lstm:
                a, lp, h, _ = ol.rollout_act(actor, torch.from_numpy(obs).float(), h, avail1,
                                             om.draw_race_noise((B, 3, 5)))
            else:
                logits = om.actor_logits(actor, torch.from_numpy(obs).float())
                a, lp = om.race_sample(logits, om.draw_race_noise(logits.shape))
        pos, vel, rew = osp.step_batched(pos, vel, lm, a.numpy())
        eps["obs"].append(obs); eps["actions"].append(a); eps["log_prob"].append(lp)
        eps["reward"].append(rew[:, 0]); eps["states"].append(raw.reshape(B, 54))
    obs = torch.from_numpy(np.stack(eps["obs"], 1)).float()
    states = torch.from_numpy(np.stack(eps["states"], 1)).float()
    actions = torch.stack(eps["actions"], 1)
    logp = torch.stack(eps["log_prob"], 1)
    reward = torch.from_numpy(np.stack(eps["reward"], 1)).float()
    mask = torch.ones(B, T_STEPS, dtype=torch.bool)
    avail = torch.ones(B, T_STEPS, 3, 5, dtype=torch.bool)
    batch = (obs, actions, logp, reward, states, avail, torch.zeros(B, T_STEPS), mask)
    ret, adv = om.td_lambda_loop(critic, obs if ippo else states, reward, mask, 0.99, 0.95, 3)      # the reference's loop form
    if lstm:
        ol.ppo_update_tbptt(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, tbptt=10)
    else:
        om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, critic_on_obs=ippo)
    return time.perf_counter() - t0


def cpu_baseline(sample_envs=64, reps=2, algo="mappo"):
    import torch
    cpu_port_iteration(8, algo=algo)
    ts = [cpu_port_iteration(sample_envs, algo=algo) for _ in range(reps)]
    t = min(ts)
    return {"value": sample_envs * T_STEPS * N_AGENTS / t, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"oracle port (reference arithmetic on torch CPU; vectorised numpy env instead of one process "
                      f"per env), one full iteration on {sample_envs} of the envs, best of {reps}: {t:.2f} s"}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch
    from oracle import ref_loader
    # torchrun exports OMP_NUM_THREADS=1 to every rank: give the reference every host core it can use, as when run alone
    torch.set_num_threads(os.cpu_count() or 1)
    B = a.ref_envs
    K, W = a.steps, a.warmup
    # keep the CPU arm inside a few minutes: one reference iteration at B=32 takes ~1.5-3 s on 8 cores
    K = min(K, 5)
    W = min(W, 1)
    # The reference forks one interpreter + one pipe per env (MME:299-319): the 4096-env configuration cannot be launched
    # as written, so this arm runs --batch_size `B` and SAYS so -- its `config` names the envs it ran, not the GPU arm's.
    cfg = workload_config(a, 1, envs_per_gpu=B,
                          note=f"; reference arm: one worker process per env caps num_envs, the GPU arm runs "
                               f"{a.envs_per_gpu}/GPU")
    cfg["parallelism"] = f"{B} env worker processes + torch CPU ({os.cpu_count()} host cores)"
    cfg["l2"] = "n/a (CPU)"
    cfg["reference_envs"] = B
    cfg["gpu_arm_envs_per_gpu"] = a.envs_per_gpu
    per_step = B * T_STEPS * N_AGENTS
    script = SCRIPTS[a.algo]
    if ref_loader.reference_dir() is not None:
        import tempfile
        import torch.utils.tensorboard as tb

        stamps = []

        class StampWriter:
            """Stands in for the TensorBoard writer (a dependency, not reference code): the reference logs
            train/num_updates exactly once per iteration (MME:612), which gives per-iteration wall clock
            without touching the script."""

            def __init__(self, *a_, **k_):
                pass

            def add_scalar(self, tag, *a_, **k_):
                if tag == "train/num_updates":
                    stamps.append(time.perf_counter())

            def add_text(self, *a_, **k_):
                pass

            def close(self):
                pass

        tb.SummaryWriter = StampWriter
        iters = max(W, 1) + K
        argv = ["--env_type", "pz", "--env_name", "simple_spread_v3", "--batch_size", str(B),
                "--total_timesteps", str(B * T_STEPS * iters), "--eval_steps", "1000000000"]
        with tempfile.TemporaryDirectory() as tmp:
            ref_loader.run_script(argv, script=script, cwd=tmp)
        assert len(stamps) == iters, (len(stamps), iters)
        dt = (stamps[-1] - stamps[-1 - K]) / K
        kind = "reference"
        sample = (f"unmodified {script} from {ref_loader.reference_dir()} (runpy, tyro CLI, one worker process per env) "
                  f"on oracle/env_stub (numpy simple_spread; PettingZoo not installed), --batch_size {B}, "
                  f"{K} iterations after {max(W, 1)} warm-up: {dt:.2f} s per iteration")
    else:
        ts = [cpu_port_iteration(B, algo=a.algo) for _ in range(W + K)][W:]
        dt = sum(ts) / len(ts)
        kind = "port"
        sample = f"oracle port, full iteration on {B} envs (reference sources not present on this box)"
    val = per_step / dt
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": K,
           "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": kind, "sample": sample,
                            "torch_threads": torch.get_num_threads()},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


# ------------------------------------------------------------------------------------------------ our arm
_JSON_OUT = None


def emit(obj):
    f = _JSON_OUT or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def trace(msg):
    if os.environ.get("CMARL_BENCH_TRACE"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def reference_math(B=4096):
    """the reference's TD(lambda) loop (MME:484-504: two batch-of-1 critic calls per (episode, step))
    and its PPO epoch loop (MME:521-594, 3 epochs incl. backward, norms and Adam) on the SURVEY 8(d) synthetic batch at
    the REAL configuration size -- the part of the reference that does run at num_envs = 4096 (no env workers involved).
    Loop-form restatement in oracle/mappo.py (the reference's loops live inline under ``__main__``).  Seconds per phase."""
    import torch
    from oracle import mappo as om
    actor, critic = om.build_networks(1)
    batch = om.synthetic_batch(B, seed=1, actor=actor)
    t0 = time.perf_counter()
    ret, adv = om.td_lambda_loop(critic, batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    t1 = time.perf_counter()
    aopt, copt = om.make_optimizers(actor, critic)
    om.ppo_update(actor, critic, aopt, copt, batch, adv, ret, epochs=3, clip=0.2, ent_coef=0.001, flat=False)
    t2 = time.perf_counter()
    n = B * T_STEPS * N_AGENTS
    return {"envs": B, "td_lambda_loop_s": t1 - t0, "ppo_3_epochs_s": t2 - t1, "torch_threads": torch.get_num_threads(),
            "agent_env_steps_per_s_math_only": n / (t2 - t0),
            "what": "loop forms of MME:484-504 and MME:521-594 (oracle port) on the synthetic [4096, 25, ...] batch; "
                    "no rollout, no env workers"}


def run_b200(a):
    import torch
    from cleanmarl_b200.mappo import MAPPO, Args, ArgsRecurrent, init_distributed
    import cleanmarl_b200 as cm

    rank, world, local = init_distributed()
    if world != a.gpus and rank == 0:
        print(f"warning: --gpus {a.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    hbm_peak, bf16_peak, sm_max, peak_src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def make_trainer(algo, envs_per_gpu, **kw):
        args = (ArgsRecurrent if algo == "mappo_lstm" else Args)(batch_size=envs_per_gpu * world, seed=1,
                                                                 critic_hidden_dim=32 if algo == "ippo" else 64)
        return MAPPO(args, device_index=local, rank=rank, world_size=world, ippo=(algo == "ippo"), **kw)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)

    def timed(fn, steps, host_visible=False):
        """K steps, each bracketed by an event pair on the launching stream; L2 flushed in between (the flush is outside
        the event pairs).  host_visible=True additionally clocks every step on the host, from after the flush has
        drained to the return of ``fn`` (which ends with a stream synchronise): the latency a caller sees, without
        the flush kernel that only the benchmark needs.  Returns (device s [max over ranks], wall s, per-step ms)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        w0 = time.perf_counter()
        host = 0.0
        for s, e in evs:
            flush.zero_()
            if host_visible:
                torch.cuda.synchronize()
                h0 = time.perf_counter()
            s.record()
            fn()
            e.record()
            if host_visible:
                host += time.perf_counter() - h0
        barrier()
        wall = host if host_visible else time.perf_counter() - w0
        ms = [s.elapsed_time(e) for s, e in evs]
        return max_over_ranks(sum(ms)) / 1e3, wall, ms

    def kernel_breakdown(tr, algo, nb=5):
        """Per-kernel device times over a few eager iterations (library-internal event pairs) + algorithmic bytes / flops
        per launch (DESIGN.md "kernels": per env-step figures x B*T)."""
        eng = tr.engine
        graph_mode = tr.use_graph
        tr.use_graph = False                                     # event pairs live in the eager launch path
        eng.timing(True)
        for _ in range(nb):
            flush.zero_()
            tr.iteration()
        kt = eng.read_timing()
        eng.timing(False)
        tr.use_graph = graph_mode
        kernels = {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / nb, "ms_per_step": v[0] / nb}
                   for k, v in kt.items()}
        total = sum(v["ms_per_step"] for v in kernels.values())
        for v in kernels.values():
            v["share"] = v["ms_per_step"] / total
        bt = tr.B * T_STEPS
        ippo, lstm = algo == "ippo", algo == "mappo_lstm"
        alg = {
            "ppo_actor_chain": (bt * (216 + 12 + 12 + (12 if ippo else 4)), bt * 3 * 9792),
            "ppo_critic_chain": (bt * (216 + (12 if ippo else 4)), bt * (27072 if ippo else 38784)),
            "critic_values": (bt * (216 + (12 if ippo else 4)), bt * (3 * 3456 if ippo else 15232)),
            "rollout": (bt * (216 + 12 + 12 + 4), bt * 3 * 3712),
            "td_lambda_scan": (bt * 16 * (3 if ippo else 1), bt * 6),
        }
        if lstm:
            # per launch = one truncated-BPTT chunk (10, 10, 5 steps: B*T/3 env-steps on average); 41 856 FLOP per
            # agent-step = forward 13 952 (fc1 21x32, gates 2x96x32, fc2 32x5) + backward 2x that, recompute not counted
            nch = len(tr.chunks)
            alg["ppo_tbptt_chunk"] = (bt // nch * (216 + 12 + 12 + 4), bt // nch * 3 * 41856)
            alg["rollout"] = (bt * (216 + 12 + 12 + 4), bt * 3 * 13952)
            alg["ppo_critic_chain"] = (bt * (216 + 4), bt * 38784)
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12            # TFLOP/s FFMA at max clock
        for k, (by, fl) in alg.items():
            if k in kernels:
                t = kernels[k]["ms_per_launch"] * 1e-3
                kernels[k].update({"alg_bytes": by, "alg_flops": fl, "gbs": by / t / 1e9, "tflops": fl / t / 1e12,
                                   "hbm_frac": by / t / 1e9 / hbm_peak, "fp32_frac": fl / t / 1e12 / fp32_peak})
        return kernels, fp32_peak

    B = a.envs_per_gpu
    lstm = a.algo == "mappo_lstm"
    tr = make_trainer(a.algo, B)
    eng = tr.engine
    trace("trainer ready")
    # ---- device-resident throughput (value): the iteration as the trainer runs it (CUDA-graph replay) ----
    for _ in range(max(a.warmup, 3)):
        tr.iteration()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = eng.launches
    secs, wall, ms = timed(tr.iteration, a.steps)
    launches = eng.launches - l0
    if tr.use_graph and tr.launches_per_iteration:
        launches = tr.launches_per_iteration * a.steps           # replayed kernel nodes (the library counter only sees eager launches)
    per_step = B * world * T_STEPS * N_AGENTS
    value = per_step * a.steps / secs

    trace("value timed")
    # ---- end to end through the public API: every step the start states of all envs come from pinned HOST memory
    #      and the step's results (per-epoch statistics, per-env episode returns) are read back to the host ----
    env_h = torch.empty(18, B, dtype=torch.float64).uniform_(-1, 1).pin_memory()
    env_h[6:12] = 0
    res_h = torch.empty(tr.results.numel(), dtype=torch.uint8).pin_memory()

    def e2e_step():
        tr.env.copy_(env_h, non_blocking=True)                   # H2D straight into the trainer's env-state buffer
        tr.iteration(env_init=tr.env)
        tr.results_to_host(res_h)                                # one D2H copy: episode returns + per-epoch statistics
        torch.cuda.current_stream().synchronize()               # the caller reads the results every step

    for _ in range(3):
        e2e_step()
    e_secs, e_wall, _ = timed(e2e_step, a.steps, host_visible=True)
    # host-visible time: the results are read on the host every step, so the host clock around each step (input copy ..
    # results on the host) is the honest number; the L2 flush between steps is outside it
    e_wall = max_over_ranks(e_wall)
    h2d = env_h.numel() * 8
    d2h = res_h.numel()

    trace("e2e timed")
    # ---- the same with the categorical race noise supplied by the host as well (what the parity tests do; eager) ----
    noise_h = torch.empty(T_STEPS, N_AGENTS, N_ACT, B).exponential_(1).pin_memory()
    noise_d = torch.empty_like(noise_h, device=dev)
    env_d = torch.empty_like(env_h, device=dev)

    def e2e_noise_step():
        env_d.copy_(env_h, non_blocking=True)
        noise_d.copy_(noise_h, non_blocking=True)
        tr.iteration(env_init=env_d, noise=noise_d)
        tr.results_to_host(res_h)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_noise_step()
    n_steps2 = max(3, a.steps // 4)
    _, e2_wall, _ = timed(e2e_noise_step, n_steps2, host_visible=True)
    e2_wall = max_over_ranks(e2_wall)
    clk = clocks.stop() if rank == 0 else None
    trace("e2e with host noise timed")

    kernels, fp32_peak = kernel_breakdown(tr, a.algo)
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    tc = bool(eng.tensor_cores)
    chain = dom in ("ppo_actor_chain", "ppo_critic_chain", "critic_values")
    tf32_peak = bf16_peak / 2.0                                  # dense tf32 = half the dense bf16 rate on tcgen05
    if tc and chain:
        # the dominant kernel runs its GEMMs on tcgen05 (kind::tf32, 3 MMAs per product for fp32-level accuracy):
        # achieved = ALGORITHMIC flops / launch time; peak = measured dense bf16, so `frac` is a deliberately conservative
        # tensor-roofline fraction.  Beside it: the same against dense TF32 (half the bf16 rate) and the ISSUED MMA rate
        # (3 tf32 MMAs per algorithmic product) -- the figure ncu's sm__pipe_tensor_cycles_active corresponds to.
        ach = kernels[dom].get("tflops")
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": (ach / bf16_peak) if ach else None, "traffic": None, "peak_source": peak_src,
                    "tf32_dense_peak": tf32_peak, "frac_of_tf32_dense": (ach / tf32_peak) if ach else None,
                    "issued_tf32_tflops": 3 * ach if ach else None,
                    "issued_frac_of_tf32_dense": (3 * ach / tf32_peak) if ach else None,
                    "hbm_gbs": kernels[dom].get("gbs"), "hbm_frac": kernels[dom].get("hbm_frac"),
                    "fp32_ffma_equiv_frac": kernels[dom].get("fp32_frac"),
                    # the two chain kernels of an epoch take about the same time; the same figures for each of them
                    "chains": {k: {"ms_per_launch": kernels[k]["ms_per_launch"], "tflops": kernels[k].get("tflops"),
                                   "frac": kernels[k]["tflops"] / bf16_peak, "frac_of_tf32_dense": kernels[k]["tflops"] / tf32_peak,
                                   "hbm_frac": kernels[k].get("hbm_frac")}
                               for k in ("ppo_actor_chain", "ppo_critic_chain") if k in kernels and kernels[k].get("tflops")},
                    "note": "tiny contractions (K 24..64, N 32..64, M = 128 samples per tile): the kernel is bound by the "
                            "CUDA-core stages between its GEMMs and by MMA issue, not by tensor-pipe throughput (ncu "
                            "summaries under profiles/); issued_frac_of_tf32_dense is the issued-MMA rate (3 tf32 MMAs per product in "
                            "the activation GEMMs, 2 stacked ones in the weight-gradient GEMMs) that sm__pipe_tensor_cycles_active "
                            "follows; hbm_frac is the same launch against the HBM roofline; the HBM-bound GAE kernel is "
                            "`gae_roofline`"}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom].get("gbs"), "peak": hbm_peak, "unit": "GB/s",
                    "frac": kernels[dom].get("hbm_frac"), "traffic": None, "peak_source": peak_src,
                    "note": "the dominant kernel is an fp32 FFMA-bound fused MLP fwd+bwd (~160 FLOP/B); its "
                            "compute-side fraction is fp32_frac in `kernels`; the HBM-bound GAE kernel is `gae_roofline`",
                    "fp32_tflops": kernels[dom].get("tflops"), "fp32_peak_tflops": fp32_peak,
                    "fp32_frac": kernels[dom].get("fp32_frac")}

    trace("per-kernel timing done")
    traffic = ncu_traffic() if (B == 4096 and not lstm) else {}
    if dom in traffic:
        roofline["traffic"] = traffic[dom]
        roofline["traffic_source"] = f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, {traffic.get('_file')}"
        roofline["alg_bytes"] = kernels[dom].get("alg_bytes")
    # ---- stand-alone GAE scan at a size that leaves L2 (the metric BASELINE.json names) ----
    gae = None
    if rank == 0 and a.gae_envs > 0:
        Bg = a.gae_envs
        e2 = cm.Engine(cm.Shapes(n_envs=Bg), local)
        v = torch.randn(T_STEPS, 1, Bg, device=dev); r = torch.randn(T_STEPS, Bg, device=dev)
        R = torch.empty_like(v); A = torch.empty_like(v)
        for _ in range(3):
            e2.td_lambda(v, r, R, A, 0.99, 0.95)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for s, e in evs:
            flush.zero_()
            s.record(); e2.td_lambda(v, r, R, A, 0.99, 0.95); e.record()
        torch.cuda.synchronize()
        tg = statistics.median(s.elapsed_time(e) for s, e in evs) * 1e-3
        by = 16 * Bg * T_STEPS
        tr_bytes = ncu_traffic().get(f"td_lambda_scan@{Bg}")
        gae = {"kernel": "td_lambda_scan", "bound": "hbm", "envs": Bg, "alg_bytes": by, "ms": tg * 1e3,
               "achieved": by / tg / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": by / tg / 1e9 / hbm_peak,
               "traffic": tr_bytes,
               "frac_dram": (tr_bytes / tg / 1e9 / hbm_peak) if tr_bytes else None,
               "note": "16 B per env-step (r, V in; R, A out), inputs larger than L2 (419 MB), L2 flushed. `frac` counts the "
                       "algorithmic bytes; `frac_dram` counts the DRAM bytes ncu saw inside the launch (reads == algorithmic "
                       "reads; part of the writes is still in L2 when the kernel ends and drains after it)"}
        e2.close()
        del v, r, R, A

    # ---- other BASELINE configurations, short runs next to the headline (driver-visible) ----
    def short_run(algo, envs_per_gpu, steps):
        t2 = make_trainer(algo, envs_per_gpu)
        for _ in range(4):
            t2.iteration()
        sc, _, _ = timed(t2.iteration, steps)
        n = envs_per_gpu * world * T_STEPS * N_AGENTS
        out = {"config": workload_config(a, world, envs_per_gpu=envs_per_gpu, algo=algo), "value": n * steps / sc, "unit": UNIT,
               "ms_per_step": sc / steps * 1e3, "steps": steps,
               "launch_mode": "CUDA graph replay of the iteration" if t2.use_graph else "eager stream launches"}
        if world == 1:
            ks, _ = kernel_breakdown(t2, algo, nb=3)
            d = max(ks, key=lambda k: ks[k]["ms_per_step"])
            out["dominant_kernel"] = {"name": d, **ks[d]}
        del t2
        return out

    variants = None
    if world == 1 and a.algo == "mappo" and not a.no_variants:
        vs = max(20, min(50, a.steps))
        variants = {"ippo": short_run("ippo", 4096, vs),                 # BASELINE configs[2]
                    "mappo_lstm": short_run("mappo_lstm", 8192, vs)}    # BASELINE configs[3]
        trace("variants timed")
    config5 = None
    if world == 8 and a.algo == "mappo" and a.envs_per_gpu != 8192:
        config5 = short_run("mappo", 8192, a.steps)                      # BASELINE configs[4]: 65 536 envs over 8 GPUs
        trace("config5 timed")

    # ---- N ranks == 1 rank on the same envs (outside every timed region) ----
    mgpu_check = None
    if world > 1:
        try:
            mgpu_check = multi_gpu_check(MAPPO, Args, rank, world, local)
        except Exception as e:       # noqa: BLE001
            mgpu_check = {"error": f"{type(e).__name__}: {e}"}
        trace("mgpu check done")

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(algo=a.algo, sample_envs=32 if lstm else 64)
        if a.algo == "mappo":
            cpu["reference_math_4096"] = reference_math(4096)

    if rank == 0:
        cfg = workload_config(a, world)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
               "ms_per_step": secs / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "gemm": "3xTF32 on tcgen05, fp32 accumulate in TMEM" if tc else "fp32 FFMA",
               "data": "synthetic", "config": cfg,
               "e2e": {"value": per_step * a.steps / e_wall, "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": e_wall / a.steps * 1e3,
                       "device_ms_per_step": e_secs / a.steps * 1e3,
                       "inputs": "start states f64 [18][B] of every env from pinned host memory every step (race noise "
                                 "drawn on the device, the trainer's default); D2H: per-epoch stats + per-env episode return",
                       "with_host_noise": {"value": per_step * n_steps2 / e2_wall, "ms_per_step": e2_wall / n_steps2 * 1e3,
                                           "h2d_bytes_per_step": h2d + noise_h.numel() * 4,
                                           "note": "Exp(1) race noise f32 [T][N][A][B] also copied from the host (eager launches)"}},
               "gpu_launches": launches,
               "launch_mode": "CUDA graph replay of the iteration" if tr.use_graph else "eager stream launches",
               "clocks": clk, "roofline": roofline, "gae_roofline": gae,
               "kernels": kernels, "cpu_baseline": cpu, "wall_ms_per_step": wall / a.steps * 1e3}
        if variants is not None:
            out["variants"] = variants
        if config5 is not None:
            out["config5"] = config5
        if mgpu_check is not None:
            out["mgpu_check"] = mgpu_check
        emit(out)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def multi_gpu_check(MAPPO, Args, rank, world, local, envs_per_rank=256, iters=3):
    """N ranks (envs sharded, gradient sums exchanged) against ONE rank on the same envs, start states and race noise
    for `iters` iterations: replicas bit-identical to each other, parameters within fp32 reassociation of the single-rank
    run.  (The driver's pytest box has one GPU, so this is where the equivalence reaches a driver record.)"""
    import torch
    dev = torch.device("cuda", local)
    B = envs_per_rank * world
    g = torch.Generator().manual_seed(11)
    env = torch.zeros(18, B, dtype=torch.float64)
    env[0:6] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    env[12:18] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    noise = torch.empty(T_STEPS, N_AGENTS, N_ACT, B).exponential_(1, generator=g)
    sl = slice(rank * envs_per_rank, (rank + 1) * envs_per_rank)
    trn = MAPPO(Args(batch_size=B, seed=3), device_index=local, rank=rank, world_size=world)
    e_n, q_n = env[:, sl].contiguous().to(dev), noise[..., sl].contiguous().to(dev)
    for _ in range(iters):
        trn.iteration(e_n, q_n)
    torch.cuda.synchronize()
    gathered = [torch.empty_like(trn.net.flat) for _ in range(world)]
    torch.distributed.all_gather(gathered, trn.net.flat)
    out = None
    if rank == 0:
        tr1 = MAPPO(Args(batch_size=B, seed=3), device_index=local, rank=0, world_size=1)
        e_1, q_1 = env.to(dev), noise.to(dev)
        for _ in range(iters):
            tr1.iteration(e_1, q_1)
        torch.cuda.synchronize()
        dp = (trn.net.flat - tr1.net.flat).abs()
        diff = dp.max().item()
        tol = 5e-6
        outside = dp > tol
        n_out = int(outside.sum().item())
        # the gradient (sum over the batch, last epoch) of the parameters outside the tolerance, relative to the largest one
        P = dp.numel()
        gabs = tr1.grads[:P].abs()
        out_grad_rel = (gabs[outside].max() / gabs.max()).item() if n_out else 0.0
        moved = (tr1.net.flat - MAPPO(Args(batch_size=B, seed=3), device_index=local).net.flat).abs().max().item()
        sd = (trn.epoch_stats - tr1.epoch_stats).abs().max().item() / max(tr1.epoch_stats.abs().max().item(), 1e-30)
        same = all(torch.equal(gathered[0], x) for x in gathered)
        frac_in = 1.0 - n_out / P
        out = {"replicas_identical": same, "max_param_diff": diff, "frac_params_within_tolerance": frac_in,
               "params_outside_tolerance": n_out, "their_max_grad_rel_to_largest_grad": out_grad_rel,
               "max_param_change_over_run": moved, "max_rel_stat_diff": sd, "comm": trn.comm, "ranks": world,
               "envs": B, "iterations": iters, "adam_steps": iters * 3, "step_counter_equal": trn.step == tr1.step,
               "tolerance": tol,
               "criterion": "replicas bit-identical, statistics within 1e-6 relative, >= 99.5 % of the parameters within the "
                            "tolerance; every parameter outside it has a gradient below 1e-3 of the largest gradient (the "
                            "regrouping moves gradients by 2-3e-6 of a tensor's largest one, i.e. by >= 0.3 % of such a gradient, "
                            "which Adam's g / sqrt(v) carries into the step) and deviates by less than 5 % of the largest "
                            "parameter change of the run",
               "note": "fp32 reassociation of the rank sums through 9 Adam steps alone: 3e-8.  The shards (256 envs: one tile "
                       "per CTA) group the tiles of the chain kernels' persistent TMEM accumulators differently from the single "
                       "rank (256 x ranks envs: some CTAs accumulate several tiles before a flush): gradients differ by 2-3e-6 "
                       "of a tensor's max, and the handful of parameters whose gradient is at that round-off level (see "
                       "their_max_grad_rel_to_largest_grad) are moved by Adam's g / sqrt(v) in a direction rounding decides "
                       "(DESIGN.md 5; CMARL_TC_FLUSH=1 removes the regrouping).  Which parameters those are depends on the "
                       "trajectories: with the round-2 rollout kernel 34 of 9 686 at 2 ranks (max 1.0e-4), before it none",
               "ok": bool(same and trn.step == tr1.step and sd < 1e-6 and frac_in >= 0.995 and out_grad_rel < 1e-3
                          and diff < 0.05 * moved)}
    torch.distributed.barrier()
    return out


def main():
    a = parse()
    # stdout carries exactly ONE line (the JSON): everything else that writes to fd 1 -- NCCL's version banner, library
    # chatter of the reference arm -- is sent to stderr for the whole run
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
