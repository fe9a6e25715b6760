"""Which launch pairs make chained (programmatic dependent) launches differ from serialised ones?  Runs whole iterations
with launch chaining switched on around selected kernels only and compares the final parameters with a plain run."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from cleanmarl_b200.mappo import MAPPO, Args

def run(mode, B=640, iters=2):
    tr = MAPPO(Args(batch_size=B, seed=7), use_graph=False)
    eng, buf, a = tr.engine, tr.buf, tr.args
    on = lambda *names: eng.set_launch_chaining(mode in names or mode == "all")
    for _ in range(iters):
        eng.set_launch_chaining(False)
        eng.env_reset(tr.env, tr.rng_key, tr.episode)
        on("rollout")
        eng.rollout(tr.net.actor, tr.env, buf["state"], buf["actions"], buf["logp"], buf["reward"], ep_return=buf["ep_return"],
                    seed=tr.rng_key, episode=tr.episode)
        tr.episode += 1
        on("critic_values")
        eng.critic_values(tr.net.critic, buf["values"], state=buf["state"])
        on("td")
        eng.td_lambda(buf["values"], buf["reward"], buf["returns"], buf["adv"], a.gamma, a.td_lambda)
        for ep in range(3):
            on("grads")        # actor chain, critic chain, reduce
            eng.ppo_epoch_grads(tr.net.flat, tr.grads, state=buf["state"], actions=buf["actions"], logp_old=buf["logp"],
                                adv=buf["adv"], returns=buf["returns"])
            on("adam")
            eng.clip_adam_step(tr.net.flat, tr.grads, tr.exp_avg, tr.exp_avg_sq, step_dev=tr.adam_step, stats_out=tr.epoch_stats[ep])
        eng.set_launch_chaining(False)
    torch.cuda.synchronize()
    return tr.net.flat.clone(), tr.epoch_stats.clone()

ref = run("none")
for mode in ("none", "rollout", "critic_values", "td", "grads", "adam", "all"):
    p, s = run(mode)
    print(f"{mode:14s} max param diff vs plain {float((p - ref[0]).abs().max()):.3e}   max stat diff {float((s - ref[1]).abs().max()):.3e}")
