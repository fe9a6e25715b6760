import sys
sys.path.insert(0, "/root/repo")
import torch
from cleanmarl_b200.mappo import MAPPO, Args
def run(mode, B=640):
    tr = MAPPO(Args(batch_size=B, seed=7), use_graph=False)
    eng, buf, a = tr.engine, tr.buf, tr.args
    tr.collect(); tr.advantages(); torch.cuda.synchronize()
    outs = []
    for ep in range(3):
        if mode == "sync_before": torch.cuda.synchronize()
        eng.set_launch_chaining(mode != "none")
        eng.ppo_epoch_grads(tr.net.flat, tr.grads, state=buf["state"], actions=buf["actions"], logp_old=buf["logp"], adv=buf["adv"], returns=buf["returns"])
        eng.set_launch_chaining(False)
        if mode != "nosync_after": torch.cuda.synchronize()
        outs.append(tr.grads.clone())
        eng.clip_adam_step(tr.net.flat, tr.grads, tr.exp_avg, tr.exp_avg_sq, step_dev=tr.adam_step, stats_out=tr.epoch_stats[ep])
    torch.cuda.synchronize()
    return outs
ref = run("none")
for mode in ("none", "sync_before", "chained", "nosync_after"):
    o = run(mode)
    for ep in range(3):
        d = (o[ep] - ref[ep]).abs()
        print(f"{mode:13s} epoch {ep}: actor grads {float(d[:1925].max()):.3e} critic grads {float(d[1925:9670].max()):.3e} stats {[round(float(x),4) for x in d[9670:]]}")
