"""Sanity run: does the device path learn?  Mean episode return of the sampling policy on device-resident
simple_spread_v3 over training iterations, for the three trainers (reference hyper-parameters, B = 4096).

usage: python profiles/tools/learning_curve.py [iterations] > profiles/learning_curve_rNN.md
"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from cleanmarl_b200.mappo import MAPPO, Args, ArgsRecurrent  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
every = max(1, iters // 15)
print(f"# learning curves (B = 4096 envs, T = 25, reference defaults: lr 8e-4, 3 epochs, clip 0.2, ent 0.001), {iters} iterations\n")
for name, args, ippo in (("mappo_multienvs", Args(batch_size=4096, seed=1), False),
                         ("ippo_multienvs", Args(batch_size=4096, seed=1, critic_hidden_dim=32), True),
                         ("mappo_lstm_multienvs", ArgsRecurrent(batch_size=4096, seed=1), False)):
    tr = MAPPO(args, ippo=ippo)
    rows = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(iters):
        tr.iteration()
        if i % every == 0 or i == iters - 1:
            rows.append((i, tr.step, tr.rollout_scalars()["ep_reward"], tr.train_scalars()))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"## {name}: {iters} iterations = {tr.step * 3:,} agent-env-steps in {dt:.1f} s wall (incl. the read-backs below)\n")
    print("| iteration | env steps | mean episode return | actor loss | critic loss | entropy | kl |")
    print("|---|---|---|---|---|---|---|")
    for i, step, r, sc in rows:
        print(f"| {i} | {step:,} | {r:.3f} | {sc['actor_loss']:.4f} | {sc['critic_loss']:.3f} | {sc['entropy']:.4f} | {sc['kl_divergence']:.5f} |")
    print()
