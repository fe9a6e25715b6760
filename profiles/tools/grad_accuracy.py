"""Gradient accuracy of cmarl_ppo_epoch_grads against an fp64 oracle, per parameter tensor, over batch sizes / GEMM
paths / persistent-grid caps (profiles/grad_accuracy_r2.md).  Usage: python profiles/tools/grad_accuracy.py [B ...]"""
import copy
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch  # noqa: E402
import cleanmarl_b200 as cm  # noqa: E402
from cleanmarl_b200 import engine as E  # noqa: E402
from oracle import mappo as om  # noqa: E402


def oracle64(actor, critic, batch, adv, ret, ippo):
    a, c = copy.deepcopy(actor).double(), copy.deepcopy(critic).double()
    obs, actions, logp, reward, states, avail, done, mask = batch
    out = om.ppo_epoch_flat(a, c, obs.double(), actions, logp.double(), (obs if ippo else states).double(), avail, mask,
                            adv.double(), ret.double(), 0.2, 0.001)
    out.actor_loss.backward(); out.critic_loss.backward()
    return [p.grad for p in a.parameters()] + [p.grad for p in c.parameters()]


def run(B, tc, ippo=False, seed=1):
    actor, critic = (om.build_networks(1, state_dim=21, critic_hidden=32) if ippo else om.build_networks(1))
    batch = list(om.synthetic_batch(B, seed=seed, actor=actor))
    ret, adv = om.td_lambda_batched(critic, batch[0] if ippo else batch[4], batch[3], batch[7], 0.99, 0.95, 3)
    if os.environ.get("GA_MASK_KINKS"):      # drop the (b, t) pairs with a hidden pre-activation within rounding of zero
        sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "tests"))
        from test_gpu_parity import _relu_kink_samples
        batch[7] = batch[7] & ~_relu_kink_samples(actor, critic, batch[0], batch[0] if ippo else batch[4])
    batch = tuple(batch)
    g64 = oracle64(actor, critic, batch, adv, ret, ippo)
    eng = cm.Engine(cm.Shapes(n_envs=B, critic_on_obs=ippo, critic_hidden=32 if ippo else 64), device=0, tensor_cores=tc)
    dev = eng.device
    d = E.to_device_layout(batch, dev, with_obs=False)
    params = torch.cat([actor.flat_params(), critic.flat_params()]).to(dev)
    grads = eng.empty(eng.n_params + 8)
    eng.ppo_epoch_grads(params, grads, state=d["state"], actions=d["actions"], logp_old=d["logp"],
                        adv=E.heads_to_device(adv, eng.n_heads, dev), returns=E.heads_to_device(ret, eng.n_heads, dev),
                        mask=d["mask"])
    g = grads.cpu().double()
    n = g[eng.n_params + 5].item()
    off, out = 0, []
    for ref in g64:
        k = ref.numel()
        a = g[off:off + k].reshape(ref.shape) / n
        out.append((a - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30))
        off += k
    eng.close()
    return out


if __name__ == "__main__":
    Bs = [int(x) for x in sys.argv[1:]] or [1024, 4096]
    names = ["aW1", "ab1", "aW2", "ab2", "aW3", "ab3", "cW1", "cb1", "cW2", "cb2", "cW3", "cb3"]
    print(f"grid cap: {os.environ.get('CMARL_DEBUG_GRID_CAP', '-')}, relu-kink samples masked: {bool(os.environ.get('GA_MASK_KINKS'))}"
          f"   max |g - g64| / max |g64| per tensor")
    print("| B | path | " + " | ".join(names) + " |")
    print("|---|---|" + "---|" * len(names))
    for B in Bs:
        for tc in ((True,) if os.environ.get("GA_TC_ONLY") else (False, True)):
            for ippo in ((False,) if os.environ.get("GA_TC_ONLY") else (False, True)):
                r = run(B, tc, ippo)
                print(f"| {B}{' ippo' if ippo else ''} | {'tc' if tc else 'ffma'} | " + " | ".join(f"{x:.1e}" for x in r) + " |", flush=True)
