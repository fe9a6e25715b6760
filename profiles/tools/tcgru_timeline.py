"""clock64 timeline of one step of tc_gru_fwd_kernel / tc_gru_bwd_kernel (CTA 0, compute thread 0, second step of its first
tile): python profiles/tools/tcgru_timeline.py [B] [tbptt]"""
import ctypes as C
import os
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "profiles" / "tools"))

import cleanmarl_b200 as cm  # noqa: E402
from cleanmarl_b200 import _lib  # noqa: E402
from cleanmarl_b200 import engine as E  # noqa: E402
from oracle import mappo as om  # noqa: E402
from oracle import mappo_lstm as ol  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    tb = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    os.environ["CMARL_TBPTT"] = "tc"
    lib = _lib.load()
    raw = C.CDLL(str(REPO / "cleanmarl_b200" / "libcmarl_b200.so"))
    actor, critic = ol.build_networks(1)
    batch = list(om.synthetic_batch(B, seed=2))
    batch[2] = ol.synthetic_old_logp(actor, batch, seed=3)
    adv = torch.randn(B, 25, 1).expand(B, 25, 3).contiguous()
    eng = cm.Engine(cm.Shapes(n_envs=B, actor_recurrent=True), device=0)
    dev = eng.device
    d = E.to_device_layout(tuple(batch), dev)
    adv_d = E.heads_to_device(adv, eng.n_heads, dev)
    flat = torch.cat([actor.flat_params(), critic.flat_params()]).to(dev).contiguous()
    na = eng.n_actor
    h_seq, stash, ga = eng.alloc_h_seq(), eng.alloc_gate_stash(), eng.empty(na + 8)
    kw = dict(state=d["state"], actions=d["actions"], logp_old=d["logp"], adv=adv_d, clip=0.2, ent_coef=0.001, stash=stash)
    for _ in range(2):
        eng.tbptt_chunk_grads(flat[:na], ga, h_seq, 0, tb, **kw)
    torch.cuda.synchronize()
    raw.cmarl_debug_tcgru_timeline(1, None)
    eng.tbptt_chunk_grads(flat[:na], ga, h_seq, 0, tb, **kw)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 32)()
    raw.cmarl_debug_tcgru_timeline(0, buf)
    v = list(buf)
    names_b = ["step start", "D_3(prev)", "operands, x1/h staged, da -> TMEM / A image", "R_1 published", "(x rows requested)", "D_1 (round rz done)",
               "n/hn staged, R_2", "D_2 (dx1|dh, round n)", "dx1/X staged, R_3", "(tile end: flush)"]
    print(f"backward step (B {B}, tbptt {tb}), cycles:")
    for i in range(1, 10):
        print(f"  {names_b[i]:34s} +{v[i] - v[i - 1]:7d}   (t = {v[i] - v[0]})")
    print(f"backward kernel (CTA 0): set-up {v[11] - v[10]}, first flush {v[13] - v[12]}, all tiles done at {v[14] - v[10]}, end at {v[15] - v[10]}")
    names_f = ["step start", "D_F1 (fc1 done)", "x1 -> A, R_X1", "X(t+1) staged", "D_G (gates done)", "epilogue, R_H"]
    print("forward step, cycles:")
    for i in range(1, 6):
        print(f"  {names_f[i]:34s} +{v[16 + i] - v[16 + i - 1]:7d}   (t = {v[16 + i] - v[16]})")


if __name__ == "__main__":
    main()
