"""tc_gru.cu against gru.cu (the fp32 FFMA kernel): hidden states, gate stash and per-chunk gradient sums of one epoch for
every mix of the two (CMARL_TBPTT = ffma | tcfwd | tcbwd | tc), plus timings of a chunk.  Usage:
    python profiles/tools/tcgru_check.py [B] [tbptt] [modes...]"""
import os
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO))

import cleanmarl_b200 as cm  # noqa: E402
from cleanmarl_b200 import engine as E  # noqa: E402
from cleanmarl_b200.mappo import tbptt_chunks  # noqa: E402
from oracle import mappo as om  # noqa: E402
from oracle import mappo_lstm as ol  # noqa: E402


def run(mode, eng, flat, d, adv_d, chunks, use_mask=True, time_it=False):
    os.environ["CMARL_TBPTT"] = mode
    dev = eng.device
    na = eng.n_actor
    h_seq = eng.alloc_h_seq(); h_seq.zero_()
    stash = eng.alloc_gate_stash(); stash.zero_()
    ga = eng.empty(na + 8)
    out = []
    for (t0, t1) in chunks:
        eng.tbptt_chunk_grads(flat[:na], ga, h_seq, t0, t1, state=d["state"], actions=d["actions"], logp_old=d["logp"],
                              adv=adv_d, mask=d["mask"] if use_mask else None, avail=d["avail"], clip=0.2, ent_coef=0.001,
                              stash=stash)
        out.append(ga.clone())
    torch.cuda.synchronize()
    ms = None
    if time_it:
        t0, t1 = chunks[0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            eng.tbptt_chunk_grads(flat[:na], ga, h_seq, t0, t1, state=d["state"], actions=d["actions"], logp_old=d["logp"],
                                  adv=adv_d, mask=d["mask"] if use_mask else None, avail=d["avail"], clip=0.2,
                                  ent_coef=0.001, stash=stash)
        ev[0].record()
        for _ in range(10):
            eng.tbptt_chunk_grads(flat[:na], ga, h_seq, t0, t1, state=d["state"], actions=d["actions"], logp_old=d["logp"],
                                  adv=adv_d, mask=d["mask"] if use_mask else None, avail=d["avail"], clip=0.2,
                                  ent_coef=0.001, stash=stash)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
    return h_seq, stash, out, ms


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    tb = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    modes = sys.argv[3:] or ["tcfwd", "tc"]
    gen = torch.Generator().manual_seed(B)
    actor, critic = ol.build_networks(B)
    batch = list(om.synthetic_batch(B, seed=B + 1))
    lengths = torch.where(torch.rand(B, generator=gen) < 0.4, torch.randint(1, 26, (B,), generator=gen), torch.full((B,), 25))
    batch[7] = torch.arange(25)[None, :] < lengths[:, None]
    avail = torch.ones(B, 25, 3, 5, dtype=torch.bool)
    avail[torch.rand(B, 25, 3, 5, generator=gen) < 0.1] = False
    avail[..., 0] = True
    batch[5] = avail
    probs = avail.float() / avail.float().sum(-1, keepdim=True)
    batch[1] = torch.multinomial(probs.reshape(-1, 5), 1, generator=gen).reshape(B, 25, 3)
    batch[2] = ol.synthetic_old_logp(actor, batch, seed=3)
    adv = torch.randn(B, 25, 1, generator=gen).expand(B, 25, 3).contiguous()
    eng = cm.Engine(cm.Shapes(n_envs=B, actor_recurrent=True), device=0)
    dev = eng.device
    d = E.to_device_layout(tuple(batch), dev)
    adv_d = E.heads_to_device(adv, eng.n_heads, dev)
    flat = torch.cat([actor.flat_params(), critic.flat_params()]).to(dev).contiguous()
    chunks = tbptt_chunks(25, tb)
    big = B >= 2048
    ref = run("ffma", eng, flat, d, adv_d, chunks, time_it=big)
    print(f"B {B} tbptt {tb}: ffma chunk {ref[3]} ms", flush=True)
    L = eng.n_actor
    names = [("w1", 32 * 21), ("b1", 32), ("wih", 3072), ("whh", 3072), ("bih", 96), ("bhh", 96), ("w2", 160), ("b2", 5)]
    for mode in modes:
        try:
            got = run(mode, eng, flat, d, adv_d, chunks, time_it=big)
        except Exception as e:  # noqa: BLE001
            print(f"mode {mode}: FAILED {e}", flush=True)
            return
        dh = (got[0] - ref[0]).abs().max().item()
        ds = (got[1] - ref[1]).abs().max().item()
        print(f"mode {mode}: chunk {got[3]} ms  max |dh_seq| {dh:.3e}  max |dstash| {ds:.3e}", flush=True)
        for ci, (gr, gg) in enumerate(zip(ref[2], got[2])):
            o = 0
            line = []
            for nm, n in names:
                r_, g_ = gr[o:o + n], gg[o:o + n]
                line.append(f"{nm} {((g_ - r_).abs().max() / (r_.abs().max() + 1e-30)).item():.1e}")
                o += n
            st = ((gg[L:] - gr[L:]).abs().max() / (gr[L:].abs().max() + 1e-30)).item()
            print(f"   chunk {ci}: " + "  ".join(line) + f"  stats {st:.1e}  all {((gg[:L] - gr[:L]).abs().max() / gr[:L].abs().max()).item():.2e}", flush=True)
    os.environ.pop("CMARL_TBPTT", None)


if __name__ == "__main__":
    main()
