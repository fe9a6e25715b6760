"""Per-phase cycle counts of tbptt_chunk_kernel (CTA 0, first tile, step 1 of each pass) from its clock64 stamps."""
import ctypes as C, sys, torch
sys.path.insert(0, "/root/repo")
from cleanmarl_b200 import _lib
from cleanmarl_b200.mappo import MAPPO, ArgsRecurrent
lib = _lib.load()
lib.cmarl_debug_gru_timeline.argtypes = [C.POINTER(C.c_longlong)]
for stash in ("1", "0"):
    import os
    os.environ["CMARL_GATE_STASH"] = stash
    tr = MAPPO(ArgsRecurrent(batch_size=8192, seed=1), use_graph=False)
    for _ in range(2):
        tr.iteration()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    lib.cmarl_debug_gru_timeline(buf)
    v = list(buf)
    names = {0: "F: step start", 1: "F: fc1 done", 2: "F: gates done", 3: "F: step end (h stored)", 4: "B: step start",
             5: "B: inputs ready (stash loaded / gates recomputed)", 6: "B: head done", 7: "B: (a)+(b) gate grads done",
             8: "B: (c) dWih/dWhh done", 9: "B: (d) dx1/dh done", 10: "B: (e) dW1 done"}
    print(f"== gate stash {stash}")
    for grp in ((0, 1, 2, 3), (4, 5, 6, 7, 8, 9, 10)):
        t0 = v[grp[0]]
        prev = t0
        for k in grp:
            print(f"{v[k] - t0:8d} (+{v[k] - prev:6d})  {names[k]}")
            prev = v[k]
