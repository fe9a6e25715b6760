import ctypes as C, sys, torch
sys.path.insert(0, "/root/repo")
from cleanmarl_b200 import _lib
from cleanmarl_b200.mappo import MAPPO, Args
lib = _lib.load()
tr = MAPPO(Args(batch_size=4096, seed=1))
for _ in range(3): tr.collect()
torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
lib.cmarl_debug_rollout_timeline(buf)
v = list(buf)
# rollout_tc_kernel (default): 2 = L1 + A images published, 1 = buffer stores issued, 3 = layer-2 MMAs complete, 4 = epilogue done;
# CMARL_ROLLOUT=ffma: the CUDA-core kernel's stages
names = {0: "S: step start", 1: "S: obs stores done", 2: "S: L1 done", 3: "S: after agent bar 1 / MMAs done", 4: "S: L2+L3 done", 5: "S: after agent bar 2",
         6: "S: sampling done", 7: "S: step end", 8: "P: step start", 9: "P: before B1", 10: "P: after B1", 11: "P: pair force done / I: operands complete",
         12: "P: after physics bar", 13: "P: integrate done", 14: "P: after B3", 15: "P: dist done / I: MMAs issued"}
ev = sorted((v[k], k) for k in names if v[k] > 0)
t0 = ev[0][0]; prev = t0
for t, k in ev:
    print(f"{t - t0:8d} (+{t - prev:6d})  {names[k]}"); prev = t

try:
    buf2 = (C.c_longlong * 104)()
    lib.cmarl_debug_rollout_timeline_mma(buf2)
    m = [x for x in buf2[:32] if x > 0]
    if m:
        print("MMA issue stamps (cycles after the first):", [x - m[0] for x in m], "first at", m[0] - t0)
        print("arrival at the barrier behind the sampling, warps 0-14:", [x - t0 for x in buf2[32:47]])
        print("arrival at the barrier behind the integration, warps 0-14:", [x - t0 for x in buf2[48:63]])
        k = list(buf2[64:70])
        print("kernel level (cycles after entry): before the dependency wait %d, predecessor complete %d, set-up done %d, 25 steps done %d, exit %d"
              % tuple(x - k[0] for x in k[1:]))
        st = [k[3]] + [x for x in buf2[72:104] if x > 0]
        print("cycles of every step (warp 0):", [b - a for a, b in zip(st, st[1:])])
except AttributeError:
    pass
