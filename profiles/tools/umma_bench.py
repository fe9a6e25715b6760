"""Issue-rate micro-benchmark of tcgen05.mma.kind::tf32 (uses libcmarl_umma_probe.so): cycles per MMA for
back-to-back MMAs on 1/2/4 accumulators, A from TMEM or shared memory, M = 64/128, small N."""
import ctypes as C
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parents[2]
lib = C.CDLL(str(REPO / "cleanmarl_b200" / "libcmarl_umma_probe.so"))
class BenchArgs(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("m", "n", "a_tmem", "count", "nacc", "lbo", "sbo", "kstep")] + [("out", C.c_void_p)]
lib.cmarl_umma_bench.argtypes = [C.POINTER(BenchArgs), C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("M N A nacc lbo  cycles/MMA(total) cycles/MMA(issue)")
for (m, n, a_tmem, lbo, sbo, kstep) in ((128, 64, 1, 128, 2048, 256), (128, 32, 1, 128, 2048, 256), (128, 64, 0, 128, 2048, 256),
                                         (64, 32, 0, 144, 4608, 288), (64, 40, 0, 144, 4608, 288), (64, 32, 0, 128, 4096, 256),
                                         (64, 64, 0, 144, 4608, 288), (128, 128, 1, 128, 2048, 256),
                                         # rollout_tc_kernel's layer 2 (SS, K-major A and B images): H = 32, H = 64, padded variants
                                         (128, 32, 0, 128, 1024, 256), (128, 64, 0, 128, 2048, 256), (128, 32, 0, 144, 4608, 288),
                                         (128, 32, 0, 144, 1152, 288), (128, 32, 0, 128, 1040, 256), (128, 32, 0, 128, 1152, 256)):
    for nacc in (1, 2, 4):
        if nacc * n > 384:
            continue
        for rep in range(2):
            a = BenchArgs(m, n, a_tmem, 96, nacc, lbo, sbo, kstep, out.data_ptr())
            rc = lib.cmarl_umma_bench(C.byref(a), None)
            torch.cuda.synchronize()
        t = out.cpu().tolist()
        print(f"{m:4d} {n:4d} {'T' if a_tmem else 'S'} {nacc} {lbo:4d}   {t[0]/96:8.1f} {t[1]/96:8.1f}", flush=True)
