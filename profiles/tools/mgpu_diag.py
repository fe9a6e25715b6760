"""2 GPUs vs 1 GPU parameter differences of tests/test_gpu_parity.py::test_two_gpus_nccl_equal_one, per flush period of the
tcgen05 chains (CMARL_TC_FLUSH) -- which parameters differ, by how much, and how large their gradients are."""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "tests"))
sys.path.insert(0, str(REPO))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "flags"
    B = 1024
    for flush in (sys.argv[2:] or ["4", "1"]):
        os.environ["CMARL_TC_FLUSH"] = flush
        tmp = Path(tempfile.mkdtemp())
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29533", str(REPO / "tests" / "mgpu_worker.py"), str(tmp), str(B), mode]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, CMARL_COMM="p2p"))
        if r.returncode:
            print(r.stderr[-2000:]); return
        two = torch.load(tmp / "mgpu.pt")
        code = (f"import sys; sys.path.insert(0, {str(REPO / 'tests')!r}); sys.path.insert(0, {str(REPO)!r}); import torch, mgpu_worker;"
                f"kw = {{'flags': {{'normalize_advantage': True, 'clip_gradients': 0.5}}, 'recurrent': {{'recurrent': True}}}}.get({mode!r}, {{}});"
                f"one = mgpu_worker.run({B}, 0, 1, 0, **kw); torch.save({{'params': one.net.flat.cpu(), 'grads': one.grads.cpu(), 'stats': one.epoch_stats.cpu()}}, {str(tmp / 'one.pt')!r})")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
        if r.returncode:
            print(r.stderr[-2000:]); return
        one = torch.load(tmp / "one.pt")
        dp = (two["params"] - one["params"]).abs()
        g = one["grads"][:dp.numel()].abs()
        worst = torch.argsort(dp, descending=True)[:6]
        print(f"mode {mode} flush {flush}: max dp {dp.max():.3e}  n(dp > 2e-6) {(dp > 2e-6).sum().item()} of {dp.numel()}  "
              f"n(dp > 1e-6) {(dp > 1e-6).sum().item()}  stats rel {((two['stats'] - one['stats']).abs().max() / one['stats'].abs().max()):.2e}")
        for i in worst.tolist():
            print(f"   param {i}: dp {dp[i]:.3e}  |last-epoch grad sum| {g[i]:.3e}  (max |grad| {g.max():.3e})")


if __name__ == "__main__":
    main()
