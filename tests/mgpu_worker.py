"""torchrun worker of tests/test_gpu_parity.py::test_two_gpus_nccl_equal_one: one rank per GPU, envs sharded, the
same global start states / race noise as the single-GPU run; rank 0 saves the replicated parameters."""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from cleanmarl_b200.mappo import MAPPO, Args, ArgsRecurrent, init_distributed  # noqa: E402


def inputs(B, seed=11):
    g = torch.Generator().manual_seed(seed)
    env = torch.zeros(18, B, dtype=torch.float64)
    env[0:6] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    env[12:18] = torch.rand(6, B, generator=g, dtype=torch.float64) * 2 - 1
    noise = torch.empty(25, 3, 5, B).exponential_(1, generator=g)
    return env, noise


def run(B, rank, world, local, iters=2, recurrent=False, **flags):
    tr = MAPPO((ArgsRecurrent if recurrent else Args)(batch_size=B, seed=3, **flags), device_index=local, rank=rank,
               world_size=world)
    env, noise = inputs(B)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    dev = tr.engine.device
    for _ in range(iters):
        tr.iteration(env[:, sl].contiguous().to(dev), noise[..., sl].contiguous().to(dev))
    torch.cuda.synchronize()
    return tr


if __name__ == "__main__":
    out, B = Path(sys.argv[1]), int(sys.argv[2])
    mode = sys.argv[3] if len(sys.argv) > 3 else "plain"
    flags = {"flags": {"normalize_advantage": True, "clip_gradients": 0.5}, "recurrent": {"recurrent": True}}.get(mode, {})
    rank, world, local = init_distributed()
    tr = run(B, rank, world, local, **flags)
    gathered = [torch.empty_like(tr.net.flat) for _ in range(world)]
    torch.distributed.all_gather(gathered, tr.net.flat)
    if rank == 0:
        assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged"
        torch.save({"params": tr.net.flat.cpu(), "stats": tr.epoch_stats.cpu(), "step": tr.step,
                    "comm": tr.comm, "graph": tr.use_graph}, out / "mgpu.pt")
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
