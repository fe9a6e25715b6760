"""Small driver for compute-sanitizer runs (profiles/sanitizer_r2.md):

    compute-sanitizer --tool {memcheck,initcheck,racecheck,synccheck} python profiles/tools/sanitize_driver.py [what]

Runs whole trainer iterations (rollout, critic forward, TD(lambda), 3 PPO epochs with the tcgen05 chain kernels, the
programmatic dependent launch of the critic chain, reduction, Adam) at a size with ragged tail tiles and several tiles per
CTA, eager launches.  `what`: mlp (default), chained (launch chaining on), ippo, recurrent, ffma.
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch  # noqa: E402
from cleanmarl_b200.mappo import MAPPO, Args, ArgsRecurrent  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "mlp"
B = int(os.environ.get("SAN_ENVS", "300"))
if what == "ffma":
    os.environ["CMARL_TENSOR_CORES"] = "0"
cls = ArgsRecurrent if what == "recurrent" else Args
tr = MAPPO(cls(batch_size=B, seed=7), ippo=(what == "ippo"), use_graph=False)
tr.chain = what == "chained"
for _ in range(2):
    tr.iteration()
torch.cuda.synchronize()
print(f"sanitize_driver {what}: B={B} launches={tr.engine.launches} params finite={bool(torch.isfinite(tr.net.flat).all())} "
      f"stats={tr.epoch_stats[-1].tolist()}")
