"""Drop-in for ``cleanmarl/mappo.py``: the single-env script collects ``--batch_size`` episodes one after another with
the same policy (mappo.py:301-344) and then runs exactly the update of ``mappo_multienvs.py``; here the ``batch_size`` episodes are
``batch_size`` parallel device envs of one rollout.  Same tyro ``Args`` (mappo.py:18-79), run directory
``runs/MAPPO-...`` (mappo.py:277-279) and TensorBoard tags."""
from __future__ import annotations

import sys
from dataclasses import dataclass
from pathlib import Path

if __package__ in (None, ""):
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))

from cleanmarl_b200.mappo import Args as _Base  # noqa: E402
from cleanmarl_b200.mappo_multienvs import main  # noqa: E402


@dataclass
class Args(_Base):
    eval_steps: int = 10
    """ Evaluate the policy each eval_steps training steps"""


if __name__ == "__main__":
    main(algo="MAPPO", ippo=False, args_cls=Args, run_prefix="MAPPO")
