"""Drop-in for ``cleanmarl/ippo.py``: the single-env script collects ``--batch_size`` episodes one after another with
the same policy (ippo.py:301-344) and then runs exactly the update of ``ippo_multienvs.py``; here the ``batch_size`` episodes are
``batch_size`` parallel device envs of one rollout.  Same tyro ``Args`` (ippo.py:18-79), run directory
``runs/IPPO-...`` (ippo.py:277-279) and TensorBoard tags."""
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
    critic_hidden_dim: int = 32
    """ Hidden dimension of critic network"""


if __name__ == "__main__":
    main(algo="IPPO", ippo=True, args_cls=Args, run_prefix="IPPO")
