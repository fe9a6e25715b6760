"""Drop-in for ``cleanmarl/ippo_lstm.py``: the single-env script collects ``--batch_size`` episodes one after another with
the same policy (ippo_lstm.py:301-344) and then runs exactly the update of ``ippo_lstm_multienvs.py``; here the ``batch_size`` episodes are
``batch_size`` parallel device envs of one rollout.  Same tyro ``Args`` (ippo_lstm.py:18-79), run directory
``runs/IPPO-lstm-...`` (ippo_lstm.py:277-279) and TensorBoard tags."""
from __future__ import annotations

import sys
from dataclasses import dataclass
from pathlib import Path

if __package__ in (None, ""):
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))

from cleanmarl_b200.mappo import ArgsRecurrentIPPO as _Base  # noqa: E402
from cleanmarl_b200.mappo_multienvs import main  # noqa: E402


@dataclass
class Args(_Base):
    optimizer: str = "Adam"
    """ The optimizer"""


if __name__ == "__main__":
    main(algo="IPPO-lstm", ippo=True, args_cls=Args, run_prefix="IPPO-lstm")
