"""Drop-in for ``cleanmarl/ippo_multienvs.py``: MAPPO's path with a decentralised critic on the
per-agent observations (``ippo_multienvs.py:34,200,336,495,503,554``)."""
from __future__ import annotations

import sys
from dataclasses import dataclass
from pathlib import Path

if __package__ in (None, ""):
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from cleanmarl_b200.mappo import Args as _Args  # noqa: E402
from cleanmarl_b200.mappo_multienvs import main  # noqa: E402


@dataclass
class Args(_Args):
    critic_hidden_dim: int = 32
    """ Hidden dimension of critic network (ippo_multienvs.py:34)"""


if __name__ == "__main__":
    main(algo="IPPO", ippo=True, args_cls=Args)
