"""Drop-in for ``cleanmarl/mappo_lstm_multienvs.py``: MAPPO with the recurrent actor (fc1 + GRUCell + fc2,
``mappo_lstm_multienvs.py:162-184``) trained by truncated BPTT (``--tbptt``, ``:603-620``).  Same tyro ``Args``
(``:18-81``), run directory ``runs/MAPPO-lstm-multienv-...`` (``:354-356``) and TensorBoard tags as the reference.

    python cleanmarl_b200/mappo_lstm_multienvs.py --batch_size 8192
"""
from __future__ import annotations

import sys
from pathlib import Path

if __package__ in (None, ""):
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from cleanmarl_b200.mappo import ArgsRecurrent as Args  # noqa: E402
from cleanmarl_b200.mappo_multienvs import main  # noqa: E402

if __name__ == "__main__":
    main(algo="MAPPO-lstm", ippo=False, args_cls=Args, run_prefix="MAPPO-lstm-multienv")
