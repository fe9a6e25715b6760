"""Drop-in for ``cleanmarl/ippo_lstm_multienvs.py``: the recurrent-actor path of ``mappo_lstm_multienvs.py`` with a
decentralised MLP critic on the per-agent observations (``ippo_lstm_multienvs.py:336, 525, 533, 623``), AdamW
(``:38``) and ``--tbptt 5`` (``:66``) as defaults; run directory ``runs/IPPO-lstm-multienvs-...`` (``:354-356``)."""
from __future__ import annotations

import sys
from pathlib import Path

if __package__ in (None, ""):
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from cleanmarl_b200.mappo import ArgsRecurrentIPPO as Args  # noqa: E402
from cleanmarl_b200.mappo_multienvs import main  # noqa: E402

if __name__ == "__main__":
    main(algo="IPPO-lstm", ippo=True, args_cls=Args, run_prefix="IPPO-lstm-multienvs")
