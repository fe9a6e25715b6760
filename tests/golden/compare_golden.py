"""Check that the committed fixtures are what the unmodified reference produces today.

    CMARL_GOLDEN_OUT=/tmp/golden_regen python tests/golden/gen_golden.py
    python tests/golden/compare_golden.py /tmp/golden_regen

    CMARL_GOLDEN_DIR=/tmp/golden_regen python -m pytest tests/test_oracle_golden.py -q

g0 / g1 / g3 / g4 are deterministic: every array must be bit-identical to the committed file.  The g8 files record whole
runs of the reference scripts, whose rollouts are NOT reproducible from run to run (every worker process re-seeds its env
from OS entropy): a regenerated g8 file is a different, equally valid run, and what must hold is that
the oracle reproduces it from its recorded batch -- the second command above (last checked 2026-10-17: 28 passed on a
fresh set).  Build container only: needs /root/reference."""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def main(other):
    other = Path(other)
    bad = 0
    for f in sorted(HERE.glob("g*.npz")):
        a, b = np.load(f), np.load(other / f.name)
        keys = sorted(set(a.files) | set(b.files))
        diff = [k for k in keys if k not in a.files or k not in b.files or a[k].dtype != b[k].dtype
                or a[k].shape != b[k].shape or not np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f")]
        if f.name.startswith("g8_"):
            print(f"{f.name}: {len(keys)} arrays, {'identical' if not diff else 'another run of the reference (expected)'}")
            continue
        print(f"{f.name}: {len(keys)} arrays, {'identical' if not diff else 'DIFFERENT: ' + ', '.join(diff)}")
        bad += bool(diff)
    same = json.loads((HERE / "g0_args.json").read_text()) == json.loads((other / "g0_args.json").read_text())
    print(f"g0_args.json: {'identical' if same else 'DIFFERENT'}")
    return bad + (not same)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
