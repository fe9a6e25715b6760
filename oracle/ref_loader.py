"""Load / run the UNMODIFIED reference script (test infrastructure).

The reference file is looked up in this order:
  1. ``$CMARL_REFERENCE_DIR`` (a directory holding ``mappo_multienvs.py``),
  2. ``/root/reference/cleanmarl`` (this container only -- absent on the GPU box),
  3. ``<repo>/baseline/_ref/cleanmarl`` (git-ignored copy made by
     ``oracle/install_reference.py``; travels to the GPU box with ``gpurun``).

``oracle/env_stub`` is put first on ``sys.path`` so the reference's
``from env.pettingzoo_wrapper import ...`` lines (MME:11-13) resolve to the numpy
stand-in (PettingZoo/gymnasium/smaclite/lbforaging are not installed).
"""
from __future__ import annotations

import importlib.util
import os
import runpy
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
STUB = Path(__file__).resolve().parent / "env_stub"


def reference_dir() -> Path | None:
    cands = [os.environ.get("CMARL_REFERENCE_DIR"), "/root/reference/cleanmarl",
             str(REPO / "baseline" / "_ref" / "cleanmarl")]
    for c in cands:
        if c and (Path(c) / "mappo_multienvs.py").is_file():
            return Path(c)
    return None


def _prepare_path():
    for p in (str(REPO), str(STUB)):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(STUB))
    # a previously imported real/other ``env`` package would shadow the stub
    for name in [m for m in sys.modules if m == "env" or m.startswith("env.")]:
        f = getattr(sys.modules[name], "__file__", "") or ""
        if str(STUB) not in f:
            del sys.modules[name]


def load_module(script: str = "mappo_multienvs.py"):
    """Import the reference file as a module (its ``__main__`` block does not run)."""
    d = reference_dir()
    if d is None:
        raise FileNotFoundError("reference sources not found (see oracle/ref_loader.py)")
    _prepare_path()
    name = "_cmarl_ref_" + script.replace(".py", "")
    spec = importlib.util.spec_from_file_location(name, d / script)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def run_script(argv: list[str], script: str = "mappo_multienvs.py", cwd: str | None = None) -> dict:
    """``python <script> argv...`` in-process; returns the script's final globals.

    The reference writes TensorBoard files under ``./runs`` -- pass a scratch ``cwd``.
    """
    d = reference_dir()
    if d is None:
        raise FileNotFoundError("reference sources not found (see oracle/ref_loader.py)")
    _prepare_path()
    old_argv, old_cwd = sys.argv, os.getcwd()
    try:
        if cwd:
            os.makedirs(cwd, exist_ok=True)
            os.chdir(cwd)
        sys.argv = [str(d / script)] + list(argv)
        return runpy.run_path(str(d / script), run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
