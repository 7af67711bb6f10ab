"""Locate and import the *reference* PyRates (harness only — test infrastructure).

Only ``tests/``, ``oracle/make_golden.py`` and bench.py's CPU legs may use this module.
The reference lives read-only at /root/reference in the build container and is NOT
present on the GPU box; everything that runs there uses committed fixtures instead.

Applies the two harness shims documented in SURVEY.md (headline 8):
  * ``ruamel.yaml`` stand-in (baseline/shims) when ruamel is not installed,
  * networkx >= 3.6 ``Graph.__new__(backend=...)`` collision with
    ``ComputeGraph.__init__(self, backend: str, ...)`` (pyrates/backend/computegraph.py:223-228).
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    for cand in (os.environ.get("PYRATES_REF"), "/root/reference",
                 os.path.join(os.path.dirname(_HERE), "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "pyrates")):
            return cand
    return None


def import_reference():
    """Return the imported reference ``pyrates`` module, or None when it is unavailable."""
    root = reference_root()
    if root is None:
        return None
    try:
        importlib.import_module("ruamel.yaml")
    except ImportError:
        shim = os.path.join(os.path.dirname(_HERE), "baseline", "shims")
        if shim not in sys.path:
            sys.path.insert(0, shim)
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        pyrates = importlib.import_module("pyrates")
        cg = importlib.import_module("pyrates.backend.computegraph")
    except Exception:
        return None
    # networkx >= 3.6 dispatch collision: a string `backend=` reaches Graph.__new__
    cg.ComputeGraph.__new__ = lambda cls, *a, **k: object.__new__(cls)
    return pyrates
