import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# tolerances of the north star (BASELINE.json): trajectories vs the reference NumPy backend
RTOL = {"float64": 1e-9, "float32": 1e-4, "complex128": 1e-9, "complex64": 1e-4}


def as_real_columns(ref):
    """Complex reference recordings -> the kernels' layout: state k as real columns (2k, 2k+1) = (re, im)."""
    ref = np.asarray(ref)
    if np.iscomplexobj(ref):
        return np.stack([ref.real, ref.imag], axis=2).reshape(ref.shape[0], -1)
    return ref


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden_names():
    """Trajectory fixtures (integration runs of the reference)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("jac_"))


def jacobian_names():
    """Jacobian fixtures (reference-generated get_jacobian_func evaluated at 16 states)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("jac_"))


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


def finite_rows(ref):
    """Number of leading rows of a reference recording that are finite (some random nets blow up)."""
    ok = np.all(np.isfinite(ref.reshape(ref.shape[0], -1)), axis=1)
    bad = np.nonzero(~ok)[0]
    return ref.shape[0] if bad.size == 0 else int(bad[0])


def traj_rel_err(got, ref):
    """Trajectory error: per state column, max |got-ref| scaled by that column's max |ref| (+tiny)."""
    n = finite_rows(ref)
    got, ref = np.asarray(got, dtype=np.float64)[:n], np.asarray(ref, dtype=np.float64)[:n]
    if n == 0:
        return 0.0
    if not np.all(np.isfinite(got)):
        return np.inf
    scale = np.max(np.abs(ref), axis=0) + 1e-300
    scale = np.maximum(scale, 1e-12 * np.max(scale))
    return float(np.max(np.abs(got - ref) / scale))


@pytest.fixture(scope="session")
def gpu():
    from pyrates_b200 import runtime as rt
    info = rt.device_info()          # raises loudly when libprb / the driver is missing
    return info
