"""Test harness (CPU, no GPU): run one of the REFERENCE's documentation scripts with every `CircuitTemplate.run` /
`grid_search` call forced to backend='cuda' and the device stage replaced by "lower + NVRTC-compile the kernels, return zeros".
Everything up to the launch is the product path: the backend hooks driven by the reference's compute graph, program
capture, scalarize / netcompile, CUDA emission and compilation for sm_100a.  Plotting is stubbed (matplotlib is not
installed).  Used by tests/test_reference_docs_on_cuda.py:  python tests/docs_harness.py <script.py> <workdir>"""
import sys, os, types, runpy, traceback, warnings
import unittest.mock as um
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors", "matplotlib.cm", "seaborn", "pycobi", "pyauto"):
    sys.modules[m] = um.MagicMock()
class Ax(um.MagicMock):
    def __iter__(self): return iter([Ax() for _ in range(16)])
    def __getitem__(self, k): return Ax()
    def flatten(self): return [Ax() for _ in range(16)]
plt = sys.modules["matplotlib.pyplot"]; sys.modules["matplotlib"].pyplot = plt
plt.subplots = lambda *a, **k: (um.MagicMock(), Ax())
plt.figure = lambda *a, **k: um.MagicMock()
import numpy as np
import pandas as pd
pd.DataFrame.plot = um.MagicMock(); pd.Series.plot = um.MagicMock()
from oracle.ref_env import import_reference
pyrates = import_reference()
import pyrates_b200.register as reg
reg.register()
from pyrates_b200 import backend as B, sweep as S, engine as E, netengine as N
from pyrates_b200.scalarize import UnsupportedModelError
stats = {"run": 0, "grid": 0, "models": []}

def fake_run(self, func, func_args, T, dt, dts, solver, **kw):
    prog = func.program(func_args)
    step = dts if dts else dt
    times = np.linspace(0.0, T, num=round(T / step), endpoint=False)
    if solver in E.ADAPTIVE_SOLVERS:
        try:
            model = E.InstanceModel(prog, solver=solver, rk_method=kw.get("method"))
        except UnsupportedModelError as e:
            if "too large" not in str(e): raise
            model = N.NetworkModel(prog, solver=solver)
    else:
        model = B.build_model(prog, solver=solver, exec_model=self.exec_model)
    model.module
    stats["run"] += 1; stats["models"].append((type(model).__name__, solver, prog.n_state, prog.float_precision))
    n = np.asarray(func_args[1]).size
    return np.zeros((len(times), n), dtype=np.asarray(func_args[1]).dtype), times
B.CudaBackend.run = fake_run

def fake_sweep_init_wrap(cls):
    orig_init = cls.__init__
    def run(self, table):
        (self.model.module if hasattr(self, "model") else None)
        stats["grid"] += 1
        ns = getattr(self, "n_samples", None)
        if ns is None:
            from pyrates_b200.engine import n_steps_for
            ns = n_steps_for(self.T, self.dt, self.dts)[2]
        return np.zeros((ns, len(getattr(self, "observers", getattr(self.model, "observers", [0]))), self.B))
    cls.run = run
    return cls
# BatchedSweep allocates device memory in __init__: replace the class by a compile-only stand-in
class FakeBatched:
    def __init__(self, prog, sweep, observers, *, solver, T, dt, dts, n_inst, **kw):
        kw.pop("chunk_samples", None); kw.pop("solver_kwargs", None); kw.pop("cutoff", None)
        self.model = E.InstanceModel(prog, solver=solver, sweep=sweep, observers=observers, reduce=kw.get("reduce", False), hist_max_delay=kw.get("hist_max_delay"))
        self.model.module
        self.B, self.T, self.dt, self.dts = n_inst, T, dt, dts
        self.h_y0 = types.SimpleNamespace(array=np.zeros((self.model.n_state, n_inst)))
        self.n_obs = len(observers)
        self.timing = {"lower_s": 0.0, "alloc_s": 0.0}
        stats["models"].append(("BatchedSweep", solver, prog.n_state, prog.float_precision))
    def run(self, table):
        stats["grid"] += 1
        ns = E.n_steps_for(self.T, self.dt, self.dts)[2]
        self._res = np.zeros((ns, self.n_obs, self.B))
        return self._res
    def detach_result(self): return self._res
    def free(self): pass
S.BatchedSweep = FakeBatched
script = sys.argv[1]
os.chdir(sys.argv[2])
# force backend
from pyrates import CircuitTemplate
import pyrates.frontend.template.circuit as fc
orig_run = fc.CircuitTemplate.run
FORCE = os.environ.get("FORCE", "1") == "1"
def run_cuda(self, *a, **k):
    if FORCE: k["backend"] = "cuda"
    k.setdefault("verbose", False)
    return orig_run(self, *a, **k)
fc.CircuitTemplate.run = run_cuda
import pyrates.utility as pu
orig_gs = pyrates.grid_search
def gs_cuda(*a, **k):
    if FORCE: k["backend"] = "cuda"
    k.setdefault("verbose", False)
    return orig_gs(*a, **k)
pyrates.grid_search = gs_cuda; pu.grid_search = gs_cuda
# get_run_func / get_jacobian_func: force the cuda backend as well; the returned callables lower + compile their evaluation
# kernel and hand back zeros of the right shape instead of launching
orig_grf, orig_gjf = fc.CircuitTemplate.get_run_func, fc.CircuitTemplate.get_jacobian_func
def grf_cuda(self, *a, **k):
    if FORCE: k["backend"] = "cuda"
    k.setdefault("verbose", False)
    return orig_grf(self, *a, **k)
def gjf_cuda(self, *a, **k):
    if FORCE: k["backend"] = "cuda"
    k.setdefault("verbose", False)
    return orig_gjf(self, *a, **k)
fc.CircuitTemplate.get_run_func, fc.CircuitTemplate.get_jacobian_func = grf_cuda, gjf_cuda
stats["evals"] = 0
def fake_evaluate(prog, t=None, y=None):
    t_arg = np.asarray(prog.arg_by_kind("t").value)
    m = E._eval_model(prog, not np.issubdtype(t_arg.dtype, np.integer))
    m.module
    stats["evals"] += 1
    n_out = m.rhs.n_out if m.rhs.n_out is not None else m.n_state
    return np.zeros((n_out, np.shape(y)[1]) if (y is not None and np.ndim(y) == 2) else n_out, dtype=m.real)
E.evaluate_rhs = fake_evaluate
try:
    runpy.run_path(script, run_name="__main__")
    print("RESULT OK", os.path.basename(script), stats["run"], "runs", stats["grid"], "sweeps", stats["evals"], "evaluations", stats["models"][:6])
except BaseException as e:
    tb = traceback.extract_tb(e.__traceback__)
    ours = [f for f in tb if "pyrates_b200" in f.filename]
    print("RESULT FAIL", os.path.basename(script), type(e).__name__, str(e)[:200].replace("\n", " "), "| runs", stats["run"], "| last frame:", (ours[-1].filename.split("/")[-1], ours[-1].lineno) if ours else (tb[-1].filename.split("/")[-1], tb[-1].lineno))
