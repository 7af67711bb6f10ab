"""Generate the golden fixtures under tests/golden/ from the REAL reference (harness only).

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden [case ...]

Each fixture is one ``.npz`` holding
  * the Program (generated ``vector_field`` assignment stream + argument arrays) exactly as the
    reference's NumPy backend produced it (captured at ``BaseBackend.generate_func`` /
    ``BaseBackend.run``, pyrates/backend/base/base_backend.py:537,596),
  * ``rec_<solver>``: the state recording returned by the reference's own ``BaseBackend.run`` for
    solvers 'euler' and 'heun' (the reference's aliased Heun), and — because the reference has no
    fixed-step RK4 / textbook Heun — ``rec_rk4`` / ``rec_heun_true`` computed by
    ``oracle.numpy_oracle`` around the reference-generated function (flagged ``oracle_defined``),
  * ``__run__``: T, dt, dts, solver list, float precision.
The GPU box has no reference; tests there compare the CUDA path against these files.
"""
from __future__ import annotations

import copy
import json
import os
import sys
import tempfile
import warnings

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle.ref_env import import_reference  # noqa: E402
from oracle import numpy_oracle as orc       # noqa: E402
from pyrates_b200.program import Program     # noqa: E402

GOLDEN = os.path.join(_ROOT, "tests", "golden")


class HistInit:
    """Stands for the reference's DDEHistory in captured argument lists: only its initial condition travels."""

    def __init__(self, y0):
        self._y = [np.array(y0, copy=True)]


def _fresh(args):
    """Copies of captured arguments for one oracle run (a new History per run)."""
    return [orc.History(a._y[0]) if isinstance(a, HistInit) else np.array(a, copy=True) for a in args]


class Capture:
    """Monkey-patches the reference NumPy backend to record what it generates and returns."""

    def __init__(self):
        import pyrates.backend.base.base_backend as bb
        import pyrates.backend.computegraph as cg
        self.bb, self.cg = bb, cg
        self.source = None
        self.func_args = None
        self.names = None
        self.state_vars = None
        self.rec = None
        self.labels = None
        self._orig = (bb.BaseBackend.generate_func, bb.BaseBackend.run, cg.ComputeGraph.to_func)
        cap = self

        def generate_func(self_, func_name, to_file=True, **kw):
            cap.source = self_.generate()
            return cap._orig[0](self_, func_name, to_file=to_file, **kw)

        def run(self_, func, func_args, T, dt, dts, solver, **kw):
            cap.func_args = [HistInit(a._y[0]) if isinstance(a, bb.DDEHistory) else np.array(a, copy=True) for a in func_args]
            out = cap._orig[1](self_, func, func_args, T, dt, dts, solver, **kw)
            cap.rec = np.array(out[0], copy=True)
            return out

        def to_func(self_, *a, **kw):
            out = cap._orig[2](self_, *a, **kw)
            cap.names = list(out[2])
            cap.state_vars = dict(out[3])
            return out

        bb.BaseBackend.generate_func = generate_func
        bb.BaseBackend.run = run
        cg.ComputeGraph.to_func = to_func

    def close(self):
        self.bb.BaseBackend.generate_func, self.bb.BaseBackend.run, self.cg.ComputeGraph.to_func = self._orig


# ----------------------------------------------------------------------------- case builders
def _yaml(path):
    def build():
        from pyrates import CircuitTemplate
        return CircuitTemplate.from_yaml(path)
    return build


def _mackey_glass():
    """Two delay-differential units in the explicit ``x(t-d)`` notation (pyrates/backend/parser.py:63-69 -> past(x, d) ->
    hist lookups, computegraph.py:1358-1366): a Mackey-Glass oscillator driving a delayed-feedback relaxation unit."""
    def build():
        from pyrates import CircuitTemplate, NodeTemplate, OperatorTemplate
        mg = OperatorTemplate(name="mg_op", equations=["x' = bet*x(t-tau)/(1 + x(t-tau)^10) - gam*x + k*x_in"],
                              variables={"x": "output(1.2)", "bet": 2.0, "gam": 1.0, "tau": 0.0173, "k": 0.3,
                                         "x_in": "input(0.0)"})
        fb = OperatorTemplate(name="fb_op", equations=["z' = -a*z(t-d) + b*z_in - z^3"],
                              variables={"z": "output(0.4)", "a": 3.0, "b": 0.5, "d": 0.0091, "z_in": "input(0.0)"})
        n1 = NodeTemplate(name="n1", operators=[mg])
        n2 = NodeTemplate(name="n2", operators=[fb])
        return CircuitTemplate(name="mg_fb", nodes={"p1": n1, "p2": n2},
                               edges=[("p1/mg_op/x", "p2/fb_op/z_in", None, {"weight": 1.0}),
                                      ("p2/fb_op/z", "p1/mg_op/x_in", None, {"weight": 1.0})])
    return build


def _jrc_pair_delayed():
    """Two Jansen-Rit circuits coupled through DISCRETE edge delays on a state variable: fixed-step solvers get ring buffers,
    adaptive ones the DDE path (pyrates/ir/circuit.py:643-667)."""
    def build():
        from pyrates import CircuitTemplate
        jrc = CircuitTemplate.from_yaml(NM + "jansenrit.JRC")
        return CircuitTemplate(name="jrc_pair_dde", circuits={"jrc1": jrc, "jrc2": jrc},
                               edges=[("jrc1/pc/rpo_e_in/v", "jrc2/pc/rpo_e_in/m_in", None, {"weight": 2000.0, "delay": 0.004}),
                                      ("jrc2/pc/rpo_e_in/v", "jrc1/pc/rpo_e_in/m_in", None, {"weight": 4000.0, "delay": 0.0075})])
    return build


def _two_qif():
    """Two coupled QIF populations, vectorised: an input to `all/...` arrives as a (steps, 2) table -> interp_rows."""
    def build():
        from pyrates import CircuitTemplate, NodeTemplate
        qif = NodeTemplate.from_yaml(NM + "qif.qif_pop")
        return CircuitTemplate("two_qif", nodes={"p1": qif, "p2": qif},
                               edges=[("p1/qif_op/r", "p2/qif_op/r_in", None, {"weight": 5.0}),
                                      ("p2/qif_op/r", "p1/qif_op/r_in", None, {"weight": 8.0})])
    return build


def _wc_pop(N, seed=42):
    def build():
        from pyrates import CircuitTemplate, NodeTemplate
        from pyrates.frontend.template.population import PopulationTemplate, Connectivity
        rng = np.random.default_rng(seed)
        W = rng.normal(0.0, 1.0 / np.sqrt(N), (N, N))
        np.fill_diagonal(W, 15.0)
        exc = NodeTemplate.from_yaml("model_templates.neural_mass_models.wilsoncowan.exc_pop")
        # the template's defaults (r=0, r_ext=0) sit exactly on a fixed point: start from random rates and
        # heterogeneous thresholds so the trajectory actually exercises the coupling
        pop = PopulationTemplate(name="e", node=exc, n=N,
                                 params={"rate_op/r": list(rng.uniform(0.05, 0.6, N)),
                                         "se_op/theta": list(rng.normal(2.0, 0.5, N))})
        conn = Connectivity(source="e/rate_op/r", target="e/se_op/r_in", weights=W)
        return CircuitTemplate(name=f"wc_n{N}", populations={"e": pop}, connections=[conn])
    return build


def _qif_sfa_pop(N, variant, seed=42):
    def build():
        from pyrates import CircuitTemplate, NodeTemplate
        from pyrates.frontend.template.population import PopulationTemplate, Connectivity
        rng = np.random.default_rng(seed)
        eta = rng.normal(-5.0, 1.0, N)
        W = rng.normal(0.0, 1.0 / np.sqrt(N), (N, N))
        qs = NodeTemplate.from_yaml("model_templates.neural_mass_models.qif.qif_sfa_pop")
        if variant in ("cascade", "ring"):
            pop = PopulationTemplate(name="p", node=qs, n=N, params={"qif_sfa_op/eta": list(eta)})
            kw = dict(delays=4e-3, spread=2e-3) if variant == "cascade" else dict(delays=4e-3)
            conn = Connectivity(source="p/qif_sfa_op/r", target="p/qif_sfa_op/r_in", weights=W, **kw)
            return CircuitTemplate(name=f"qsfa_{variant}", populations={"p": pop}, connections=[conn])
        p1 = PopulationTemplate(name="p1", node=qs, n=N, params={"qif_sfa_op/eta": list(eta)})
        p2 = PopulationTemplate(name="p2", node=qs, n=N, params={"qif_sfa_op/eta": list(eta[::-1])})
        W2 = rng.normal(0.0, 1.0 / np.sqrt(N), (N, N))
        c12 = Connectivity(source="p1/qif_sfa_op/r", target="p2/qif_sfa_op/r_in", weights=W, delays=4e-3, spread=2e-3)
        c21 = Connectivity(source="p2/qif_sfa_op/r", target="p1/qif_sfa_op/r_in", weights=W2, delays=3e-3)
        return CircuitTemplate(name="qsfa_both", populations={"p1": p1, "p2": p2}, connections=[c12, c21])
    return build


def _kuramoto(N, dense=True):
    def build():
        from pyrates import CircuitTemplate, NodeTemplate, EdgeTemplate
        from pyrates.frontend.template.population import PopulationTemplate, Connectivity
        phase_pop = NodeTemplate.from_yaml("model_templates.oscillators.kuramoto.phase_pop")
        sin_edge = EdgeTemplate.from_yaml("model_templates.oscillators.kuramoto.sin_edge")
        theta = np.random.default_rng(1).uniform(0, 2 * np.pi, N)
        omega = np.random.default_rng(1).normal(10, 1, N)
        pop = PopulationTemplate(name="e", node=phase_pop, n=N,
                                 params={"phase_op/theta": list(theta), "phase_op/omega": list(omega)})
        W = np.random.default_rng(2).uniform(0, 2.0 / N, (N, N)) if dense else 1.0 / N
        conn = Connectivity(source="e/phase_op/theta", target="e/phase_op/s_in", weights=W, edge=sin_edge,
                            edge_var_map={"theta_s": "source", "theta_t": "e/phase_op/theta"})
        return CircuitTemplate(name="kuramoto", populations={"e": pop}, connections=[conn])
    return build


def _matrix_delay_test(kind):
    """The reference's own known-answer set-ups, tests/test_population_connectivity.py:361-476."""
    def build():
        from pyrates import CircuitTemplate, NodeTemplate, OperatorTemplate
        from pyrates.frontend.template.population import PopulationTemplate, Connectivity
        node_op = OperatorTemplate(name="rate_op_d", equations=["r' = (-r + eta + s_in) / tau_r"],
                                   variables={"r": "output", "eta": 5.0, "tau_r": 0.1, "s_in": "input"})
        node = NodeTemplate(name="rate_node_d", operators=[node_op])
        pop = PopulationTemplate(name="p", node=node, n=2,
                                 params={"rate_op_d/eta": [5.0, 8.0], "rate_op_d/r": [1.0, 2.0]})
        W = np.array([[0., 0.5], [0.5, 0.]])
        kw = dict(delays=4e-3) if kind == "ring" else dict(delays=5e-3, spread=2e-3)
        conn = Connectivity(source="p/rate_op_d/r", target="p/rate_op_d/s_in", weights=W, **kw)
        return CircuitTemplate(name="dtest", populations={"p": pop}, connections=[conn])
    return build


def _jrc_grid(n_c, n_u, T, dt, dts):
    """reference grid_search over (C, u) like documentation/model_analysis/parameter_sweeps.py:44-90."""
    def run(solver, precision):
        from pyrates import grid_search
        param_grid = {"C": np.linspace(68.0, 1350.0, n_c), "u": np.linspace(120.0, 320.0, n_u)}
        param_map = {
            "C": {"vars": ["weight"], "edges": [("pc/pro/m", "ein/rpo_e/m_in")]},
            "C2": {"vars": ["weight"], "edges": [("ein/pro/m", "pc/rpo_e_in/m_in")]},
            "C4": {"vars": ["weight"], "edges": [("pc/pro/m", "iin/rpo_e/m_in"), ("iin/pro/m", "pc/rpo_i/m_in")]},
            "u": {"vars": ["rpo_e_in/u"], "nodes": ["pc"]},
        }
        import pandas as pd
        from pyrates.utility import linearize_grid
        grid = linearize_grid(param_grid, permute=True)
        grid["C2"] = 0.8 * grid["C"]
        grid["C4"] = 0.25 * grid["C"]
        res, pmap = grid_search(NM + "jansenrit.JRC", grid, param_map, step_size=dt, simulation_time=T,
                                outputs={"v": "pc/rpo_e_in/v"}, sampling_step_size=dts, solver=solver,
                                backend="default", float_precision=precision, verbose=False, clear=True)
        return {"grid_C": np.asarray(pmap["C"].values, dtype=np.float64), "grid_u": np.asarray(pmap["u"].values, dtype=np.float64),
                f"frame_{solver}": np.asarray(res.values, dtype=np.float64),
                "frame_columns": np.frombuffer(json.dumps([list(map(str, c)) for c in res.columns]).encode(), dtype=np.uint8)}
    return run


def _sin_input(n, cols=None):
    t = np.arange(n) * 1e-3
    if cols is None:
        return np.sin(2 * np.pi * 5.0 * t) * 3.0
    return np.stack([np.sin(2 * np.pi * (3.0 + k) * t) * (1.0 + k) for k in range(cols)], axis=1)


TB = "model_templates.test_resources.test_backend."
NM = "model_templates.neural_mass_models."

# name -> dict(build, T, dt, dts, outputs (any valid output, only used to drive run()), inputs, precisions)
CASES = {
    "qif":        dict(build=_yaml(NM + "qif.qif"), T=1.0, dt=1e-3, out="p/qif_op/r"),
    "qif_f32":    dict(build=_yaml(NM + "qif.qif"), T=1.0, dt=1e-3, out="p/qif_op/r", precision="float32"),
    "qif_dt2":    dict(build=_yaml(NM + "qif.qif"), T=2.0, dt=1e-2, dts=2e-2, out="p/qif_op/r"),
    "qif_input":  dict(build=_yaml(NM + "qif.qif"), T=0.5, dt=1e-3, out="p/qif_op/r",
                       inputs=lambda: {"p/qif_op/I_ext": _sin_input(500)}),
    "qif_sfa":    dict(build=_yaml(NM + "qif.qif_sfa"), T=1.0, dt=1e-3, dts=5e-3, out="p/qif_sfa_op/r"),
    "jrc":        dict(build=_yaml(NM + "jansenrit.JRC"), T=0.2, dt=1e-4, dts=1e-3, out="pc/rpo_e_in/v"),
    "jrc_2delay": dict(build=_yaml(NM + "jansenrit.JRC_2delaycoupled"), T=0.1, dt=1e-4, dts=1e-3,
                       out="jrc1/pc/rpo_e_in/v"),
    "net0":  dict(build=_yaml(TB + "net0"), T=0.2, dt=1e-3, out="pop0/op0/a"),
    "net1":  dict(build=_yaml(TB + "net1"), T=0.2, dt=1e-3, out="pop0/op1/a",
                  inputs=lambda: {"pop0/op1/u": _sin_input(200)}),
    "net2":  dict(build=_yaml(TB + "net2"), T=0.2, dt=1e-3, out="pop0/op2/a"),
    "net3":  dict(build=_yaml(TB + "net3"), T=0.2, dt=1e-3, out="pop0/op3/a",
                  inputs=lambda: {"pop0/op3/u": _sin_input(200)}),
    "net6":  dict(build=_yaml(TB + "net6"), T=0.2, dt=1e-3, out="pop0/op1/a"),
    "net7":  dict(build=_yaml(TB + "net7"), T=0.2, dt=1e-3, out="pop0/op1/a"),
    "net8":  dict(build=_yaml(TB + "net8"), T=0.2, dt=1e-3, out="pop0/op0/a"),
    "net9":  dict(build=_yaml(TB + "net9"), T=0.2, dt=1e-3, out="pop0/op1/a",
                  inputs=lambda: {"pop1/op7/inp": _sin_input(200)}),
    "net10": dict(build=_yaml(TB + "net10"), T=1.5, dt=1e-3, dts=1e-2, out="pop0/op8/a"),
    "net11": dict(build=_yaml(TB + "net11"), T=0.6, dt=1e-3, dts=2e-3, out="pop0/op1/a",
                  inputs=lambda: {"pop1/op7/inp": _sin_input(600)}),
    "net12": dict(build=_yaml(TB + "net12"), T=0.4, dt=1e-3, dts=2e-3, out="pop0/op7/a",
                  inputs=lambda: {"pop0/op7/inp": _sin_input(400, cols=2)}),
    "net13": dict(build=_yaml(TB + "net13"), T=0.2, dt=1e-3, out="p1/op9/a"),
    "ring_kat":    dict(build=_matrix_delay_test("ring"), T=10e-3, dt=1e-3, out="p/rate_op_d/r"),
    "gamma_kat":   dict(build=_matrix_delay_test("gamma"), T=20e-3, dt=1e-3, out="p/rate_op_d/r"),
    "wc_n8":       dict(build=_wc_pop(8), T=0.2, dt=1e-3, out="e/rate_op/r"),
    "wc_n128":     dict(build=_wc_pop(128), T=0.1, dt=1e-3, dts=2e-3, out="e/rate_op/r"),
    "wc_n128_f32": dict(build=_wc_pop(128), T=0.1, dt=1e-3, dts=2e-3, out="e/rate_op/r", precision="float32"),
    # float32 networks for the tcgen05 (3xTF32) sweep-batched coupling: padded rows / K, several K chunks, two matrices
    "wc_n300_f32": dict(build=_wc_pop(300), T=0.05, dt=1e-3, dts=2e-3, out="e/rate_op/r", precision="float32"),
    "qsfa_both_n96_f32": dict(build=_qif_sfa_pop(96, "both"), T=0.05, dt=1e-3, out="p1/qif_sfa_op/r", precision="float32"),
    "kuramoto_n128_f32": dict(build=_kuramoto(128), T=0.05, dt=1e-3, out="e/phase_op/theta", precision="float32"),
    # complex precisions (Ott-Antonsen mean fields: complex state z, conjugate, i*omega)
    "kmo_mf_c128": dict(build=_yaml("model_templates.oscillators.kuramoto.kmo_mf"), T=2.0, dt=1e-3, dts=1e-2, out="p/kmo_op/z",
                        precision="complex128"),
    "kmo_mf2_c64": dict(build=_yaml("model_templates.oscillators.kuramoto.kmo_mf_2coupled"), T=2.0, dt=1e-3, dts=1e-2,
                        out="p1/kmo_op/z", precision="complex64"),
    "kmo_theta_c128": dict(build=_yaml(NM + "theta.kmo_theta"), T=2.0, dt=1e-3, dts=1e-2, out="p/theta_op/z",
                           precision="complex128"),
    # delay-differential models, fixed-step solvers: `x(t-d)` notation -> hist(t*dt - d)[idx] lookups into the state history
    # (DDEHistory, base_backend.py:88-177; the reference's own test: tests/test_backend_simulations.py:551-583)
    "dde_net16":     dict(build=_yaml(TB + "net16"), T=0.5, dt=1e-3, out="p1/op_dde/x", run_kw=dict(vectorize=False)),
    "dde_net16_f32": dict(build=_yaml(TB + "net16"), T=0.5, dt=1e-3, dts=2e-3, out="p1/op_dde/x", precision="float32",
                          run_kw=dict(vectorize=False)),
    "dde_mg":        dict(build=_mackey_glass(), T=0.3, dt=1e-4, dts=1e-3, out="p1/mg_op/x", run_kw=dict(vectorize=False)),
    "dde_mg_f32":    dict(build=_mackey_glass(), T=0.3, dt=1e-4, dts=1e-3, out="p1/mg_op/x", precision="float32",
                          run_kw=dict(vectorize=False)),
    "qsfa_cascade_n8":   dict(build=_qif_sfa_pop(8, "cascade"), T=0.1, dt=1e-3, out="p/qif_sfa_op/r"),
    "qsfa_ring_n8":      dict(build=_qif_sfa_pop(8, "ring"), T=0.1, dt=1e-3, out="p/qif_sfa_op/r"),
    "qsfa_both_n8":      dict(build=_qif_sfa_pop(8, "both"), T=0.1, dt=1e-3, out="p1/qif_sfa_op/r"),
    "qsfa_both_n96":     dict(build=_qif_sfa_pop(96, "both"), T=0.05, dt=1e-3, out="p1/qif_sfa_op/r"),
    "kuramoto_n8":       dict(build=_kuramoto(8), T=0.2, dt=1e-3, out="e/phase_op/theta"),
    "kuramoto_n128":     dict(build=_kuramoto(128), T=0.05, dt=1e-3, out="e/phase_op/theta"),
    "kuramoto_scalar_n128": dict(build=_kuramoto(128, dense=False), T=0.05, dt=1e-3, out="e/phase_op/theta"),
    # reference grid_search (B deep copies vectorised into one circuit, pyrates/utility.py:189-275)
    "jrc_grid_b16":   dict(run=_jrc_grid(4, 4, 0.05, 1e-4, 1e-3), T=0.05, dt=1e-4, dts=1e-3, out="pc/rpo_e_in/v"),
    "jrc_grid_b1024": dict(run=_jrc_grid(32, 32, 2e-3, 1e-4, 1e-3), T=2e-3, dt=1e-4, dts=1e-3, out="pc/rpo_e_in/v"),
}


def make_case(name: str, spec: dict, cap: Capture) -> str:
    from pyrates import clear
    from pyrates.ir.node import clear_ir_caches
    precision = spec.get("precision", "float64")
    T, dt, dts = spec["T"], spec["dt"], spec.get("dts")
    recs = {}
    prog = None
    first_args = None
    cwd = os.getcwd()
    extra = {}
    for solver in ("euler", "heun"):
        clear_ir_caches()
        circuit = None
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                if "run" in spec:
                    extra.update(spec["run"](solver, precision))
                else:
                    circuit = spec["build"]()
                    inputs = spec["inputs"]() if "inputs" in spec else None
                    circuit.run(simulation_time=T, step_size=dt, sampling_step_size=dts, solver=solver,
                                outputs={"o": spec["out"]}, inputs=inputs, backend="default",
                                float_precision=precision, verbose=False, clear=True, **spec.get("run_kw", {}))
            finally:
                os.chdir(cwd)
        recs[solver] = cap.rec
        if prog is None:
            first_args = cap.func_args
            prog = Program.from_source(cap.source, list(zip(cap.names, cap.func_args)),
                                       state_vars=cap.state_vars, float_precision=precision)
        try:
            if circuit is not None:
                clear(circuit)
        except Exception:
            pass
    # oracle-defined solvers (no reference implementation): run around the reference-generated text
    func = orc.build_vector_field(prog.source())
    for solver in ("heun_true", "rk4"):
        args = _fresh(first_args)
        rec, _ = orc.run(func, args, T, dt, dts, solver)
        recs[solver] = rec
    # cross-check: the oracle must reproduce the reference bit-for-bit on the reference's solvers
    for solver in ("euler", "heun"):
        args = _fresh(first_args)
        rec, _ = orc.run(func, args, T, dt, dts, solver)
        diff = float(np.max(np.abs(rec - recs[solver]))) if rec.size else 0.0
        assert diff == 0.0 or not np.isfinite(diff), f"{name}/{solver}: oracle != reference (max abs diff {diff})"
    run_meta = {"T": T, "dt": dt, "dts": dts, "solver": "euler", "precision": precision,
                "reference_solvers": ["euler", "heun"], "oracle_defined": ["heun_true", "rk4"],
                "reference_version": "1.2.3", "out": spec["out"]}
    path = os.path.join(GOLDEN, f"{name}.npz")
    prog.save(path, __run__=np.frombuffer(json.dumps(run_meta).encode(), dtype=np.uint8),
              **{f"rec_{k}": v for k, v in recs.items()}, **extra)
    return path


# adaptive cases: the reference's solver='scipy' (solve_ivp RK45, base_backend.py:802-812); float t, no ring buffers
ADAPTIVE_CASES = {
    "qif_rk45":     dict(build=_yaml(NM + "qif.qif"), T=2.0, dt=1e-3, dts=1e-2, out="p/qif_op/r"),
    "qif_rk45_tight": dict(build=_yaml(NM + "qif.qif"), T=2.0, dt=1e-3, dts=1e-2, out="p/qif_op/r",
                           kw=dict(rtol=1e-8, atol=1e-10)),
    "jrc_rk45":     dict(build=_yaml(NM + "jansenrit.JRC"), T=0.5, dt=1e-4, dts=1e-3, out="pc/rpo_e_in/v"),
    "jrc_rk45_tight": dict(build=_yaml(NM + "jansenrit.JRC"), T=0.5, dt=1e-4, dts=1e-3, out="pc/rpo_e_in/v",
                           kw=dict(rtol=1e-7, atol=1e-9)),
    "qif_sfa_rk45": dict(build=_yaml(NM + "qif.qif_sfa"), T=2.0, dt=1e-3, dts=5e-3, out="p/qif_sfa_op/r",
                         kw=dict(rtol=1e-6, atol=1e-8)),
    "wc_n8_rk45":   dict(build=_wc_pop(8), T=0.5, dt=1e-3, dts=5e-3, out="e/rate_op/r", kw=dict(rtol=1e-6, atol=1e-9)),
    # extrinsic input with an adaptive solver: the frontend emits interp(t, time, u_input) (circuit.py:1692-1709)
    "qif_input_rk45": dict(build=_yaml(NM + "qif.qif"), T=2.0, dt=1e-3, dts=1e-2, out="p/qif_op/r",
                           inputs=lambda: {"p/qif_op/I_ext": _sin_input(2000) * 3.0}, kw=dict(rtol=1e-7, atol=1e-9)),
    # vector-valued input with an adaptive solver: interp_rows(t, time, u_input) over a (steps, 2) table (circuit.py:1696-1700)
    "qif2_input_rk45": dict(build=_two_qif(), T=1.0, dt=1e-3, dts=1e-2, out="all/qif_op/r",
                            inputs=lambda: {"all/qif_op/I_ext": _sin_input(1000, cols=2)}, kw=dict(rtol=1e-7, atol=1e-9)),
    # SciPy's RK23 (the reference's own solver test uses it, tests/test_backend_simulations.py:370)
    "jrc_rk23":     dict(build=_yaml(NM + "jansenrit.JRC"), T=0.3, dt=1e-4, dts=1e-3, out="pc/rpo_e_in/v",
                         kw=dict(method="RK23", rtol=1e-6, atol=1e-9)),
    "qif_sfa_rk23": dict(build=_yaml(NM + "qif.qif_sfa"), T=2.0, dt=1e-3, dts=5e-3, out="p/qif_sfa_op/r", kw=dict(method="RK23")),
    # delay-differential models with solver='scipy': scipy.integrate.ode('dopri5') restarted per sample interval, accepted
    # steps appended to the DDEHistory (BaseBackend._solve_scipy_dde, base_backend.py:771-800)
    "dde_net16_dopri5":     dict(build=_yaml(TB + "net16"), T=2.0, dt=1e-3, dts=1e-2, out="p1/op_dde/x", run_kw=dict(vectorize=False)),
    "dde_net16_dopri5_f32": dict(build=_yaml(TB + "net16"), T=2.0, dt=1e-3, dts=1e-2, out="p1/op_dde/x", precision="float32",
                                 run_kw=dict(vectorize=False)),
    "dde_net10_dopri5":     dict(build=_yaml(TB + "net10"), T=3.0, dt=1e-3, dts=5e-2, out="pop0/op8/a", run_kw=dict(vectorize=False)),
    "dde_mg_dopri5":        dict(build=_mackey_glass(), T=0.3, dt=1e-4, dts=1e-3, out="p1/mg_op/x", run_kw=dict(vectorize=False)),
    "dde_jrc_pair_dopri5":  dict(build=_jrc_pair_delayed(), T=0.2, dt=1e-4, dts=1e-3, out="jrc1/pc/rpo_e_in/v",
                                 run_kw=dict(vectorize=False)),
    # SciPy's DOP853 (8th order; the reference's documentation runs it: documentation/model_analysis/entrainment.py:56)
    "jrc_dop853":   dict(build=_yaml(NM + "jansenrit.JRC"), T=0.5, dt=1e-4, dts=1e-3, out="pc/rpo_e_in/v",
                         kw=dict(method="DOP853", rtol=1e-8, atol=1e-10)),
    "qif_sfa_dop853": dict(build=_yaml(NM + "qif.qif_sfa"), T=2.0, dt=1e-3, dts=5e-3, out="p/qif_sfa_op/r", kw=dict(method="DOP853")),
    "vdp_dop853":   dict(build=_yaml("model_templates.oscillators.vanderpol.vdp"), T=20.0, dt=1e-3, dts=5e-2, out="p/vdp_op/x",
                         kw=dict(method="DOP853", rtol=1e-6, atol=1e-9)),
    # networks too large for one thread: adaptive integration in the persistent network kernel (grid-wide error norm)
    "wc_n128_rk45": dict(build=_wc_pop(128), T=60.0, dt=1e-3, dts=0.5, out="e/rate_op/r", kw=dict(rtol=1e-7, atol=1e-10)),
    "kuramoto_n128_rk45": dict(build=_kuramoto(128), T=10.0, dt=1e-3, dts=0.05, out="e/phase_op/theta",
                               kw=dict(rtol=1e-8, atol=1e-10)),
}


def make_adaptive_case(name: str, spec: dict, cap: Capture) -> str:
    from pyrates import clear
    from pyrates.ir.node import clear_ir_caches
    T, dt, dts, kw = spec["T"], spec["dt"], spec.get("dts"), spec.get("kw", {})
    precision = spec.get("precision", "float64")
    clear_ir_caches()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            circuit = spec["build"]()
            circuit.run(simulation_time=T, step_size=dt, sampling_step_size=dts, solver="scipy", outputs={"o": spec["out"]},
                        inputs=spec["inputs"]() if "inputs" in spec else None,
                        backend="default", float_precision=precision, verbose=False, clear=True, **kw,
                        **spec.get("run_kw", {}))
        finally:
            os.chdir(cwd)
    prog = Program.from_source(cap.source, list(zip(cap.names, cap.func_args)), state_vars=cap.state_vars,
                               float_precision=precision)
    rec = cap.rec
    # the restatement (oracle.solve_rk45) must reproduce reference + SciPy bit for bit
    func = orc.build_vector_field(prog.source())
    args = _fresh(cap.func_args)
    args[0] = args[0][()]
    mine, _ = orc.run_adaptive(func, args, T, dt, dts, **kw)
    diff = float(np.max(np.abs(mine - rec)))
    assert mine.shape == rec.shape and diff == 0.0, f"{name}: oracle RK45 != reference+scipy (max abs diff {diff})"
    try:
        clear(circuit)
    except Exception:
        pass
    run_meta = {"T": T, "dt": dt, "dts": dts, "solver": "scipy", "method": "dopri5" if prog.is_dde else kw.get("method", "RK45"), "precision": precision,
                "rtol": kw.get("rtol", 1e-3), "atol": kw.get("atol", 1e-6), "reference_solvers": ["scipy"],
                "reference_version": "1.2.3", "scipy_version": __import__("scipy").__version__, "out": spec["out"]}
    path = os.path.join(GOLDEN, f"{name}.npz")
    prog.save(path, __run__=np.frombuffer(json.dumps(run_meta).encode(), dtype=np.uint8), rec_scipy=rec)
    return path


# Jacobian functions (ComputeGraph.get_jacobian_func, computegraph.py:516-763): the reference differentiates the vector field
# symbolically and has the backend print `J0 = zeros((n, n)); J0[i, j] = expr; return J0`
def _transcendental():
    def build():
        from pyrates import CircuitTemplate, NodeTemplate, OperatorTemplate
        op = OperatorTemplate(name="trans_op", equations=["u' = -u + exp(-a*u) + cos(w)", "w' = -w + sin(u)"],
                              variables={"u": "output(0.1)", "w": "variable(0.2)", "a": 1.5})      # tests/test_jacobian.py:185-192
        return CircuitTemplate(name="trans_circ", nodes={"p": NodeTemplate(name="trans_pop", operators=[op])})
    return build


JACOBIAN_CASES = {
    "jac_vdp":   dict(build=_yaml("model_templates.oscillators.vanderpol.vdp"), scale=1.0),
    "jac_trans": dict(build=_transcendental(), scale=0.5),
    "jac_jrc":   dict(build=_yaml(NM + "jansenrit.JRC"), scale=5e-3),
    "jac_qif_sfa_f32": dict(build=_yaml(NM + "qif.qif_sfa"), scale=0.3, precision="float32"),
}


def make_jacobian_case(name: str, spec: dict, cap: Capture) -> str:
    import ast
    from pyrates import clear
    from pyrates.ir.node import clear_ir_caches
    from pyrates_b200.program import Arg, Stmt
    precision = spec.get("precision", "float64")
    clear_ir_caches()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            circuit = spec["build"]()
            func, args, names, sv = circuit.get_jacobian_func(func_name="jacobian", step_size=1e-3, solver="euler", in_place=False,
                                                              clear=False, vectorize=False, backend="default",
                                                              float_precision=precision, verbose=False)
        finally:
            os.chdir(cwd)
    fdef = next(n for n in ast.parse(cap.source.replace("\t", "    ")).body if isinstance(n, ast.FunctionDef))
    sig = [a.arg for a in fdef.args.args]
    assert len(sig) == len(args) and sig[:2] == ["t", "y"]
    stmts, jname, shape = [], None, None
    for node in fdef.body:
        if isinstance(node, ast.Return):
            assert isinstance(node.value, ast.Name) and node.value.id == jname
        elif isinstance(node.targets[0], ast.Name):              # J0 = zeros((n, n), dtype=...)
            jname, shape = node.targets[0].id, tuple(ast.literal_eval(node.value.args[0]))
        else:
            full = ast.unparse(node.targets[0])
            stmts.append(Stmt(jname, ast.unparse(node.value), full[len(jname) + 1:-1]))
    n = shape[0]
    pargs = [Arg("t", "t", np.asarray(0.0)), Arg("y", "y", np.asarray(args[1], dtype=precision)),
             Arg(jname, "dy", np.zeros(shape, dtype=precision))]
    pargs += [Arg(s_, "const", np.array(a, copy=True), lab) for s_, a, lab in zip(sig[2:], args[2:], names[2:])]
    prog = Program(pargs, stmts, dict(sv), precision, "jacobian")
    # the reference function at its own state and at 15 perturbed states (float64 states, as users pass them)
    rng = np.random.default_rng(7)
    y0 = np.asarray(args[1], dtype=np.float64)
    pts = np.vstack([y0[None, :], y0[None, :] + spec["scale"] * rng.normal(size=(15, n))])
    ref = np.stack([np.asarray(func(0.0, y.copy(), *args[2:])) for y in pts])
    # the oracle executes the regenerated text (output buffer passed in instead of allocated inside): identical numbers
    f2 = orc.build_vector_field(prog.source(), "jacobian")
    mine = np.stack([np.array(f2(0.0, y.copy(), np.zeros(shape, dtype=precision), *args[2:])) for y in pts])
    assert ref.dtype == mine.dtype and np.array_equal(ref, mine, equal_nan=True), f"{name}: oracle != reference Jacobian"
    try:
        clear(circuit)
    except Exception:
        pass
    run_meta = {"kind": "jacobian", "precision": precision, "reference_version": "1.2.3", "solver": "euler", "T": 0.0, "dt": 1e-3,
                "dts": None}
    path = os.path.join(GOLDEN, f"{name}.npz")
    prog.save(path, __run__=np.frombuffer(json.dumps(run_meta).encode(), dtype=np.uint8), jac_points=pts, jac_ref=ref)
    return path


def main(argv):
    warnings.filterwarnings("ignore")
    if import_reference() is None:
        raise SystemExit("reference PyRates not importable here; fixtures can only be regenerated in the build container")
    os.makedirs(GOLDEN, exist_ok=True)
    names = argv or (list(CASES) + list(ADAPTIVE_CASES) + list(JACOBIAN_CASES))
    cap = Capture()
    try:
        for n in names:
            if n in JACOBIAN_CASES:
                p = make_jacobian_case(n, JACOBIAN_CASES[n], cap)
            else:
                p = make_adaptive_case(n, ADAPTIVE_CASES[n], cap) if n in ADAPTIVE_CASES else make_case(n, CASES[n], cap)
            print(f"{n:24s} -> {os.path.relpath(p, _ROOT)}  ({os.path.getsize(p) / 1024:.1f} KiB)")
    finally:
        cap.close()


if __name__ == "__main__":
    main(sys.argv[1:])
