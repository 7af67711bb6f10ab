"""Front end on CPU: the scalarised DAG (what the instance kernels are printed from) evaluated with NumPy
must reproduce the reference's Euler trajectories for every fixture small enough to unroll."""
import numpy as np
import pytest

from conftest import as_real_columns, golden_names, golden_path, jacobian_names, traj_rel_err
from dag_eval import DagMachine
from oracle import numpy_oracle as orc
from pyrates_b200.program import Program
from pyrates_b200.scalarize import UnsupportedModelError, scalarize

LARGE = {"kuramoto_n128", "qsfa_both_n96", "wc_n128", "wc_n128_f32", "jrc_grid_b1024", "wc_n300_f32", "qsfa_both_n96_f32",
         "kuramoto_n128_f32"}


@pytest.mark.parametrize("name", [n for n in golden_names() if n not in LARGE and "_rk45" not in n and "_rk23" not in n
                                  and "_dopri5" not in n and "_dop853" not in n])
def test_dag_matches_reference_euler(name):
    fx = orc.load_fixture(golden_path(name))
    prog = Program.load(golden_path(name))
    rhs = scalarize(prog)
    r = fx["run"]
    ss = round((r["dts"] or r["dt"]) / r["dt"])
    n_steps = min(round(r["T"] / r["dt"]), 40 * ss)
    hist = orc.History(next(a.value for a in prog.args if a.kind == "hist")) if prog.is_dde else None
    rec = DagMachine(rhs, prog.real_dtype).euler(prog.y0, int(prog.arg_by_kind("t").value), n_steps, r["dt"], ss, hist=hist)
    ref = as_real_columns(fx["extra"]["rec_euler"])[: rec.shape[0]]
    tol = 1e-12 if prog.float_precision in ("float64", "complex128") else 1e-5
    assert traj_rel_err(rec, ref) <= tol


@pytest.mark.parametrize("swept", [False, True])
@pytest.mark.parametrize("name", jacobian_names())
def test_jacobian_dag_matches_reference(name, swept):
    """Jacobian programs (`J0[i, j] = expr` into an n x n output buffer): n*n outputs from n state inputs, unassigned
    entries zero; with every scalar constant turned into a runtime parameter (what evaluate_rhs does) the values agree."""
    fx = orc.load_fixture(golden_path(name))
    prog = Program.load(golden_path(name))
    sweep = [(a.name, 0) for a in prog.args if a.kind == "const" and a.value.size == 1] if swept else []
    rhs = scalarize(prog, sweep)
    n = prog.n_state
    assert rhs.n_state == n and rhs.n_out == n * n and len(rhs.dy) == n * n and len(rhs.params) == len(sweep)
    mach = DagMachine(rhs, prog.real_dtype)
    tol = 1e-12 if prog.float_precision == "float64" else 1e-5
    for y, ref in zip(fx["extra"]["jac_points"], fx["extra"]["jac_ref"]):
        J = mach.eval(0.0, y.astype(prog.real_dtype)).reshape(n, n)
        assert np.allclose(J, ref, rtol=tol, atol=tol * np.max(np.abs(ref)))


def test_complex_build_sweeps_real_parameters():
    """Complex precisions: every float constant arrives complex; a swept one (zero imaginary part) becomes a real runtime
    parameter, one with an imaginary part (the imaginary unit `i` of the Ott-Antonsen templates) cannot be swept."""
    fx = orc.load_fixture(golden_path("kmo_mf_c128"))
    prog = Program.load(golden_path("kmo_mf_c128"))
    rhs = scalarize(prog, [("Delta", 0), ("omega", 0)])
    assert sorted(p.arg for p in rhs.params) == ["Delta", "omega"] and rhs.n_state == 2
    with pytest.raises(UnsupportedModelError, match="real-valued"):
        scalarize(prog, [("i", 0)])
    mach = DagMachine(rhs, np.float64)
    vals = {"Delta": 0.7, "omega": 12.0}
    mach.params = np.array([vals[p.arg] for p in rhs.params])
    f = orc.build_vector_field(fx["source"])
    args = orc.fixture_args(fx)
    for k, v in vals.items():
        args[fx["names"].index(k)] = np.complex128(v)
    for y in (0.3 + 0.2j, -0.1 + 0.6j):
        ref = f(0, np.array([y]), *args[2:]).copy()
        got = mach.eval(0, np.array([y.real, y.imag]))
        assert np.allclose(got, [ref[0].real, ref[0].imag], rtol=1e-13)


def test_swept_history_delay_is_a_runtime_parameter():
    """grid_search over the delay of a delay-differential model: the lookup `hist(t*dt - d)` reads d from the instance's
    parameters; the history depth is sized from the caller's bound; the DAG reproduces the oracle for another delay."""
    from pyrates_b200.engine import InstanceModel
    fx = orc.load_fixture(golden_path("dde_mg"))
    prog = Program.load(golden_path("dde_mg"))
    rhs = scalarize(prog, [("tau", 0), ("bet", 0)])
    assert rhs.hist_delay_params == [0] and [p.arg for p in rhs.params] == ["tau", "bet"]
    with pytest.raises(UnsupportedModelError, match="hist_max_delay"):
        InstanceModel(prog, solver="euler", sweep=[("tau", 0)])
    m = InstanceModel(prog, solver="euler", sweep=[("tau", 0)], hist_max_delay=0.05)
    assert m.kernel.hist_depth == 512 and "T.p[0]" in m.kernel.source          # ceil(0.05 / 1e-4) + 3 = 503 -> 512
    mach = DagMachine(rhs, np.float64)
    mach.params = np.array([0.0123, 2.2])
    r = fx["run"]
    rec = mach.euler(prog.y0, 0, 300, r["dt"], 10, hist=orc.History(prog.arg("hist").value))
    args = orc.fixture_args(fx)
    args[fx["names"].index("tau")] = np.float64(0.0123)
    args[fx["names"].index("bet")] = np.float64(2.2)
    ref, _ = orc.run(orc.build_vector_field(fx["source"]), args, 300 * r["dt"], r["dt"], 10 * r["dt"], "euler")
    assert traj_rel_err(rec, ref) <= 1e-12


@pytest.mark.parametrize("name", sorted(LARGE))
def test_large_models_refuse_to_unroll(name):
    with pytest.raises(UnsupportedModelError):
        scalarize(Program.load(golden_path(name)))


def test_program_roundtrip_and_source():
    prog = Program.load(golden_path("net12"))
    p2 = prog.clone()
    assert p2.source() == prog.source()
    assert [a.kind for a in prog.args][:3] == ["t", "y", "dy"]
    assert prog.arg("a_buffer").kind == "var" and prog.arg("u").kind == "var"
    p3 = Program.from_source(prog.source(), [(a.name, a.value) for a in prog.args], state_vars=prog.state_vars)
    assert [s.text() for s in p3.stmts] == [s.text() for s in prog.stmts]


def test_sweep_params_become_leaves():
    prog = Program.load(golden_path("jrc"))
    rhs = scalarize(prog, sweep=[("u", 0), ("weight", 0), ("weight_v2", 1)])
    assert [(p.arg, p.elem) for p in rhs.params] == [("u", 0), ("weight_v2", 1), ("weight", 0)] or len(rhs.params) == 3
    with pytest.raises(KeyError):
        scalarize(prog, sweep=[("source_idx", 0)])


def test_unsupported_ops_raise():
    src = "def vector_field(t,y,dy,c):\n\ta = y[0]\n\tdy[0] = past(a, c)\n\treturn dy\n"
    prog = Program.from_source(src, [("t", np.int32(0)), ("y", np.zeros(1)), ("dy", np.zeros(1)), ("c", np.float64(1))])
    with pytest.raises(UnsupportedModelError):
        scalarize(prog)
