"""Drop-in boundary: with the reference importable, `backend='cuda'` makes PyRates' own compute graph drive
CudaBackend, which must record exactly the program the NumPy backend would have printed (same statements,
same argument arrays).  CPU part: capture only.  GPU part (needs baseline/_ref on the box): the public
CircuitTemplate.run / grid_search calls end to end against the reference NumPy backend."""
import ast
import json
import os
import warnings

import numpy as np
import pytest

from conftest import RTOL, golden_path, traj_rel_err
from oracle import numpy_oracle as orc
from oracle.ref_env import import_reference
from pyrates_b200 import runtime as rt
from pyrates_b200.program import Program

pyrates = import_reference()
needs_ref = pytest.mark.skipif(pyrates is None, reason="reference PyRates not importable on this machine")
warnings.filterwarnings("ignore")


def _fresh(path):
    """from_yaml returns a cached template object that run() mutates in place: work on private copies."""
    from copy import deepcopy
    from pyrates import CircuitTemplate
    return deepcopy(CircuitTemplate.from_yaml(path))


def _norm(stmts):
    return [ast.dump(ast.parse(s.text())) for s in stmts]


@pytest.fixture()
def registered(tmp_path, monkeypatch):
    import pyrates_b200.register as reg
    assert reg.register()
    monkeypatch.chdir(tmp_path)
    yield reg
    from pyrates.ir.node import clear_ir_caches
    clear_ir_caches()


@needs_ref
@pytest.mark.parametrize("name,path,solver", [
    ("jrc", "model_templates.neural_mass_models.jansenrit.JRC", "rk4"),
    ("qif", "model_templates.neural_mass_models.qif.qif", "euler"),
    ("net10", "model_templates.test_resources.test_backend.net10", "heun"),
    ("net13", "model_templates.test_resources.test_backend.net13", "heun_true"),
    ("jrc_2delay", "model_templates.neural_mass_models.jansenrit.JRC_2delaycoupled", "euler"),
])
def test_cuda_backend_records_same_program_as_numpy_backend(registered, name, path, solver):
    from pyrates import clear
    c = _fresh(path)
    fx = orc.load_fixture(golden_path(name))
    func, args, names, svmap = c.get_run_func("vector_field", step_size=fx["run"]["dt"], backend="cuda", solver=solver,
                                              float_precision="float64", verbose=False)
    prog = func.program(args)
    gold = Program.load(golden_path(name))
    assert _norm(prog.stmts) == _norm(gold.stmts)
    assert [(a.name, a.kind) for a in prog.args] == [(a.name, a.kind) for a in gold.args]
    for a, b in zip(prog.args, gold.args):
        assert a.value.dtype == b.value.dtype and np.array_equal(a.value, b.value), a.name
    clear(c)


@needs_ref
@pytest.mark.parametrize("name,solver", [("dde_net16", "euler"), ("dde_net16_dopri5", "scipy"), ("dde_net10_dopri5", "scipy")])
def test_cuda_backend_records_delay_differential_programs(registered, name, solver):
    """`x(t-d)` terms / delayed edges under an adaptive solver: ComputeGraph.to_func drives add_var_hist and asks for the
    `hist` argument (computegraph.py:424-450, 471-478); the recorded program (history lookups first, `(t, y, hist, dy, ..)`
    signature, float64 initial condition as the history's value) is the one the NumPy backend printed."""
    from pyrates import clear
    path = "model_templates.test_resources.test_backend." + name.split("_")[1]
    c = _fresh(path)
    fx = orc.load_fixture(golden_path(name))
    func, args, names, svmap = c.get_run_func("vector_field", step_size=fx["run"]["dt"], backend="cuda", solver=solver,
                                              float_precision="float64", verbose=False, vectorize=False)
    prog = func.program(args)
    gold = Program.load(golden_path(name))
    assert prog.is_dde and names[2] == "hist"
    assert _norm(prog.stmts) == _norm(gold.stmts)
    assert [(a.name, a.kind) for a in prog.args] == [(a.name, a.kind) for a in gold.args]
    for a, b in zip(prog.args, gold.args):
        assert a.value.dtype == b.value.dtype and np.array_equal(a.value, b.value), a.name
    with pytest.raises(Exception, match="lives on the device"):
        args[2](0.0)
    clear(c)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("path,out,solver,precision,T,dts", [
    ("net16", "p1/op_dde/x", "euler", "float64", 0.5, 1e-3),       # the reference's own DDE test (test_2_8_dde)
    ("net16", "p1/op_dde/x", "heun", "float32", 0.5, 2e-3),
    ("net16", "p1/op_dde/x", "scipy", "float64", 2.0, 1e-2),
    ("net16", "p1/op_dde/x", "scipy", "float32", 2.0, 1e-2),
    ("net10", "pop0/op8/a", "scipy", "float64", 3.0, 5e-2),        # delayed edges under an adaptive solver -> DDE
])
def test_delay_differential_models_match_numpy_backend(gpu, registered, path, out, solver, precision, T, dts):
    """CircuitTemplate.run on delay-differential models: `x(t-d)` notation with fixed-step solvers (history ring on the
    device) and solver='scipy' (dopri5 + growing history, BaseBackend._solve_scipy_dde) vs the reference NumPy backend."""
    kw = dict(simulation_time=T, step_size=1e-3, sampling_step_size=dts, solver=solver, outputs={"o": out}, verbose=False,
              float_precision=precision, clear=True, vectorize=False)
    model = "model_templates.test_resources.test_backend." + path
    ref = _fresh(model).run(backend="default", **kw)
    got = _fresh(model).run(backend="cuda", **kw)
    assert got.shape == ref.shape and got.values.dtype == ref.values.dtype
    assert traj_rel_err(got.values, ref.values) <= RTOL[precision]
    if path == "net16" and solver == "euler":
        x = got["o"].values.squeeze()                 # the assertions of tests/test_backend_simulations.py:568-583
        np.testing.assert_allclose(x[:100], 1.0 - np.arange(100) * 1e-3, atol=5e-3)
        assert x[0] == pytest.approx(1.0, abs=1e-5) and np.all(np.abs(x) < 5.0)


@needs_ref
@pytest.mark.parametrize("name,path,precision", [("jac_vdp", "model_templates.oscillators.vanderpol.vdp", "float64"),
                                                 ("jac_jrc", "model_templates.neural_mass_models.jansenrit.JRC", "float64"),
                                                 ("jac_qif_sfa_f32", "model_templates.neural_mass_models.qif.qif_sfa", "float32")])
def test_cuda_backend_records_jacobian_functions(registered, name, path, precision):
    """CircuitTemplate.get_jacobian_func(backend='cuda'): the graph differentiates symbolically and drives
    emit_local_array_alloc / emit_local_array_assign (computegraph.py:704-729); the recorded `J0[i, j] = expr` stream and
    the parameter arguments are the ones the NumPy backend printed.  sparse=True is refused like upstream's array backends."""
    from pyrates import clear
    c = _fresh(path)
    func, args, names, svmap = c.get_jacobian_func(func_name="jacobian", step_size=1e-3, solver="euler", in_place=False,
                                                   clear=False, vectorize=False, backend="cuda", float_precision=precision,
                                                   verbose=False)
    gold = Program.load(golden_path(name))
    prog = func.program(args)
    assert names[:2] == ("t", "y") and "def jacobian(" in func.source()
    assert _norm(prog.stmts) == _norm(gold.stmts)
    assert [(a.name, a.kind, a.value.shape) for a in prog.args] == [(a.name, a.kind, a.value.shape) for a in gold.args]
    for a, b in zip(prog.args[3:], gold.args[3:]):
        assert a.value.dtype == b.value.dtype and np.array_equal(a.value, b.value), a.name
    clear(c)
    with pytest.raises(NotImplementedError, match="sparse"):
        _fresh(path).get_jacobian_func(func_name="jac_sp", step_size=1e-3, solver="euler", in_place=False, clear=True,
                                       vectorize=False, backend="cuda", sparse=True, verbose=False)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("solver", ["euler", "scipy"])
def test_grid_search_over_a_history_delay_matches_reference_runs(gpu, registered, solver):
    """pyrates.grid_search(backend='cuda') over the delay d of x' = -x(t-d) (the reference's DDE test model): one batch, every
    grid point with its own delay and history; checked against the reference NumPy backend run once per grid point."""
    import pyrates
    from pyrates.utility import adapt_circuit
    path = "model_templates.test_resources.test_backend.net16"
    grid = {"d": np.linspace(0.05, 0.4, 8)}
    pm = {"d": {"vars": ["op_dde/d"], "nodes": ["p1"]}}
    kw = dict(step_size=1e-3, simulation_time=1.0, sampling_step_size=1e-2, solver=solver, float_precision="float64")
    res, pmap = pyrates.grid_search(path, grid, pm, outputs={"x": "p1/op_dde/x"}, backend="cuda", verbose=False, **kw)
    assert res.shape == (100, 8)
    for i in (0, 3, 7):
        c = adapt_circuit(path, {"d": float(grid["d"][i])}, pm)
        ref = c.run(outputs={"x": "p1/op_dde/x"}, backend="default", verbose=False, clear=True, vectorize=False, **kw)
        assert traj_rel_err(res.iloc[:, [i]].values, ref.values) <= 1e-9, (solver, i)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("precision", ["complex128", "complex64"])
def test_grid_search_over_a_complex_valued_model_matches_reference_runs(gpu, registered, precision):
    """grid_search over the Ott-Antonsen mean field (complex state z, float_precision='complex64/128'): swept parameters are
    the real parts of complex-typed constants, results come back as complex frames; vs the reference run per grid point."""
    import pyrates
    from pyrates.utility import adapt_circuit, linearize_grid
    path = "model_templates.oscillators.kuramoto.kmo_mf"
    grid = linearize_grid({"Delta": np.linspace(0.5, 2.0, 4), "omega": np.linspace(5.0, 15.0, 3)}, permute=True)
    pm = {"Delta": {"vars": ["kmo_op/Delta"], "nodes": ["p"]}, "omega": {"vars": ["kmo_op/omega"], "nodes": ["p"]}}
    kw = dict(step_size=1e-3, simulation_time=1.0, sampling_step_size=1e-2, solver="heun", float_precision=precision)
    res, pmap = pyrates.grid_search(path, grid, pm, outputs={"z": "p/kmo_op/z"}, backend="cuda", verbose=False, **kw)
    assert res.shape == (100, 12) and np.iscomplexobj(res.values)
    from conftest import as_real_columns
    for i in (0, 5, 11):
        c = adapt_circuit(path, {k: float(grid[k].values[i]) for k in pm}, pm)
        ref = c.run(outputs={"z": "p/kmo_op/z"}, backend="default", verbose=False, clear=True, **kw)
        assert res.values.dtype == ref.values.dtype
        assert traj_rel_err(as_real_columns(res.iloc[:, [i]].values), as_real_columns(ref.values)) <= RTOL[precision], i


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("precision", ["float64", "float32"])
def test_documentation_parameter_sweep_example_runs_unchanged_on_cuda(gpu, registered, precision):
    """The reference's own gallery example (documentation/model_analysis/parameter_sweeps.py:44-88): a sweep over the
    Jansen-Rit connectivity scaling C with a shared noise input, two outputs and a cutoff — the call is repeated verbatim with
    backend='cuda' (shorter horizon) and must return the same frames (labels, index, values) as backend='default'."""
    import pyrates
    T, dt, dts, cutoff = 0.5, 1e-4, 1e-3, 0.1
    param_grid = {'C': [68., 128., 135., 270., 675., 1350.]}
    param_map = {'C': {'vars': ['jrc_op/c'], 'nodes': ['jrc']}}
    noise = np.random.default_rng(0).uniform(120.0, 320.0, size=(int(np.round(T / dt, decimals=0)), 1))
    kw = dict(circuit_template="model_templates.neural_mass_models.jansenrit.JRC2", param_grid=param_grid, param_map=param_map,
              simulation_time=T, step_size=dt, sampling_step_size=dts, inputs={'jrc/jrc_op/u': noise},
              outputs={'V_pce': 'jrc/jrc_op/V_e', 'V_pci': 'jrc/jrc_op/V_i'}, solver="euler", cutoff=cutoff, verbose=False,
              float_precision=precision)
    ref, ref_map = pyrates.grid_search(backend="default", **kw)
    from pyrates.ir.node import clear_ir_caches
    clear_ir_caches()
    got, got_map = pyrates.grid_search(backend="cuda", **kw)
    assert [tuple(map(str, c)) for c in got.columns] == [tuple(map(str, c)) for c in ref.columns]
    assert got.shape == ref.shape and np.allclose(got.index.values, ref.index.values)
    assert list(got_map.index) == list(ref_map.index) and np.array_equal(got_map["C"].values, ref_map["C"].values)
    assert traj_rel_err(got.values, ref.values) <= RTOL[precision]


@pytest.mark.gpu
@needs_ref
def test_documentation_entrainment_example_dop853_on_cuda(gpu, registered):
    """The reference's gallery example documentation/model_analysis/entrainment.py:28-58,128-145: a Van der Pol oscillator
    driven by a Kuramoto oscillator, integrated with solver='scipy', method='DOP853' — the single run and a (smaller)
    omega x J grid_search, repeated with backend='cuda' and compared with the reference (SciPy's DOP853)."""
    import pyrates
    from pyrates import CircuitTemplate, NodeTemplate
    from pyrates.utility import adapt_circuit
    from copy import deepcopy

    def make():
        vpo = NodeTemplate.from_yaml("model_templates.oscillators.vanderpol.vdp_pop")
        ko = NodeTemplate.from_yaml("model_templates.oscillators.kuramoto.sin_pop")
        net = CircuitTemplate(name="VPO_forced", nodes={'VPO': vpo, 'KO': ko},
                              edges=[('KO/sin_op/s', 'VPO/vdp_op/inp', None, {'weight': 1.0})])
        net.update_var(node_vars={'KO/phase_op/omega': 0.5, 'VPO/vdp_op/mu': 2.0})
        return net
    # tight solver tolerances: at SciPy's defaults (rtol 1e-3) two valid runs that differ in one accept / reject decision —
    # e.g. parameters folded as constants vs passed per grid point — differ by the solver tolerance itself
    kw = dict(simulation_time=60.0, step_size=1e-3, solver='scipy', method='DOP853', cutoff=10.0, sampling_step_size=1e-2,
              outputs={'VPO': 'VPO/vdp_op/x', 'KO': 'KO/phase_op/theta'}, float_precision="float64", verbose=False,
              rtol=1e-9, atol=1e-11)
    ref = make().run(backend="default", in_place=False, **kw)
    got = make().run(backend="cuda", in_place=False, **kw)
    assert got.shape == ref.shape and list(got.columns) == list(ref.columns)
    # tolerance: DOP853's dense output under the aliased rhs buffer depends on the step sequence (see
    # test_adaptive_rk45_matches_reference_scipy); the REFERENCE against itself with mu + 1 ulp differs by 6.7e-4 on this run
    assert traj_rel_err(got.values, ref.values) <= 5e-3
    # (the gallery's J = 0 column is left out: the unforced oscillator starts on its unstable fixed point and only leaves it
    # through rounding noise, so its trajectory is not reproducible between any two builds)
    params = {'omega': np.linspace(0.3, 0.5, 3), 'J': np.linspace(0.5, 2.0, 4)}
    pm = {'omega': {'vars': ['phase_op/omega'], 'nodes': ['KO']},
          'J': {'vars': ['weight'], 'edges': [('KO/sin_op/s', 'VPO/vdp_op/inp')]}}
    res, res_map = pyrates.grid_search(circuit_template=make(), param_grid=params, param_map=pm, inputs=None, vectorize=True,
                                       permute_grid=True, backend="cuda", **kw)
    assert res.shape[1] == 2 * 12
    for i in (0, 7, 11):
        c = adapt_circuit(make(), {k: float(res_map[k].values[i]) for k in pm}, pm)
        one = c.run(backend="default", in_place=False, **kw)
        col = res_map.index[i]
        for key in ("VPO", "KO"):
            # per-point reference runs fold the parameters as constants, the grid passes them per instance: rounding differs
            # in the last bit, one accept / reject decision flips, and DOP853's aliasing-compromised dense output moves the
            # samples — the REFERENCE against itself with mu + 1 ulp differs by 6.7e-4 on this very run (measured)
            a, b = np.ravel(res[key][col].values), np.ravel(one[key].values)      # (n, 1) frame vs (n,) series
            assert a.shape == b.shape and traj_rel_err(a[:, None], b[:, None]) <= 5e-3, (i, key)


@needs_ref
def test_generated_function_file_is_written_and_cleared_like_the_reference(registered, tmp_path):
    """`file_name=` / `clear=`: like the NumPy backend (base_backend.py:316-324, 537-578, 613-627) the recorded vector field is
    written to `<file_name>.py` (users and the reference's documentation print it) and removed again by clear()."""
    from pyrates import clear
    c = _fresh("model_templates.neural_mass_models.qif.qif")
    func, args, names, _ = c.get_run_func("vector_field", step_size=1e-3, backend="cuda", solver="euler", verbose=False,
                                          clear=False, file_name="qif_generated")
    path = tmp_path / "qif_generated.py"
    assert path.exists()
    text = path.read_text()
    assert "def vector_field(t,y,dy" in text and "dy[0] =" in text and "return dy" in text
    src = func.source()
    assert src.startswith("def vector_field(") and src.rstrip().endswith("return dy")
    assert all(st.text() in src for st in func.stmts)
    clear(c)
    assert not path.exists()


@needs_ref
def test_new_solver_names_are_fixed_step(registered):
    import pyrates.frontend.template.circuit as fc
    assert not fc.is_integration_adaptive("rk4") and not fc.is_integration_adaptive("heun_true")
    assert not fc.is_integration_adaptive("euler") and fc.is_integration_adaptive("scipy")


@needs_ref
def test_unsupported_solver_and_no_device_raise(registered):
    from pyrates import CircuitTemplate
    c = _fresh("model_templates.neural_mass_models.qif.qif")
    with pytest.raises(Exception, match="does not support solver"):
        c.run(simulation_time=0.01, step_size=1e-3, solver="julia_dp5", outputs={"r": "p/qif_op/r"}, backend="cuda",
              verbose=False, in_place=False)
    try:
        rt.device_count()
    except rt.CudaBackendError:
        c = _fresh("model_templates.neural_mass_models.qif.qif")
        with pytest.raises(rt.CudaBackendError, match="no CPU execution path"):
            c.run(simulation_time=0.01, step_size=1e-3, solver="euler", outputs={"r": "p/qif_op/r"}, backend="cuda",
                  verbose=False)


# ------------------------------------------------------------------------------------------- GPU end to end
@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("path,out,solver,T,dt,precision", [
    ("model_templates.neural_mass_models.qif.qif", "p/qif_op/r", "euler", 1.0, 1e-3, "float64"),
    ("model_templates.neural_mass_models.qif.qif", "p/qif_op/v", "heun", 1.0, 1e-3, "float32"),
    ("model_templates.neural_mass_models.jansenrit.JRC", "pc/rpo_e_in/v", "heun", 0.2, 1e-4, "float64"),
    ("model_templates.test_resources.test_backend.net10", "pop0/op8/a", "euler", 1.5, 1e-3, "float64"),
    ("model_templates.neural_mass_models.qif.qif_sfa", "p/qif_sfa_op/r", "euler", 1.0, 1e-3, "float64"),
])
def test_circuit_run_cuda_matches_numpy_backend(gpu, registered, path, out, solver, T, dt, precision):
    """tests/test_backend_simulations.py:527 pattern (cross-backend parity) with the north-star tolerance."""
    from pyrates import CircuitTemplate
    kw = dict(simulation_time=T, step_size=dt, sampling_step_size=dt * 5, solver=solver, outputs={"o": out},
              verbose=False, float_precision=precision, clear=True)
    ref = _fresh(path).run(backend="default", **kw)
    got = _fresh(path).run(backend="cuda", **kw)
    assert list(got.columns) == list(ref.columns) and np.allclose(got.index, ref.index)
    assert traj_rel_err(got.values, ref.values) <= RTOL[precision]


@pytest.mark.gpu
@needs_ref
def test_circuit_run_with_inputs_and_state_writeback(gpu, registered):
    from pyrates import CircuitTemplate
    inp = np.sin(np.arange(500) * 0.02) * 3.0
    kw = dict(simulation_time=0.5, step_size=1e-3, solver="heun", outputs={"r": "p/qif_op/r", "v": "p/qif_op/v"},
              inputs={"p/qif_op/I_ext": inp}, verbose=False, float_precision="float64", clear=True)
    c_ref, c_gpu = (_fresh("model_templates.neural_mass_models.qif.qif") for _ in range(2))
    ref = c_ref.run(backend="default", **kw)
    got = c_gpu.run(backend="cuda", **kw)
    assert traj_rel_err(got.values, ref.values) <= 1e-9
    for key, val in c_ref.state.items():                       # final stored sample written back (computegraph.py:503-507)
        np.testing.assert_allclose(np.asarray(c_gpu.state[key], dtype=float), np.asarray(val, dtype=float), rtol=1e-9)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("solver", ["euler", "heun"])
def test_grid_search_cuda_matches_reference_grid_search(gpu, registered, solver):
    """New compile-once grid_search vs the frames the reference's own grid_search produced (golden)."""
    import pyrates
    from pyrates.utility import linearize_grid
    fx = orc.load_fixture(golden_path("jrc_grid_b16"))
    grid = linearize_grid({"C": np.linspace(68.0, 1350.0, 4), "u": np.linspace(120.0, 320.0, 4)}, permute=True)
    grid["C2"], grid["C4"] = 0.8 * grid["C"], 0.25 * grid["C"]
    param_map = {"C": {"vars": ["weight"], "edges": [("pc/pro/m", "ein/rpo_e/m_in")]},
                 "C2": {"vars": ["weight"], "edges": [("ein/pro/m", "pc/rpo_e_in/m_in")]},
                 "C4": {"vars": ["weight"], "edges": [("pc/pro/m", "iin/rpo_e/m_in"), ("iin/pro/m", "pc/rpo_i/m_in")]},
                 "u": {"vars": ["rpo_e_in/u"], "nodes": ["pc"]}}
    res, pmap = pyrates.grid_search("model_templates.neural_mass_models.jansenrit.JRC", grid, param_map, step_size=1e-4,
                                    simulation_time=0.05, outputs={"v": "pc/rpo_e_in/v"}, sampling_step_size=1e-3,
                                    solver=solver, backend="cuda", float_precision="float64", verbose=False)
    ref_cols = [tuple(c) for c in json.loads(bytes(fx["extra"]["frame_columns"]).decode())]
    assert [tuple(map(str, c)) for c in res.columns] == ref_cols
    assert list(pmap.index) == [f"JRC_{i}" for i in range(16)]
    np.testing.assert_array_equal(pmap["C"].values, fx["extra"]["grid_C"])
    assert traj_rel_err(res.values, fx["extra"][f"frame_{solver}"]) <= 1e-9


@pytest.mark.gpu
@needs_ref
def test_grid_search_reduce_equals_statistics_of_the_reference_frames(gpu, registered):
    """grid_search(reduce=True, cutoff=...): per-grid-point mean / var / min / max / last computed on the device == the same
    statistics of the frame the REFERENCE's grid_search produced (golden), same column labels."""
    import pyrates
    from pyrates.utility import linearize_grid
    fx = orc.load_fixture(golden_path("jrc_grid_b16"))
    grid = linearize_grid({"C": np.linspace(68.0, 1350.0, 4), "u": np.linspace(120.0, 320.0, 4)}, permute=True)
    grid["C2"], grid["C4"] = 0.8 * grid["C"], 0.25 * grid["C"]
    param_map = {"C": {"vars": ["weight"], "edges": [("pc/pro/m", "ein/rpo_e/m_in")]},
                 "C2": {"vars": ["weight"], "edges": [("ein/pro/m", "pc/rpo_e_in/m_in")]},
                 "C4": {"vars": ["weight"], "edges": [("pc/pro/m", "iin/rpo_e/m_in"), ("iin/pro/m", "pc/rpo_i/m_in")]},
                 "u": {"vars": ["rpo_e_in/u"], "nodes": ["pc"]}}
    cutoff = 0.02
    res, pmap = pyrates.grid_search("model_templates.neural_mass_models.jansenrit.JRC", grid, param_map, step_size=1e-4,
                                    simulation_time=0.05, outputs={"v": "pc/rpo_e_in/v"}, sampling_step_size=1e-3,
                                    solver="heun", backend="cuda", float_precision="float64", verbose=False, reduce=True,
                                    cutoff=cutoff)
    ref_cols = [tuple(c) for c in json.loads(bytes(fx["extra"]["frame_columns"]).decode())]
    assert [tuple(map(str, c)) for c in res.columns] == ref_cols and list(res.index) == ["mean", "var", "min", "max", "last"]
    frame = fx["extra"]["frame_heun"]
    times = np.linspace(0.0, 0.05, num=frame.shape[0], endpoint=False)
    keep = frame[times >= cutoff]
    want = np.stack([keep.mean(axis=0), keep.var(axis=0), keep.min(axis=0), keep.max(axis=0), keep[-1]])
    scale = np.max(np.abs(keep), axis=0)
    assert np.max(np.abs(res.values - want) / np.stack([scale, scale ** 2, scale, scale, scale])) <= 1e-9


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("precision,solver", [("float64", "heun"), ("float32", "euler")])
def test_grid_search_over_population_network_matches_reference_runs(gpu, registered, precision, solver):
    """grid_search over a *network* (WC population, dense Connectivity shared by all grid points): the grid steps as one
    batch in the persistent network kernel, couplings as one GEMM per evaluation (DMMA in fp64, tcgen05 3xTF32 in
    fp32).  Checked against the reference NumPy backend run once per sampled grid point."""
    import pyrates
    from copy import deepcopy
    from pyrates.utility import linearize_grid
    from oracle.make_golden import _wc_pop
    tpl = _wc_pop(96)()
    grid = linearize_grid({"tau": np.linspace(8.0, 12.0, 8), "k": np.linspace(0.8, 1.2, 5)}, permute=True)     # 40 points
    pm = {"tau": {"vars": ["rate_op/tau"], "nodes": ["e"]}, "k": {"vars": ["rate_op/k"], "nodes": ["e"]}}
    kw = dict(step_size=1e-3, simulation_time=0.05, sampling_step_size=2e-3, solver=solver, float_precision=precision)
    before = rt.launch_count()
    res, pmap = pyrates.grid_search(deepcopy(tpl), grid, pm, outputs={"r": "e/rate_op/r"}, backend="cuda",
                                    exec_model="network", verbose=False, **kw)
    assert rt.launch_count() == before + 1                      # the whole grid, the whole simulation: one launch
    assert res.shape == (25, 40 * 96) and list(pmap.index)[:2] == ["wc_n96_0", "wc_n96_1"]
    for i in (0, 17, 39):
        c = deepcopy(tpl)
        c.populations["e"].params["rate_op/tau"] = float(grid["tau"].values[i])
        c.populations["e"].params["rate_op/k"] = float(grid["k"].values[i])
        ref = c.run(outputs={"r": "e/rate_op/r"}, backend="default", verbose=False, clear=True, **kw)
        got = res.xs(f"wc_n96_{i}", axis=1, level=1)
        assert got.shape == ref.shape
        assert traj_rel_err(got.values, ref.values) <= RTOL[precision], (precision, i)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("path,out,T,dt,kw", [
    ("model_templates.neural_mass_models.qif.qif", "p/qif_op/r", 2.0, 1e-3, {}),
    ("model_templates.neural_mass_models.jansenrit.JRC", "pc/rpo_e_in/v", 0.5, 1e-4, dict(rtol=1e-7, atol=1e-9)),
    ("model_templates.neural_mass_models.qif.qif_sfa", "p/qif_sfa_op/r", 2.0, 1e-3, dict(method="RK45", rtol=1e-6, atol=1e-8)),
])
def test_circuit_run_scipy_solver_on_cuda_matches_solve_ivp(gpu, registered, path, out, T, dt, kw):
    """solver='scipy' end to end: the reference hands the vector field to scipy.solve_ivp (RK45); the CUDA backend runs
    the same step controller and dense output on the device."""
    run = dict(simulation_time=T, step_size=dt, sampling_step_size=dt * 10, solver="scipy", outputs={"o": out},
               verbose=False, float_precision="float64", clear=True, **kw)
    ref = _fresh(path).run(backend="default", **run)
    got = _fresh(path).run(backend="cuda", **run)
    assert got.shape == ref.shape and np.allclose(got.index, ref.index)
    assert traj_rel_err(got.values, ref.values) <= 1e-9


@pytest.mark.gpu
@needs_ref
def test_grid_search_with_adaptive_solver_matches_reference_runs(gpu, registered):
    """grid_search(solver='scipy'): every grid point runs its own RK45 step sequence in its own thread; checked against
    the reference's solver='scipy' run (SciPy solve_ivp) of individual grid points."""
    import pyrates
    from pyrates.utility import linearize_grid, adapt_circuit
    path = "model_templates.neural_mass_models.jansenrit.JRC"
    grid = linearize_grid({"u": np.linspace(120.0, 320.0, 12)}, permute=True)
    pm = {"u": {"vars": ["rpo_e_in/u"], "nodes": ["pc"]}}
    kw = dict(step_size=1e-4, simulation_time=0.3, sampling_step_size=1e-3, solver="scipy", float_precision="float64",
              rtol=1e-7, atol=1e-9)
    res, pmap = pyrates.grid_search(path, grid, pm, outputs={"v": "pc/rpo_e_in/v"}, backend="cuda", verbose=False, **kw)
    assert res.shape == (300, 12)
    for i in (0, 7, 11):
        c = adapt_circuit(path, {"u": float(grid["u"].values[i])}, pm)
        ref = c.run(outputs={"v": "pc/rpo_e_in/v"}, backend="default", verbose=False, clear=True, **kw)
        assert traj_rel_err(res.iloc[:, [i]].values, ref.values) <= 1e-7, i


@pytest.mark.gpu
@needs_ref
def test_population_network_run_with_scipy_solver_uses_the_adaptive_network_kernel(gpu, registered):
    """CircuitTemplate.run(solver='scipy') on a population with dense Connectivity: too large for one thread, so the
    whole network integrates adaptively in the persistent kernel; compared with the reference's solve_ivp run."""
    from copy import deepcopy
    from oracle.make_golden import _wc_pop
    from pyrates_b200.netengine import NetworkModel
    import pyrates_b200.backend as be
    tpl = _wc_pop(128)()
    kw = dict(simulation_time=40.0, step_size=1e-3, sampling_step_size=0.5, solver="scipy", outputs={"r": "e/rate_op/r"},
              verbose=False, float_precision="float64", clear=True, rtol=1e-7, atol=1e-10)
    ref = deepcopy(tpl).run(backend="default", **kw)
    seen = []
    orig = NetworkModel.__init__

    def spy(self, *a, **k):
        seen.append(k.get("solver"))
        return orig(self, *a, **k)
    NetworkModel.__init__ = spy
    try:
        got = deepcopy(tpl).run(backend="cuda", **kw)
    finally:
        NetworkModel.__init__ = orig
    assert seen == ["scipy"]
    assert got.shape == ref.shape == (80, 128)
    assert traj_rel_err(got.values, ref.values) <= 1e-9


NM, OSC = "model_templates.neural_mass_models.", "model_templates.oscillators."
ZOO = [  # every circuit of the reference's model_templates that the reference itself can build (3 have template errors);
         # theta.kmo_theta / kuramoto.kmo_mf* are built real-valued here (float_precision='float64': i = 0, conjugate = identity)
    (NM + "ik.ik_eta", "p/ik_eta_op/r", 30.0, 1e-2), (NM + "ik.ik_nodim", "p/ik_nodim_op/r", 1.0, 1e-3),
    (NM + "ik.ik_theta", "p/ik_theta_op/r", 30.0, 1e-2), (NM + "jansenrit.JRC2", "jrc/jrc_op/V_e", 0.2, 1e-4),
    (NM + "qif.qif_conduct", "p/qif_conduct_op/r", 1.0, 1e-3), (NM + "qif.qif_sd", "p/qif_op/r", 1.0, 1e-3),
    (NM + "qif.qif_tsodyks", "p/qif_op/v", 1.0, 1e-3), (NM + "wilsoncowan.WC", "e/rate_op/r", 20.0, 1e-2),
    (NM + "wilsoncowan.WC_stp", "e/rate_op/r", 20.0, 1e-2), (OSC + "kuramoto.kmo", "p/phase_op/theta", 2.0, 1e-3),
    (OSC + "kuramoto.kmo_2coupled", "p1/phase_op/theta", 2.0, 1e-3), (OSC + "stuartlandau.sl", "p/sl_op/x", 5.0, 1e-3),
    (OSC + "vanderpol.vdp", "p/vdp_op/x", 5.0, 1e-3), (NM + "theta.kmo_theta", "p/theta_op/z", 2.0, 1e-3),
    (OSC + "kuramoto.kmo_mf", "p/kmo_op/z", 2.0, 1e-3), (OSC + "kuramoto.kmo_mf_2coupled", "p1/kmo_op/z", 2.0, 1e-3),
]


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("solver", ["heun", "scipy"])
@pytest.mark.parametrize("path,out,T,dt", ZOO)
def test_model_zoo_runs_on_cuda_like_on_numpy(gpu, registered, path, out, T, dt, solver):
    """The reference's own model templates through the public run() call, both backends, fixed-step and adaptive."""
    kw = dict(simulation_time=T, step_size=dt, sampling_step_size=dt * 10, solver=solver, outputs={"o": out}, verbose=False,
              float_precision="float64", clear=True)
    if solver == "scipy":
        kw.update(rtol=1e-7, atol=1e-10)
    ref = _fresh(path).run(backend="default", **kw)
    got = _fresh(path).run(backend="cuda", **kw)
    assert got.shape == ref.shape
    if not np.all(np.isfinite(ref.values)):
        pytest.skip("the reference trajectory itself is not finite for the template defaults")
    assert traj_rel_err(got.values, ref.values) <= (1e-9 if solver == "heun" else 1e-7)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("path,out,precision,solver", [
    (OSC + "kuramoto.kmo_mf", "p/kmo_op/z", "complex128", "heun"),
    (NM + "theta.kmo_theta", "p/theta_op/z", "complex128", "euler"),
    (OSC + "kuramoto.kmo_mf_2coupled", "p1/kmo_op/z", "complex64", "euler"),
    (OSC + "kuramoto.kmo_mf_2coupled", "p2/kmo_op/z", "complex128", "scipy"),
])
def test_complex_precision_models_match_numpy_backend(gpu, registered, path, out, precision, solver):
    """float_precision='complex64/128': the complex state is carried as (re, im) pairs of real state variables, complex
    arithmetic is lowered to real operations; run() hands back complex frames like the reference."""
    kw = dict(simulation_time=2.0, step_size=1e-3, sampling_step_size=1e-2, solver=solver, outputs={"z": out}, verbose=False,
              float_precision=precision, clear=True)
    if solver == "scipy":
        kw.update(rtol=1e-8, atol=1e-10)
    ref = _fresh(path).run(backend="default", **kw)
    got = _fresh(path).run(backend="cuda", **kw)
    assert got.shape == ref.shape and np.iscomplexobj(got.values) and got.values.dtype == ref.values.dtype
    from conftest import as_real_columns
    # adaptive: SciPy's error norm uses complex moduli, the kernel's treats re / im as separate components
    tol = {"complex128": 1e-9, "complex64": 1e-4}[precision] if solver != "scipy" else 1e-6
    assert traj_rel_err(as_real_columns(got.values), as_real_columns(ref.values)) <= tol
