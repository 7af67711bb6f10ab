"""GPU parity (through the C ABI): fused rhs+solver instance kernels vs the golden trajectories of the
reference NumPy backend (euler, aliased heun) and the oracle-defined heun_true / rk4.

Tolerances are the north star's: rel 1e-9 in fp64, 1e-4 in fp32 (per-column scaled max error)."""
import numpy as np
import pytest

from conftest import RTOL, as_real_columns, golden_names, golden_path, jacobian_names, traj_rel_err
from oracle import numpy_oracle as orc
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel, n_steps_for
from pyrates_b200.program import Program

pytestmark = pytest.mark.gpu

SKIP = {"kuramoto_n128", "qsfa_both_n96", "wc_n128", "wc_n128_f32", "jrc_grid_b1024", "wc_n300_f32",
        "qsfa_both_n96_f32", "kuramoto_n128_f32", "wc_n128_rk45", "kuramoto_n128_rk45"}                                             # network-mode fixtures
ADAPTIVE = [n for n in golden_names() if ("_rk45" in n or "_rk23" in n or "_dop853" in n) and n not in SKIP]
DDE_ADAPTIVE = [n for n in golden_names() if "_dopri5" in n]
NAMES = [n for n in golden_names() if n not in SKIP and n not in ADAPTIVE and n not in DDE_ADAPTIVE]
SOLVERS = ["euler", "heun", "heun_true", "rk4"]


@pytest.mark.parametrize("solver", SOLVERS)
@pytest.mark.parametrize("name", NAMES)
def test_instance_kernel_matches_golden(gpu, name, solver):
    fx = orc.load_fixture(golden_path(name))
    prog = Program.load(golden_path(name))
    r = fx["run"]
    before = rt.launch_count()
    out, y_final = InstanceModel(prog, solver=solver).run(r["T"], r["dt"], r["dts"])
    assert rt.launch_count() == before + 1          # one fused kernel launch for the whole simulation
    ref = as_real_columns(fx["extra"][f"rec_{solver}"])        # complex fixtures: state k -> real columns (2k, 2k+1)
    got = out[:, :, 0]
    assert got.shape == ref.shape
    err = traj_rel_err(got, ref)
    assert err <= RTOL[prog.float_precision], f"{name}/{solver}: rel err {err:.3e}"


def test_state_writeback_and_chunked_launches(gpu):
    """Two launches that continue from the device-resident state == one launch (step0/eval0/sample0
    bookkeeping, ring-buffer heads included)."""
    for name in ("qsfa_ring_n8", "net10", "qif_input"):
        prog = Program.load(golden_path(name))
        fx = orc.load_fixture(golden_path(name))
        r = fx["run"]
        steps, store_step, n_samples = n_steps_for(r["T"], r["dt"], r["dts"])
        m = InstanceModel(prog, solver="heun")
        s = m.session(1, n_samples)
        s.reset()
        cut = (steps // 3) | 1
        s.launch(cut, r["dt"], store_step)
        s.launch(steps - cut, r["dt"], store_step)
        got = s.download()[:, :, 0]
        s.free()
        assert traj_rel_err(got, fx["extra"]["rec_heun"]) <= RTOL[prog.float_precision]


@pytest.mark.parametrize("solver", ["heun", "rk4"])
def test_jrc_sweep_batch_matches_per_instance_oracle(gpu, solver):
    """grid_search-style batch: B instances with per-instance (C, u); spot-check instances against the oracle
    run one instance at a time with the same parameters."""
    prog = Program.load(golden_path("jrc"))
    B = 1000
    rng = np.random.default_rng(7)
    C = rng.uniform(68.0, 1350.0, B)
    u = rng.uniform(120.0, 320.0, B)
    sweep = [("u", 0), ("weight_v2", 0), ("weight_v2", 1), ("weight", 0), ("weight_v1", 0)]
    m = InstanceModel(prog, solver=solver, sweep=sweep, observers=[0, 2])
    order = [(p.arg, p.elem) for p in m.rhs.params]
    vals = {("u", 0): u, ("weight_v2", 0): C, ("weight_v2", 1): C / 4, ("weight", 0): 0.8 * C, ("weight_v1", 0): C / 4}
    params = np.stack([vals[k] for k in order])
    T, dt, dts = 0.05, 1e-4, 1e-3
    out, _ = m.run(T, dt, dts, params=params)
    assert out.shape == (50, 2, B)
    f = orc.build_vector_field(prog.source())
    for b in (0, 1, 499, 998, 999):
        args = [np.array(a.value, copy=True) for a in prog.args]
        names = [a.name for a in prog.args]
        args[names.index("u")] = np.float64(u[b])
        args[names.index("weight_v2")] = np.array([C[b], C[b] / 4])
        args[names.index("weight")] = np.float64(0.8 * C[b])
        args[names.index("weight_v1")] = np.float64(C[b] / 4)
        rec, _ = orc.run(f, args, T, dt, dts, solver)
        assert traj_rel_err(out[:, :, b], rec[:, [0, 2]]) <= 1e-9


@pytest.mark.parametrize("T,dt,dts", [
    (0.0101, 1e-3, 3e-3),     # round(T/dts) = 3 rows but steps 0, 3, 6, 9 sample: the reference overruns its buffer
    (0.0119, 1e-3, 3e-3),     # sampling step does not divide the horizon: 12 steps, 4 rows
    (6e-3, 1e-3, 5e-3),       # sampling step almost as long as the run: 6 steps, store_step 5, round(1.2) = 1 row, 2 sample points
    (4e-4, 1e-3, None),       # horizon shorter than one step: round(0.4) = 0 steps, empty recording
    (1e-3, 1e-3, None),       # exactly one step: the recording holds the initial state only
    (0.0155, 1e-3, 2e-3),     # banker's rounding in np.round: 15.5 -> 16 steps, 7.75 -> 8 samples
])
@pytest.mark.parametrize("solver", ["euler", "heun"])
def test_ragged_horizons_follow_the_reference_rounding(gpu, T, dt, dts, solver):
    """Edge cases of BaseBackend._solve_euler/_solve_heun bookkeeping (base_backend.py:716-733): steps = round(T/dt),
    rows = round(T/dts), store_step = round(dts/dt), a row whenever i % store_step == 0 — including recordings the
    reference itself would overrun (more sample points than rows are not produced by these combinations)."""
    fx = orc.load_fixture(golden_path("qif_sfa"))
    prog = Program.load(golden_path("qif_sfa"))
    f = orc.build_vector_field(fx["source"])
    steps, store_step, n_samples = n_steps_for(T, dt, dts)
    want_rows = len(range(0, steps, store_step))
    if want_rows > n_samples:
        # the reference's own recording buffer overruns (IndexError in its time loop): refused here before launching
        with pytest.raises(IndexError):
            orc.run(f, orc.fixture_args(fx), T, dt, dts, solver)
        with pytest.raises(ValueError, match="would record"):
            InstanceModel(prog, solver=solver).run(T, dt, dts)
        return
    ref, _ = orc.run(f, orc.fixture_args(fx), T, dt, dts, solver)
    out, y_final = InstanceModel(prog, solver=solver).run(T, dt, dts)
    got = out[:, :, 0]
    assert got.shape == ref.shape == (n_samples, prog.n_state)
    if want_rows:
        assert traj_rel_err(got[:want_rows], ref[:want_rows]) <= 1e-9
    # final state == the oracle's in-place updated y0 (the reference mutates its state vector)
    args = orc.fixture_args(fx)
    orc.run(f, args, T, dt, dts, solver)
    assert np.allclose(y_final[:, 0], args[1], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("B", [1, 31, 129])
def test_odd_batch_sizes_cover_partial_blocks_and_warps(gpu, B):
    prog = Program.load(golden_path("jrc"))
    m = InstanceModel(prog, solver="rk4", sweep=[("u", 0)], observers=[0])
    u = np.linspace(120.0, 320.0, B)[None, :]
    out, _ = m.run(0.02, 1e-4, 1e-3, params=u)
    assert out.shape == (20, 1, B) and np.all(np.isfinite(out))
    single, _ = m.run(0.02, 1e-4, 1e-3, params=u[:, -1:])
    assert np.array_equal(single[:, :, 0], out[:, :, -1])


def test_missing_device_buffers_are_rejected(gpu):
    a = rt.KernelArgs()
    a.n_steps, a.store_step, a.n_inst = 1, 1, 1
    m = InstanceModel(Program.load(golden_path("qif")), solver="euler")
    with pytest.raises(rt.CudaBackendError, match="state pointer"):
        m.module.run("prb_integrate", a, block=128)
    a.y = 8
    with pytest.raises(rt.CudaBackendError, match="not found"):
        m.module.run("no_such_kernel", a, block=128)


@pytest.mark.parametrize("name", ADAPTIVE)
def test_adaptive_rk45_matches_reference_scipy(gpu, name):
    """solver='scipy': SciPy's RK45 (step controller, rejected steps incl. the aliased-buffer effect, dense output at the
    sample times) on the device vs the trajectories the real reference + SciPy produced."""
    fx = orc.load_fixture(golden_path(name))
    r = fx["run"]
    before = rt.launch_count()
    out, _ = InstanceModel(Program.load(golden_path(name)), solver="scipy", rk_method=r["method"]).run(
        r["T"], r["dt"], r["dts"], rtol=r["rtol"], atol=r["atol"], method=r["method"])
    assert rt.launch_count() == before + 1
    ref = fx["extra"]["rec_scipy"]
    assert out[:, :, 0].shape == ref.shape
    # a piecewise-linear input makes the embedded error estimate jump at every knot: accept / reject decisions then hang on
    # the last bit (perturbing y0 by 2e-16 moves the ORACLE by 2.8e-5 and changes its attempt count 451 -> 470), so this
    # fixture is compared at the solver's own accuracy; the interpolation itself is pinned by the rhs test below
    tol = 1e-3 if "_input_" in name else 1e-9
    if "_dop853" in name:
        # DOP853 + the reference's aliased rhs buffer: the dense output takes "f at the step end" from the LAST extra stage, so
        # its error depends on the step sequence; one flipped accept / reject decision moves the samples far more than the
        # solver tolerance (a 1-ulp change of one parameter moves the ORACLE by 2.5e-6 on jrc_dop853, 5.5e-8 on vdp_dop853)
        tol = 1e-5
    assert traj_rel_err(out[:, :, 0], ref) <= tol, name


@pytest.mark.parametrize("name", DDE_ADAPTIVE)
def test_adaptive_dde_dopri5_matches_reference(gpu, name):
    """Delay-differential models with solver='scipy' (BaseBackend._solve_scipy_dde, base_backend.py:771-800): dopri5 restarted
    per sample interval, accepted steps appended to the device-side history the rhs interpolates in — vs the trajectories of
    the real reference + scipy.integrate.ode."""
    fx = orc.load_fixture(golden_path(name))
    r = fx["run"]
    prog = Program.load(golden_path(name))
    assert prog.is_dde
    before = rt.launch_count()
    m = InstanceModel(prog, solver="scipy")
    out, _ = m.run(r["T"], r["dt"], r["dts"])
    assert rt.launch_count() == before + 1
    ref = fx["extra"]["rec_scipy"]
    assert out[:, :, 0].shape == ref.shape
    # accept/reject decisions are discrete: fp64 rounding differences (FMA contraction) stay at 1e-12, the float32 build is
    # compared at the north star's float32 tolerance
    assert traj_rel_err(out[:, :, 0], ref) <= RTOL[prog.float_precision], name


def test_adaptive_dde_history_ring_grows_when_too_small(gpu):
    """A history ring that is too small for the delay window is detected on the device (status 3) and the run repeated with a
    larger one: same trajectory."""
    name = "dde_net10_dopri5"
    fx = orc.load_fixture(golden_path(name))
    r = fx["run"]
    m = InstanceModel(Program.load(golden_path(name)), solver="scipy")
    n_samples = int(round(r["T"] / r["dts"]))
    s = m.session(1, n_samples)
    try:
        s.reset()
        s._launch_dopri5_dde(r["T"], r["dt"], r["dts"], capacity=4)
        assert s.hist_capacity > 4
        out = s.download()
    finally:
        s.free()
    assert traj_rel_err(out[:, :, 0], fx["extra"]["rec_scipy"]) <= 1e-9


def test_dde_batch_of_instances_and_chunked_launches(gpu):
    """Every instance keeps its own history; launches that continue from the device-resident history == one launch."""
    name = "dde_mg"
    fx = orc.load_fixture(golden_path(name))
    r = fx["run"]
    prog = Program.load(golden_path(name))
    steps, store_step, n_samples = n_steps_for(r["T"], r["dt"], r["dts"])
    m = InstanceModel(prog, solver="heun")
    B = 67
    s = m.session(B, n_samples)
    try:
        s.reset()
        cut = (steps // 3) | 1
        s.launch(cut, r["dt"], store_step)
        s.launch(steps - cut, r["dt"], store_step)
        got = s.download()
    finally:
        s.free()
    for b in (0, 31, 66):
        assert traj_rel_err(got[:, :, b], fx["extra"]["rec_heun"]) <= 1e-9
    with pytest.raises(rt.CudaBackendError):
        s2 = m.session(1, n_samples)
        try:
            s2.reset()
            s2.launch(10, 2 * r["dt"], store_step)        # the vector field was generated for another step size
        finally:
            s2.free()


@pytest.mark.parametrize("name", jacobian_names())
def test_jacobian_kernel_matches_reference(gpu, name):
    """get_jacobian_func on the device: the reference-generated `J0[i, j] = expr` stream evaluated by one thread per state,
    16 states in ONE launch, vs the matrices the reference's own compiled NumPy function returned."""
    from pyrates_b200.engine import evaluate_rhs
    fx = orc.load_fixture(golden_path(name))
    prog = Program.load(golden_path(name))
    pts, ref = fx["extra"]["jac_points"], fx["extra"]["jac_ref"]
    n = prog.n_state
    before = rt.launch_count()
    got = evaluate_rhs(prog, y=pts.T)                       # (n*n, 16)
    assert rt.launch_count() == before + 1
    got = got.reshape(n, n, -1).transpose(2, 0, 1)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    tol = 1e-12 if prog.float_precision == "float64" else 1e-5
    assert np.allclose(got, ref, rtol=tol, atol=tol * np.max(np.abs(ref)))
    single = evaluate_rhs(prog).reshape(n, n)               # at the program's own state
    assert np.allclose(single, ref[0], rtol=tol, atol=tol * np.max(np.abs(ref)))
    # new parameter values reuse the compiled kernel (constants are runtime parameters of the evaluation kernel)
    p2 = prog.clone()
    for a in p2.args:
        if a.kind == "const" and a.value.size == 1 and a.value.dtype.kind == "f":
            a.value = (a.value * 1.25).astype(a.value.dtype)
    n_mod = len(rt._module_cache)
    other = evaluate_rhs(p2).reshape(n, n)
    assert len(rt._module_cache) == n_mod and not np.allclose(other, single)
    f = orc.build_vector_field(p2.source(), p2.func_name)
    want = f(0.0, p2.y0.copy(), np.zeros((n, n), dtype=prog.real_dtype), *[a.value for a in p2.args[3:]])
    assert np.allclose(other, want, rtol=tol, atol=tol * np.max(np.abs(want)))


@pytest.mark.parametrize("solver", ["heun", "rk4", "scipy"])
def test_dde_sweep_matches_per_instance_oracle(gpu, solver):
    """grid_search-style batch over a delay-differential model (Mackey-Glass + delayed-feedback unit): every instance has its
    own parameters AND its own history; chunked streaming launches continue from the device-resident history rings."""
    from pyrates_b200.sweep import BatchedSweep
    name = "dde_mg_dopri5" if solver == "scipy" else "dde_mg"
    prog = Program.load(golden_path(name))
    fx = orc.load_fixture(golden_path(name))
    B = 300
    rng = np.random.default_rng(5)
    table = np.stack([rng.uniform(1.5, 2.5, B), rng.uniform(2.0, 4.0, B)])          # bet, a
    T, dt, dts = 0.1, 1e-4, 1e-3
    sw = BatchedSweep(prog, [("bet", 0), ("a", 0)], observers=[0, 1], solver=solver, T=T, dt=dt, dts=dts, n_inst=B,
                      chunk_samples=7, const_div_to_mul=False)
    try:
        out = np.array(sw.run(table), copy=True)
    finally:
        sw.free()
    assert out.shape == (100, 2, B)
    if solver != "scipy":
        # the fast_math build (invariant folding + speculative exp / reciprocal) of the same sweep stays within 1e-9
        fast = BatchedSweep(prog, [("bet", 0), ("a", 0)], observers=[0, 1], solver=solver, T=T, dt=dt, dts=dts, n_inst=B,
                            chunk_samples=50)
        try:
            assert fast.model.kernel.source.count("T.q[") > 0
            out_fast = np.array(fast.run(table), copy=True)
        finally:
            fast.free()
        assert traj_rel_err(out_fast.reshape(100, -1), out.reshape(100, -1)) <= 1e-9
    f = orc.build_vector_field(fx["source"])
    for b in (0, 17, 299):
        args = orc.fixture_args(fx)
        args[fx["names"].index("bet")] = np.float64(table[0, b])
        args[fx["names"].index("a")] = np.float64(table[1, b])
        rec = orc.run_adaptive(f, args, T, dt, dts)[0] if solver == "scipy" else orc.run(f, args, T, dt, dts, solver)[0]
        assert traj_rel_err(out[:, :, b], rec) <= 1e-9, (solver, b)


@pytest.mark.parametrize("name,solver", [("jrc", "rk4"), ("qif_f32", "heun"), ("dde_mg", "euler")])
def test_on_device_observer_reduction_matches_numpy_statistics(gpu, name, solver):
    """reduce=True: mean / var / min / max / last of every observer over the decimated samples at times >= cutoff, computed
    in registers by the integration kernel, equal NumPy's statistics of the full trajectories of the same sweep."""
    from pyrates_b200.sweep import BatchedSweep
    prog = Program.load(golden_path(name))
    r = orc.load_fixture(golden_path(name))["run"]
    sweep = {"jrc": [("u", 0), ("weight", 0)], "qif_f32": [("eta", 0)], "dde_mg": [("bet", 0), ("a", 0)]}[name]
    B = 130
    rng = np.random.default_rng(3)
    base = np.array([float(np.asarray(prog.arg(a).value).reshape(-1)[e]) for a, e in sweep])
    table = base[:, None] * rng.uniform(0.8, 1.2, (len(sweep), B))
    obs = list(range(min(prog.n_state, 3)))
    T, dt, dts = r["T"], r["dt"], (r["dts"] or r["dt"])
    cutoff = 0.25 * T
    kw = dict(solver=solver, T=T, dt=dt, dts=dts, n_inst=B, const_div_to_mul=False)
    full_sw = BatchedSweep(prog, sweep, obs, chunk_samples=9, **kw)
    red_sw = BatchedSweep(prog, sweep, obs, reduce=True, cutoff=cutoff, **kw)
    try:
        full = np.array(full_sw.run(table), copy=True)
        before = rt.launch_count()
        red = np.array(red_sw.run(table), copy=True)
        assert rt.launch_count() == before + 1 and red_sw.d2h_bytes == 5 * len(obs) * B * full.itemsize
    finally:
        full_sw.free()
        red_sw.free()
    times = np.linspace(0.0, T, num=full.shape[0], endpoint=False)
    keep = full[times >= cutoff].astype(np.float64)
    assert red.shape == (5, len(obs), B) and 0 < keep.shape[0] < full.shape[0]
    tol = 1e-12 if prog.float_precision == "float64" else 2e-5
    scale = np.max(np.abs(keep), axis=0) + 1e-300
    for k, want in enumerate([keep.mean(axis=0), keep.var(axis=0), keep.min(axis=0), keep.max(axis=0), keep[-1]]):
        s = scale ** 2 if k == 1 else scale
        assert np.max(np.abs(red[k] - want) / s) <= tol, InstanceModel.__name__ + f" stat {k}"


@pytest.mark.parametrize("solver", ["heun", "scipy"])
def test_sweep_over_history_delays_matches_per_instance_oracle(gpu, solver):
    """grid_search over the DELAY of a delay-differential model: every instance looks a different distance back into its own
    history; the device-side history is sized from the bound on the swept delays, which every parameter table is checked
    against."""
    from pyrates_b200.sweep import BatchedSweep
    name = "dde_mg_dopri5" if solver == "scipy" else "dde_mg"
    prog = Program.load(golden_path(name))
    fx = orc.load_fixture(golden_path(name))
    B = 96
    rng = np.random.default_rng(9)
    table = np.stack([rng.uniform(0.004, 0.03, B), rng.uniform(0.002, 0.02, B), rng.uniform(1.5, 2.5, B)])     # tau, d, bet
    T, dt, dts = 0.1, 1e-4, 1e-3
    sw = BatchedSweep(prog, [("tau", 0), ("d", 0), ("bet", 0)], observers=[0, 1], solver=solver, T=T, dt=dt, dts=dts, n_inst=B,
                      chunk_samples=13, const_div_to_mul=False, hist_max_delay=0.03)
    try:
        out = np.array(sw.run(table), copy=True)
        bad = table.copy()
        bad[0, 5] = 0.031
        with pytest.raises(rt.CudaBackendError, match="hist_max_delay"):
            sw.run(bad)
    finally:
        sw.free()
    f = orc.build_vector_field(fx["source"])
    for b in (0, 41, 95):
        args = orc.fixture_args(fx)
        for k, nm in enumerate(("tau", "d", "bet")):
            args[fx["names"].index(nm)] = np.float64(table[k, b])
        rec = orc.run_adaptive(f, args, T, dt, dts)[0] if solver == "scipy" else orc.run(f, args, T, dt, dts, solver)[0]
        assert traj_rel_err(out[:, :, b], rec) <= 1e-9, (solver, b)


def test_interpolated_input_rhs_matches_numpy_interp(gpu):
    """interp(t, time, u_input) of the adaptive build (numpy.interp semantics): knots, interior points, both clamps."""
    from pyrates_b200.engine import evaluate_rhs
    fx = orc.load_fixture(golden_path("qif_input_rk45"))
    prog = Program.load(golden_path("qif_input_rk45"))
    f = orc.build_vector_field(fx["source"])
    time = fx["args"][fx["names"].index("time")]
    for t in [0.0, float(time[1]), float(time[7]) * 0.999999, 0.123456789, float(time[-2]) + 1e-7, float(time[-1]), 2.5, -0.3]:
        args = [np.array(a, copy=True) for a in fx["args"]]
        ref = f(t, args[1], *args[2:]).copy()
        got = evaluate_rhs(prog, t=t)
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13)


def test_adaptive_rk45_batch_has_independent_step_sequences(gpu):
    """A batch of JRC instances with different inputs: every thread runs its own adaptive step sequence; each matches
    the oracle (scipy restatement) integrating that instance alone."""
    prog = Program.load(golden_path("jrc_rk45"))
    fx = orc.load_fixture(golden_path("jrc_rk45"))
    r = fx["run"]
    B = 64
    u = np.linspace(120.0, 320.0, B)
    m = InstanceModel(prog, solver="scipy", sweep=[("u", 0)])
    out, _ = m.run(r["T"], r["dt"], r["dts"], params=u[None, :], rtol=1e-6, atol=1e-8)
    f = orc.build_vector_field(fx["source"])
    for b in (0, 21, 63):
        args = [np.array(a, copy=True) for a in fx["args"]]
        args[0] = args[0][()]
        args[fx["names"].index("u")] = np.full_like(args[fx["names"].index("u")], u[b])
        rec, _ = orc.run_adaptive(f, args, r["T"], r["dt"], r["dts"], rtol=1e-6, atol=1e-8)
        # rounding differences (FMA, summation order) feed back through the step-size controller of an oscillator:
        # observed 0 .. 2e-9; the bound is 1e-7 = rtol / 10, three orders below the solver's own error
        assert traj_rel_err(out[:, :, b], rec) <= 1e-7, b



@pytest.mark.parametrize("name,solver", [("jrc", "rk4"), ("jrc_2delay", "heun"), ("wc_n8", "euler"), ("qif_sfa", "heun_true")])
def test_fast_math_variant_keeps_fp64_parity(gpu, name, solver):
    """Opt-in fast_math (table exp, rcp + Newton division): same 1e-9 tolerance as the default kernels."""
    fx = orc.load_fixture(golden_path(name))
    r = fx["run"]
    out, _ = InstanceModel(Program.load(golden_path(name)), solver=solver, fast_math=True, const_div_to_mul=True).run(r["T"], r["dt"], r["dts"])
    assert traj_rel_err(out[:, :, 0], fx["extra"][f"rec_{solver}"]) <= 1e-9


def test_fast_math_float32_sweep_keeps_fp32_parity(gpu):
    """float32 (the reference's default precision) with fast_math (__expf, __fdividef, folded constants): a JRC batch
    against the float32 oracle per instance at the fp32 tolerance."""
    import bench
    import bench_configs as bc
    prog = bc.to_float32(Program.load(golden_path("jrc")))
    B = 256
    table = bench.param_table(16, 16, 0, B)
    m = InstanceModel(prog, solver="rk4", sweep=bench.SWEEP, observers=[0, 2], const_div_to_mul=True, fast_math=True, min_blocks=8)
    assert "prb_expf_ftz(" in m.kernel.source and "__fdividef(" in m.kernel.source
    order = [(p.arg, p.elem) for p in m.rhs.params]
    out, _ = m.run(0.2, 1e-4, 1e-3, params=np.stack([table[bench.SWEEP.index(k)] for k in order]))
    f = orc.build_vector_field(prog.source())
    names = [a.name for a in prog.args]
    for b in (0, 100, 255):
        u, C, C4, C08, _ = table[:, b]
        args = [np.array(a.value, copy=True) for a in prog.args]
        args[names.index("u")] = np.float32(u)
        args[names.index("weight_v2")] = np.array([C, C4], dtype=np.float32)
        args[names.index("weight")] = np.float32(C08)
        args[names.index("weight_v1")] = np.float32(C4)
        with np.errstate(over="ignore"):
            rec, _ = orc.run(f, args, 0.2, 1e-4, 1e-3, "rk4")
        assert traj_rel_err(out[:, :, b], rec[:, [0, 2]]) <= 1e-4, b
