"""The oracle (oracle/numpy_oracle.py) pinned against the reference: golden trajectories produced by the
real reference NumPy backend (bit-exact) and the reference's own hand-computed known-answer loops."""
import numpy as np
import pytest

from conftest import golden_names, golden_path, jacobian_names
from oracle import numpy_oracle as orc


ADAPTIVE = [n for n in golden_names() if "_rk45" in n or "_rk23" in n or "_dop853" in n]
DDE_ADAPTIVE = [n for n in golden_names() if "_dopri5" in n]


@pytest.mark.parametrize("name", DDE_ADAPTIVE)
def test_oracle_dopri5_dde_reproduces_reference_plus_scipy_bit_exact(name):
    """solver='scipy' on a delay-differential model: the restated Hairer DOPRI5 core, its restart per sample interval and the
    history updates (BaseBackend._solve_scipy_dde + DDEHistory) give exactly what the real reference + scipy.integrate.ode
    produced in the build container."""
    fx = orc.load_fixture(golden_path(name))
    rec, times = orc.run_fixture(fx, solver="scipy")
    ref = fx["extra"]["rec_scipy"]
    assert rec.shape == ref.shape and rec.dtype == ref.dtype and np.array_equal(rec, ref)


def test_oracle_dopri5_matches_scipy_ode_on_a_plain_function():
    """Without a history the restated core is scipy.integrate.ode('dopri5') itself, solout calls included."""
    integ = pytest.importorskip("scipy.integrate")
    f = lambda t, y: np.array([y[1], -4.0 * y[0] - 0.1 * y[1] + np.sin(3 * t)])
    seen_ref, seen = [], []
    solver = integ.ode(f).set_integrator("dopri5", first_step=1e-3, nsteps=50000)
    solver.set_initial_value(np.array([1.0, 0.0]), 0.0)
    solver.set_solout(lambda t, y: seen_ref.append((t, y.copy())) and None)
    x, y = 0.0, np.array([1.0, 0.0])
    for t_out in np.linspace(0.0, 3.0, 7):
        solver.integrate(t_out)
        x, y, ok = orc.dopri5_call(f, x, y, float(t_out), 1e-3, lambda t, y_: seen.append((t, y_.copy())))
        assert ok and x == solver.t and np.array_equal(y, solver.y)
    assert len(seen) == len(seen_ref) and all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(seen, seen_ref))


def test_oracle_history_is_the_references_ddehistory():
    """Known answers of DDEHistory.__call__ (base_backend.py:165-177): initial condition before t0, latest state after the
    last stored time, linear interpolation in between, the LAST of duplicated times as the left bracket; and the reference's
    own DDE test (tests/test_backend_simulations.py:551-583): x' = -x(t-d) decays linearly while t < d."""
    h = orc.History(np.array([1.0, -2.0]))
    h.update(0.0, np.array([1.0, -2.0]))
    h.update(0.5, np.array([2.0, 0.0]))
    h.update(0.5, np.array([4.0, 4.0]))
    h.update(1.0, np.array([6.0, 8.0]))
    assert np.array_equal(h(-3.0), [1.0, -2.0]) and np.array_equal(h(0.0), [1.0, -2.0])
    assert np.array_equal(h(0.25), [1.5, -1.0])
    assert np.array_equal(h(0.5), [4.0, 4.0]) and np.array_equal(h(0.75), [5.0, 6.0])
    assert np.array_equal(h(1.0), [6.0, 8.0]) and np.array_equal(h(7.0), [6.0, 8.0])
    fx = orc.load_fixture(golden_path("dde_net16"))
    rec, _ = orc.run_fixture(fx, solver="euler")
    x = rec[:, 0]
    np.testing.assert_allclose(x[:100], 1.0 - np.arange(100) * 1e-3, atol=5e-3)
    assert np.all(np.abs(x) < 5.0)


@pytest.mark.parametrize("name", ADAPTIVE)
def test_oracle_rk45_reproduces_reference_plus_scipy_bit_exact(name):
    """solver='scipy': the restated Dormand-Prince stepper, step controller, dense output and t_eval loop (plus the
    aliased rhs buffer) give exactly what the real reference + SciPy produced in the build container."""
    fx = orc.load_fixture(golden_path(name))
    r = fx["run"]
    f = orc.build_vector_field(fx["source"])
    args = [np.array(a, copy=True) for a in fx["args"]]
    args[0] = args[0][()]
    rec, times = orc.run_adaptive(f, args, r["T"], r["dt"], r["dts"], rtol=r["rtol"], atol=r["atol"], method=r["method"])
    ref = fx["extra"]["rec_scipy"]
    assert rec.shape == ref.shape and np.array_equal(rec, ref)


def test_oracle_rk45_coefficients_are_scipys():
    rk = pytest.importorskip("scipy.integrate._ivp.rk")
    for mine, theirs in ((orc.RK45_A, rk.RK45.A), (orc.RK45_B, rk.RK45.B), (orc.RK45_C, rk.RK45.C), (orc.RK45_E, rk.RK45.E),
                         (orc.RK45_P, rk.RK45.P), (orc.RK23_A, rk.RK23.A), (orc.RK23_B, rk.RK23.B), (orc.RK23_C, rk.RK23.C),
                         (orc.RK23_E, rk.RK23.E), (orc.RK23_P, rk.RK23.P)):
        assert np.array_equal(mine, theirs)


def test_dop853_coefficients_are_scipys_and_oracle_matches_solve_ivp():
    """pyrates_b200/dop853.py carries the doubles of SciPy's dop853_coefficients; the restated DOP853 (error estimate from
    the 5th / 3rd order pair, extra stages + 7-term dense output) is scipy.solve_ivp itself on a function that returns fresh
    arrays — apart from the aliasing of `f`, which a plain function does not have: emulate plain SciPy by handing the oracle a
    function whose result is copied."""
    integ = pytest.importorskip("scipy.integrate")
    from scipy.integrate._ivp import dop853_coefficients as d
    from pyrates_b200 import dop853 as mine
    for name in ("A", "B", "C", "E3", "E5", "D"):
        assert np.array_equal(np.array(getattr(mine, name)), getattr(d, name)), name
    assert (mine.N_STAGES, mine.N_STAGES_EXTENDED, mine.INTERPOLATOR_POWER) == (d.N_STAGES, d.N_STAGES_EXTENDED, d.INTERPOLATOR_POWER)
    buf = np.zeros(2)

    def f_alias(t, y):                       # like the reference's generated rhs: returns its in-place buffer
        buf[0] = y[1]
        buf[1] = -4.0 * y[0] - 0.1 * y[1] + np.sin(3 * t)
        return buf
    times = np.linspace(0.0, 5.0, 100, endpoint=False)
    ref = integ.solve_ivp(f_alias, (0.0, 5.0), np.array([1.0, 0.0]), first_step=1e-3, t_eval=times, rtol=1e-8, atol=1e-10,
                          method="DOP853").y.T
    buf[:] = 0
    rec = orc.solve_dop853(f_alias, (), 5.0, 1e-3, np.array([1.0, 0.0]), 0.0, times, rtol=1e-8, atol=1e-10)
    assert np.array_equal(rec, ref)


def test_oracle_rk45_matches_solve_ivp_on_a_plain_function():
    """Without the aliased buffer (a function returning fresh arrays) the restatement is scipy.solve_ivp itself."""
    integ = pytest.importorskip("scipy.integrate")
    f = lambda t, y: np.array([y[1], -4.0 * y[0] - 0.1 * y[1] + np.sin(3 * t)])
    times = np.linspace(0.0, 5.0, 100, endpoint=False)
    ref = integ.solve_ivp(f, (0.0, 5.0), np.array([1.0, 0.0]), first_step=1e-3, t_eval=times, rtol=1e-6, atol=1e-9).y.T
    rec = orc.solve_rk45(f, (), 5.0, 1e-3, np.array([1.0, 0.0]), 0.0, times, rtol=1e-6, atol=1e-9, aliased=False)
    assert np.array_equal(rec, ref)
    # the reference's in-place rhs buffer changes what a rejected attempt restarts from: measurably different numbers
    quirk = orc.solve_rk45(f, (), 5.0, 1e-3, np.array([1.0, 0.0]), 0.0, times, rtol=1e-6, atol=1e-9, aliased=True)
    assert 1e-6 < np.max(np.abs(quirk - ref)) < 1e-2


@pytest.mark.parametrize("name", [n for n in golden_names() if n not in ADAPTIVE and n not in DDE_ADAPTIVE])
@pytest.mark.parametrize("solver", ["euler", "heun"])
def test_oracle_reproduces_reference_bit_exact(name, solver):
    fx = orc.load_fixture(golden_path(name))
    rec, times = orc.run_fixture(fx, solver=solver)
    ref = fx["extra"][f"rec_{solver}"]
    assert rec.shape == ref.shape and rec.dtype == ref.dtype
    assert np.array_equal(rec, ref, equal_nan=True), f"max abs diff {np.nanmax(np.abs(rec - ref))}"
    r = fx["run"]
    assert times.shape[0] == ref.shape[0] == round(r["T"] / (r["dts"] or r["dt"]))


def test_reference_heun_is_aliased_not_textbook():
    """SURVEY headline 4: `heun` == y + dt*f(y + dt*f(y)); textbook Heun differs measurably."""
    fx = orc.load_fixture(golden_path("qif_dt2"))
    ref = fx["extra"]["rec_heun"]
    true_heun = fx["extra"]["rec_heun_true"]
    f = orc.build_vector_field(fx["source"])
    args = [np.array(a, copy=True) for a in fx["args"]]
    y, dt = args[1].copy(), fx["run"]["dt"]
    rec = []
    ss = round(fx["run"]["dts"] / dt)
    for i in range(round(fx["run"]["T"] / dt)):
        if i % ss == 0:
            rec.append(y.copy())
        k1 = f(i, y, *args[2:]).copy()
        k2 = f(i, y + dt * k1, *args[2:]).copy()
        y = y + dt / 2 * (k2 + k2)
    assert np.array_equal(np.array(rec), ref)
    assert np.max(np.abs(true_heun - ref)) > 1e-4


def test_ring_buffer_known_answer():
    """reference tests/test_population_connectivity.py:361-415 (roll -> write -> read, d=4 steps)."""
    fx = orc.load_fixture(golden_path("ring_kat"))
    rec, _ = orc.run_fixture(fx, solver="euler")
    N, d_steps, dt, tau_r = 2, 4, 1e-3, 0.1
    W = np.array([[0., 0.5], [0.5, 0.]])
    r, eta = np.array([1.0, 2.0]), np.array([5.0, 8.0])
    buf = np.zeros((N, d_steps + 1))
    ref = np.zeros((10, N))
    for k in range(10):
        ref[k] = r
        buf = np.roll(buf, 1, axis=1)
        buf[:, 0] = r
        r = r + dt * ((-r + eta + W @ buf[:, d_steps]) / tau_r)
    np.testing.assert_allclose(rec, ref, rtol=1e-12)
    np.testing.assert_allclose(fx["extra"]["rec_euler"], ref, rtol=1e-5)   # the reference's own bar


def test_gamma_cascade_known_answer():
    """reference tests/test_population_connectivity.py:422-476 (n=6 stage ODE cascade)."""
    fx = orc.load_fixture(golden_path("gamma_kat"))
    rec, _ = orc.run_fixture(fx, solver="euler")
    delay, spread, dt, tau_r = 5e-3, 2e-3, 1e-3, 0.1
    n = max(1, int(round((delay / spread) ** 2)))
    a = n / delay
    W = np.array([[0., 0.5], [0.5, 0.]])
    r, eta = np.array([1.0, 2.0]), np.array([5.0, 8.0])
    z = np.zeros((n, 2))
    ref = np.zeros((20, 2))
    for k in range(20):
        ref[k] = r
        dz = np.zeros_like(z)
        dz[0] = a * (r - z[0])
        for j in range(1, n):
            dz[j] = a * (z[j - 1] - z[j])
        r, z = r + dt * ((-r + eta + W @ z[-1]) / tau_r), z + dt * dz
    np.testing.assert_allclose(rec[:, :2], ref, rtol=1e-10)


def test_kuramoto_first_euler_step_known_answer():
    """reference tests/test_population_connectivity.py:238-283, restated through the oracle's op table."""
    src = ("def vector_field(t,y,dy,K,s_ext,omega,weight):\n"
           "\ttheta = y[0:2]\n"
           "\ts_in = wsum(weight, -sin(broadcast_post(theta) - broadcast_pre(theta)))\n"
           "\tdy[0:2] = K*s_in + omega + s_ext\n\treturn dy\n")
    f = orc.build_vector_field(src)
    y = np.array([0.0, np.pi / 2])
    rec, _ = orc.run(f, [0, y, np.zeros(2), np.ones(2), 0.0, np.full(2, 10.0), np.array([[0., 1.], [1., 0.]])],
                     2e-3, 1e-3, None, "euler")
    np.testing.assert_allclose(rec[1], [0.011, np.pi / 2 + 0.009], rtol=1e-12)


@pytest.mark.parametrize("solver,order", [("euler", 1), ("heun_true", 2), ("rk4", 4), ("heun", 1)])
def test_solver_convergence_order(solver, order):
    """RK4 / textbook Heun have no reference implementation: pin them on y' = -y + sin(t-free) problem
    (analytic solution) by their convergence order.  The aliased reference `heun` is only 1st order."""
    src = "def vector_field(t,y,dy,lam):\n\ta = y[0]\n\tdy[0] = lam*a - a**3\n\treturn dy\n"
    f = orc.build_vector_field(src)
    lam, y0, T = -1.5, 0.8, 1.0
    # analytic solution of y' = lam*y - y^3 (Bernoulli): y^2 = lam / (1 - (1 - lam/y0^2) exp(-2 lam t))... use fine RK4 as truth
    def final(dt, s):
        y = np.array([y0])
        n = round(T / dt)
        rec, _ = orc.run(f, [0, y, np.zeros(1), lam], T + dt, dt, None, s)
        return rec[n, 0]
    truth = final(1e-4, "rk4")
    e1, e2 = abs(final(2e-2, solver) - truth), abs(final(1e-2, solver) - truth)
    assert abs(np.log2(e1 / e2) - order) < 0.35, (e1, e2)


@pytest.mark.parametrize("name", jacobian_names())
def test_oracle_jacobian_reproduces_reference(name):
    """get_jacobian_func (computegraph.py:516-763): the oracle executes the reference-generated assignment stream
    `J0[i, j] = expr` and must return exactly what the reference's own compiled function returned at 16 states."""
    fx = orc.load_fixture(golden_path(name))
    f = orc.build_vector_field(fx["source"], fx["meta"]["func_name"])
    shape = fx["args"][2].shape
    for y, ref in zip(fx["extra"]["jac_points"], fx["extra"]["jac_ref"]):
        J = f(0.0, y.copy(), np.zeros(shape, dtype=fx["args"][2].dtype), *fx["args"][3:])
        assert J.dtype == ref.dtype and np.array_equal(J, ref, equal_nan=True)


def test_jacobian_fixture_known_answers():
    """The reference's own analytical check (tests/test_jacobian.py:121-158): van der Pol at (x, z) = (0, 1), mu = 1 has
    J = [[0, 1], [-1, 1]]; and the transcendental model's Jacobian agrees with central differences of its vector field."""
    fx = orc.load_fixture(golden_path("jac_vdp"))
    f = orc.build_vector_field(fx["source"], fx["meta"]["func_name"])
    J = f(0.0, np.array([0.0, 1.0]), np.zeros((2, 2)), *fx["args"][3:])
    assert np.allclose(J, [[0.0, 1.0], [-1.0, 1.0]], atol=1e-8)
    fx = orc.load_fixture(golden_path("jac_trans"))
    f = orc.build_vector_field(fx["source"], fx["meta"]["func_name"])
    a = float(fx["args"][fx["names"].index("a")])
    rhs = lambda y: np.array([-y[0] + np.exp(-a * y[0]) + np.cos(y[1]), -y[1] + np.sin(y[0])])
    for y in fx["extra"]["jac_points"][:4]:
        J = f(0.0, y.copy(), np.zeros((2, 2)), *fx["args"][3:])
        fd = np.stack([(rhs(y + 1e-6 * e) - rhs(y - 1e-6 * e)) / 2e-6 for e in np.eye(2)], axis=1)
        assert np.allclose(J, fd, atol=1e-6)
