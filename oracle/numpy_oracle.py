"""CPU oracle for the PyRates numerical-integration hot path  —  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs (``cpu_baseline`` and
``--impl reference``) may import this module.  The product (``pyrates_b200``) never does: it fails
loudly when its CUDA library is missing.

This is a restatement, in plain NumPy, of the reference NumPy backend's algorithm for the path
"generated vector field + fixed-step solver + state recording":

* ``build_vector_field``  <- ``BaseBackend.generate_func`` (pyrates/backend/base/base_backend.py:537-578):
  the generated ``vector_field(t, y, dy, *consts)`` text is exec'd against the reference's op table
  (pyrates/backend/base/base_funcs.py:135-177: ``dot``, ``roll``, ``sum``, ``einsum``-based ``wsum``,
  ``sigmoid``, ``broadcast_pre/post`` ...).  The function writes ``dy`` in place and returns it
  (pyrates/backend/computegraph.py:1220-1228, base_backend.py:532-535).
* ``solve_euler``  <- ``BaseBackend._solve_euler`` (base_backend.py:712-740).
* ``solve_heun``   <- ``BaseBackend._solve_heun``  (base_backend.py:742-769) INCLUDING its buffer
  aliasing: ``rhs`` is the in-place ``dy`` buffer, the second call overwrites it, so the update is
  ``y += dt/2*(k2 + k2)`` (SURVEY.md headline 4).  ``solve_heun_true`` is the textbook scheme.
* ``solve_rk4``: the reference has NO fixed-step RK4 (``SUPPORTED_SOLVERS``, base_backend.py:241).
  Definition used by this project: classic RK4 around the reference rhs, every stage
  called with the same integer ``step`` (inputs piecewise constant over a step, like the reference's
  Heun which passes ``step`` to both calls, base_backend.py:763-765), each stage result copied because
  the rhs returns its in-place buffer.  Mutable buffers (delay rings) advance once per *evaluation*.
* ``run``          <- ``BaseBackend.run`` (base_backend.py:596-611): sample times and dispatch.

Pinning: ``tests/test_oracle.py`` checks this module against (a) trajectories produced by the real
reference in the build container (``tests/golden/*.npz``, made by ``oracle/make_golden.py``) with
max-abs-diff 0.0 for euler/heun, and (b) the reference's own hand-computed known-answer loops
(tests/test_population_connectivity.py:238-283, 361-415, 422-476; tests/test_backend_simulations.py).
RK4 has no reference implementation, so it is pinned only against the analytic solution of linear
test problems and against 4th-order convergence ("parity unpinned" w.r.t. the reference for rk4).
"""
from __future__ import annotations

import bisect
import json
from typing import Callable, List, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------- op table
# restates pyrates/backend/base/base_funcs.py:42-177 (names as they appear in generated code)


def _sigmoid(x):
    return 1. / (1. + np.exp(-x))


def _wsum(weight, coupling):
    return np.einsum('ij,ij->i', weight, coupling)


def _interp_rows(t, time, inp):
    return np.array([np.interp(t, time, inp[:, k]) for k in range(inp.shape[1])])


OPS = {
    'pi': np.pi, 'sqrt': np.sqrt, 'maximum': np.maximum, 'minimum': np.minimum, 'round': np.round,
    'sum': np.sum, 'mean': np.mean, 'dot': np.dot, 'roll': np.roll, 'tanh': np.tanh, 'sinh': np.sinh,
    'cosh': np.cosh, 'arctan': np.arctan, 'arcsin': np.arcsin, 'arccos': np.arccos, 'sin': np.sin,
    'cos': np.cos, 'tan': np.tan, 'exp': np.exp, 'interp': np.interp, 'interp_rows': _interp_rows,
    'sigmoid': _sigmoid, 'broadcast_pre': lambda x: x[None, :], 'broadcast_post': lambda x: x[:, None],
    'wsum': _wsum, 'einsum': np.einsum, 'reshape2d': lambda x, n, m: x.reshape(n, m),
    'flatten1d': lambda x: x.reshape(-1), 'identity': lambda x: x, 'abs': np.abs, 'sign': np.sign,
    'log': np.log, 'concatenate': np.concatenate, 'real': np.real, 'imag': np.imag,
    'conjugate': np.conjugate, 'array': np.array,
}


def build_vector_field(source: str, func_name: str = "vector_field") -> Callable:
    """exec the generated function text (base_backend.py:537-578) against the op table."""
    ns = dict(OPS)
    body = source[source.index(f"def {func_name}"):]
    exec(compile(body, f"<oracle:{func_name}>", "exec"), ns)
    return ns[func_name]


# ----------------------------------------------------------------------------- delay-differential histories
class History:
    """Restatement of ``DDEHistory`` (base_backend.py:88-177): (time, state) pairs, linear interpolation between the stored
    points, the initial condition for t <= t0, the latest state for t >= the last stored time.  The generated rhs reads it
    as ``hist(t*dt - d)[idx]`` (fixed-step solvers: ``t`` is the integer step counter) or ``hist(t - d)[idx]`` (adaptive),
    base_backend.py:437-444."""

    def __init__(self, y0, t0: float = 0.0):
        self._t = [float(t0)]
        self._y = [np.array(y0, copy=True)]

    def update(self, t, y):
        self._t.append(float(t))
        self._y.append(np.array(y, dtype=self._y[0].dtype, copy=True))      # rows of a buffer with y0's dtype (:160)

    def __call__(self, t):
        t = float(t)
        if t <= self._t[0]:
            return self._y[0]
        if t >= self._t[-1]:
            return self._y[-1]
        idx = bisect.bisect_right(self._t, t) - 1
        t0_, t1_ = self._t[idx], self._t[idx + 1]
        alpha = (t - t0_) / (t1_ - t0_)
        return self._y[idx] + alpha * (self._y[idx + 1] - self._y[idx])


def _dde(args):
    """base_backend.py:724 / :755: ``has_dde = len(args) > 0 and isinstance(args[0], DDEHistory)``."""
    return args[0] if len(args) > 0 and isinstance(args[0], History) else None


# ----------------------------------------------------------------------------- solvers
def _prep(T, dt, dts, y):
    steps = int(np.round(T / dt))
    store_steps = int(np.round(T / dts))
    store_step = int(np.round(dts / dt))
    rec = np.empty((store_steps, y.shape[0]) if y.shape else (store_steps, 1), dtype=y.dtype)
    return steps, store_step, rec


def solve_euler(func, args, T, dt, dts, y, t0):
    """base_backend.py:712-740 (sample before update; integer step passed as t)."""
    steps, store_step, rec = _prep(T, dt, dts, y)
    hist = _dde(args)
    idx = 0
    for i in range(steps):
        if i % store_step == 0:
            rec[idx, :] = y
            idx += 1
        step = i + t0
        rhs = func(step, y, *args)
        y += dt * rhs
        if hist is not None:
            hist.update((i + 1) * dt, y)
    return rec


def solve_heun(func, args, T, dt, dts, y, t0):
    """base_backend.py:742-769 verbatim semantics, including the in-place ``dy`` aliasing."""
    steps, store_step, rec = _prep(T, dt, dts, y)
    hist = _dde(args)
    idx = 0
    for i in range(steps):
        if i % store_step == 0:
            rec[idx, :] = y
            idx += 1
        step = i + t0
        rhs = func(step, y, *args)
        y_0 = y + dt * rhs
        y += dt / 2 * (rhs + func(step, y_0, *args))
        if hist is not None:
            hist.update((i + 1) * dt, y)
    return rec


def solve_heun_true(func, args, T, dt, dts, y, t0):
    """Textbook Heun (explicit trapezoid): the first stage is copied before the second call."""
    steps, store_step, rec = _prep(T, dt, dts, y)
    hist = _dde(args)
    idx = 0
    for i in range(steps):
        if i % store_step == 0:
            rec[idx, :] = y
            idx += 1
        step = i + t0
        k1 = func(step, y, *args).copy()
        y_0 = y + dt * k1
        k2 = func(step, y_0, *args)
        y += dt / 2 * (k1 + k2)
        if hist is not None:
            hist.update((i + 1) * dt, y)
    return rec


def solve_rk4(func, args, T, dt, dts, y, t0):
    """Classic RK4 around the reference rhs (project definition, see module docstring)."""
    steps, store_step, rec = _prep(T, dt, dts, y)
    hist = _dde(args)
    idx = 0
    for i in range(steps):
        if i % store_step == 0:
            rec[idx, :] = y
            idx += 1
        step = i + t0
        k1 = func(step, y, *args).copy()
        k2 = func(step, y + (dt / 2) * k1, *args).copy()
        k3 = func(step, y + (dt / 2) * k2, *args).copy()
        k4 = func(step, y + dt * k3, *args)
        y += (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
        if hist is not None:
            hist.update((i + 1) * dt, y)
    return rec


SOLVERS = {"euler": solve_euler, "heun": solve_heun, "heun_true": solve_heun_true, "rk4": solve_rk4}


# ----------------------------------------------------------------------------- adaptive: solver='scipy' (RK45)
# Restatement of what BaseBackend._solve_scipy (base_backend.py:802-812) executes:
#   solve_ivp(fun=func, t_span=(t0, T), y0=y, first_step=dt, args=args, t_eval=times, **kwargs)
# with SciPy's default method RK45 (Dormand-Prince 5(4), scipy/integrate/_ivp/rk.py: RungeKutta._step_impl, rk_step,
# RkDenseOutput; scipy/integrate/_ivp/ivp.py: the t_eval loop).  SciPy is a third-party dependency of the reference
# (setup.py: 'scipy', unpinned; 1.18.1 installed here); its published algorithm is restated below and pinned bit-exact
# against the real reference + SciPy in tests/golden/*_rk45.npz.
# One reference-specific effect is part of the semantics: the generated rhs returns its in-place ``dy`` buffer
# (computegraph.py:1220-1228) and SciPy keeps ``self.f = f_new`` by reference, so ``K[0]`` of every attempt is simply
# the MOST RECENT rhs evaluation — after a rejected attempt that is f(t + h_rejected, y_rejected), not f(t, y).
RK45_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
RK45_A = np.array([[0, 0, 0, 0, 0], [1 / 5, 0, 0, 0, 0], [3 / 40, 9 / 40, 0, 0, 0], [44 / 45, -56 / 15, 32 / 9, 0, 0],
                   [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0],
                   [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]])
RK45_B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
RK45_E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])
RK45_P = np.array([[1, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432], [0, 0, 0, 0],
                   [0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799],
                   [0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072],
                   [0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632],
                   [0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844],
                   [0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423]])


# SciPy's RK23 (Bogacki-Shampine 3(2), cubic Hermite dense output): same driver, other tableau
RK23_C = np.array([0, 1 / 2, 3 / 4])
RK23_A = np.array([[0, 0, 0], [1 / 2, 0, 0], [0, 3 / 4, 0]])
RK23_B = np.array([2 / 9, 1 / 3, 4 / 9])
RK23_E = np.array([5 / 72, -1 / 12, -1 / 9, 1 / 8])
RK23_P = np.array([[1, -4 / 3, 5 / 9], [0, 1, -2 / 3], [0, 4 / 3, -8 / 9], [0, -1, 1]])
TABLEAUS = {"RK45": (RK45_C, RK45_A, RK45_B, RK45_E, RK45_P, -1 / 5), "RK23": (RK23_C, RK23_A, RK23_B, RK23_E, RK23_P, -1 / 3)}


def solve_rk45(func, args, T, dt, y0, t0, times, rtol=1e-3, atol=1e-6, max_step=np.inf, aliased=True, method="RK45"):
    """solve_ivp(method='RK45', first_step=dt, t_eval=times) around the reference rhs; returns state_rec[len(times), n].
    The state is integrated in float64 whatever the model precision (scipy/integrate/_ivp/base.py: check_arguments).
    ``aliased=True`` is the reference's behaviour (rhs returns its in-place buffer, see above); ``aliased=False`` is plain
    SciPy around a function that returns fresh arrays (a rejected attempt restarts from the true f(t, y))."""
    C_, A_, B_, E_, P_, expo = TABLEAUS[method]
    nst = B_.size
    y = np.asarray(y0).astype(float, copy=False)
    n = y.size
    t, t_bound = float(t0), float(T)
    f = np.asarray(func(t, y, *args), dtype=float)           # alias of the rhs buffer
    h_abs = float(dt)
    K = np.empty((nst + 1, n))
    rec = np.empty((len(times), n))
    i_eval = 0
    while t != t_bound:
        min_step = 10 * np.abs(np.nextafter(t, np.inf) - t)
        if h_abs > max_step:
            h_abs = max_step
        elif h_abs < min_step:
            h_abs = min_step
        rejected = False
        while True:
            if h_abs < min_step:
                raise RuntimeError("Required step size is less than spacing between numbers.")
            t_new = t + h_abs
            if t_new - t_bound > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = np.abs(h)
            K[0] = f
            for s in range(1, nst):
                dy = np.dot(K[:s].T, A_[s, :s]) * h
                K[s] = func(t + C_[s] * h, y + dy, *args)
            y_new = y + h * np.dot(K[:-1].T, B_)
            f_new = np.asarray(func(t + h, y_new, *args), dtype=float)
            K[-1] = f_new
            if aliased:
                f = f_new
            scale = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
            err = np.dot(K.T, E_) * h / scale
            error_norm = np.linalg.norm(err) / err.size ** 0.5
            if error_norm < 1:
                factor = 10.0 if error_norm == 0 else min(10.0, 0.9 * error_norm ** expo)
                if rejected:
                    factor = min(1, factor)
                h_abs *= factor
                f = f_new
                break
            h_abs *= max(0.2, 0.9 * error_norm ** expo)
            rejected = True
        t_old, y_old = t, y
        t, y = t_new, y_new
        i_new = int(np.searchsorted(times, t, side="right"))
        if i_new > i_eval:
            Q = K.T.dot(P_)
            x = (times[i_eval:i_new] - t_old) / h
            p = np.cumprod(np.tile(x, (P_.shape[1], 1)), axis=0)
            rec[i_eval:i_new] = (h * np.dot(Q, p) + y_old[:, None]).T
            i_eval = i_new
    return rec[:i_eval]


# ----------------------------------------------------------------------------- adaptive: solve_ivp(method='DOP853')
# The reference's documentation uses it (documentation/model_analysis/entrainment.py:56,142).  Restatement of SciPy's DOP853
# (scipy/integrate/_ivp/rk.py: class DOP853, rk_step, RungeKutta._step_impl, Dop853DenseOutput; coefficients of Hairer,
# Norsett & Wanner's DOP853 in scipy/integrate/_ivp/dop853_coefficients.py — third-party, read from the installed SciPy):
# 12 stages, error estimate from the 5th- and 3rd-order embedded formulas, controller exponent -1/8, and a 7th-order dense
# output that costs three EXTRA rhs evaluations for every step that contains sample times.  The reference's aliased rhs
# buffer matters twice here: `self.f` is the MOST RECENT evaluation — after a rejected attempt f(t + h_rej, y_rej), and
# after a dense output the last extra stage — and it is what the dense output uses as "f at the step end" and what the next
# attempt starts from.
def solve_dop853(func, args, T, dt, y0, t0, times, rtol=1e-3, atol=1e-6, max_step=np.inf, coeffs=None):
    if coeffs is None:
        from scipy.integrate._ivp import dop853_coefficients as coeffs
    NS, NX = coeffs.N_STAGES, coeffs.N_STAGES_EXTENDED
    A_, C_, B_ = coeffs.A[:NS, :NS], coeffs.C[:NS], coeffs.B
    E3, E5, D_ = coeffs.E3, coeffs.E5, coeffs.D
    A_EXTRA, C_EXTRA = coeffs.A[NS + 1:], coeffs.C[NS + 1:]
    expo = -1 / 8
    y = np.asarray(y0).astype(float, copy=False)
    n = y.size
    t, t_bound = float(t0), float(T)
    f = np.asarray(func(t, y, *args), dtype=float)           # alias of the rhs buffer
    h_abs = float(dt)
    KX = np.empty((NX, n))
    K = KX[:NS + 1]
    rec = np.empty((len(times), n))
    i_eval = 0
    while t != t_bound:
        min_step = 10 * np.abs(np.nextafter(t, np.inf) - t)
        if h_abs > max_step:
            h_abs = max_step
        elif h_abs < min_step:
            h_abs = min_step
        rejected = False
        while True:
            if h_abs < min_step:
                raise RuntimeError("Required step size is less than spacing between numbers.")
            t_new = t + h_abs
            if t_new - t_bound > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = np.abs(h)
            K[0] = f
            for s in range(1, NS):
                dy = np.dot(K[:s].T, A_[s, :s]) * h
                K[s] = func(t + C_[s] * h, y + dy, *args)
            y_new = y + h * np.dot(K[:-1].T, B_)
            f_new = np.asarray(func(t + h, y_new, *args), dtype=float)
            K[-1] = f_new
            f = f_new                                             # aliased buffer
            scale = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
            err5 = np.dot(K.T, E5) / scale
            err3 = np.dot(K.T, E3) / scale
            err5_norm_2 = np.linalg.norm(err5) ** 2
            err3_norm_2 = np.linalg.norm(err3) ** 2
            if err5_norm_2 == 0 and err3_norm_2 == 0:
                error_norm = 0.0
            else:
                denom = err5_norm_2 + 0.01 * err3_norm_2
                error_norm = np.abs(h) * err5_norm_2 / np.sqrt(denom * len(scale))
            if error_norm < 1:
                factor = 10.0 if error_norm == 0 else min(10.0, 0.9 * error_norm ** expo)
                if rejected:
                    factor = min(1, factor)
                h_abs *= factor
                break
            h_abs *= max(0.2, 0.9 * error_norm ** expo)
            rejected = True
        t_old, y_old = t, y
        t, y = t_new, y_new
        i_new = int(np.searchsorted(times, t, side="right"))
        if i_new > i_eval:
            # Dop853 dense output: three extra stages (they overwrite the aliased buffer f), then the 7-term polynomial
            for s, (a, c) in enumerate(zip(A_EXTRA, C_EXTRA), start=NS + 1):
                dy = np.dot(KX[:s].T, a[:s]) * h
                KX[s] = func(t_old + c * h, y_old + dy, *args)
                f = np.asarray(KX[s], dtype=float)                # what the aliased buffer holds now
            F = np.empty((coeffs.INTERPOLATOR_POWER, n))
            f_old = KX[0]
            delta_y = y - y_old
            F[0] = delta_y
            F[1] = h * f_old - delta_y
            F[2] = 2 * delta_y - h * (f + f_old)
            F[3:] = h * np.dot(D_, KX)
            x = ((times[i_eval:i_new] - t_old) / h)[:, None]
            out = np.zeros((len(x), n))
            for i, fi in enumerate(reversed(F)):
                out += fi
                if i % 2 == 0:
                    out *= x
                else:
                    out *= 1 - x
            out += y_old
            rec[i_eval:i_new] = out
            i_eval = i_new
    return rec[:i_eval]


# ----------------------------------------------------------------------------- adaptive DDE: solver='scipy' with a history
# Restatement of BaseBackend._solve_scipy_dde (base_backend.py:771-800):
#   solver = scipy.integrate.ode(rhs).set_integrator('dopri5', first_step=dt, nsteps=50000); solver.set_solout(hist.update)
#   for t_out in times: solver.integrate(t_out); state_rec[i] = solver.y
# i.e. Hairer's DOPRI5 (third-party: SciPy's _dop extension, 1.18.1 installed here; rtol=1e-6, atol=1e-12, safety 0.9,
# factors 0.2 / 10, beta 0.04 — the reference forwards none of the caller's tolerances) restarted for every sample interval
# with h = first_step, and the accepted steps of every call (plus its starting point) appended to the history.
D5_C = (0.2, 0.3, 0.8, 8.0 / 9.0)
D5_A = ((0.2,), (3.0 / 40.0, 9.0 / 40.0), (44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0),
        (19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0),
        (9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0),
        (35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0))
D5_E = (71.0 / 57600.0, 0.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0, -1.0 / 40.0)


def dopri5_call(f, x, y, xend, h, solout, rtol=1e-6, atol=1e-12, nmax=50000, safe=0.9, fac1=0.2, fac2=10.0, beta=0.04):
    """One ``dopri5`` call (Hairer, Norsett, Wanner: DOPCOR) from x to xend with initial step h.  Returns (x, y, ok)."""
    uround = 2.3e-16
    n = y.size
    facold, expo1 = 1e-4, 0.2 - beta * 0.75
    facc1, facc2 = 1.0 / fac1, 1.0 / fac2
    posneg = 1.0 if xend - x >= 0 else -1.0
    hmax = abs(xend - x)
    last, reject = False, False
    k1 = np.array(f(x, y), dtype=float)
    nstep = naccpt = 0
    solout(x, y)
    while True:
        if nstep > nmax:
            return x, y, False
        if 0.1 * abs(h) <= abs(x) * uround:
            return x, y, False
        if (x + 1.01 * h - xend) * posneg > 0.0:
            h = xend - x
            last = True
        nstep += 1
        a = D5_A
        k2 = np.array(f(x + D5_C[0] * h, y + h * a[0][0] * k1), dtype=float)
        k3 = np.array(f(x + D5_C[1] * h, y + h * (a[1][0] * k1 + a[1][1] * k2)), dtype=float)
        k4 = np.array(f(x + D5_C[2] * h, y + h * (a[2][0] * k1 + a[2][1] * k2 + a[2][2] * k3)), dtype=float)
        k5 = np.array(f(x + D5_C[3] * h, y + h * (a[3][0] * k1 + a[3][1] * k2 + a[3][2] * k3 + a[3][3] * k4)), dtype=float)
        ysti = y + h * (a[4][0] * k1 + a[4][1] * k2 + a[4][2] * k3 + a[4][3] * k4 + a[4][4] * k5)
        xph = x + h
        k6 = np.array(f(xph, ysti), dtype=float)
        y1 = y + h * (a[5][0] * k1 + a[5][2] * k3 + a[5][3] * k4 + a[5][4] * k5 + a[5][5] * k6)
        k7 = np.array(f(xph, y1), dtype=float)
        e = D5_E
        k4 = (e[0] * k1 + e[2] * k3 + e[3] * k4 + e[4] * k5 + e[5] * k6 + e[6] * k7) * h
        sk = atol + rtol * np.maximum(np.abs(y), np.abs(y1))
        err = 0.0
        for i in range(n):
            err += (k4[i] / sk[i]) ** 2
        err = np.sqrt(err / n)
        fac11 = err ** expo1
        fac = fac11 / facold ** beta
        fac = max(facc2, min(facc1, fac / safe))
        hnew = h / fac
        if err <= 1.0:
            facold = max(err, 1e-4)
            naccpt += 1
            k1, y, x = k7, y1, xph
            solout(x, y)
            if last:
                return x, y, True
            if abs(hnew) > hmax:
                hnew = posneg * hmax
            if reject:
                hnew = posneg * min(abs(hnew), abs(h))
            reject = False
        else:
            hnew = h / min(facc1, fac11 / safe)
            reject = True
            last = False
        h = hnew


def solve_dopri5_dde(func, args, T, dt, y0, t0, times, **kwargs):
    """base_backend.py:771-800; ``args[0]`` is the History.  Returns state_rec[len(times), n] (rows after a failed call stay
    zero, like the reference's ``if not solver.successful(): break``)."""
    hist = args[0]
    y = np.asarray(y0).astype(float)          # scipy.integrate.ode works on float64 copies
    x = float(t0)
    rec = np.zeros((len(times), y.shape[0]), dtype=np.asarray(y0).dtype)

    def rhs(t, y_):
        # the extension hands the callback a Python float: `t - d` with a float32 constant d is then a float32 operation
        return func(float(t), y_, *args)

    ok = True
    for i, t_out in enumerate(times):
        if not ok:
            break
        x, y, ok = dopri5_call(rhs, x, y, float(t_out), float(dt), hist.update)
        rec[i, :] = y
    return rec


def run_adaptive(func, func_args: Sequence, T: float, dt: float, dts: float = None, **kw):
    """base_backend.py:596-611 with solver='scipy' (method RK45)."""
    t0, y0 = func_args[0], func_args[1]
    step = dts if dts else dt
    times = np.linspace(0.0, T, num=round(T / step), endpoint=False)
    if _dde(func_args[2:]) is not None:       # base_backend.py:704-706
        return solve_dopri5_dde(func, tuple(func_args[2:]), T, dt, y0, t0, times, **kw), times
    if kw.get("method") == "DOP853":
        kw = {k: v for k, v in kw.items() if k != "method"}
        return solve_dop853(func, tuple(func_args[2:]), T, dt, y0, t0, times, **kw), times
    return solve_rk45(func, tuple(func_args[2:]), T, dt, y0, t0, times, **kw), times


def run(func, func_args: Sequence, T: float, dt: float, dts: float = None, solver: str = "euler"):
    """base_backend.py:596-611: returns ``(state_rec[round(T/dts), n_state], times)``.
    ``func_args = (t0, y0, dy, *consts)``; ``y0`` and mutable buffers are modified in place like
    the reference does."""
    t0, y0 = func_args[0], func_args[1]
    step = dts if dts else dt
    times = np.linspace(0.0, T, num=round(T / step), endpoint=False)
    rec = SOLVERS[solver](func, tuple(func_args[2:]), T, dt, step, y0, t0)
    return rec, times


# ----------------------------------------------------------------------------- fixtures
def load_fixture(path: str) -> dict:
    """Load a golden fixture written by oracle/make_golden.py (format of pyrates_b200.Program.save
    plus reference outputs).  Independent of the product package on purpose."""
    with np.load(path, allow_pickle=False) as d:
        meta = json.loads(bytes(np.asarray(d["__program__"]).tobytes()).decode())
        args = [np.array(d[f"arg{i}"]) for i in range(len(meta["args"]))]
        extra = {k: np.array(d[k]) for k in d.files if not k.startswith("arg") and k != "__program__"}
    names = [a["name"] for a in meta["args"]]
    lines = [f"def {meta['func_name']}({','.join(names)}):"]
    for lhs, rhs, idx in meta["stmts"]:
        lines.append(f"\t{lhs}[{idx}] = {rhs}" if idx is not None else f"\t{lhs} = {rhs}")
    ret = next(a["name"] for a in meta["args"] if a["kind"] == "dy")
    lines.append(f"\treturn {ret}")
    run_meta = json.loads(bytes(extra.pop("__run__").tobytes()).decode()) if "__run__" in extra else {}
    return {"meta": meta, "args": args, "names": names, "source": "\n".join(lines) + "\n",
            "run": run_meta, "extra": extra}


def fixture_args(fx: dict) -> list:
    """Fresh argument list of a loaded fixture; the 'hist' argument of delay-differential fixtures (stored as the float64
    initial state the reference built its DDEHistory from) becomes a new History."""
    args = [np.array(a, copy=True) for a in fx["args"]]
    args[0] = args[0][()] if args[0].ndim == 0 else args[0]
    for k, m in enumerate(fx["meta"]["args"]):
        if m["kind"] == "hist":
            args[k] = History(args[k])
    return args


def run_fixture(fx: dict, solver: str = None, T: float = None, dt: float = None, dts: float = None):
    """Integrate a loaded fixture with the oracle; returns ``(state_rec, times)``."""
    r = fx["run"]
    func = build_vector_field(fx["source"], fx["meta"]["func_name"])
    args = fixture_args(fx)
    if (solver or r["solver"]) == "scipy":
        return run_adaptive(func, args, T if T is not None else r["T"], dt if dt is not None else r["dt"],
                            dts if dts is not None else r.get("dts"))
    return run(func, args, T if T is not None else r["T"], dt if dt is not None else r["dt"],
               dts if dts is not None else r.get("dts"), solver or r["solver"])
