"""Host-side driver of the instance execution model: Program -> JIT kernel -> device session -> results.

This is the CUDA counterpart of ``BaseBackend.run`` + ``_solve_*`` (pyrates/backend/base/base_backend.py:
596-611, 693-769): ``InstanceModel`` owns the compiled kernel of one model, ``InstanceSession`` the
device buffers of one batch of B sweep instances.  Everything numerical happens in the generated kernel;
the host only allocates, copies and launches through the C ABI (``pyrates_b200.runtime``).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from . import runtime as rt
from .emit_inst import InstKernel, emit_inst_kernel
from .program import Program
from .scalarize import ScalarRhs, UnsupportedModelError, scalarize

__all__ = ["InstanceModel", "InstanceSession", "n_steps_for", "FIXED_STEP_SOLVERS", "ADAPTIVE_SOLVERS", "evaluate_rhs"]

FIXED_STEP_SOLVERS = ("euler", "heun", "heun_true", "rk4")
ADAPTIVE_SOLVERS = ("scipy", "rk45")        # the reference's solver='scipy' with SciPy's default method RK45


def n_steps_for(T: float, dt: float, dts: Optional[float]) -> Tuple[int, int, int]:
    """(steps, store_step, n_samples) exactly as the reference computes them (base_backend.py:716-719)."""
    dts = dts if dts else dt
    steps = int(np.round(T / dt))
    n_samples = int(np.round(T / dts))
    store_step = int(np.round(dts / dt))
    return steps, max(store_step, 1), n_samples


class InstanceModel:
    """One model compiled for the instance execution model (one thread per sweep instance)."""

    def __init__(self, prog: Program, *, solver: str = "euler", sweep: Sequence[Tuple[str, int]] = (),
                 observers: Optional[Sequence[int]] = None, block: int = 128, fmad: bool = True,
                 const_div_to_mul: bool = False, max_nodes: int = 6000, min_blocks: int = 1, fast_math: bool = False,
                 rk_method: Optional[str] = None, reduce: bool = False, hist_max_delay: Optional[float] = None,
                 y_smem: bool = False, burst: Optional[bool] = None):
        if solver not in FIXED_STEP_SOLVERS + ADAPTIVE_SOLVERS:
            raise rt.CudaBackendError(f"Backend `CudaBackend` does not support solver `{solver}`. "
                                      f"Supported solvers: {list(FIXED_STEP_SOLVERS + ADAPTIVE_SOLVERS)}.")
        self.prog = prog
        self.solver = solver
        self.adaptive = solver in ADAPTIVE_SOLVERS
        # delay-differential models: the reference integrates them with scipy.integrate.ode('dopri5') and ignores `method`
        # and the tolerances (BaseBackend._solve_scipy_dde, base_backend.py:771-800)
        self.dde = prog.is_dde
        self.rk_method = (rk_method or "RK45") if (self.adaptive and not self.dde) else None     # solve_ivp's `method`
        if self.adaptive and not self.dde and self.rk_method not in ("RK45", "RK23", "DOP853"):
            raise rt.CudaBackendError(f"solver='scipy' on the CUDA backend implements SciPy's explicit Runge-Kutta methods "
                                      f"'RK45' (default), 'RK23' and 'DOP853'; got method={self.rk_method!r}")
        self.precision = "float32" if prog.real_dtype == np.float32 else "float64"      # complex64/128 -> real pairs
        self.real = prog.real_dtype
        self.rhs: ScalarRhs = scalarize(prog, sweep, max_nodes=max_nodes)
        self.observers = list(range(prog.n_state)) if observers is None else [int(o) for o in observers]
        self.kernel_solver = "dopri5_dde" if (self.adaptive and self.dde) else solver
        self.kernel: InstKernel = emit_inst_kernel(self.rhs, solver=self.kernel_solver, precision=self.precision,
                                                   observers=self.observers, block=block,
                                                   const_div_to_mul=const_div_to_mul, min_blocks=min_blocks,
                                                   fast_math=fast_math, rk_method=self.rk_method or "RK45", reduce=reduce,
                                                   hist_max_delay=hist_max_delay, y_smem=y_smem, burst=burst)
        # swept delays of delay-differential models: the bound the history was sized for (checked against every table)
        self.hist_max_delay = None if hist_max_delay is None else float(hist_max_delay)
        self.reduce = bool(reduce)
        self.fmad = fmad
        self._module: Optional[rt.Module] = None
        t0 = np.asarray(prog.arg_by_kind("t").value)
        self.t0 = float(t0) if self.adaptive else int(t0)
        # history of delay-differential models: entry 0 is the float64 initial state the reference builds its DDEHistory
        # from (computegraph.py:471-478), which is NOT rounded to the model precision
        self.hist_y0 = np.zeros(0)
        if self.kernel.n_hist:
            h0 = np.asarray(next(a.value for a in prog.args if a.kind == "hist"), dtype=np.float64).reshape(-1)
            self.hist_y0 = h0[self.kernel.hist_vars]

    @property
    def module(self) -> rt.Module:
        return self.compile()

    def compile(self, shared: bool = False) -> rt.Module:
        """NVRTC-compile the kernel (cached).  ``shared``: under an initialised torch.distributed group rank 0 compiles and
        the other ranks receive the cubin — every rank of a sharded sweep runs the same source (collective call)."""
        if self._module is None:
            self._module = rt.compile_module(self.kernel.source, f"prb_inst_{self.kernel_solver}.cu", fmad=self.fmad, shared=shared)
        return self._module

    @property
    def n_state(self) -> int:
        return self.kernel.n_state

    @property
    def n_params(self) -> int:
        return self.kernel.n_params

    def session(self, n_inst: int = 1, n_samples: int = 0, stream=None) -> "InstanceSession":
        return InstanceSession(self, n_inst, n_samples, stream)

    # convenience: the drop-in path of BaseBackend.run for a single instance
    def run(self, T: float, dt: float, dts: Optional[float] = None, *, params: Optional[np.ndarray] = None,
            y0: Optional[np.ndarray] = None, n_inst: Optional[int] = None, **solver_kwargs):
        """Fixed-step solvers: ``(out[n_samples, n_obs, B], y_final)``.  ``solver='scipy'`` accepts SciPy's ``rtol``,
        ``atol``, ``max_step`` and ``method='RK45'`` keywords (base_backend.py:802-812 forwards them to solve_ivp)."""
        steps, store_step, n_samples = n_steps_for(T, dt, dts)
        if n_inst is None:
            n_inst = 1 if params is None else int(np.asarray(params).shape[-1])
        if self.adaptive:
            n_samples = int(round(T / (dts if dts else dt)))
        s = self.session(n_inst, n_samples)
        try:
            s.reset(y0=y0, params=params)
            if self.adaptive:
                s.launch_adaptive(T, dt, dts, **solver_kwargs)
            else:
                if solver_kwargs:
                    raise TypeError(f"unexpected solver keywords for a fixed-step solver: {sorted(solver_kwargs)}")
                s.launch(steps, dt, store_step)
            out = s.download()
            y_final = s.download_state()
        finally:
            s.free()
        return out, y_final


class InstanceSession:
    """Device buffers of a batch of B instances of one ``InstanceModel``."""

    def __init__(self, model: InstanceModel, n_inst: int, n_samples: int, stream=None):
        self.m = model
        k = model.kernel
        self.B = int(n_inst)
        self.n_samples = int(n_samples)
        self.stream = stream
        r = np.dtype(model.real).itemsize
        B = self.B
        self.d_y = rt.DeviceBuffer(k.n_state * B * r)
        self.d_out = rt.DeviceBuffer(max(self.n_samples * k.n_obs * B * r, 8))
        self.d_params = rt.DeviceBuffer(max(k.n_params * B * r, 8))
        self.d_slots = rt.DeviceBuffer(max(k.n_slots * B * r, 8))
        self.d_rings = rt.DeviceBuffer(max(k.ring_elems * B * r, 8))
        self.d_tables = rt.DeviceBuffer.from_host(np.asarray(k.table_values, dtype=model.real)) \
            if k.table_values.size else rt.DeviceBuffer(8)
        # delay-differential models, fixed-step solvers: ring [hist_depth][n_hist][B] of past states (the adaptive path
        # sizes its (time, state) ring per run, launch_adaptive)
        self.d_hist = rt.DeviceBuffer(k.hist_depth * k.n_hist * B * r) if (k.n_hist and k.hist_depth) else None
        self.steps_done = 0
        self.samples_done = 0

    # -- state management
    def reset(self, y0: Optional[np.ndarray] = None, params: Optional[np.ndarray] = None):
        """(Re-)initialise state, parameters, slots and ring buffers; host->device copies."""
        k, B, real = self.m.kernel, self.B, self.m.real
        y0_given = None if y0 is None else np.asarray(y0)
        y0 = self.m.prog.y0 if y0 is None else np.asarray(y0)
        if y0_given is not None and y0_given.ndim == 1:
            y0_given = np.repeat(y0_given[:, None], B, axis=1)
        if y0.ndim == 1:
            y0 = np.repeat(y0.astype(real)[:, None], B, axis=1)
        if y0.shape != (k.n_state, B):
            raise ValueError(f"y0 must have shape ({k.n_state},) or ({k.n_state}, {B})")
        self.d_y.upload(np.ascontiguousarray(y0, dtype=real), self.stream)
        if k.n_params:
            if params is None:
                params = np.repeat(k.param_defaults.astype(real)[:, None], B, axis=1)
            self.upload_params(params)
        if k.n_slots:
            self.d_slots.upload(np.repeat(k.slot_init.astype(real)[:, None], B, axis=1), self.stream)
        if k.ring_elems:
            self.d_rings.upload(np.repeat(k.ring_init.astype(real)[:, None], B, axis=1), self.stream)
        if self.d_hist is not None:
            h = np.zeros((k.hist_depth, k.n_hist, B), dtype=real)
            h[0] = (self.m.hist_y0.astype(real)[:, None] if y0_given is None
                    else np.ascontiguousarray(y0_given, dtype=real)[k.hist_vars])          # entry 0: the state at t0
            self.d_hist.upload(h, self.stream)
        self.steps_done = 0
        self.samples_done = 0

    def upload_params(self, params: np.ndarray):
        k, B = self.m.kernel, self.B
        params = np.ascontiguousarray(params, dtype=self.m.real)
        if params.shape != (k.n_params, B):
            raise ValueError(f"params must have shape ({k.n_params}, {B}), got {params.shape}")
        self.check_delays(params)
        self.d_params.upload(params, self.stream)

    def check_delays(self, params: np.ndarray):
        """Swept delays must stay inside [0, hist_max_delay]: the history was sized for that bound."""
        rows = self.m.rhs.hist_delay_params
        if rows:
            d = np.asarray(params)[rows]
            bound = self.m.hist_max_delay if self.m.hist_max_delay is not None else max(self.m.rhs.hist_delays)
            if not np.all((d >= 0) & (d <= bound * (1 + 1e-12))):
                raise rt.CudaBackendError(f"swept history delays must lie in [0, {bound}] (hist_max_delay); got "
                                          f"[{float(d.min())}, {float(d.max())}]")
        self.delay_max = float(np.asarray(params)[rows].max()) if rows else None

    # -- launch
    def samples_in(self, n_steps: int, store_step: int) -> int:
        """Samples the next ``n_steps`` steps will record (sampling happens before a step when
        step % store_step == 0, base_backend.py:730-733)."""
        if n_steps <= 0:
            return 0
        first = (-self.steps_done) % store_step
        return 0 if first >= n_steps else (n_steps - 1 - first) // store_step + 1

    def launch(self, n_steps: int, dt: float, store_step: int = 1, out: Optional[rt.DeviceBuffer] = None,
               out_capacity: Optional[int] = None):
        """Advance all instances by ``n_steps`` solver steps (asynchronous on the session's stream).
        With ``out`` the samples of this launch go to rows 0.. of that buffer (chunked/streamed runs)
        instead of being appended to the session's own recording."""
        k = self.m.kernel
        if self.m.reduce:
            raise rt.CudaBackendError("this model reduces its observers on the device: use launch_reduce")
        new_samples = self.samples_in(n_steps, store_step)
        if out is not None:
            if new_samples > (out_capacity if out_capacity is not None else new_samples):
                raise ValueError("chunk buffer too small for this launch")
        elif self.samples_done + new_samples > self.n_samples:
            raise ValueError(f"launch would record {self.samples_done + new_samples} samples but the session "
                             f"holds only {self.n_samples}")
        a = rt.KernelArgs()
        a.n_steps, a.store_step, a.step0 = int(n_steps), int(store_step), int(self.steps_done)
        a.t0, a.eval0 = self.m.t0, int(self.steps_done) * k.evals_per_step
        a.n_inst, a.sample0, a.dt = self.B, (0 if out is not None else int(self.samples_done)), float(dt)
        a.y, a.out, a.params = self.d_y.ptr, (out.ptr if out is not None else self.d_out.ptr), self.d_params.ptr
        a.slots, a.rings, a.tables = self.d_slots.ptr, self.d_rings.ptr, self.d_tables.ptr
        a.consts = a.iconsts = a.work = a.sync = None
        if self.d_hist is not None:
            self._check_hist_dt(dt)
            a.work = self.d_hist.ptr
        self.m.module.run(k.kernel_name, a, block=k.block, stream=self.stream)
        self.steps_done += int(n_steps)
        self.samples_done += new_samples
        return new_samples

    def _check_hist_dt(self, dt: float):
        """Delay-differential models, fixed-step solvers: the generated history lookups `hist(t*dt - d)` carry the step size
        the vector field was built for; integrating with another one would silently interpolate a wrong history."""
        hdt = self.m.rhs.hist_dt
        if abs(float(dt) - hdt) > 1e-9 * hdt:
            raise rt.CudaBackendError(f"this delay-differential vector field was generated for step_size={hdt!r} (its history "
                                      f"lookups are `hist(t*{hdt!r} - d)`); got dt={dt!r}")

    REDUCE_STATS = ("mean", "var", "min", "max", "last")

    def launch_reduce(self, n_steps: int, dt: float, store_step: int = 1, skip_samples: int = 0,
                      out: Optional[rt.DeviceBuffer] = None):
        """Models built with ``reduce=True``: integrate ``n_steps`` steps in ONE launch and leave
        ``[len(REDUCE_STATS), n_obs, B]`` running statistics of the decimated samples ``>= skip_samples`` in the output
        buffer (numpy.mean / var / min / max over the sample axis, and the last sample)."""
        if not self.m.reduce:
            raise rt.CudaBackendError("launch_reduce needs a model built with reduce=True")
        k = self.m.kernel
        need = len(self.REDUCE_STATS) * k.n_obs * self.B * np.dtype(self.m.real).itemsize
        if out is None and self.d_out.nbytes < need:
            self.d_out.free()
            self.d_out = rt.DeviceBuffer(need)
        a = rt.KernelArgs()
        a.n_steps, a.store_step, a.step0 = int(n_steps), int(store_step), 0
        a.t0, a.eval0 = self.m.t0, 0
        a.n_inst, a.sample0, a.dt = self.B, int(skip_samples), float(dt)
        a.y, a.out, a.params = self.d_y.ptr, (out.ptr if out is not None else self.d_out.ptr), self.d_params.ptr
        a.slots, a.rings, a.tables = self.d_slots.ptr, self.d_rings.ptr, self.d_tables.ptr
        a.consts = a.iconsts = a.work = a.sync = None
        if self.d_hist is not None:
            self._check_hist_dt(dt)
            a.work = self.d_hist.ptr
        self.m.module.run(k.kernel_name, a, block=k.block, stream=self.stream)
        self.steps_done = int(n_steps)

    def download_reduced(self) -> np.ndarray:
        out = np.empty((len(self.REDUCE_STATS), self.m.kernel.n_obs, self.B), dtype=self.m.real)
        self.d_out.download(out, self.stream)
        if self.stream is not None:
            self.stream.synchronize() if hasattr(self.stream, "synchronize") else rt.synchronize()
        return out

    def launch_adaptive(self, T: float, dt: float, dts: Optional[float] = None, *, rtol: float = 1e-3, atol: float = 1e-6,
                        max_step: float = np.inf, method: Optional[str] = None, max_attempts: int = 50_000_000, **unknown):
        """The whole adaptive integration of every instance in one launch: solve_ivp(method='RK45', first_step=dt,
        t_eval=linspace(0, T, round(T/dts), endpoint=False)) per instance, each thread with its own step sequence.
        Raises when an instance fails (step size underflow), like a failed solve_ivp would break the reference's run."""
        if self.m.dde:
            return self._launch_dopri5_dde(T, dt, dts)
        if method is not None and method != self.m.rk_method:
            raise rt.CudaBackendError(f"this model was compiled for method={self.m.rk_method!r}; build it with "
                                      f"rk_method={method!r} (the tableau is a compile-time constant of the kernel)")
        if unknown:
            raise TypeError(f"unsupported solve_ivp options on the CUDA backend: {sorted(unknown)}")
        if np.ndim(atol) or np.ndim(rtol):
            raise rt.CudaBackendError("per-component tolerances are not supported; pass scalar rtol / atol")
        k = self.m.kernel
        step = dts if dts else dt
        n_samples = int(round(T / step))
        if n_samples > self.n_samples:
            raise ValueError(f"session holds {self.n_samples} samples, the run records {n_samples}")
        rtol = max(float(rtol), 100 * np.finfo(float).eps)          # validate_tol, scipy/integrate/_ivp/common.py
        te_step = (float(T) - 0.0) / n_samples if n_samples else 0.0   # numpy.linspace: step = delta / div
        opts = np.array([rtol, float(atol), self.m.t0, float(T), float(dt), te_step, float(max_step), float(max_attempts)])
        for name in ("d_opts", "d_status"):                     # buffers of an earlier launch_adaptive call
            old = getattr(self, name, None)
            if old is not None:
                old.free()
        self.d_opts = rt.DeviceBuffer.from_host(opts)
        self.d_status = rt.DeviceBuffer(4 * self.B)
        self.d_status.zero(self.stream)
        a = rt.KernelArgs()
        a.n_steps, a.store_step, a.step0, a.t0, a.eval0 = n_samples, 1, 0, 0, 0
        a.n_inst, a.sample0, a.dt = self.B, 0, float(dt)
        a.y, a.out, a.params = self.d_y.ptr, self.d_out.ptr, self.d_params.ptr
        a.slots, a.rings, a.tables = self.d_slots.ptr, self.d_rings.ptr, self.d_tables.ptr
        a.consts, a.sync = self.d_opts.ptr, self.d_status.ptr
        a.iconsts = a.work = None
        self.m.module.run(k.kernel_name, a, block=k.block, stream=self.stream)
        status = np.zeros(self.B, dtype=np.uint32)
        self.d_status.download(status, self.stream)
        if self.stream is not None:
            self.stream.synchronize()
        self.samples_done = n_samples
        self.attempts = (status >> 4).astype(np.int64)        # RK45 step attempts per instance (6 rhs evaluations each)
        status = status & 15
        if status.any():
            bad = np.nonzero(status)[0]
            why = {1: "Required step size is less than spacing between numbers.", 2: "attempt budget exhausted"}
            raise rt.CudaBackendError(f"adaptive integration failed for {bad.size} instance(s), first {int(bad[0])}: "
                                      f"{why.get(int(status[bad[0]]), 'unknown')}")
        return n_samples

    def _launch_dopri5_dde(self, T: float, dt: float, dts: Optional[float] = None, *, nsteps: int = 50000,
                           capacity: Optional[int] = None, max_capacity_bytes: int = 8 << 30):
        """BaseBackend._solve_scipy_dde (base_backend.py:771-800) for every instance in one launch: dopri5 (rtol 1e-6, atol
        1e-12, first_step=dt, nsteps=50000) restarted per sample interval, accepted steps appended to the device history.
        The history is a ring of ``capacity`` (time, state) entries per instance; a run that needs an entry that was
        already overwritten reports status 3 and is repeated with a ring twice as large."""
        k, B = self.m.kernel, self.B
        step = dts if dts else dt
        n_samples = int(round(T / step))
        if n_samples > self.n_samples:
            raise ValueError(f"session holds {self.n_samples} samples, the run records {n_samples}")
        te_step = (float(T) - 0.0) / n_samples if n_samples else 0.0
        dmax = max(self.m.rhs.hist_delays + ([self.m.hist_max_delay] if self.m.hist_max_delay is not None else []))
        if capacity is None:
            per_interval = 2 + min(int(np.ceil(te_step / dt)) if dt > 0 else 1, 8)
            capacity = max(256, int(2 * (np.ceil(dmax / max(te_step, 1e-300)) + 2) * per_interval))
        y_start = np.empty((k.n_state, B), dtype=self.m.real)
        self.d_y.download(y_start, self.stream)
        if self.stream is not None:
            self.stream.synchronize() if hasattr(self.stream, "synchronize") else rt.synchronize()
        if getattr(self, "d_status", None) is not None:
            self.d_status.free()
        self.d_status = rt.DeviceBuffer(4 * B)
        while True:
            if capacity * (1 + k.n_hist) * B * 8 > max_capacity_bytes:
                raise rt.CudaBackendError(f"the state history of this delay-differential run needs more than "
                                          f"{max_capacity_bytes >> 20} MiB (capacity {capacity} entries x {B} instances)")
            opts = np.concatenate([[1e-6, 1e-12, self.m.t0, float(T), float(dt), te_step, 0.0, float(nsteps), float(capacity)],
                                   self.m.hist_y0])
            d_opts = rt.DeviceBuffer.from_host(opts)
            d_work = rt.DeviceBuffer(capacity * (1 + k.n_hist) * B * 8)
            try:
                self.d_status.zero(self.stream)
                a = rt.KernelArgs()
                a.n_steps, a.store_step, a.step0, a.t0, a.eval0 = n_samples, 1, 0, 0, 0
                a.n_inst, a.sample0, a.dt = B, 0, float(dt)
                a.y, a.out, a.params = self.d_y.ptr, self.d_out.ptr, self.d_params.ptr
                a.slots, a.rings, a.tables = self.d_slots.ptr, self.d_rings.ptr, self.d_tables.ptr
                a.consts, a.sync, a.work = d_opts.ptr, self.d_status.ptr, d_work.ptr
                a.iconsts = None
                self.m.module.run(k.kernel_name, a, block=k.block, stream=self.stream)
                status = np.zeros(B, dtype=np.uint32)
                self.d_status.download(status, self.stream)
                if self.stream is not None:
                    self.stream.synchronize() if hasattr(self.stream, "synchronize") else rt.synchronize()
            finally:
                d_opts.free()
                d_work.free()
            self.hist_capacity = capacity
            if not np.any((status & 15) == 3):
                break
            capacity *= 2                                   # history ring too small: restart from the initial state
            self.d_y.upload(y_start, self.stream)
        self.samples_done = n_samples
        self.attempts = (status >> 4).astype(np.int64)        # dopri5 steps (accepted + rejected) per instance
        status = status & 15
        if status.any():
            bad = np.nonzero(status)[0]
            why = {1: "step size becomes too small", 2: "larger nsteps is needed"}
            raise rt.CudaBackendError(f"adaptive integration (dopri5) failed for {bad.size} instance(s), first {int(bad[0])}: "
                                      f"{why.get(int(status[bad[0]]), 'unknown')}")
        return n_samples

    # -- results
    def download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Recorded observers as ``[n_samples, n_obs, B]`` (row k = state at step k*store_step)."""
        k = self.m.kernel
        shape = (self.n_samples, k.n_obs, self.B)
        if out is None:
            out = np.empty(shape, dtype=self.m.real)
        self.d_out.download(out.reshape(-1)[: int(np.prod(shape))].reshape(shape) if out.shape != shape else out,
                            self.stream)
        if self.stream is not None:
            self.stream.synchronize() if hasattr(self.stream, "synchronize") else rt.synchronize()
        return out

    def download_state(self) -> np.ndarray:
        y = np.empty((self.m.kernel.n_state, self.B), dtype=self.m.real)
        self.d_y.download(y, self.stream)
        if self.stream is not None:
            self.stream.synchronize() if hasattr(self.stream, "synchronize") else rt.synchronize()
        return y

    def free(self):
        for n in ("d_y", "d_out", "d_params", "d_slots", "d_rings", "d_tables", "d_opts", "d_status", "d_hist"):
            b = getattr(self, n, None)
            if b is not None:
                b.free()


_EVAL_MODELS: "dict" = {}          # structure key -> InstanceModel of the single-evaluation kernels (insertion-ordered, bounded)
_EVAL_MODELS_MAX = 64


def _eval_model(prog: Program, float_time: bool) -> InstanceModel:
    """InstanceModel for ONE evaluation of ``prog``; cached by program structure.  Scalar float constants become runtime
    parameters, so calls that only change parameter values (fitting / continuation loops around get_run_func /
    get_jacobian_func, finite-difference probes) reuse both the lowered model and the compiled kernel."""
    solver = "scipy" if float_time else "euler"
    cand = [(a.name, e) for a in prog.args if a.kind == "const" and np.issubdtype(a.value.dtype, np.floating)
            and a.value.size <= 8 for e in range(a.value.size)]
    if not 0 < len(cand) <= 96:
        cand = []
    swept = {n for n, _ in cand}

    def key(with_values_of_swept: bool):
        sig = []
        for a in prog.args:
            v = np.asarray(a.value)
            baked = a.kind in ("const", "var", "hist") and (with_values_of_swept or a.name not in swept)
            sig.append((a.name, a.kind, v.shape, str(v.dtype), v.tobytes() if baked else None))
        return (prog.source(), prog.float_precision, solver, tuple(sig))

    for k, sweep in ((key(False), cand), (key(True), [])):
        if k in _EVAL_MODELS:
            m = _EVAL_MODELS[k]
            if m is not None:
                return m
            continue                                # swept form known not to lower: try the baked form
        try:
            m = InstanceModel(prog, solver=solver, sweep=sweep)
        except (KeyError, UnsupportedModelError):
            if not sweep:
                raise
            m = None                                # a constant that must stay compile-time (unused, delay, table ...)
        while len(_EVAL_MODELS) >= _EVAL_MODELS_MAX:
            _EVAL_MODELS.pop(next(iter(_EVAL_MODELS)))
        _EVAL_MODELS[k] = m
        if m is not None:
            return m
    raise AssertionError("unreachable")


def evaluate_rhs(prog: Program, t: Optional[float] = None, y: Optional[np.ndarray] = None) -> np.ndarray:
    """One evaluation ``dy = f(t, y, *args)`` of the vector field on the GPU (kernel ``prb_rhs_eval``).  Programs built
    for an adaptive solver carry a float ``t`` (optionally overridden by ``t``).  ``y`` of shape ``(n_state, B)`` evaluates a
    batch of B states in one launch (one thread each) and returns ``(n_out, B)``; Jacobian programs
    (``CudaBackend`` recording ``get_jacobian_func``) return their ``n*n`` entries row-major."""
    t_arg = np.asarray(prog.arg_by_kind("t").value)
    float_time = not np.issubdtype(t_arg.dtype, np.integer)
    m = _eval_model(prog, float_time)
    n_out = m.rhs.n_out if m.rhs.n_out is not None else m.n_state
    batch = None if y is None or np.ndim(y) < 2 else int(np.shape(y)[1])
    B = batch or 1
    # the cached model may have been built from other values: state, parameters and time come from THIS program
    y0 = prog.y0 if y is None else np.asarray(y, dtype=m.real).reshape(m.n_state, B)
    params = None
    if m.kernel.n_params:
        vals = np.array([np.asarray(prog.arg(ps.arg).value).reshape(-1)[ps.elem] for ps in m.rhs.params], dtype=m.real)
        params = np.repeat(vals[:, None], B, axis=1)
    hist_y0 = m.hist_y0
    if m.kernel.n_hist:
        h0 = np.asarray(next(a.value for a in prog.args if a.kind == "hist"), dtype=np.float64).reshape(-1)
        hist_y0 = h0[m.kernel.hist_vars]
    s = m.session(B, 0)
    d_dy = rt.DeviceBuffer(max(n_out, 1) * B * np.dtype(m.real).itemsize)
    d_extra = []
    try:
        s.reset(y0=y0, params=params)
        a = rt.KernelArgs()
        a.n_steps, a.store_step, a.step0, a.eval0, a.n_inst, a.sample0 = 0, 1, 0, 0, B, 0
        a.t0, a.dt = (0, float(t_arg if t is None else t)) if float_time else (int(t_arg), 0.0)
        a.y, a.out, a.params = s.d_y.ptr, d_dy.ptr, s.d_params.ptr
        a.slots, a.rings, a.tables = s.d_slots.ptr, s.d_rings.ptr, s.d_tables.ptr
        if m.kernel.n_hist and float_time:          # history = the initial condition only (a fresh DDEHistory)
            if B != 1:
                raise rt.CudaBackendError("batched evaluation of delay-differential vector fields is not supported")
            d_extra.append(rt.DeviceBuffer.from_host(np.concatenate([[0, 0, float(t_arg), 0, 0, 0, 0, 0, 1.0], hist_y0])))
            d_extra.append(rt.DeviceBuffer((1 + m.kernel.n_hist) * 8))
            a.consts, a.work = d_extra[0].ptr, d_extra[1].ptr
        elif m.kernel.n_hist:
            h = np.zeros((m.kernel.hist_depth, m.kernel.n_hist, B), dtype=m.real)
            h[0] = hist_y0.astype(m.real)[:, None]
            s.d_hist.upload(h)
            a.work, a.dt = s.d_hist.ptr, float(m.rhs.hist_dt)
        m.module.run("prb_rhs_eval", a, block=m.kernel.block)
        dy = np.empty((n_out, B), dtype=m.real)
        d_dy.download(dy)
        return dy if batch else dy[:, 0]
    finally:
        d_dy.free()
        for d in d_extra:
            d.free()
        s.free()
