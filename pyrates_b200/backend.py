"""``CudaBackend`` — the object PyRates' ``ComputeGraph`` drives when ``backend='cuda'``.

It mirrors, name for name, the interface the reference's compute graph calls on a backend object
(reference implementation: ``BaseBackend``, pyrates/backend/base/base_backend.py:229-812):

  ctor kwargs ............ float_precision / int_precision / file_name / add_hist_arg   (:271-324)
  get_var ................ argument arrays in the backend's precision                    (:326-340)
  get_op ................. op table used while the graph is built / constant-folded       (:342-384)
  add_var_update ......... one assignment of the generated vector field                  (:386-393)
  generate_func_head/tail, generate_func, create_index_str, finalize_idx_str, expr_to_str,
  register_vars, add_var_hist, get_hist_func, clear, CodeGen line helpers               (:180-222, 437-660)
  run .................... fixed-step integration + state recording                      (:596-611, 693-769)

Instead of printing Python source and exec'ing it, the assignment stream is recorded into a
``pyrates_b200.Program``; ``run`` lowers that program to CUDA C for sm_100a (instance model:
``engine.InstanceModel``; network model: ``netengine.NetworkModel``), JIT-builds it through the C ABI and
integrates on the GPU.  It does not import the reference, so it also works from recorded programs
(tests/golden) on machines without PyRates.

There is no CPU fallback: unsupported ops / solvers raise, a missing GPU raises.
"""
from __future__ import annotations

import os
import re
from typing import Callable, Dict, List, Optional, Tuple, Union

import warnings

import numpy as np

from . import runtime as rt
from .engine import ADAPTIVE_SOLVERS, FIXED_STEP_SOLVERS, InstanceModel, n_steps_for
from .program import Arg, Program, Stmt
from .scalarize import UnsupportedModelError

__all__ = ["CudaBackend", "CudaVectorField"]


def _backend_error(msg: str) -> Exception:
    """PyRatesException when the reference is importable (same error type users see from other backends)."""
    try:
        from pyrates.backend import PyRatesException      # pyrates/backend/__init__.py:42
        return PyRatesException(msg)
    except Exception:
        return rt.CudaBackendError(msg)


# ----------------------------------------------------------------------------- build-time op table
# Used only while the reference's parser builds / constant-folds the compute graph on the host
# (ComputeGraph.eval_node, computegraph.py:297-312) — never during integration.  Same names and call
# strings as pyrates/backend/base/base_funcs.py:135-177 so the generated text is identical.
def _sig(x):
    return 1. / (1. + np.exp(-x))


_OPS: Dict[str, Tuple[str, Callable]] = {
    'maxi': ('maximum', np.maximum), 'mini': ('minimum', np.minimum), 'round': ('round', np.round),
    'vsum': ('sum', np.sum), 'mean': ('mean', np.mean), 'matmul': ('dot', np.dot), 'matvec': ('dot', np.dot),
    'roll': ('roll', np.roll), 'tanh': ('tanh', np.tanh), 'sinh': ('sinh', np.sinh), 'cosh': ('cosh', np.cosh),
    'arctan': ('arctan', np.arctan), 'arcsin': ('arcsin', np.arcsin), 'arccos': ('arccos', np.arccos),
    'sin': ('sin', np.sin), 'cos': ('cos', np.cos), 'tan': ('tan', np.tan), 'exp': ('exp', np.exp),
    'sigmoid': ('sigmoid', _sig), 'broadcast_pre': ('broadcast_pre', lambda x: x[None, :]),
    'broadcast_post': ('broadcast_post', lambda x: x[:, None]),
    'wsum': ('wsum', lambda w, c: np.einsum('ij,ij->i', w, c)),
    'reshape2d': ('reshape2d', lambda x, n, m: x.reshape(n, m)), 'flatten1d': ('flatten1d', lambda x: x.reshape(-1)),
    'no_op': ('identity', lambda x: x), 'index': ('index_1d', lambda x, i: x[i]),
    'index_range': ('index_range', lambda x, a, b: x[a:b]), 'index_2d': ('index_2d', lambda x, i, j: x[i, j]),
    'index_axis': ('index_axis', lambda x, idx=None, axis=0: (x[idx] if axis == 0 else x[:, idx]) if idx is not None else x[:]),
    'absv': ('abs', np.abs), 'sign': ('sign', np.sign), 'log': ('log', np.log),
    'interp': ('interp', np.interp),
    'interp_rows': ('interp_rows', lambda t, time, inp: np.array([np.interp(t, time, inp[:, k]) for k in range(inp.shape[1])])),
    'past': ('past', lambda x, tau: x), 'concatenate': ('concatenate', np.concatenate),
    'real': ('real', np.real), 'imag': ('imag', np.imag), 'conj': ('conjugate', np.conjugate),
    'randn': ('randn', np.random.randn),
}


_YHIST = re.compile(r"^(_yhist_\w+) = hist\((.*)\)$")


class CudaHistory:
    """Placeholder for ``DDEHistory`` (base_backend.py:88-177) in the argument tuple of a delay-differential vector field:
    carries the initial condition only.  It cannot be called on the host — there is no CPU execution path."""

    def __init__(self, y0, t0: float = 0.0):
        self.y0 = np.array(y0, dtype=np.float64, copy=True)
        self.t0 = float(t0)

    def __call__(self, t):
        raise _backend_error("the state history of a CUDA vector field lives on the device; it cannot be queried on the host")


class CudaVectorField:
    """Handle returned by ``generate_func``: the recorded vector field.  Calling it evaluates the rhs ONCE
    on the GPU (``func(t, y, *args) -> dy``), matching what ``get_run_func`` hands to users
    (pyrates/frontend/template/circuit.py:552-650)."""

    def __init__(self, backend: "CudaBackend", func_name: str, arg_names: List[str], stmts: List[Stmt],
                 state_var: str, return_var: str, jacobian: Optional[Tuple[str, tuple]] = None):
        self.backend = backend
        self.__name__ = func_name
        self.arg_names = list(arg_names)
        self.stmts = list(stmts)
        self.state_var, self.return_var = state_var, return_var
        # get_jacobian_func (computegraph.py:516-763): `J0 = zeros((n, n)); J0[i, j] = expr; return J0` — the signature has
        # no output buffer, the (name, shape) of the locally allocated matrix is kept here instead
        self.jacobian = jacobian

    def program(self, func_args) -> Program:
        if self.jacobian is not None:
            return self._jacobian_program(func_args)
        if len(func_args) != len(self.arg_names):
            raise ValueError(f"vector field takes {len(self.arg_names)} arguments, got {len(func_args)}")
        assigned = {s.lhs for s in self.stmts}
        args = []
        for i, (name, val) in enumerate(zip(self.arg_names, func_args)):
            if name == "hist" and i == 2:
                # the history object only carries the initial condition (DDEHistory.__init__, base_backend.py:130-146);
                # the history itself lives on the device (csrc/templates/inst_kernel.cuh)
                y_init = getattr(val, "y0", getattr(val, "_y", [func_args[1]])[0])
                args.append(Arg(name, "hist", np.array(y_init, dtype=np.float64, copy=True)))
                continue
            if callable(val):
                raise UnsupportedModelError(f"callable argument `{name}` of the vector field")
            kind = "t" if i == 0 else ("y" if name == self.state_var else ("dy" if name == self.return_var else
                                                                           ("var" if name in assigned else "const")))
            args.append(Arg(name, kind, np.array(val, copy=True)))
        return Program(args, self.stmts, {}, self.backend._float_precision, self.__name__)

    def _jacobian_program(self, func_args) -> Program:
        name, shape, _ = self.jacobian
        if len(func_args) != len(self.arg_names):
            raise ValueError(f"Jacobian function takes {len(self.arg_names)} arguments, got {len(func_args)}")
        prec = self.backend._float_precision
        if "complex" in str(prec):
            raise UnsupportedModelError("Jacobian functions of complex-valued builds")
        args = []
        for i, (n, val) in enumerate(zip(self.arg_names, func_args)):
            if n == "hist" and i == 2:
                y_init = getattr(val, "y0", getattr(val, "_y", [func_args[1]])[0])
                args.append(Arg(n, "hist", np.array(y_init, dtype=np.float64, copy=True)))
                continue
            if callable(val):
                raise UnsupportedModelError(f"callable argument `{n}` of the Jacobian function")
            if i == 0:
                args.append(Arg(n, "t", np.asarray(float(np.asarray(val)))))
            elif n == self.state_var:
                args.append(Arg(n, "y", np.array(val, dtype=prec).reshape(-1)))      # the caller's float64 state, model precision
            else:
                args.append(Arg(n, "const", np.array(val, copy=True)))
        args.insert(3 if len(args) > 2 and args[2].kind == "hist" else 2, Arg(name, "dy", np.zeros(shape, dtype=prec)))
        return Program(args, self.stmts, {}, prec, self.__name__)

    def source(self) -> str:
        if self.jacobian is not None:
            name, shape, _ = self.jacobian
            lines = [f"def {self.__name__}({','.join(self.arg_names)}):", f"\t{name} = zeros({tuple(shape)})"]
            lines += ["\t" + st.text() for st in self.stmts] + [f"\treturn {name}"]
            return "\n".join(lines) + "\n"
        lines = [f"def {self.__name__}({','.join(self.arg_names)}):"] + ["\t" + st.text() for st in self.stmts]
        return "\n".join(lines + [f"\treturn {self.return_var}"]) + "\n"

    def __call__(self, t, y, *args):
        from .engine import evaluate_rhs
        if self.jacobian is not None:
            # J(t, y, *params) -> ndarray (n, n); a batch of states y[n, B] gives (n, n, B) from one launch
            y = np.asarray(y)
            prog = self.program([t, y if y.ndim == 1 else y[:, 0], *args])
            shape, n_hist = tuple(self.jacobian[1]), self.jacobian[2]
            out = evaluate_rhs(prog, t=float(np.asarray(t)), y=None if y.ndim == 1 else y)
            out = out.reshape(shape if y.ndim == 1 else shape + (y.shape[1],))
            # delay-differential models: (J0, [J_tau1, ...]), computegraph.py:521-524
            return out if n_hist is None else (out[0], [out[k] for k in range(1, n_hist + 1)])
        names = self.arg_names
        dy = np.zeros_like(np.asarray(y))
        full = [np.asarray(t), np.asarray(y)] + list(args)
        if len(full) == len(names) - 1:              # dy omitted by the caller
            full.insert(2, dy)
        prog = self.program(full)
        from .engine import evaluate_rhs
        return evaluate_rhs(prog)


class CudaBackend:
    """Code generator + runner with the reference backend interface, targeting B200 CUDA kernels."""

    SUPPORTED_SOLVERS: Tuple[str, ...] = FIXED_STEP_SOLVERS + ('scipy',)   # 'scipy' = solve_ivp RK45 semantics on device
    SUPPORTS_SPARSE_JACOBIAN: bool = False
    SUPPORTS_EDGE_DELAY_BUFFER: bool = True
    _no_funcs: Tuple[str, ...] = ("identity", "index_1d", "index_2d", "index_range", "index_axis")

    def __init__(self, ops: Optional[Dict[str, str]] = None, imports: Optional[List[str]] = None, **kwargs) -> None:
        self.code: List[str] = []
        self.lvl = 0
        self._stmts: List[Stmt] = []
        self._n_hist_stmts = 0
        self._local_arrays: Dict[str, tuple] = {}
        self._yhist_stmts: List[Stmt] = []
        self._return_expr: Optional[str] = None
        self._imports: List[str] = []
        self._helper_funcs: List[str] = []
        self.add_hist_arg = kwargs.pop('add_hist_arg', True)                  # reference default (:295)
        self.lags = {}
        self.idx_dummy_var = "temporary_pyrates_var_index"
        self._float_precision = kwargs.pop('float_precision', 'float32')      # reference default (:300)
        self._int_precision = kwargs.pop('int_precision', 'int32')
        if self._float_precision not in ('float32', 'float64', 'complex64', 'complex128'):
            raise _backend_error(f"CudaBackend supports float_precision 'float32', 'float64', 'complex64' or 'complex128', "
                                 f"got `{self._float_precision}`")
        self._idx_left, self._idx_right = '[', ']'
        self._start_idx = 0
        kwargs.pop('idx_left', None), kwargs.pop('idx_right', None), kwargs.pop('start_idx', None)
        # like the reference (base_backend.py:316-324) the generated function is written to `<file_name>.py` for inspection
        # — here in the NumPy-syntax form the compute graph printed, i.e. the text the CUDA kernels are lowered from
        fname = kwargs.pop('file_name', 'pyrates_run')
        self.fdir = os.path.dirname(fname)
        self._fname = os.path.basename(fname).split('.')[0]
        self._fend = ".py"
        self._written: Optional[str] = None
        # CUDA-specific options, popped so they never leak into solver kwargs (SURVEY §5 "config / flags")
        self.device = kwargs.pop('device', None)
        self.fmad = kwargs.pop('fmad', True)
        self.const_div_to_mul = kwargs.pop('const_div_to_mul', False)
        self.fast_math = kwargs.pop('fast_math', False)
        self.exec_model = kwargs.pop('exec_model', 'auto')        # 'auto' | 'instance' | 'network'
        self.last_model = None
        self._head: Optional[Tuple[str, str, str, List[str]]] = None

    # ------------------------------------------------------------------ CodeGen line helpers (:180-222)
    def generate(self) -> str:
        return '\n'.join(self.code)

    def add_code_line(self, code_str: str):
        for code in code_str.split('\n'):
            self.code.append("\t" * self.lvl + code)
        # get_jacobian_func of a delay-differential model prints `_yhist_<d> = hist(t - <d>)` through this generic hook
        # (computegraph.py:697-701): kept as a lookup of the whole delayed state
        m = _YHIST.match(code_str.strip())
        if m:
            self._yhist_stmts.append(Stmt(m.group(1), f"hist({m.group(2)})[:]", None))

    def add_linebreak(self):
        self.code.append("")

    def add_indent(self):
        self.lvl += 1

    def remove_indent(self):
        if self.lvl == 0:
            raise SyntaxError("Error in generation of network function file: A net negative indentation was requested.")
        self.lvl -= 1

    # ------------------------------------------------------------------ variables and ops
    def get_var(self, v):
        """base_backend.py:326-340: value in backend precision; (1,) constants squeezed to 0-d."""
        if v.is_float or v.is_complex:
            dtype = self._float_precision
            if 'complex' in str(dtype) and v.name in ['t', 'time']:
                dtype = f'float{str(dtype)[7:]}'          # base_backend.py:329-330: time stays real in complex builds
        else:
            dtype = self._int_precision
        with warnings.catch_warnings():
            # like the reference, a complex template variable built with a real float_precision keeps its real part
            warnings.simplefilter("ignore")
            result = np.asarray(v.value, dtype=dtype)
        if result.shape == (1,) and v.vtype == 'constant':
            result = result.squeeze()
        return result

    def get_op(self, name: str, **kwargs) -> dict:
        call, func = _OPS[name]              # KeyError for unknown ops, like the reference's table lookup
        return {'func': func, 'call': call}

    @staticmethod
    def register_vars(variables: list):
        pass

    # ------------------------------------------------------------------ index strings (:453-475, 643-660)
    def create_index_str(self, idx, separator: str = ',', apply: bool = True, **kwargs) -> Tuple[str, dict]:
        if type(idx) is str and separator in idx:
            idx = tuple(idx.split(separator))
        if type(idx) is tuple:
            parts = tuple(f"{self._process_idx(i)}" for i in idx)
            body = separator.join(parts)
            return (f"[{body}]" if apply else body), dict()
        body = self._process_idx(idx)
        return (f"[{body}]" if apply else body), dict()

    def _process_idx(self, idx) -> str:
        if hasattr(idx, "vtype") and hasattr(idx, "name"):        # ComputeVar holding an index array
            return idx.name
        if type(idx) is tuple:
            return f"{idx[0]}:{idx[1]}"
        if type(idx) is int:
            return f"{idx}"
        try:
            return f"{int(idx)}"
        except (TypeError, ValueError):
            return idx

    @staticmethod
    def finalize_idx_str(var, idx: str):
        return f"{var.name}{idx}"

    @staticmethod
    def expr_to_str(expr: str, args: tuple):
        return expr

    # ------------------------------------------------------------------ assignment stream
    def add_var_update(self, lhs, rhs: str, lhs_idx: Optional[str] = None, rhs_shape: Optional[tuple] = ()):
        idx_str = None
        if lhs_idx:
            idx_str, _ = self.create_index_str(lhs_idx, apply=False)
        self._stmts.append(Stmt(lhs.name, rhs, idx_str))
        self.add_code_line(f"{lhs.name}[{idx_str}] = {rhs}" if idx_str is not None else f"{lhs.name} = {rhs}")

    def add_var_hist(self, lhs: str, delay, state_idx, dt: Optional[float] = None, dt_adapt: bool = True, **kwargs):
        """base_backend.py:437-444: ``lhs = hist(t*dt - d)[idx]`` for fixed-step solvers (t = integer step counter),
        ``hist(t - d)[idx]`` for adaptive ones.  Recorded in the reference's text form; lowered by scalarize._hist_read."""
        idx = self._process_idx(state_idx)
        is_var = hasattr(delay, "vtype") and hasattr(delay, "name")                       # ComputeVar (:675-676)
        d = f"{delay.name}[{self._start_idx}]" if is_var and delay.shape else (delay.name if is_var else f"{delay}")
        rhs = f"hist(t*{dt:.10e}-{d})[{idx}]" if (dt is not None and not dt_adapt) else f"hist(t-{d})[{idx}]"
        # ComputeGraph.to_func (computegraph.py:424-450) prints the body first, then the head and the history lookups, and
        # re-inserts the body after them: the lookups precede every recorded body statement
        self._stmts.insert(self._n_hist_stmts, Stmt(lhs, rhs, None))
        self._n_hist_stmts += 1
        self.add_code_line(f"{lhs} = {rhs}")

    def add_import(self, line: str):
        if line not in self._imports:
            self._imports.append(line)

    # ------------------------------------------------------------------ Jacobian assembly (base_backend.py:417-435)
    def declare_local_array_imports(self) -> None:
        pass

    def emit_local_array_alloc(self, name: str, shape: tuple, dtype: str) -> None:
        """``name = zeros(shape)``: the matrix a Jacobian function fills and returns (computegraph.py:704-729)."""
        if not self._local_arrays:
            # the body of a Jacobian function starts with its first allocation; whole-state history lookups
            # (`_yhist_d = hist(t - d)`, recorded by add_code_line) precede it
            self._stmts, self._n_hist_stmts = list(self._yhist_stmts), 0
        self._local_arrays[name] = tuple(int(x) for x in shape)
        self.add_code_line(f"{name} = zeros({tuple(shape)}, dtype='{dtype}')")

    def emit_local_array_assign(self, name: str, indices: tuple, expr: str) -> None:
        idx = ', '.join(str(i) for i in indices)
        self._stmts.append(Stmt(name, str(expr), idx))
        self.add_code_line(f"{name}[{idx}] = {expr}")

    def generate_func_head(self, func_name: str, state_var: str = 'y', return_var: str = 'dy', func_args: list = None,
                           add_hist_func: Optional[bool] = None):
        """base_backend.py:483-528: signature order = t, y, [hist], then unique args in first-seen order."""
        names = [a.name for a in func_args] if func_args else []
        seen, uniq = set(), []
        for n in names:
            if n not in seen:
                seen.add(n)
                uniq.append(n)
        args = ['t', state_var] + (['hist'] if add_hist_func else []) + uniq
        self._head = (func_name, state_var, return_var, args)
        self.add_linebreak()
        self.add_code_line(f"def {func_name}({','.join(args)}):")
        self.add_indent()
        return args

    def generate_func_tail(self, rhs_var: str = 'dy'):
        self._return_expr = rhs_var
        self.add_code_line(f"return {rhs_var}")
        self.remove_indent()

    def generate_func(self, func_name: str, to_file: bool = True, **kwargs):
        if self._head is None:
            raise _backend_error("generate_func called before generate_func_head")
        _, state_var, return_var, args = self._head
        jac, stmts = None, self._stmts
        if self._local_arrays:
            # `return J0` or, for delay-differential models, `return J0, [J_hist_<d>, ...]` (computegraph.py:731-743)
            names = re.findall(r"[A-Za-z_]\w*", self._return_expr or return_var)
            if not names or any(n not in self._local_arrays for n in names) or "csr_matrix" in names:
                raise _backend_error(f"generated function returns `{self._return_expr}`, expected its local arrays")
            shapes = {self._local_arrays[n] for n in names}
            if len(shapes) != 1:
                raise _backend_error("Jacobian blocks of different shapes")
            if len(names) > 1:
                # several matrices: one stacked output buffer [K + 1, n, n]
                stmts = [Stmt("_jac_blocks", st.rhs, f"{names.index(st.lhs)}, {st.lhs_idx}") if st.lhs in names else st
                         for st in stmts]
                jac = ("_jac_blocks", (len(names),) + next(iter(shapes)), len(names) - 1)
            else:
                jac = (names[0], self._local_arrays[names[0]], None)
        func = CudaVectorField(self, func_name, args, stmts, state_var, return_var, jacobian=jac)
        if to_file:
            try:
                if self.fdir:
                    os.makedirs(self.fdir, exist_ok=True)
                path = os.path.join(self.fdir, f"{self._fname}{self._fend}")
                with open(path, "w") as f:
                    f.write("# vector field recorded by pyrates_b200.CudaBackend (NumPy-syntax form; lowered to CUDA C at run time)\n")
                    f.write(self.generate() + "\n")
                self._written = path
            except OSError:
                self._written = None
        decorator = kwargs.pop('decorator', None)
        if decorator:
            raise _backend_error("CudaBackend does not support python decorators on the vector field")
        return func

    @staticmethod
    def get_hist_func(y: np.ndarray, t0: float = 0.0):
        """base_backend.py:648-649.  The returned object stands for the state history in the argument tuple; the history
        itself is kept on the device."""
        return CudaHistory(y, t0)

    @staticmethod
    def to_file(fn: str, **kwargs):
        np.savez(fn, **kwargs)

    def clear(self):
        if getattr(self, "_written", None):             # base_backend.py:613-627: generated files go with the model
            try:
                os.remove(self._written)
            except OSError:
                pass
            self._written = None
        self.code.clear()
        self._stmts = []
        self._n_hist_stmts = 0
        self._local_arrays = {}
        self._yhist_stmts = []
        self._head = None

    # ------------------------------------------------------------------ integration
    def _validate_solver(self, solver: str) -> None:
        if solver not in self.SUPPORTED_SOLVERS:
            raise _backend_error(f"Backend `{type(self).__name__}` does not support solver `{solver}`. "
                                 f"Supported solvers: {list(self.SUPPORTED_SOLVERS)}.")

    def run(self, func: CudaVectorField, func_args: tuple, T: float, dt: float, dts: float, solver: str,
            **kwargs) -> tuple:
        """Counterpart of BaseBackend.run (:596-611): returns ``(state_rec[round(T/dts), n_state], times)``
        with row k = state at step k*store_step sampled before the update."""
        self._validate_solver(solver)
        if not isinstance(func, CudaVectorField):
            raise _backend_error("CudaBackend.run needs the vector-field handle returned by its own generate_func")
        if self.device is not None:
            rt.set_device(int(self.device))
        prog = func.program(func_args)
        step = dts if dts else dt
        times = np.linspace(0.0, T, num=round(T / step), endpoint=False)
        if solver in ADAPTIVE_SOLVERS:
            # BaseBackend._solve_scipy (:802-812): solve_ivp(first_step=dt, t_eval=times, **kwargs); one thread integrates the
            # circuit with SciPy's RK45 step controller and dense output (csrc/templates/inst_kernel.cuh)
            try:
                if self.exec_model == "network":
                    raise UnsupportedModelError("model too large (network execution model requested)")
                model = InstanceModel(prog, solver=solver, fmad=self.fmad, max_nodes=4000 if self.exec_model == "auto" else 20000,
                                      rk_method=kwargs.get("method"))
            except UnsupportedModelError as e:
                if self.exec_model == "instance" or "too large" not in str(e):
                    raise
                # populations / dense coupling: the whole grid steps with one adaptive step size (net_kernel.cuh)
                from .netengine import NetworkModel
                model = NetworkModel(prog, solver=solver, fmad=self.fmad)
            self.last_model = model
            # like solve_ivp, ignore options that have no effect for RK45 (the reference forwards e.g. float_precision)
            opts = {k: v for k, v in kwargs.items() if k in ("rtol", "atol", "max_step", "method")}
            out, y_final = model.run(T, dt, dts, **opts)
        else:
            model = build_model(prog, solver=solver, exec_model=self.exec_model, fmad=self.fmad,
                                const_div_to_mul=self.const_div_to_mul, fast_math=self.fast_math)
            self.last_model = model
            out, y_final = model.run(T, dt, dts)
        rec = np.ascontiguousarray(out[:, :, 0])
        yf = y_final[:, 0]
        if prog.is_complex:
            # complex state k lives in real state variables (2k, 2k+1): hand complex arrays back like the reference
            cdt = np.complex64 if prog.float_precision == "complex64" else np.complex128
            rec = (rec[:, 0::2] + 1j * rec[:, 1::2]).astype(cdt)
            yf = (yf[0::2] + 1j * yf[1::2]).astype(cdt)
        # mirror the reference's in-place semantics: y0 (func_args[1]) ends up holding the final state
        try:
            np.asarray(func_args[1])[...] = yf
        except (ValueError, TypeError):
            pass
        return rec, times


def build_model(prog: Program, *, solver: str, exec_model: str = "auto", **kw):
    """Pick the execution model: 'instance' (whole state in one thread's registers — small circuits and
    parameter sweeps) or 'network' (persistent cooperative grid, one thread per node — large populations)."""
    if exec_model not in ("auto", "instance", "network"):
        raise ValueError("exec_model must be 'auto', 'instance' or 'network'")
    if exec_model in ("auto", "instance"):
        try:
            limit = 4000 if exec_model == "auto" else 20000
            return InstanceModel(prog, solver=solver, max_nodes=limit, **kw)
        except UnsupportedModelError as e:
            if exec_model == "instance" or "too large" not in str(e):
                raise
    from .netengine import NetworkModel
    kw.pop("const_div_to_mul", None)
    kw.pop("fast_math", None)
    if prog.is_complex:
        raise UnsupportedModelError("complex-valued models are supported by the instance execution model only")
    return NetworkModel(prog, solver=solver, **kw)
