"""Scalarising front end: Program (NumPy-syntax assignment stream) -> scalar expression DAG.

Used for the *instance* execution model (one CUDA thread owns the whole state vector of one
sweep instance): every array element of the generated vector field becomes one node of a
hash-consed scalar DAG, which ``emit_inst`` prints as straight-line CUDA C held in registers.

The statements are interpreted with the reference's semantics — sequential execution, NumPy
broadcasting / fancy indexing, in-place mutable buffers (pyrates/backend/base/base_backend.py:483-535
generated-function shape; op vocabulary pyrates/backend/base/base_funcs.py:135-177) — by running
them over ``numpy`` *object* arrays whose elements are symbolic ``Sym`` nodes, so NumPy itself
supplies the indexing/broadcast rules and the result is exact for any model small enough to unroll.

Special objects
* state vector ``y`` / vector field ``dy``  -> 'state' leaves / per-element outputs
* parameters: baked as constants, or 'param' leaves when they vary per sweep instance
* ``x_input[t]`` timed inputs (pyrates/frontend/template/circuit.py:1711-1719) -> 'table' reads
* delay ring buffers ``buf[:] = roll(buf, 1, axis); buf[:,0] = src; buf[:,d]``
  (pyrates/ir/circuit.py:400-425, 598-641) -> circular addressing, advanced once per rhs
  *evaluation* exactly like the reference (SURVEY.md headline 5)
* other mutable buffers (``u[target_idx] = ...``) -> persistent 'slot' leaves, stored back only
  if some evaluation can observe the previous evaluation's value.
"""
from __future__ import annotations

import ast
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .program import Program

__all__ = ["Graph", "Sym", "ScalarRhs", "scalarize", "UnsupportedModelError"]


class UnsupportedModelError(NotImplementedError):
    """Raised (loudly, never a CPU fallback) when a model uses an op the CUDA backend lacks."""


# ----------------------------------------------------------------------------- DAG
UNARY_FUNCS = ("exp", "log", "sin", "cos", "tan", "tanh", "sinh", "cosh", "arctan", "arcsin", "arccos",
               "sqrt", "abs", "sign", "round")
BINARY_FUNCS = ("maximum", "minimum", "pow")


class Graph:
    """Hash-consed scalar expression DAG.  Node = (op, args, payload)."""

    def __init__(self, real_dtype=np.float64):
        self.real = np.dtype(real_dtype).type
        self.ops: List[str] = []
        self.args: List[tuple] = []
        self.payload: List[object] = []
        self.isint: List[bool] = []
        self._memo: Dict[tuple, int] = {}

    def _node(self, op: str, args: tuple = (), payload=None, isint=False) -> "Sym":
        key = (op, args, payload)
        i = self._memo.get(key)
        if i is None:
            i = len(self.ops)
            self.ops.append(op)
            self.args.append(args)
            self.payload.append(payload)
            self.isint.append(isint)
            self._memo[key] = i
        return Sym(self, i)

    # leaves ---------------------------------------------------------------
    def const(self, value, weak=False, isint=False) -> "Sym":
        if isint:
            return self._node("const", (), ("i", int(value)), True)
        v = float(value)
        if not weak:
            v = float(self.real(v))
        # key on the bit pattern so that -0.0 / nan stay distinct and hashable
        return self._node("const", (), ("w" if weak else "s", v.hex()), False)

    def leaf(self, op: str, payload, isint=False) -> "Sym":
        return self._node(op, (), payload, isint)

    # helpers --------------------------------------------------------------
    def is_const(self, i: int) -> bool:
        return self.ops[i] == "const"

    def const_value(self, i: int):
        kind, v = self.payload[i]
        return v if kind == "i" else float.fromhex(v)

    def const_weak(self, i: int) -> bool:
        return self.payload[i][0] == "w"

    def lift(self, x) -> "Sym":
        if isinstance(x, Sym):
            return x
        if isinstance(x, (bool, np.bool_)):
            return self.const(float(x), weak=True)
        if isinstance(x, (int, np.integer)):
            return self.const(int(x), isint=True)
        if isinstance(x, (float, np.floating)):
            return self.const(float(x), weak=isinstance(x, float) and not isinstance(x, np.floating))
        if isinstance(x, complex):
            raise UnsupportedModelError("complex-valued models are not supported by the CUDA backend")
        raise TypeError(f"cannot lift {type(x)} into the scalar DAG")

    # arithmetic with folding ---------------------------------------------------
    def _fold(self, op: str, a: int, b: Optional[int]) -> Optional["Sym"]:
        if not self.is_const(a) or (b is not None and not self.is_const(b)):
            return None
        va = self.const_value(a)
        vb = self.const_value(b) if b is not None else None
        ints = self.isint[a] and (b is None or self.isint[b])
        weak = (self.isint[a] or self.const_weak(a)) and (b is None or self.isint[b] or self.const_weak(b))
        try:
            if ints and op in ("add", "sub", "mul", "neg"):
                r = {"add": lambda: va + vb, "sub": lambda: va - vb, "mul": lambda: va * vb, "neg": lambda: -va}[op]()
                return self.const(r, isint=True)
            if weak:
                fa, fb = float(va), (float(vb) if vb is not None else None)
                cast = float
            else:
                cast = self.real
                fa, fb = cast(va), (cast(vb) if vb is not None else None)
            with np.errstate(all="ignore"):
                if op == "add":
                    r = fa + fb
                elif op == "sub":
                    r = fa - fb
                elif op == "mul":
                    r = fa * fb
                elif op == "div":
                    r = np.divide(np.float64(fa), np.float64(fb)) if weak else np.divide(fa, fb)
                elif op == "neg":
                    r = -fa
                elif op == "pow":
                    r = np.power(np.float64(fa), np.float64(fb)) if weak else np.power(fa, fb)
                elif op in UNARY_FUNCS:
                    fn = {"abs": np.abs, "round": np.round}.get(op) or getattr(np, op)
                    r = fn(np.float64(fa)) if weak else fn(fa)
                elif op == "maximum":
                    r = np.maximum(fa, fb)
                elif op == "minimum":
                    r = np.minimum(fa, fb)
                else:
                    return None
            return self.const(float(r), weak=weak)
        except (OverflowError, ZeroDivisionError, ValueError):
            return None

    def _is_val(self, i: int, v: float) -> bool:
        return self.is_const(i) and float(self.const_value(i)) == v

    def binop(self, op: str, a: "Sym", b: "Sym") -> "Sym":
        a, b = self.lift(a), self.lift(b)
        f = self._fold(op, a.id, b.id)
        if f is not None:
            return f
        ia, ib = a.id, b.id
        if op == "add":
            if self._is_val(ia, 0.0):
                return self._as_real(b)
            if self._is_val(ib, 0.0):
                return self._as_real(a)
        elif op == "sub":
            if self._is_val(ib, 0.0):
                return self._as_real(a)
            if self._is_val(ia, 0.0):
                return self.unop("neg", b)
        elif op == "mul":
            if self._is_val(ia, 1.0):
                return self._as_real(b)
            if self._is_val(ib, 1.0):
                return self._as_real(a)
            if self._is_val(ia, 0.0) or self._is_val(ib, 0.0):
                if not (self.isint[ia] and self.isint[ib]):
                    return self.const(0.0)
            if self._is_val(ia, -1.0):
                return self.unop("neg", b)
            if self._is_val(ib, -1.0):
                return self.unop("neg", a)
        elif op == "div":
            if self._is_val(ib, 1.0):
                return self._as_real(a)
            if self._is_val(ia, 0.0):
                return self.const(0.0)
        elif op == "pow":
            if self.is_const(ib):
                e = float(self.const_value(ib))
                # mirrors numpy's scalar-exponent fast paths (square / sqrt / reciprocal / positive)
                if e == 2.0:
                    return self.binop("mul", a, a)
                if e == 1.0:
                    return self._as_real(a)
                if e == 0.5:
                    return self.unop("sqrt", a)
                if e == -1.0:
                    return self.binop("div", self.const(1.0, weak=True), a)
                if e == 0.0:
                    return self.const(1.0)
                if e == 3.0:
                    return self.binop("mul", self.binop("mul", a, a), a)
                if e == 4.0:
                    sq = self.binop("mul", a, a)
                    return self.binop("mul", sq, sq)
        isint = self.isint[ia] and self.isint[ib] and op in ("add", "sub", "mul")
        if op in ("add", "mul", "maximum", "minimum") and ia > ib:
            ia, ib = ib, ia         # canonical operand order for commutative ops (exact in IEEE)
        return self._node(op, (ia, ib), None, isint)

    def _as_real(self, a: "Sym") -> "Sym":
        return a

    def unop(self, op: str, a: "Sym") -> "Sym":
        a = self.lift(a)
        f = self._fold(op, a.id, None)
        if f is not None:
            return f
        if op == "neg" and self.ops[a.id] == "neg":
            return Sym(self, self.args[a.id][0])
        return self._node(op, (a.id,), None, self.isint[a.id] and op == "neg")


class Sym:
    """Handle to a DAG node; overloads arithmetic so NumPy object arrays do the broadcasting."""
    __slots__ = ("g", "id")

    def __init__(self, g: Graph, i: int):
        self.g = g
        self.id = i

    def __hash__(self):
        return hash(self.id)

    def __eq__(self, other):      # identity of DAG node (NOT a symbolic comparison)
        return isinstance(other, Sym) and other.id == self.id

    def _bin(self, op, a, b):
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            return NotImplemented            # let the ndarray operand broadcast element-wise
        if isinstance(a, (CSym, complex)) or isinstance(b, (CSym, complex)):
            return CSym.binop(self.g, op, a, b)
        return self.g.binop(op, a, b)

    def __add__(self, o): return self._bin("add", self, o)
    def __radd__(self, o): return self._bin("add", o, self)
    def __sub__(self, o): return self._bin("sub", self, o)
    def __rsub__(self, o): return self._bin("sub", o, self)
    def __mul__(self, o): return self._bin("mul", self, o)
    def __rmul__(self, o): return self._bin("mul", o, self)
    def __truediv__(self, o): return self._bin("div", self, o)
    def __rtruediv__(self, o): return self._bin("div", o, self)
    def __pow__(self, o): return self._bin("pow", self, o)
    def __rpow__(self, o): return self._bin("pow", o, self)
    def __neg__(self): return self.g.unop("neg", self)
    def __pos__(self): return self

    def __repr__(self):
        g = self.g
        return f"Sym#{self.id}<{g.ops[self.id]}>"


class CSym:
    """A complex value as a pair of real DAG nodes (float_precision='complex64' / 'complex128' builds: every float
    argument arrives complex, pyrates/backend/base/base_backend.py:326-331).  Arithmetic is lowered to real operations
    — (a+bi)(c+di) = (ac-bd) + (ad+bc)i, division through the conjugate — so constants with a zero imaginary part fold
    away in the graph and the kernels stay real; the complex state occupies two consecutive real state variables."""
    __slots__ = ("g", "re", "im")

    def __init__(self, g: Graph, re, im):
        self.g, self.re, self.im = g, g.lift(re), g.lift(im)

    @staticmethod
    def of(g: Graph, x) -> "CSym":
        if isinstance(x, CSym):
            return x
        if isinstance(x, complex):
            return CSym(g, g.const(float(x.real), weak=True), g.const(float(x.imag), weak=True))
        if isinstance(x, (bool, np.bool_)):
            raise TypeError("bool in complex arithmetic")
        return CSym(g, x, g.const(0.0, weak=True))

    @staticmethod
    def binop(g: Graph, op: str, a, b):
        a, b = CSym.of(g, a), CSym.of(g, b)
        if op == "add":
            return CSym(g, a.re + b.re, a.im + b.im)
        if op == "sub":
            return CSym(g, a.re - b.re, a.im - b.im)
        if op == "mul":
            return CSym(g, a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re)
        if op == "div":
            den = b.re * b.re + b.im * b.im
            return CSym(g, (a.re * b.re + a.im * b.im) / den, (a.im * b.re - a.re * b.im) / den)
        if op == "pow":
            if g.is_const(b.im.id) and float(g.const_value(b.im.id)) == 0.0 and g.is_const(b.re.id):
                e = float(g.const_value(b.re.id))
                if e == int(e) and 0 <= int(e) <= 8:
                    out = CSym.of(g, 1.0)
                    for _ in range(int(e)):
                        out = CSym.binop(g, "mul", out, a)
                    return out
            raise UnsupportedModelError("complex powers other than small non-negative integers")
        raise UnsupportedModelError(f"`{op}` on complex values")

    def _bin(self, op, a, b):
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            return NotImplemented
        return CSym.binop(self.g, op, a, b)

    def __add__(self, o): return self._bin("add", self, o)
    def __radd__(self, o): return self._bin("add", o, self)
    def __sub__(self, o): return self._bin("sub", self, o)
    def __rsub__(self, o): return self._bin("sub", o, self)
    def __mul__(self, o): return self._bin("mul", self, o)
    def __rmul__(self, o): return self._bin("mul", o, self)
    def __truediv__(self, o): return self._bin("div", self, o)
    def __rtruediv__(self, o): return self._bin("div", o, self)
    def __pow__(self, o): return self._bin("pow", self, o)
    def __neg__(self): return CSym(self.g, -self.re, -self.im)
    def __pos__(self): return self

    def conj(self): return CSym(self.g, self.re, -self.im)

    def abs(self):
        return self.g.unop("sqrt", self.re * self.re + self.im * self.im)

    def exp(self):
        e = self.g.unop("exp", self.re)
        return CSym(self.g, e * self.g.unop("cos", self.im), e * self.g.unop("sin", self.im))

    def __repr__(self):
        return f"CSym({self.re!r}, {self.im!r})"


# ----------------------------------------------------------------------------- result container
@dataclass
class RingSpec:
    arg: str                 # argument name
    shape: Tuple[int, ...]   # logical shape (R, L) or (L,)
    rows: int
    length: int
    rolls_per_eval: int = 0
    init: Optional[np.ndarray] = None   # initial contents, (rows, length)


@dataclass
class TableSpec:
    arg: str
    shape: Tuple[int, ...]
    value: Optional[np.ndarray] = None


@dataclass
class ParamSpec:
    arg: str
    elem: int                # flat element index into the argument array
    default: float


@dataclass
class ScalarRhs:
    graph: Graph
    n_state: int
    dy: List[int]                              # node id per state element
    params: List[ParamSpec]
    tables: List[TableSpec]
    rings: List[RingSpec]
    ring_writes: List[Tuple[int, int, int, int]]   # (ring, row, logical_col, node) applied at eval end
    slots: List[Tuple[str, int, float]]            # persistent slots: (arg, elem, init)
    slot_writes: List[Tuple[int, int]]             # (slot, node)
    uses_t: bool = False
    # delay-differential models (`hist(t*dt - d)[idx]` lookups, base_backend.py:437-444): state elements whose history is
    # kept, every delay that is looked up, and the step size baked into the fixed-step query text (None: adaptive form)
    n_out: Optional[int] = None                    # outputs when they are not the state derivatives (Jacobian programs: n*n)
    hist_vars: List[int] = field(default_factory=list)
    hist_delays: List[float] = field(default_factory=list)
    hist_dt: Optional[float] = None
    hist_delay_params: List[int] = field(default_factory=list)      # swept delays: parameter rows that are looked up as delays

    def live_nodes(self) -> List[int]:
        """Topologically ordered ids of all nodes reachable from the outputs."""
        g = self.graph
        roots = list(self.dy) + [n for *_, n in self.ring_writes] + [n for _, n in self.slot_writes]
        seen = set()
        stack = list(roots)
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            stack.extend(g.args[i])
            if g.ops[i] == "table":
                stack.append(g.payload[i][1])
        return sorted(seen)          # ids are created in dependency order => sorted == topological


# ----------------------------------------------------------------------------- interpreter
class _Ring:
    def __init__(self, idx: int, spec: RingSpec):
        self.idx = idx
        self.spec = spec
        self.shift = 0
        self.writes: Dict[Tuple[int, int], Sym] = {}
        n = spec.rows * spec.length
        self.grid = np.arange(n).reshape(spec.shape)


class _Rolled:
    def __init__(self, ring: _Ring, shift: int, axis: int):
        self.ring, self.shift, self.axis = ring, shift, axis


class _Table:
    def __init__(self, idx: int, spec: TableSpec):
        self.idx, self.spec = idx, spec


def _real_table(name: str, v) -> np.ndarray:
    """Input tables of complex builds arrive complex with a zero imaginary part (base_backend.py:326-331 casts every float
    argument): the kernels read them as real tables."""
    v = np.asarray(v)
    if np.iscomplexobj(v):
        if np.any(v.imag != 0):
            raise UnsupportedModelError(f"input table `{name}` has a non-zero imaginary part")
        v = np.ascontiguousarray(v.real)
    return v


def _obj(x) -> np.ndarray:
    a = np.empty((), dtype=object)
    a[()] = x
    return a


def _wrap(r):
    return r if isinstance(r, np.ndarray) else _obj(r)


class _Interp:
    def __init__(self, prog: Program, sweep: Sequence[Tuple[str, int]], max_nodes: int):
        self.p = prog
        self.complex = "complex" in str(prog.float_precision)
        self.g = Graph(prog.real_dtype)
        self.max_nodes = max_nodes
        self.env: Dict[str, object] = {}
        self.params: List[ParamSpec] = []
        self.tables: List[TableSpec] = []
        self.rings: List[_Ring] = []
        self.slots: List[Tuple[str, int, float]] = []
        self.slot_arrays: Dict[str, np.ndarray] = {}
        self.uses_t = False
        self.hist_vars: List[int] = []
        self.hist_delays: List[float] = []
        self.hist_dt: Optional[float] = None
        self.hist_delay_params: List[int] = []
        self.hist_name = next((a.name for a in prog.args if a.kind == "hist"), None)
        self.jacobian = False
        self.tname = prog.arg_by_kind("t").name
        self.yname = prog.arg_by_kind("y").name
        self.dyname = prog.arg_by_kind("dy").name
        sweep = list(sweep)
        sweep_by_arg: Dict[str, Dict[int, int]] = {}
        for name, elem in sweep:
            sweep_by_arg.setdefault(name, {})

        trees = [(s, ast.parse(s.rhs, mode="eval").body,
                  ast.parse(f"_[{s.lhs_idx}]", mode="eval").body.slice if s.lhs_idx is not None else None)
                 for s in prog.stmts]
        self.trees = trees
        self.interp_args = set()
        self.slot_arrays_complex = set()
        t_indexed, rolled = self._prescan(trees)

        for a in prog.args:
            v = np.asarray(a.value)
            if a.kind == "t":
                self.env[a.name] = _obj(self.g.leaf("t", None, isint=bool(np.issubdtype(np.asarray(a.value).dtype, np.integer))))
            elif a.kind == "y":
                arr = np.empty(v.size, dtype=object)
                for k in range(v.size):
                    if self.complex:          # complex state k = real state variables (2k, 2k+1)
                        arr[k] = CSym(self.g, self.g.leaf("state", 2 * k), self.g.leaf("state", 2 * k + 1))
                    else:
                        arr[k] = self.g.leaf("state", k)
                self.env[a.name] = arr
            elif a.kind == "dy":
                # 2-d output buffer: a Jacobian program (`J0 = zeros((n, n)); J0[i, j] = expr`, computegraph.py:704-717),
                # entries that are never assigned stay zero
                self.jacobian = v.ndim >= 2
                arr = np.empty(v.shape if self.jacobian else v.size, dtype=object)
                arr[...] = self.g.const(0.0) if self.jacobian else None
                self.env[a.name] = arr
            elif a.kind == "hist":
                self.env[a.name] = None                          # only ever called: hist(t - d)[idx]
            elif a.name in self.interp_args:
                if a.kind != "const" or any(n == a.name for n, _ in sweep):
                    raise UnsupportedModelError(f"interpolation table `{a.name}` must be a constant that is not swept")
                self.env[a.name] = None                          # only ever used as an argument of interp()
            elif a.name in t_indexed:
                if a.kind != "const":
                    raise UnsupportedModelError(f"time-indexed buffer `{a.name}` must be constant")
                self.tables.append(TableSpec(a.name, v.shape, _real_table(a.name, v)))
                self.env[a.name] = _Table(len(self.tables) - 1, self.tables[-1])
            elif a.name in rolled:
                if v.ndim not in (1, 2):
                    raise UnsupportedModelError(f"ring buffer `{a.name}` with ndim={v.ndim}")
                rows = 1 if v.ndim == 1 else v.shape[0]
                spec = RingSpec(a.name, v.shape, rows, v.shape[-1], 0, v.reshape(rows, v.shape[-1]).copy())
                self.rings.append(_Ring(len(self.rings), spec))
                self.env[a.name] = self.rings[-1]
            elif a.is_int:
                if a.kind == "var":
                    raise UnsupportedModelError(f"mutable integer buffer `{a.name}`")
                self.env[a.name] = v.copy()           # concrete index array (never swept)
            elif np.iscomplexobj(v):
                cswept = {e for n, e in sweep if n == a.name}
                if cswept and (a.kind != "const" or np.any(v.reshape(-1)[sorted(cswept)].imag != 0)):
                    raise UnsupportedModelError(f"complex argument `{a.name}` can only be swept as a real-valued constant")
                if a.kind == "var":
                    # mutable complex buffers are the vectorised frontend's input routing (`u[target_idx] = ...`, rewritten
                    # in every evaluation before it is read): tracked symbolically, never stored back
                    self.slot_arrays_complex.add(a.name)
                arr = np.empty(v.shape, dtype=object)
                flat = arr.reshape(-1) if v.ndim else None
                vals = v.reshape(-1)
                for k in range(vals.size):
                    if k in cswept:
                        # a real parameter of a complex build (every float argument arrives complex, base_backend.py:326-331):
                        # the sweep varies its real part
                        self.params.append(ParamSpec(a.name, k, float(vals[k].real)))
                        c = CSym(self.g, self.g.leaf("param", len(self.params) - 1), self.g.const(0.0))
                    else:
                        c = CSym(self.g, self.g.const(vals[k].real), self.g.const(vals[k].imag))
                    if v.ndim:
                        flat[k] = c
                    else:
                        arr[()] = c
                self.env[a.name] = arr
            else:
                arr = np.empty(v.shape, dtype=object)
                flat = arr.reshape(-1) if v.ndim else None
                vals = v.reshape(-1)
                swept = {e for n, e in sweep if n == a.name}
                for k in range(vals.size):
                    if a.kind == "var":
                        self.slots.append((a.name, k, float(vals[k])))
                        s = self.g.leaf("slot", len(self.slots) - 1)
                    elif k in swept:
                        self.params.append(ParamSpec(a.name, k, float(vals[k])))
                        s = self.g.leaf("param", len(self.params) - 1)
                    else:
                        s = self.g.const(vals[k])
                    if v.ndim:
                        flat[k] = s
                    else:
                        arr[()] = s
                if a.kind == "var":
                    self.slot_arrays[a.name] = arr
                self.env[a.name] = arr
        for name, elem in sweep:
            if not any(ps.arg == name and ps.elem == elem for ps in self.params):
                raise KeyError(f"sweep target {name}[{elem}] is not a float constant of the program")

    # -- prescan for time-indexed tables and rolled rings
    def _prescan(self, trees):
        t_indexed, rolled = set(), set()
        for s, rhs, _ in trees:
            for node in ast.walk(rhs):
                if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name):
                    names = {n.id for n in ast.walk(node.slice) if isinstance(n, ast.Name)}
                    if self.tname in names:
                        t_indexed.add(node.value.id)
                if isinstance(node, ast.Call) and getattr(node.func, "id", None) == "roll":
                    if node.args and isinstance(node.args[0], ast.Name):
                        rolled.add(node.args[0].id)
                if isinstance(node, ast.Call) and getattr(node.func, "id", None) in ("interp", "interp_rows"):
                    for arg in node.args[1:]:
                        if isinstance(arg, ast.Name):
                            self.interp_args.add(arg.id)       # (time, value) tables: read from memory, never unrolled
        return t_indexed, rolled

    # -- expression evaluation
    def ev(self, n):
        g = self.g
        if len(g.ops) > self.max_nodes:
            raise UnsupportedModelError(f"model too large to unroll into a per-thread kernel (> {self.max_nodes} "
                                        f"scalar operations); use the network execution model")
        if isinstance(n, ast.Constant):
            if isinstance(n.value, complex):
                raise UnsupportedModelError("complex-valued models are not supported by the CUDA backend")
            if isinstance(n.value, bool) or not isinstance(n.value, (int, float)):
                raise UnsupportedModelError(f"literal {n.value!r}")
            if isinstance(n.value, int):
                return n.value
            return _obj(g.const(n.value, weak=True))
        if isinstance(n, ast.Name):
            if n.id == "pi":
                return _obj(g.const(math.pi, weak=True))
            if n.id not in self.env:
                raise UnsupportedModelError(f"unknown name `{n.id}` in generated code")
            if n.id == self.tname:
                self.uses_t = True
            return self.env[n.id]
        if isinstance(n, ast.UnaryOp):
            v = self.ev(n.operand)
            if isinstance(n.op, ast.USub):
                return _wrap(-self._arr(v))
            if isinstance(n.op, ast.UAdd):
                return v
            raise UnsupportedModelError(ast.unparse(n))
        if isinstance(n, ast.BinOp):
            a, b = self._arr(self.ev(n.left)), self._arr(self.ev(n.right))
            if isinstance(n.op, ast.Add):
                return _wrap(a + b)
            if isinstance(n.op, ast.Sub):
                return _wrap(a - b)
            if isinstance(n.op, ast.Mult):
                return _wrap(a * b)
            if isinstance(n.op, ast.Div):
                return _wrap(a / b)
            if isinstance(n.op, ast.Pow):
                return _wrap(a ** b)
            raise UnsupportedModelError(f"operator in `{ast.unparse(n)}`")
        if isinstance(n, ast.Call):
            return self.call(n)
        if isinstance(n, ast.Subscript):
            return self.subscript(n)
        raise UnsupportedModelError(f"expression `{ast.unparse(n)}`")

    def _arr(self, v):
        """Normalise a value for arithmetic: ints/ints arrays become Sym consts."""
        if isinstance(v, (_Ring, _Table, _Rolled)):
            raise UnsupportedModelError("ring buffers / input tables can only be indexed, rolled or assigned")
        if isinstance(v, (int, np.integer)):
            return _obj(self.g.const(int(v), isint=True))
        if isinstance(v, np.ndarray) and v.dtype != object:
            out = np.empty(v.shape, dtype=object)
            it = np.nditer(v, flags=["multi_index", "zerosize_ok"])
            isint = np.issubdtype(v.dtype, np.integer)
            for x in it:
                out[it.multi_index] = self.g.const(x.item(), isint=isint)
            return out
        return v

    def _map1(self, fname, v):
        g = self.g

        def one(x):
            if isinstance(x, CSym):
                if fname == "abs":
                    return x.abs()
                if fname == "exp":
                    return x.exp()
                raise UnsupportedModelError(f"`{fname}` of a complex value")
            return g.unop(fname, x)
        f = np.frompyfunc(one, 1, 1)
        r = f(self._arr(v))
        return r if isinstance(r, np.ndarray) else _obj(r)

    def _map2(self, fname, a, b):
        g = self.g
        f = np.frompyfunc(lambda x, y: g.binop(fname, x, y), 2, 1)
        r = f(self._arr(a), self._arr(b))
        return r if isinstance(r, np.ndarray) else _obj(r)

    def _sumseq(self, items):
        acc = None
        for it in items:
            acc = it if acc is None else acc + it
        return acc if acc is not None else self.g.const(0.0)

    def call(self, n: ast.Call):
        g = self.g
        fname = getattr(n.func, "id", None)
        if fname is None:
            raise UnsupportedModelError(ast.unparse(n))
        if n.keywords:
            raise UnsupportedModelError(f"keyword arguments in `{ast.unparse(n)}`")
        if fname == "roll":
            base = self.ev(n.args[0])
            if not isinstance(base, _Ring):
                raise UnsupportedModelError("`roll` is only supported as the delay ring-buffer idiom")
            shift = self._const_int(n.args[1])
            axis = self._const_int(n.args[2]) if len(n.args) > 2 else (0 if base.spec.rows == 1 and len(base.spec.shape) == 1 else None)
            nd = len(base.spec.shape)
            if axis is None or (axis % nd) != nd - 1:
                raise UnsupportedModelError("ring buffers must be rolled along their last axis")
            return _Rolled(base, shift, axis)
        if fname == "interp":
            # adaptive solvers: timed inputs are interpolated, `interp(t, time, u_input)` (frontend/template/circuit.py:
            # 1692-1709; numpy.interp semantics) -> one 'interp' node over two constant tables (time grid, values)
            if len(n.args) != 3 or not all(isinstance(a, ast.Name) for a in n.args[1:]):
                raise UnsupportedModelError(f"unsupported interp call `{ast.unparse(n)}`")
            tv = self._arr(self.ev(n.args[0]))
            xp = np.asarray(_real_table(n.args[1].id, self.p.arg(n.args[1].id).value), dtype=np.float64)
            fp = np.asarray(_real_table(n.args[2].id, self.p.arg(n.args[2].id).value), dtype=np.float64)
            if tv.ndim != 0 or xp.ndim != 1 or fp.shape != xp.shape or xp.size < 2:
                raise UnsupportedModelError("interp is supported for a scalar time over 1-d (time, value) tables")
            self.tables.append(TableSpec(n.args[1].id, xp.shape, xp))
            self.tables.append(TableSpec(n.args[2].id, fp.shape, fp))
            return _obj(g._node("interp", (tv[()].id,), (len(self.tables) - 2, len(self.tables) - 1, int(xp.size))))
        if fname == "interp_rows":
            # vector-valued timed inputs of adaptive runs: `interp_rows(t, time, u_input)` interpolates every column of the
            # (steps, n_cols) table (frontend/template/circuit.py:1696-1700) -> one 'interp' node per column
            if len(n.args) != 3 or not all(isinstance(a, ast.Name) for a in n.args[1:]):
                raise UnsupportedModelError(f"unsupported interp_rows call `{ast.unparse(n)}`")
            tv = self._arr(self.ev(n.args[0]))
            xp = np.asarray(_real_table(n.args[1].id, self.p.arg(n.args[1].id).value), dtype=np.float64)
            fp = np.asarray(_real_table(n.args[2].id, self.p.arg(n.args[2].id).value), dtype=np.float64)
            if tv.ndim != 0 or xp.ndim != 1 or fp.ndim != 2 or fp.shape[0] != xp.size or xp.size < 2:
                raise UnsupportedModelError("interp_rows is supported for a scalar time over (time[n], values[n, cols]) tables")
            self.tables.append(TableSpec(n.args[1].id, xp.shape, xp))
            tx = len(self.tables) - 1
            out = np.empty(fp.shape[1], dtype=object)
            for c in range(fp.shape[1]):
                col = np.ascontiguousarray(fp[:, c])
                self.tables.append(TableSpec(f"{n.args[2].id}[:, {c}]", col.shape, col))
                out[c] = g._node("interp", (tv[()].id,), (tx, len(self.tables) - 1, int(xp.size)))
            return out
        args = [self.ev(a) for a in n.args]
        if fname in UNARY_FUNCS:
            return self._map1(fname, args[0])
        if fname in ("conjugate", "real", "imag"):
            pick = {"conjugate": lambda x: x.conj() if isinstance(x, CSym) else x,
                    "real": lambda x: x.re if isinstance(x, CSym) else x,
                    "imag": lambda x: x.im if isinstance(x, CSym) else g.const(0.0)}[fname]
            r = np.frompyfunc(pick, 1, 1)(self._arr(args[0]))
            return r if isinstance(r, np.ndarray) else _obj(r)
        if fname == "sigmoid":
            x = self._arr(args[0])
            one = _obj(g.const(1.0, weak=True))
            return _wrap(one / _wrap(one + self._map1("exp", _wrap(-x))))
        if fname in ("maximum", "minimum"):
            return self._map2(fname, args[0], args[1])
        if fname == "identity":
            return args[0]
        if fname in ("dot", "matmul"):
            a, b = self._arr(args[0]), self._arr(args[1])
            if a.ndim == 2 and b.ndim == 1:
                out = np.empty(a.shape[0], dtype=object)
                for i in range(a.shape[0]):
                    out[i] = self._sumseq(a[i, j] * b[j] for j in range(a.shape[1]))
                return out
            if a.ndim == 1 and b.ndim == 1:
                return _obj(self._sumseq(a[j] * b[j] for j in range(a.shape[0])))
            if a.ndim == 2 and b.ndim == 2:
                out = np.empty((a.shape[0], b.shape[1]), dtype=object)
                for i in range(a.shape[0]):
                    for k in range(b.shape[1]):
                        out[i, k] = self._sumseq(a[i, j] * b[j, k] for j in range(a.shape[1]))
                return out
            if a.ndim == 0 or b.ndim == 0:
                return _wrap(a * b)
            raise UnsupportedModelError(f"dot with shapes {a.shape} x {b.shape}")
        if fname == "sum":
            a = self._arr(args[0])
            return _obj(self._sumseq(a.reshape(-1)))
        if fname == "mean":
            a = self._arr(args[0])
            return _obj(self._sumseq(a.reshape(-1)) / g.const(float(a.size), weak=True))
        if fname == "wsum":
            w, c = self._arr(args[0]), self._arr(args[1])
            prod = w * c
            if prod.ndim != 2:
                raise UnsupportedModelError("wsum expects 2-d operands")
            out = np.empty(prod.shape[0], dtype=object)
            for i in range(prod.shape[0]):
                out[i] = self._sumseq(prod[i, :])
            return out
        if fname == "broadcast_pre":
            return self._arr(args[0])[None, :]
        if fname == "broadcast_post":
            return self._arr(args[0])[:, None]
        if fname == "reshape2d":
            return self._arr(args[0]).reshape(int(args[1]), int(args[2]))
        if fname == "flatten1d":
            return self._arr(args[0]).reshape(-1)
        if fname == "concatenate":
            raise UnsupportedModelError("concatenate")
        if fname == "index_1d":
            return self._arr(args[0])[self._index_value(args[1])]
        if fname in ("past", "hist"):
            raise UnsupportedModelError(f"unsupported delay-differential term `{ast.unparse(n)}` (expected the generated form "
                                        f"`hist(t - d)[idx]`)")
        if fname == "randn":
            raise UnsupportedModelError("stochastic `randn` terms are not supported by the CUDA backend")
        raise UnsupportedModelError(f"function `{fname}` is not supported by the CUDA backend")

    def _const_int(self, node) -> int:
        v = self.ev(node)
        if isinstance(v, (int, np.integer)):
            return int(v)
        if isinstance(v, np.ndarray) and v.dtype != object and v.size == 1:
            return int(v.reshape(-1)[0])
        raise UnsupportedModelError(f"expected a constant integer, got `{ast.unparse(node)}`")

    # -- indexing
    def _index_value(self, v):
        if isinstance(v, np.ndarray) and v.dtype != object:
            return int(v) if v.ndim == 0 else v
        if isinstance(v, (int, np.integer)):
            return int(v)
        raise UnsupportedModelError("dynamic (state-dependent) indices are not supported")

    def _index(self, sl):
        """AST slice -> concrete python index (ints, slices, int arrays, tuples thereof)."""
        if isinstance(sl, ast.Tuple):
            return tuple(self._index(e) for e in sl.elts)
        if isinstance(sl, ast.Slice):
            f = lambda x: None if x is None else self._const_int(x)
            return slice(f(sl.lower), f(sl.upper), f(sl.step))
        if isinstance(sl, ast.Constant) and sl.value is None:
            return None
        return self._index_value(self.ev(sl))

    def _hist_read(self, n: ast.Subscript):
        """``hist(t*dt - d)[idx]`` (fixed-step solvers, integer step counter t) / ``hist(t - d)[idx]`` (adaptive):
        the delayed value of state element(s) idx, linearly interpolated in the recorded history (DDEHistory.__call__,
        base_backend.py:165-177).  One 'hist' leaf per (state element, delay); the delay must be a compile-time constant."""
        g = self.g
        call = n.value
        if self.complex:
            raise UnsupportedModelError("delay-differential terms in complex-valued builds")
        if len(call.args) != 1 or call.keywords or not isinstance(call.args[0], ast.BinOp) or not isinstance(call.args[0].op, ast.Sub):
            raise UnsupportedModelError(f"unsupported history lookup `{ast.unparse(n)}`")
        tq, dl = call.args[0].left, call.args[0].right
        scale = None
        if isinstance(tq, ast.BinOp) and isinstance(tq.op, ast.Mult) and isinstance(tq.left, ast.Name) \
                and isinstance(tq.right, ast.Constant) and isinstance(tq.right.value, float):
            scale, tq = float(tq.right.value), tq.left
        if not (isinstance(tq, ast.Name) and tq.id == self.tname):
            raise UnsupportedModelError(f"unsupported history lookup `{ast.unparse(n)}`")
        t_is_int = bool(g.isint[self.env[self.tname][()].id])
        if (scale is None) == t_is_int:
            raise UnsupportedModelError(f"history lookup `{ast.unparse(n)}` does not match the solver's time argument")
        self.uses_t = True
        # NumPy promotion of `t - d` in the reference: a named (array-typed) delay is strong, a literal is a Python float
        strong = not isinstance(dl, ast.Constant)
        try:
            dv = self.ev(dl)
        except IndexError:
            # vectorised builds print `d[0]` for a (1,)-shaped delay that reaches the backend squeezed to 0-d
            # (base_backend.py:338-339 vs :675-676): take the only element
            if not isinstance(dl, ast.Subscript):
                raise
            dv = self.ev(dl.value)
        dparam = None
        if isinstance(dv, (int, np.integer)):
            dval = float(dv)
        else:
            dv = self._arr(dv)
            dv = dv[()] if isinstance(dv, np.ndarray) and dv.ndim == 0 else (dv.reshape(-1)[0] if isinstance(dv, np.ndarray) and dv.size == 1 else dv)
            if isinstance(dv, Sym) and g.ops[dv.id] == "param":
                # swept delay (grid_search over `d`): a per-instance parameter; the history depth comes from the caller's
                # bound on the table (InstanceModel(hist_max_delay=...))
                dparam = int(g.payload[dv.id])
                dval = float(self.params[dparam].default)
            elif not isinstance(dv, Sym) or not g.is_const(dv.id):
                raise UnsupportedModelError("history delays must be constants or swept parameters (not state-dependent)")
            else:
                dval = float(g.const_value(dv.id))
        if not (dval >= 0.0) or not math.isfinite(dval):
            raise UnsupportedModelError(f"history delay {dval!r}")
        if scale is not None:
            if self.hist_dt is not None and self.hist_dt != scale:
                raise UnsupportedModelError("history lookups with different step sizes")
            self.hist_dt = scale
        idx = self._index(n.slice)
        n_state = int(np.asarray(self.p.arg_by_kind("y").value).size)
        elems = np.arange(n_state)[idx]

        def one(e):
            e = int(e)
            if e not in self.hist_vars:
                self.hist_vars.append(e)
            if dval not in self.hist_delays:
                self.hist_delays.append(dval)
            if dparam is not None and dparam not in self.hist_delay_params:
                self.hist_delay_params.append(dparam)
            dkey = float(dval).hex() if dparam is None else ("p", dparam)
            return g._node("hist", (), (self.hist_vars.index(e), scale, dkey, strong))
        if np.ndim(elems) == 0:
            return _obj(one(elems))
        out = np.empty(elems.shape, dtype=object)
        for k, e in enumerate(elems.reshape(-1)):
            out.reshape(-1)[k] = one(e)
        return out

    def subscript(self, n: ast.Subscript):
        if isinstance(n.value, ast.Call) and getattr(n.value.func, "id", None) == "hist" and self.hist_name == "hist":
            return self._hist_read(n)
        base = self.ev(n.value)
        if isinstance(base, _Table):
            return self._table_read(base, n.slice)
        idx = self._index(n.slice)
        if isinstance(base, _Ring):
            pos = base.grid[idx]
            return self._ring_read(base, pos)
        base = self._arr(base)
        r = base[idx]
        return r if isinstance(r, np.ndarray) else _obj(r)

    def _table_read(self, tab: _Table, sl):
        g = self.g
        shape = tab.spec.shape
        if isinstance(sl, ast.Tuple):
            raise UnsupportedModelError("input tables may only be indexed as `table[t]`")
        tv = self.ev(sl)
        tv = tv[()] if isinstance(tv, np.ndarray) and tv.dtype == object and tv.ndim == 0 else tv
        if not isinstance(tv, Sym) or not g.isint[tv.id]:
            raise UnsupportedModelError("input tables must be indexed with the integer step counter")
        if len(shape) == 1:
            return _obj(g._node("table", (), (tab.idx, tv.id, 0, 1)))
        if len(shape) == 2:
            out = np.empty(shape[1], dtype=object)
            for c in range(shape[1]):
                out[c] = g._node("table", (), (tab.idx, tv.id, c, shape[1]))
            return out
        raise UnsupportedModelError("input tables with more than 2 dimensions")

    def _ring_read(self, ring: _Ring, pos):
        g = self.g
        L = ring.spec.length

        def one(p):
            p = int(p)
            r, c = divmod(p, L)
            w = ring.writes.get((r, c))
            if w is not None:
                return w
            return g._node("ring", (), (ring.idx, r, c, ring.shift))
        if isinstance(pos, np.ndarray):
            out = np.empty(pos.shape, dtype=object)
            flat_o, flat_p = out.reshape(-1), pos.reshape(-1)
            for k in range(flat_p.size):
                flat_o[k] = one(flat_p[k])
            return out
        return _obj(one(pos))

    # -- statements
    def run(self) -> ScalarRhs:
        g = self.g
        dy = self.env[self.dyname]
        for s, rhs, lidx in self.trees:
            val = self.ev(rhs)
            if lidx is None:
                if isinstance(val, _Rolled):
                    raise UnsupportedModelError("`roll` result must be written back into its ring buffer")
                if s.lhs in (self.yname, self.dyname, self.tname):
                    raise UnsupportedModelError(f"rebinding `{s.lhs}`")
                self.env[s.lhs] = val
                continue
            target = self.env.get(s.lhs)
            if isinstance(target, _Ring):
                if isinstance(val, _Rolled):
                    full = isinstance(lidx, ast.Slice) and lidx.lower is None and lidx.upper is None
                    if val.ring is not target or not full or target.writes:
                        raise UnsupportedModelError("unsupported ring-buffer roll pattern")
                    target.shift += val.shift
                    continue
                pos = target.grid[self._index(lidx)]
                v = self._arr(val)
                if isinstance(pos, np.ndarray):
                    vb = np.broadcast_to(v, pos.shape)
                    for p, x in zip(pos.reshape(-1), vb.reshape(-1)):
                        target.writes[divmod(int(p), target.spec.length)] = x
                else:
                    target.writes[divmod(int(pos), target.spec.length)] = v.reshape(-1)[0] if v.size == 1 else None
                continue
            if isinstance(val, _Rolled):
                raise UnsupportedModelError("`roll` result must be written back into its ring buffer")
            if not isinstance(target, np.ndarray) or target.dtype != object:
                raise UnsupportedModelError(f"indexed assignment to `{s.lhs}`")
            if s.lhs != self.dyname and s.lhs not in self.slot_arrays and s.lhs not in self.slot_arrays_complex:
                # indexed write into a local intermediate: copy-on-write keeps earlier aliases intact
                target = target.copy()
                self.env[s.lhs] = target
            v = self._arr(val)
            idx = self._index(lidx)
            if v.ndim == 1 and v.size == 1 and np.ndim(target[idx]) == 0:
                v = v.reshape(())            # dy[k] = (1,)-shaped rhs (numpy<2.3 semantics of the reference)
            if v.ndim == 0:
                v = v[()]                    # store the node itself, not a 0-d object array wrapping it
                if np.ndim(target[idx]) > 0:
                    target[idx] = np.full(np.shape(target[idx]), v, dtype=object)
                    continue
            target[idx] = v
        dy = dy.reshape(-1)
        if any(x is None for x in dy):
            raise UnsupportedModelError("vector field leaves some state derivatives unassigned")
        # ring bookkeeping
        ring_writes = []
        for r in self.rings:
            r.spec.rolls_per_eval = r.shift
            for (row, col), node in r.writes.items():
                ring_writes.append((r.idx, row, col, node.id))
        if self.complex:
            # complex state -> interleaved (re, im) real derivatives; real-valued right-hand sides get im = 0
            flat = []
            for x in dy:
                c = CSym.of(g, x)
                flat += [c.re, c.im]
            n_out, dy_ids = len(flat), [x.id for x in flat]
        else:
            if any(isinstance(x, CSym) for x in dy):
                raise UnsupportedModelError("complex derivative in a real-valued build")
            n_out, dy_ids = dy.size, [g.lift(x).id for x in dy]
        n_in = int(np.asarray(self.p.arg_by_kind("y").value).size) * (2 if self.complex else 1)
        if n_out != n_in and not self.jacobian:
            raise UnsupportedModelError(f"vector field returns {n_out} derivatives for {n_in} state variables")
        out = ScalarRhs(g, n_in, dy_ids, self.params, self.tables, [r.spec for r in self.rings],
                        ring_writes, self.slots, [], self.uses_t, n_out if self.jacobian else None,
                        list(self.hist_vars), list(self.hist_delays), self.hist_dt, list(self.hist_delay_params))
        # persistent slots: store back only if any evaluation can observe an incoming value
        slot_final = {}
        for name, arr in self.slot_arrays.items():
            base = [k for k, s in enumerate(self.slots) if s[0] == name]
            for k, x in zip(base, arr.reshape(-1)):
                slot_final[k] = x.id
        out.slot_writes = [(k, n) for k, n in slot_final.items()]
        live = set(out.live_nodes())
        read_slots = {g.payload[i] for i in live if g.ops[i] == "slot"}
        changed = [(k, n) for k, n in slot_final.items()
                   if not (g.ops[n] == "slot" and g.payload[n] == k)]
        if read_slots:
            # conservative: once any slot of a buffer is observed across evaluations keep all its writes
            names = {self.slots[k][0] for k in read_slots}
            out.slot_writes = [(k, n) for k, n in changed if self.slots[k][0] in names]
        else:
            out.slot_writes = []
        return out


def scalarize(prog: Program, sweep: Sequence[Tuple[str, int]] = (), max_nodes: int = 6000) -> ScalarRhs:
    """Lower ``prog`` to a scalar DAG.  ``sweep`` lists ``(arg_name, flat_element)`` pairs that vary
    per sweep instance (they become runtime 'param' leaves instead of baked constants)."""
    return _Interp(prog, sweep, max_nodes).run()
