"""Network front end: Program -> phased kernel IR for the *network* execution model.

Used when the state vector is too large to live in one thread (populations with N units, dense / pair-space
coupling).  One rhs evaluation becomes a short sequence of *phases*; inside a phase every statement is
either

* an element-wise store over an index space of length n (thread-per-element, or fused behind the row
  reductions of the same length: warp-per-row), or
* a row reduction  ``out[i] = sum_j W[i,j] * g(i, j)``  covering the reference's dense ``dot(W, x)``
  (pyrates/ir/circuit.py:765-782, base_funcs.py:141-142), the pair-space ``wsum(W, g(post, pre))``
  (circuit.py:784-876, base_funcs.py:75-100) and the global ``sum(x)`` (circuit.py:752-761), or
* a full reduction to a scalar (computed redundantly per block, no grid barrier).

A grid-wide barrier separates phases only where a value produced at one index is consumed at another
(gather edges ``m[source_idx]``, scatter ``u[target_idx] = ...``, reductions over computed vectors).  Delay
ring buffers (circuit.py:400-425) use circular addressing and advance once per evaluation.

Expressions are kept as a scalar DAG (``scalarize.Graph``) whose leaves are *indexed loads*; element-wise
arithmetic reuses exactly the folding rules of the instance model.
"""
from __future__ import annotations

import ast
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .devgen import is_device_generated
from .program import Program
from .scalarize import Graph, Sym, UnsupportedModelError, UNARY_FUNCS

__all__ = ["NetIR", "netcompile", "MemArray", "Reduction", "Store"]

SCALAR, VEC, PAIR = "scalar", "vec", "pair"


@dataclass
class MemArray:
    """A named array in device memory.  space: 'y' (stage input), 'const', 'iconst', 'work', 'ring', 'table'."""
    name: str
    space: str
    shape: Tuple[int, ...]
    offset: int = 0                 # element offset inside the packed buffer of its space
    init: Optional[np.ndarray] = None
    rolls: int = 0                  # rings: rolls per evaluation
    persistent: bool = False        # work arrays that keep their value across evaluations ('var' args)
    replicate: bool = False         # 2-d constants every rank holds completely (interpolation tables), not row-sharded

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1


@dataclass
class Reduction:
    id: int
    kind: str                        # 'row' (sum over j of expr(i,j)) | 'full' (sum over j of expr(j))
    n: int                           # rows (1 for 'full')
    m: int                           # reduced length
    node: int                        # DAG node of the summand (roles i / j / ij)
    phase: int
    simple: Optional[Tuple[str, str, int]] = None   # (W array, x array, x offset) when summand == W[i,j]*x[j]


@dataclass
class Store:
    n: int                            # index-space length (1 for scalars)
    target: Tuple                     # ('work', array) | ('dy', state_offset) | ('ring', array, col) | ('scatter', array, idx_array)
    node: int
    phase: int
    scalar: bool = False              # value is a scalar broadcast to n elements


@dataclass
class NetIR:
    graph: Graph
    n_state: int
    arrays: Dict[str, MemArray]
    reductions: List[Reduction]
    stores: List[Store]
    n_phases: int
    const_size: int
    iconst_size: int
    work_size: int
    ring_size: int
    table_size: int
    uses_t: bool
    params: List[Tuple[str, float]] = field(default_factory=list)    # swept per-instance scalars (arg name, default)


class _V:
    """Value in the network interpreter."""
    __slots__ = ("kind", "n", "m", "sym", "idx")

    def __init__(self, kind, sym=None, n=1, m=1, idx=None):
        self.kind, self.sym, self.n, self.m, self.idx = kind, sym, n, m, idx


class _NetInterp:
    def __init__(self, prog: Program, separable: bool = True, sweep: Sequence[str] = ()):
        self.p = prog
        self.separable = separable
        sweep = list(sweep)
        self.params: List[Tuple[str, float]] = [(n, 0.0) for n in sweep]   # rows of the params array, caller's order
        self._role_memo: Dict[int, frozenset] = {}
        self.g = Graph(prog.real_dtype)
        self.arrays: Dict[str, MemArray] = {}
        self.env: Dict[str, object] = {}
        self.reductions: List[Reduction] = []
        self.stores: List[Store] = []
        self.avail: Dict[str, int] = {}            # array -> first phase in which its new content is readable
        self.last_read: Dict[str, int] = {}        # persistent arrays: latest phase that read the OLD content
        self.red_phase: Dict[int, int] = {}
        self.off = {"const": 0, "iconst": 0, "work": 0, "ring": 0, "table": 0}
        self.uses_t = False
        self.tname = prog.arg_by_kind("t").name
        self.yname = prog.arg_by_kind("y").name
        self.dyname = prog.arg_by_kind("dy").name
        self.n_state = prog.n_state
        self.dy_done = np.zeros(self.n_state, dtype=bool)
        self._tmp = 0
        trees = [(s, ast.parse(s.rhs, mode="eval").body,
                  ast.parse(f"_[{s.lhs_idx}]", mode="eval").body.slice if s.lhs_idx is not None else None)
                 for s in prog.stmts]
        self.trees = trees
        t_indexed, rolled = set(), set()
        for s, rhs, _ in trees:
            for node in ast.walk(rhs):
                if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name):
                    if self.tname in {n.id for n in ast.walk(node.slice) if isinstance(n, ast.Name)}:
                        t_indexed.add(node.value.id)
                if isinstance(node, ast.Call) and getattr(node.func, "id", None) == "roll" and node.args \
                        and isinstance(node.args[0], ast.Name):
                    rolled.add(node.args[0].id)
        unknown = [n for n in sweep if n not in {a.name for a in prog.args}]
        if unknown:
            raise KeyError(f"sweep targets {unknown} are not arguments of the program")
        bound = set()
        for a in prog.args:
            # device-generated constants (devgen.DeviceUniform) own no host memory: they only flow to the constant space
            lazy = is_device_generated(a.value)
            if lazy and (a.kind != "const" or a.value.ndim != 2 or a.name in sweep):
                raise UnsupportedModelError(f"device-generated argument `{a.name}` must be a 2-d read-only constant")
            v = a.value if lazy else np.asarray(a.value)
            if a.kind == "t":
                self.env[a.name] = _V(SCALAR, self.g.leaf("t", None, isint=True))
            elif a.kind == "y":
                self._add_array(a.name, "y", (v.size,), None)
                self.env[a.name] = ("array", a.name)
            elif a.kind == "dy":
                self.env[a.name] = ("dy",)
            elif a.name in t_indexed:
                self._add_array(a.name, "table", v.shape, v)
                self.env[a.name] = ("table", a.name)
            elif a.name in rolled:
                if v.ndim != 2:
                    raise UnsupportedModelError("network model supports 2-d (rows x delay) ring buffers only")
                arr = self._add_array(a.name, "ring", v.shape, v)
                self.env[a.name] = ("ring", a.name)
            elif not lazy and a.is_int:
                if v.ndim == 0:
                    self.env[a.name] = int(v)
                else:
                    self._add_array(a.name, "iconst", v.shape, v.astype(np.int32))
                    self.env[a.name] = ("index", a.name)
            elif not lazy and np.iscomplexobj(v):
                raise UnsupportedModelError("complex-valued models are not supported by the CUDA backend")
            elif a.name in sweep:
                # per-instance scalar of a batched sweep (a 0-d parameter, or a uniform per-node vector swept as a whole)
                if a.kind != "const" or not (v.ndim == 0 or (v.ndim == 1 and np.all(v == v.reshape(-1)[0]))):
                    raise UnsupportedModelError(f"swept argument `{a.name}` must be a scalar or a uniform per-node vector")
                leaf = self.g.leaf("param", sweep.index(a.name))
                self.params[sweep.index(a.name)] = (a.name, float(v.reshape(-1)[0]))
                bound.add(a.name)
                self.env[a.name] = _V(SCALAR, leaf) if v.ndim == 0 else _V(VEC, leaf, n=v.size)
            elif a.kind == "var":
                arr = self._add_array(a.name, "work", v.shape, v)
                arr.persistent = True
                self.env[a.name] = ("array", a.name)
            elif v.ndim == 0:
                self.env[a.name] = _V(SCALAR, self.g.const(float(v)))
            elif v.ndim == 1 and v.size > 1 and np.all(v == v[0]):
                # uniform per-node parameter vectors fold to immediates, but keep their length
                self.env[a.name] = _V(VEC, self.g.const(float(v[0])), n=v.size)
            else:
                self._add_array(a.name, "const", v.shape, v)
                self.env[a.name] = ("array", a.name)
        unbound = [n for n in sweep if n not in bound]
        if unbound:
            raise UnsupportedModelError(f"swept arguments {unbound} are not real-valued constants of the program")

    # ------------------------------------------------------------------ arrays
    def _add_array(self, name, space, shape, init) -> MemArray:
        size = int(np.prod(shape)) if shape else 1
        key = "work" if space == "work" else space
        off = 0
        if space in self.off:
            off = self.off[key]
            self.off[key] = off + ((size + 31) // 32) * 32       # 256-byte alignment for fp64 vector loads
        # read-only constants are referenced, not copied (a dense connectivity can be tens of GiB: C5 at N=65536 is 32 GiB)
        if is_device_generated(init):
            arr = MemArray(name, space, tuple(shape), off, init)
            self.arrays[name] = arr
            self.avail[name] = 0
            return arr
        keep = space in ("const", "iconst", "table") and init is not None and np.asarray(init).nbytes >= (1 << 20)
        arr = MemArray(name, space, tuple(shape), off, None if init is None else (np.asarray(init) if keep else np.array(init, copy=True)))
        self.arrays[name] = arr
        self.avail[name] = 0
        return arr

    def _interp_table(self, name: str) -> MemArray:
        """A (time or value) table of an interp call as a replicated constant array (never row-sharded, never folded)."""
        arr = self.arrays.get(name)
        if arr is None:
            a = self.p.arg(name)
            if a.kind != "const" or a.is_int or name in [n_ for n_, _ in self.params]:
                raise UnsupportedModelError(f"interpolation table `{name}` must be a real-valued constant that is not swept")
            arr = self._add_array(name, "const", np.asarray(a.value).shape, np.asarray(a.value))
        if arr.space != "const":
            raise UnsupportedModelError(f"interpolation table `{name}` must be a constant")
        arr.replicate = True
        return arr

    def _new_work(self, shape, hint="tmp") -> MemArray:
        self._tmp += 1
        return self._add_array(f"__{hint}{self._tmp}", "work", shape, np.zeros(shape))

    # ------------------------------------------------------------------ helpers
    def _leaf(self, op, payload):
        return self.g._node(op, (), payload)

    def _load(self, name: str, role: str, off: int = 0, n: Optional[int] = None) -> _V:
        arr = self.arrays[name]
        if n is None:
            n = arr.size
        if n == 1 and not (arr.shape and arr.shape[0] > 1):
            return _V(SCALAR, self._leaf("lds", (name, off)))
        return _V(VEC, self._leaf("ld", (name, off, "i")), n=n)

    def node_phase(self, node: int) -> int:
        """Earliest phase in which `node` can be evaluated (all loaded arrays readable)."""
        g = self.g
        ph, seen, stack = 0, set(), [node]
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            op = g.ops[i]
            if op in ("ld", "lds", "ldm", "gather"):
                ph = max(ph, self.avail.get(g.payload[i][0], 0))
            elif op == "ldring":
                # only logical column 0 can hold a value written during this evaluation
                if g.payload[i][1] == 0:
                    ph = max(ph, self.avail.get(g.payload[i][0] + "@0", 0))
            elif op == "ldringg":
                if g.payload[i][4]:                       # some edge has delay 0: it reads this evaluation's write
                    ph = max(ph, self.avail.get(g.payload[i][0] + "@0", 0))
            elif op == "red":
                ph = max(ph, self.red_phase[g.payload[i][0]])
            elif op == "table":
                stack.append(g.payload[i][1])
            stack.extend(g.args[i])
        return ph

    def _mark_reads(self, node: int, phase: int):
        g = self.g
        seen, stack = set(), [node]
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            if g.ops[i] in ("ld", "lds", "gather"):
                name = g.payload[i][0]
                a = self.arrays.get(name)
                if a is not None and a.persistent and self.avail.get(name, 0) == 0:
                    self.last_read[name] = max(self.last_read.get(name, -1), phase)
            stack.extend(g.args[i])

    def _role_swap(self, node: int, frm: str, to: str) -> Sym:
        """Rebuild `node` with every load of role `frm` turned into role `to` (i <-> j)."""
        g = self.g
        memo: Dict[int, Sym] = {}

        def rec(i: int) -> Sym:
            if i in memo:
                return memo[i]
            op, a, pl = g.ops[i], g.args[i], g.payload[i]
            if op == "ld" and pl[2] == frm:
                r = self._leaf("ld", (pl[0], pl[1], to))
            elif op == "gather" and pl[3] == frm:
                r = self._leaf("gather", (pl[0], pl[1], pl[2], to))
            elif op == "ldring" and pl[2] == frm:
                r = self._leaf("ldring", (pl[0], pl[1], to))
            elif op == "ldringg" and pl[3] == frm:
                r = self._leaf("ldringg", (pl[0], pl[1], pl[2], to, pl[4]))
            elif op == "interprow" and pl[4] == frm:
                r = g._node("interprow", a, (pl[0], pl[1], pl[2], pl[3], to))
            elif op == "red" and frm == "i":
                raise UnsupportedModelError("internal: reduction result used as reduction source without materialisation")
            elif not a:
                r = Sym(g, i)
            elif len(a) == 1:
                r = g.unop(op, rec(a[0]))
            else:
                r = g.binop(op, rec(a[0]), rec(a[1]))
            memo[i] = r
            return r
        return rec(node)

    def _has(self, node: int, ops: Tuple[str, ...]) -> bool:
        g = self.g
        seen, stack = set(), [node]
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            if g.ops[i] in ops:
                return True
            stack.extend(g.args[i])
        return False

    def _materialize(self, v: _V, hint="m") -> Tuple[str, int]:
        """Make vector value `v` addressable as (array, offset) for non-local consumers."""
        g = self.g
        if v.kind != VEC:
            raise UnsupportedModelError("internal: only vectors can be materialised")
        op = g.ops[v.sym.id]
        if op == "ld" and g.payload[v.sym.id][2] == "i":
            name, off, _ = g.payload[v.sym.id]
            return name, off
        arr = self._new_work((v.n,), hint)
        ph = self.node_phase(v.sym.id)
        self._mark_reads(v.sym.id, ph)
        self.stores.append(Store(v.n, ("work", arr.name), v.sym.id, ph))
        self.avail[arr.name] = ph + 1
        return arr.name, 0

    # ------------------------------------------------------------------ expressions
    def as_val(self, x) -> _V:
        if isinstance(x, _V):
            return x
        if isinstance(x, tuple):
            if x[0] == "array":
                arr = self.arrays[x[1]]
                if len(arr.shape) == 2:
                    return _V(PAIR, self._leaf("ldm", (x[1], 0, arr.shape[1])), n=arr.shape[0], m=arr.shape[1])
                if arr.size == 1:
                    return _V(SCALAR, self._leaf("lds", (x[1], 0)))
                return _V(VEC, self._leaf("ld", (x[1], 0, "i")), n=arr.size)
            raise UnsupportedModelError(f"`{x[0]}` object used in arithmetic")
        if isinstance(x, bool):
            raise UnsupportedModelError("boolean literal")
        if isinstance(x, int):
            return _V(SCALAR, self.g.const(int(x), isint=True))
        if isinstance(x, float):
            return _V(SCALAR, self.g.const(x, weak=True))
        raise UnsupportedModelError(f"cannot use {type(x).__name__} in arithmetic")

    def _binary(self, op: str, a, b) -> _V:
        a, b = self.as_val(a), self.as_val(b)
        sym = self.g.binop(op, a.sym, b.sym)
        return self._shape_of(sym, a, b)

    def _shape_of(self, sym, *vals) -> _V:
        kinds = [v.kind for v in vals]
        if PAIR in kinds:
            n = max([v.n for v in vals if v.kind == PAIR] + [1])
            m = max([v.m for v in vals if v.kind == PAIR] + [1])
            for v in vals:
                if v.kind == VEC:
                    raise UnsupportedModelError("vector x matrix broadcasting outside wsum/dot")
            return _V(PAIR, sym, n=n, m=m)
        if VEC in kinds:
            ns = {v.n for v in vals if v.kind == VEC}
            if len(ns) != 1:
                raise UnsupportedModelError(f"element-wise operation on vectors of different lengths {sorted(ns)}")
            return _V(VEC, sym, n=ns.pop())
        return _V(SCALAR, sym)

    def ev(self, n):
        g = self.g
        if isinstance(n, ast.Constant):
            if isinstance(n.value, (int, float)) and not isinstance(n.value, bool):
                return n.value
            raise UnsupportedModelError(f"literal {n.value!r}")
        if isinstance(n, ast.Name):
            if n.id == "pi":
                return math.pi
            if n.id not in self.env:
                raise UnsupportedModelError(f"unknown name `{n.id}` in generated code")
            if n.id == self.tname:
                self.uses_t = True
            return self.env[n.id]
        if isinstance(n, ast.UnaryOp):
            v = self.as_val(self.ev(n.operand))
            if isinstance(n.op, ast.USub):
                return self._shape_of(g.unop("neg", v.sym), v)
            if isinstance(n.op, ast.UAdd):
                return v
            raise UnsupportedModelError(ast.unparse(n))
        if isinstance(n, ast.BinOp):
            ops = {ast.Add: "add", ast.Sub: "sub", ast.Mult: "mul", ast.Div: "div", ast.Pow: "pow"}
            for k, name in ops.items():
                if isinstance(n.op, k):
                    return self._binary(name, self.ev(n.left), self.ev(n.right))
            raise UnsupportedModelError(f"operator in `{ast.unparse(n)}`")
        if isinstance(n, ast.Call):
            return self.call(n)
        if isinstance(n, ast.Subscript):
            return self.subscript(n)
        raise UnsupportedModelError(f"expression `{ast.unparse(n)}`")

    def call(self, n: ast.Call):
        g = self.g
        fname = getattr(n.func, "id", None)
        if fname is None or n.keywords:
            raise UnsupportedModelError(ast.unparse(n))
        if fname == "roll":
            base = self.ev(n.args[0])
            if not (isinstance(base, tuple) and base[0] == "ring"):
                raise UnsupportedModelError("`roll` is only supported as the delay ring-buffer idiom")
            shift = self.ev(n.args[1])
            axis = self.ev(n.args[2]) if len(n.args) > 2 else None
            if not isinstance(shift, int) or axis not in (1, -1):
                raise UnsupportedModelError("ring buffers must be rolled along their last axis")
            return ("rolled", base[1], shift)
        args = [self.ev(a) for a in n.args]
        if fname in UNARY_FUNCS:
            v = self.as_val(args[0])
            return self._shape_of(g.unop(fname, v.sym), v)
        if fname == "sigmoid":
            v = self.as_val(args[0])
            one = g.const(1.0, weak=True)
            s = g.binop("div", one, g.binop("add", one, g.unop("exp", g.unop("neg", v.sym))))
            return self._shape_of(s, v)
        if fname in ("maximum", "minimum"):
            a, b = self.as_val(args[0]), self.as_val(args[1])
            return self._shape_of(g.binop(fname, a.sym, b.sym), a, b)
        if fname == "identity":
            return args[0]
        if fname == "broadcast_post":                 # x[:, None]  -> depends on the row index i
            v = self.as_val(args[0])
            if v.kind != VEC:
                raise UnsupportedModelError("broadcast_post of a non-vector")
            return _V(PAIR, v.sym, n=v.n, m=1)
        if fname == "broadcast_pre":                  # x[None, :]  -> depends on the column index j
            v = self.as_val(args[0])
            if v.kind != VEC:
                raise UnsupportedModelError("broadcast_pre of a non-vector")
            src = self._as_source(v)
            return _V(PAIR, src, n=1, m=v.n)
        if fname in ("dot", "matmul"):
            w, x = self.as_val(args[0]), self.as_val(args[1])
            if w.kind == SCALAR or x.kind == SCALAR:
                return self._binary("mul", w, x)
            if w.kind != PAIR or g.ops[w.sym.id] != "ldm" or x.kind != VEC or x.n != w.m:
                raise UnsupportedModelError("network model supports dot(W, x) with a constant 2-d W and a vector x")
            src = self._as_source(x)
            summand = g.binop("mul", w.sym, src)
            simple = None
            if g.ops[src.id] == "ld":
                simple = (g.payload[w.sym.id][0], g.payload[src.id][0], g.payload[src.id][1])
            elif g.ops[src.id] == "ldring":          # delayed source: one contiguous column of the ring buffer
                simple = (g.payload[w.sym.id][0], g.payload[src.id][0], ("ring", g.payload[src.id][1]))
            return self._reduce_rows(summand, w.n, w.m, simple)
        if fname == "wsum":
            w, c = self.as_val(args[0]), self.as_val(args[1])
            if w.kind != PAIR or c.kind != PAIR:
                raise UnsupportedModelError("wsum expects 2-d operands")
            nrow, ncol = max(w.n, c.n), max(w.m, c.m)
            if self.separable and g.ops[w.sym.id] == "ldm" and (w.n, w.m) == (nrow, ncol):
                terms = self._separate(c.sym.id)
                if terms is not None and 0 < len(terms) <= 4:
                    return self._wsum_separated(w, terms, nrow, ncol)
            return self._reduce_rows(g.binop("mul", w.sym, c.sym), nrow, ncol, None)
        if fname in ("sum", "mean"):
            v = self.as_val(args[0])
            if v.kind == SCALAR:
                return v
            if v.kind != VEC:
                raise UnsupportedModelError("sum of a 2-d expression")
            src = self._as_source(v)
            ph = self.node_phase(src.id)
            self._mark_reads(src.id, ph)
            rid = len(self.reductions)
            self.reductions.append(Reduction(rid, "full", 1, v.n, src.id, ph))
            self.red_phase[rid] = ph
            out = _V(SCALAR, self._leaf("red", (rid, "s")))
            if fname == "mean":
                out = self._binary("div", out, float(v.n))
            return out
        if fname in ("past", "hist"):
            raise UnsupportedModelError("delay-differential `past()` terms are not supported by the CUDA backend")
        if fname in ("interp", "interp_rows"):
            # timed inputs of adaptive runs: `interp(t, time, u_input)` / `interp_rows(t, time, u_input[steps, cols])`
            # (frontend/template/circuit.py:1692-1709; numpy.interp semantics) -> one node over two constant tables
            if len(n.args) != 3 or not all(isinstance(a, ast.Name) for a in n.args[1:]):
                raise UnsupportedModelError(f"unsupported {fname} call `{ast.unparse(n)}`")
            tv = self.as_val(self.ev(n.args[0]))
            xa, fa = (self._interp_table(a.id) for a in n.args[1:])
            if tv.kind != SCALAR or len(xa.shape) != 1:
                raise UnsupportedModelError("interp is supported for a scalar time over a 1-d time grid")
            if fname == "interp":
                if fa.shape != xa.shape:
                    raise UnsupportedModelError("interp tables of different lengths")
                return _V(SCALAR, g._node("interp", (tv.sym.id,), (xa.name, fa.name, int(xa.size))))
            if len(fa.shape) != 2 or fa.shape[0] != xa.size:
                raise UnsupportedModelError("interp_rows expects values of shape (len(time), n_cols)")
            return _V(VEC, g._node("interprow", (tv.sym.id,), (xa.name, fa.name, int(xa.size), int(fa.shape[1]), "i")),
                      n=int(fa.shape[1]))
        raise UnsupportedModelError(f"function `{fname}` is not supported by the CUDA network model")

    # ------------------------------------------------------------------ separable pair expressions
    def _roles(self, i: int, memo=None) -> frozenset:
        g = self.g
        memo = self._role_memo
        r = memo.get(i)
        if r is None:
            op, pl = g.ops[i], g.payload[i]
            if op in ("ld", "ldring"):
                r = frozenset(pl[2])
            elif op in ("gather", "ldringg"):
                r = frozenset(pl[3])
            elif op == "ldm":
                r = frozenset("ij")
            elif op == "red":
                r = frozenset("i") if pl[1] == "i" else frozenset()
            elif op == "table":
                r = frozenset("i") if pl[2] == "i" else frozenset()
            elif op == "interprow":
                r = frozenset(pl[4])
            else:
                r = frozenset()
                for a in g.args[i]:
                    r = r | self._roles(a)
            memo[i] = r
        return r

    def _separate(self, i: int):
        """Try to write pair expression i as sum_k f_k(i) * g_k(j); returns [(f_k | None, g_k | None)] or None.
        Rules: linearity, products, and the angle-addition / exponential identities
        sin(u+v) = sin u cos v + cos u sin v,  cos(u+v) = cos u cos v - sin u sin v,  exp(u+v) = exp u exp v,
        which turn e.g. the Kuramoto coupling W_ij sin(theta_j - theta_i) into two GEMVs over one pass of W."""
        g = self.g
        role = self._roles(i)
        if "j" not in role:
            return [(Sym(g, i), None)]
        if "i" not in role:
            return [(None, Sym(g, i))]
        op, a = g.ops[i], g.args[i]
        neg = lambda ts: [(g.unop("neg", f), gj) if f is not None else (None, g.unop("neg", gj)) for f, gj in ts]
        if op in ("add", "sub"):
            l, r = self._separate(a[0]), self._separate(a[1])
            if l is None or r is None:
                return None
            return l + (neg(r) if op == "sub" else r)
        if op == "neg":
            t = self._separate(a[0])
            return None if t is None else neg(t)
        if op == "mul":
            l, r = self._separate(a[0]), self._separate(a[1])
            if l is None or r is None or len(l) * len(r) > 4:
                return None
            mul = lambda x, y: y if x is None else (x if y is None else g.binop("mul", x, y))
            return [(mul(f1, f2), mul(g1, g2)) for f1, g1 in l for f2, g2 in r]
        if op in ("sin", "cos", "exp"):
            t = self._separate(a[0])
            if t is None or any(f is not None and gj is not None for f, gj in t):
                return None
            one = g.const(1.0, weak=True)
            u = v = None
            for f, gj in t:
                if gj is None:
                    u = f if u is None else g.binop("add", u, f)
                else:
                    v = gj if v is None else g.binop("add", v, gj)
            if u is None or v is None:
                return None
            if op == "exp":
                return [(g.unop("exp", u), g.unop("exp", v))]
            su, cu, sv, cv = g.unop("sin", u), g.unop("cos", u), g.unop("sin", v), g.unop("cos", v)
            if op == "sin":
                return [(su, cv), (cu, sv)]
            return [(cu, cv), (g.unop("neg", su), sv)]
        return None

    def _wsum_separated(self, w: _V, terms, nrow: int, ncol: int) -> _V:
        g = self.g
        out = None
        wname = g.payload[w.sym.id][0]
        for f, gj in terms:
            src = _V(VEC, self._role_swap(gj.id, "j", "i") if gj is not None else g.const(1.0), n=ncol)
            name, off = self._materialize(src, "sep")
            summand = g.binop("mul", w.sym, self._leaf("ld", (name, off, "j")))
            red = self._reduce_rows(summand, nrow, ncol, (wname, name, off))
            term = red if f is None else self._shape_of(g.binop("mul", f, red.sym), _V(VEC, f, n=nrow), red)
            out = term if out is None else self._shape_of(g.binop("add", out.sym, term.sym), out, term)
        return out

    def _as_source(self, v: _V) -> Sym:
        """Vector value as a j-indexed summand source.  Cheap expressions are inlined (recomputed at j),
        expensive ones (transcendentals, reductions) are materialised once and loaded."""
        g = self.g
        heavy = self._has(v.sym.id, ("red", "interp", "interprow") + tuple(UNARY_FUNCS) + ("pow", "div"))
        if heavy and g.ops[v.sym.id] != "ld":
            name, off = self._materialize(v, "src")
            return self._leaf("ld", (name, off, "j"))
        return self._role_swap(v.sym.id, "i", "j")

    def _reduce_rows(self, summand: Sym, n: int, m: int, simple) -> _V:
        ph = self.node_phase(summand.id)
        self._mark_reads(summand.id, ph)
        rid = len(self.reductions)
        self.reductions.append(Reduction(rid, "row", n, m, summand.id, ph, simple))
        self.red_phase[rid] = ph
        return _V(VEC, self._leaf("red", (rid, "i")), n=n)

    # ------------------------------------------------------------------ indexing
    def _const_index(self, sl):
        if isinstance(sl, ast.Slice):
            f = lambda x: None if x is None else self._int(self.ev(x))
            return slice(f(sl.lower), f(sl.upper), f(sl.step))
        if isinstance(sl, ast.Tuple):
            return tuple(self._const_index(e) for e in sl.elts)
        v = self.ev(sl)
        if isinstance(v, int):
            return v
        if isinstance(v, tuple) and v[0] == "index":
            return v
        if isinstance(v, _V) and v.kind == SCALAR and self.g.isint[v.sym.id] and self.g.ops[v.sym.id] == "t":
            return "t"
        raise UnsupportedModelError(f"index expression `{ast.unparse(sl)}`")

    @staticmethod
    def _int(v) -> int:
        if isinstance(v, int):
            return v
        raise UnsupportedModelError("non-constant slice bound")

    def subscript(self, n: ast.Subscript):
        g = self.g
        base = self.ev(n.value)
        idx = self._const_index(n.slice)
        if isinstance(base, tuple) and base[0] == "table":
            arr = self.arrays[base[1]]
            if idx != "t":
                raise UnsupportedModelError("input tables may only be indexed as `table[t]`")
            tnode = self.env[self.tname].sym.id
            self.uses_t = True
            if len(arr.shape) == 1:
                return _V(SCALAR, g._node("table", (), (base[1], tnode, 0, 1)))
            return _V(VEC, g._node("table", (), (base[1], tnode, "i", arr.shape[1])), n=arr.shape[1])
        if isinstance(base, tuple) and base[0] == "ring":
            arr = self.arrays[base[1]]
            if isinstance(idx, tuple) and len(idx) == 2 and idx[0] == slice(None, None, None) and isinstance(idx[1], int):
                return _V(VEC, self._leaf("ldring", (base[1], idx[1] % arr.shape[1], "i")), n=arr.shape[0])
            if isinstance(idx, tuple) and len(idx) == 2 and all(isinstance(x, tuple) and x[0] == "index" for x in idx):
                # legacy per-edge delays (pyrates/ir/circuit.py:598-641): buf[source_idx, delays] gathers, for every edge e,
                # the value its source wrote delays[e] evaluations ago
                rows_a, cols_a = self.arrays[idx[0][1]], self.arrays[idx[1][1]]
                if rows_a.size != cols_a.size:
                    raise UnsupportedModelError("ring gather with index arrays of different lengths")
                cols = np.asarray(cols_a.init).reshape(-1)
                if cols.min() < 0 or cols.max() >= arr.shape[1] or np.asarray(rows_a.init).min() < 0 \
                        or np.asarray(rows_a.init).max() >= arr.shape[0]:
                    raise UnsupportedModelError("ring gather indices out of range")
                return _V(VEC, self._leaf("ldringg", (base[1], idx[0][1], idx[1][1], "i", bool((cols == 0).any()))), n=rows_a.size)
            raise UnsupportedModelError("network model supports ring reads of the form `buf[:, d]` and `buf[idx, delays]` only")
        if isinstance(base, tuple) and base[0] == "array":
            arr = self.arrays[base[1]]
            if len(arr.shape) != 1:
                raise UnsupportedModelError(f"indexing the 2-d array `{base[1]}`")
            v = _V(VEC, self._leaf("ld", (base[1], 0, "i")), n=arr.size)
        else:
            v = self.as_val(base)
        if v.kind != VEC:
            raise UnsupportedModelError(f"indexing a {v.kind} value")
        if isinstance(idx, int):
            name, off = self._materialize(v, "elem")
            return _V(SCALAR, self._leaf("lds", (name, off + (idx % v.n))))
        if isinstance(idx, slice):
            start, stop, step = idx.indices(v.n)
            if step != 1:
                raise UnsupportedModelError("strided slices")
            name, off = self._materialize(v, "slice")
            if stop - start == 1:
                return _V(VEC, self._leaf("ld", (name, off + start, "i")), n=1)
            return _V(VEC, self._leaf("ld", (name, off + start, "i")), n=stop - start)
        if isinstance(idx, tuple) and idx[0] == "index":
            name, off = self._materialize(v, "gsrc")
            iarr = self.arrays[idx[1]]
            return _V(VEC, self._leaf("gather", (name, off, idx[1], "i")), n=iarr.size)
        raise UnsupportedModelError(f"index `{ast.unparse(n.slice)}`")

    # ------------------------------------------------------------------ statements
    def run(self) -> NetIR:
        g = self.g
        for s, rhs, lidx in self.trees:
            val = self.ev(rhs)
            if lidx is None:
                if isinstance(val, tuple) and val[0] == "rolled":
                    raise UnsupportedModelError("`roll` result must be written back into its ring buffer")
                self.env[s.lhs] = val
                continue
            target = self.env.get(s.lhs)
            if isinstance(target, tuple) and target[0] == "ring":
                arr = self.arrays[target[1]]
                if isinstance(val, tuple) and val[0] == "rolled":
                    if val[1] != target[1]:
                        raise UnsupportedModelError("unsupported ring-buffer roll pattern")
                    arr.rolls += val[2]
                    continue
                idx = self._const_index(lidx)
                if not (isinstance(idx, tuple) and len(idx) == 2 and idx[0] == slice(None, None, None) and idx[1] == 0):
                    raise UnsupportedModelError("network model supports ring writes of the form `buf[:, 0] = x` only")
                v = self.as_val(val)
                ph = self.node_phase(v.sym.id)
                self._mark_reads(v.sym.id, ph)
                self.stores.append(Store(arr.shape[0], ("ring", arr.name, 0), v.sym.id, ph, scalar=v.kind == SCALAR))
                self.avail[arr.name + "@0"] = ph + 1
                continue
            if isinstance(target, tuple) and target[0] == "dy":
                idx = self._const_index(lidx)
                if isinstance(idx, int):
                    a, b = idx, idx + 1
                elif isinstance(idx, slice):
                    a, b, st = idx.indices(self.n_state)
                else:
                    raise UnsupportedModelError("dy index")
                v = self.as_val(val)
                if v.kind == PAIR or (v.kind == VEC and v.n != b - a and v.n != 1):
                    raise UnsupportedModelError(f"dy[{a}:{b}] assigned a value of length {v.n}")
                ph = self.node_phase(v.sym.id)
                self._mark_reads(v.sym.id, ph)
                self.stores.append(Store(b - a, ("dy", a), v.sym.id, ph,
                                         scalar=v.kind == SCALAR or (v.n == 1 and b - a > 1)))
                self.dy_done[a:b] = True
                continue
            if isinstance(target, tuple) and target[0] == "array" and self.arrays[target[1]].space == "work":
                arr = self.arrays[target[1]]
                idx = self._const_index(lidx)
                v = self.as_val(val)
                ph = self.node_phase(v.sym.id)
                ph = max(ph, self.last_read.get(arr.name, -1) + 1)      # WAR: earlier readers of the old content
                self._mark_reads(v.sym.id, ph)
                if isinstance(idx, tuple) and idx[0] == "index":
                    iarr = self.arrays[idx[1]]
                    self.stores.append(Store(iarr.size, ("scatter", arr.name, idx[1]), v.sym.id, ph, scalar=v.kind == SCALAR))
                elif isinstance(idx, slice) and idx.indices(arr.size) == (0, arr.size, 1):
                    self.stores.append(Store(arr.size, ("work", arr.name), v.sym.id, ph, scalar=v.kind == SCALAR))
                elif isinstance(idx, int):
                    self.stores.append(Store(1, ("work_at", arr.name, idx % arr.size), v.sym.id, ph, scalar=True))
                else:
                    raise UnsupportedModelError(f"assignment `{s.text()}`")
                self.avail[arr.name] = max(self.avail.get(arr.name, 0), ph + 1)
                continue
            raise UnsupportedModelError(f"assignment `{s.text()}` in the network model")
        if not self.dy_done.all():
            raise UnsupportedModelError("vector field leaves some state derivatives unassigned")
        # ring-buffer reads of logical column 0 after this evaluation's write must wait for the write
        for st in self.stores:
            pass
        n_phases = 1 + max([s.phase for s in self.stores] + [r.phase for r in self.reductions] + [0])
        return NetIR(g, self.n_state, self.arrays, self.reductions, self.stores, n_phases,
                     self.off["const"], self.off["iconst"], self.off["work"], self.off["ring"], self.off["table"],
                     self.uses_t, self.params)


def netcompile(prog: Program, separable: bool = True, sweep: Sequence[str] = ()) -> NetIR:
    """``separable``: rewrite pair-space couplings that factor as sum_k f_k(i) g_k(j) (e.g. sin(theta_j - theta_i))
    into GEMVs over materialised source vectors — algebraically equal, not bit-equal, to the direct evaluation."""
    return _NetInterp(prog, separable, sweep).run()
