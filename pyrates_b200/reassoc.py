"""fast_math reassociation pass of the instance model: group thread-invariant factors.

An FP64-pipe-bound kernel pays for every multiply.  The reference-generated expressions interleave per-instance
parameters, literals and state-dependent terms, e.g. the Jansen-Rit PSP drive

    h * (weight * (m_max / (1 + exp(..))) + u) / tau          ->   mul, div(4), mul, add, mul      (8 fp64 instructions)

In a sweep ``weight`` and ``u`` are per-thread constants, so the same value is  ``K0 + K1 * rcp(1 + exp(..))``  with
``K0 = h*u/tau`` and ``K1 = h*weight*m_max/tau`` computed ONCE per thread: reciprocal (3) + one FMA.  This pass rewrites
the scalar DAG into that form:

* every real node is normalised to a linear combination  ``inv + sum_j K_j * B_j``  with thread-invariant coefficients
  ``inv, K_j`` (expressions over literals and swept parameters only) and state-dependent bases ``B_j``;
* multiplication / division by an invariant distributes over the combination (the coefficients absorb it),
  ``a / b`` with a state-dependent ``b`` becomes ``K_a * (B_a * (1 / b))`` so that the numerator's factor joins the
  coefficient and the division is a bare reciprocal;
* materialisation is a multiply-add chain ``((inv + B_1*K_1) + B_2*K_2) ...`` which the compiler contracts to FMAs;
* invariant sub-expressions are emitted once per thread (``emit_inst``: ``T.q[k]`` in ``prb_thread_init``).

This is algebraically exact and changes rounding by a few ulp per folded factor — inside ``fast_math``'s contract (the
trajectories stay within 1e-9 of the exact kernel; ``tests/test_gpu_fullsize.py``), never applied to the exact kernels.
Only pure vector fields (no ring buffers / mutable slots) are rewritten (float64: speculative exp / reciprocal kernels;
float32: `__expf` / `__fdividef` kernels — JRC RK4 sweep 3.7e11 -> 4.0e11 node-steps/s).
"""
from __future__ import annotations

import copy
import os
from typing import Dict, List, Optional, Tuple

from .scalarize import Graph, ScalarRhs, Sym

_LEAVES = ("state", "t", "slot", "table", "ring", "interp", "hist")


class _LC:
    """inv + sum_j K_j * B_j; ``inv`` / ``K_j`` are Syms of thread-invariant nodes (None = 0 resp. 1)."""
    __slots__ = ("inv", "terms")

    def __init__(self, inv: Optional[Sym], terms: List[Tuple[Optional[Sym], Sym]]):
        self.inv, self.terms = inv, terms

    @property
    def invariant(self) -> bool:
        return not self.terms


def fold_invariants(rhs: ScalarRhs, max_terms: int = 6) -> ScalarRhs:
    """Return an equivalent ScalarRhs whose graph groups thread-invariant factors (see module docstring)."""
    g = rhs.graph
    ng = Graph(g.real)
    import numpy as _np
    exp2_f32 = _np.dtype(g.real) == _np.float32 and os.environ.get("PRB_EXPF_FTZ", "1") == "1"
    one = lambda: ng.const(1.0, weak=True)
    memo: Dict[int, _LC] = {}
    imemo: Dict[int, Sym] = {}

    def icopy(i: int) -> Sym:
        """Integer-valued nodes (step counter arithmetic of table lookups) are copied structurally."""
        if i in imemo:
            return imemo[i]
        op, a = g.ops[i], g.args[i]
        if op == "const":
            r = ng.const(g.const_value(i), isint=True)
        elif op == "t":
            r = ng.leaf("t", None, isint=True)
        elif op == "neg":
            r = ng.unop("neg", icopy(a[0]))
        else:
            r = ng.binop(op, icopy(a[0]), icopy(a[1]))
        imemo[i] = r
        return r

    def kmul(a: Optional[Sym], b: Optional[Sym]) -> Optional[Sym]:
        if a is None:
            return b
        if b is None:
            return a
        return ng.binop("mul", a, b)

    depth_memo: Dict[int, int] = {}
    cost = {"exp": 8, "exp2": 8, "div": 5, "log": 8, "sin": 12, "cos": 12, "tan": 14, "tanh": 12, "pow": 16, "sqrt": 5}

    def depth(i: int) -> int:
        """Length of the dependency chain that ends in node i of the new graph (rough instruction latencies)."""
        if i not in depth_memo:
            a = ng.args[i]
            depth_memo[i] = (cost.get(ng.ops[i], 1) + max(depth(x) for x in a)) if a else 0
        return depth_memo[i]

    def mat(lc: _LC) -> Sym:
        # shallow bases first: the term that hangs on the longest chain (a sigmoid's reciprocal) joins the sum LAST, so
        # only one FMA follows it on the critical path of the evaluation (the kernels are bound by dependency latency)
        acc = lc.inv
        terms = sorted(lc.terms, key=lambda kb: depth(kb[1].id))
        if acc is None and os.environ.get("PRB_MAT_REGFIRST", "1") == "1":
            # register-file reads (B200: an FP64 instruction with three distinct register operands issues in 3 cycles instead
            # of 2, profiles/r02_issue_probe.txt): literal coefficients are instruction immediates, per-thread coefficients
            # sit in registers — a sum without an invariant offset starts with ONE such product as a plain multiply (two
            # reads) and adds the literal-coefficient terms by FMA (two reads each), at the price of a longer chain behind it
            reg = [kb for kb in terms if kb[0] is not None and not ng.is_const(kb[0].id)]
            if len(reg) >= 1 and len(terms) >= 2:
                terms = [reg[-1]] + [kb for kb in terms if kb is not reg[-1]]
        for k, b in terms:
            t = b if k is None else ng.binop("mul", b, k)
            acc = t if acc is None else ng.binop("add", acc, t)
        return acc if acc is not None else ng.const(0.0, weak=True)

    def single(lc: _LC) -> Optional[Tuple[Optional[Sym], Sym]]:
        """(K, B) when the combination is one scaled base without an invariant offset."""
        if lc.inv is None and len(lc.terms) == 1:
            return lc.terms[0]
        return None

    def scale(lc: _LC, m: Sym) -> _LC:
        return _LC(None if lc.inv is None else ng.binop("mul", lc.inv, m), [(kmul(k, m), b) for k, b in lc.terms])

    def as_dynamic(s: Sym) -> _LC:
        return _LC(None, [(None, s)])

    def form(i: int) -> _LC:
        if i in memo:
            return memo[i]
        op, a = g.ops[i], g.args[i]
        if g.isint[i]:
            s = icopy(i)
            r = _LC(s, []) if g.is_const(i) else as_dynamic(s)
        elif op == "const":
            r = _LC(ng.const(g.const_value(i), weak=True), [])
        elif op == "param":
            r = _LC(ng.leaf("param", g.payload[i]), [])
        elif op in _LEAVES:
            payload = g.payload[i]
            if op == "table":                      # (table, step-counter node, column, n_cols): remap the index node
                payload = (payload[0], icopy(payload[1]).id, payload[2], payload[3])
            args = tuple(mat(form(x)).id for x in a)
            r = as_dynamic(ng._node(op, args, payload))
        elif op in ("add", "sub"):
            x, y = form(a[0]), form(a[1])
            if op == "sub":
                y = _LC(None if y.inv is None else ng.unop("neg", y.inv),
                        [(ng.const(-1.0, weak=True) if k is None else ng.unop("neg", k), b) for k, b in y.terms])
            inv = y.inv if x.inv is None else (x.inv if y.inv is None else ng.binop("add", x.inv, y.inv))
            terms = list(x.terms)
            for k, b in y.terms:
                for n, (k0, b0) in enumerate(terms):
                    if b0.id == b.id:              # same base: add the invariant coefficients
                        terms[n] = (ng.binop("add", k0 if k0 is not None else one(), k if k is not None else one()), b)
                        break
                else:
                    terms.append((k, b))
            r = _LC(inv, terms)
            if len(terms) > max_terms:             # keep combinations short: cut here
                r = as_dynamic(mat(r))
        elif op == "neg":
            x = form(a[0])
            r = _LC(None if x.inv is None else ng.unop("neg", x.inv),
                    [(ng.const(-1.0, weak=True) if k is None else ng.unop("neg", k), b) for k, b in x.terms])
        elif op == "mul":
            x, y = form(a[0]), form(a[1])
            if x.invariant and y.invariant:
                r = _LC(ng.binop("mul", x.inv, y.inv) if x.inv is not None and y.inv is not None else None, [])
            elif x.invariant:
                r = scale(y, x.inv) if x.inv is not None else _LC(None, [])
            elif y.invariant:
                r = scale(x, y.inv) if y.inv is not None else _LC(None, [])
            else:
                sx, sy = single(x), single(y)
                if sx is not None and sy is not None:
                    r = _LC(None, [(kmul(sx[0], sy[0]), ng.binop("mul", sx[1], sy[1]))])
                elif sx is not None:
                    r = _LC(None, [(sx[0], ng.binop("mul", sx[1], mat(y)))])
                elif sy is not None:
                    r = _LC(None, [(sy[0], ng.binop("mul", mat(x), sy[1]))])
                else:
                    r = as_dynamic(ng.binop("mul", mat(x), mat(y)))
        elif op == "div":
            x, y = form(a[0]), form(a[1])
            if y.invariant:
                d = y.inv if y.inv is not None else ng.const(0.0, weak=True)
                if x.invariant:
                    r = _LC(ng.binop("div", x.inv if x.inv is not None else ng.const(0.0, weak=True), d), [])
                else:
                    r = scale(x, ng.binop("div", one(), d))
            else:
                sy = single(y)
                den = mat(y) if sy is None else sy[1]
                kden = None if sy is None else sy[0]           # a / (K*B) = (a / K) * (1 / B)
                rcp = ng.binop("div", one(), den)
                if x.invariant:
                    k = x.inv if x.inv is not None else ng.const(0.0, weak=True)
                    r = _LC(None, [(k if kden is None else ng.binop("div", k, kden), rcp)])
                else:
                    sx = single(x)
                    if sx is not None:
                        k = sx[0] if kden is None else ng.binop("div", sx[0] if sx[0] is not None else one(), kden)
                        r = _LC(None, [(k, ng.binop("mul", sx[1], rcp))])
                    else:
                        k = None if kden is None else ng.binop("div", one(), kden)
                        r = _LC(None, [(k, ng.binop("mul", mat(x), rcp))])
        elif op in ("pow", "maximum", "minimum"):
            x, y = form(a[0]), form(a[1])
            s = ng.binop(op, mat(x), mat(y))
            r = _LC(s, []) if (x.invariant and y.invariant) else as_dynamic(s)
        else:                                       # unary functions
            x = form(a[0])
            if op == "exp" and exp2_f32 and not x.invariant:
                # float32 fast_math: e^x = 2^(x log2 e) on MUFU.EX2 — the factor log2 e joins the coefficients of the
                # argument's combination instead of costing a multiply per call
                s = ng.unop("exp2", mat(scale(x, ng.const(1.4426950408889634, weak=True))))
            else:
                s = ng.unop(op, mat(x))
            r = _LC(s, []) if x.invariant else as_dynamic(s)
        memo[i] = r
        return r

    for i in rhs.live_nodes():            # ids are topologically ordered: bottom-up, no deep recursion
        form(i)
    out = copy.copy(rhs)
    out.graph = ng
    out.dy = [mat(form(n)).id for n in rhs.dy]
    out.ring_writes = [(r_, row, col, mat(form(n)).id) for r_, row, col, n in rhs.ring_writes]
    out.slot_writes = [(s_, mat(form(n)).id) for s_, n in rhs.slot_writes]
    return out


def invariant_nodes(rhs: ScalarRhs) -> Dict[int, bool]:
    """Node id -> True when the node depends on literals and swept parameters only (thread-invariant)."""
    g = rhs.graph
    inv: Dict[int, bool] = {}
    for i in rhs.live_nodes():
        op = g.ops[i]
        if op in ("const", "param"):
            inv[i] = True
        elif op in _LEAVES:
            inv[i] = False
        else:
            inv[i] = all(inv.get(x, False) for x in g.args[i])
    return inv
