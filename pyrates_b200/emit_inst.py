"""CUDA C emitter for the *instance* execution model (one thread = one sweep instance).

Input: the scalar DAG produced by ``scalarize`` (walks the same compute-graph assignment stream the
reference's NumPy backend prints, pyrates/backend/computegraph.py:1109-1228).  Output: one CUDA C
translation unit = kernel ABI + generated ``prb_rhs`` / ``prb_record`` + the hand-written solver
template ``csrc/templates/inst_kernel.cuh`` (fused rhs + Euler/Heun/RK4 update + decimated observer store).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .scalarize import ScalarRhs, UnsupportedModelError

_TPL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "templates")

SOLVER_IDS = {"euler": 0, "heun": 1, "heun_ref": 1, "heun_true": 2, "rk4": 3, "scipy": 4, "rk45": 4, "dopri5_dde": 5}
ADAPTIVE_SOLVERS = ("scipy", "rk45", "dopri5_dde")

_FUNC = {"exp": "exp", "log": "log", "sin": "sin", "cos": "cos", "tan": "tan", "tanh": "tanh", "sinh": "sinh",
         "cosh": "cosh", "arctan": "atan", "arcsin": "asin", "arccos": "acos", "sqrt": "sqrt", "abs": "fabs",
         "round": "rint"}


ETAB_BITS = 11
# Integer side of the fast exp (round 2): scaled shared-memory index, pre-compensated table, min / max range checks — 4 instead
# of 8 non-fp64 instructions per call (profiles/r02_issue_probe.txt: instructions of other pipes are NOT free next to a
# saturated FP64 pipe on B200).  PRB_EXP_LEAN=0 restores the round-1 form.
def exp_lean() -> int:
    """0: round-1 integer side; 1: pre-compensated table, 2^k applied to the table value BEFORE the final FMA (3 integer
    instructions per call, but the first consumer of the shared-memory load is an early integer instruction: the load latency
    is exposed); 2: plain table, the exponent bits (n & ~2047) << 9 are prepared early and added to the FMA RESULT (4 integer
    instructions, the load's first consumer is the final FMA behind the polynomial)."""
    return int(os.environ.get("PRB_EXP_LEAN", "1"))


def exp_tables(zmax: float = 700.0) -> str:
    """Constants of prb_fast_exp / prb_spec_exp: 2^(j/2048) rounded from extended precision, 2048/ln2, and -ln2/2048 split
    into a high part whose LOW WORD IS ZERO (21 significant bits: n * hi is exact for |x| <= 700, and the value is an
    instruction immediate instead of a constant-bank load) and a low part.  The degree-3 polynomial r + r^2/2 + c3 r^3 on
    |r| <= ln2/4096 truncates at 0.3 ulp; c3 is 1/6 cut to its high word (the cubic term is < 1e-12, so 2e-7 relative on its
    coefficient is invisible) — again an immediate."""
    n = 1 << ETAB_BITS
    ld = np.longdouble
    tab = [float(np.exp2(ld(j) / ld(n))) for j in range(n)]
    ln2 = np.log(ld(2))
    c = -ln2 / ld(n)

    def high_word(x: float) -> float:
        return float(np.array([np.float64(x).view(np.uint64) & ~np.uint64(0xFFFFFFFF)], dtype=np.uint64).view(np.float64)[0])
    hi = high_word(float(c))
    lo = float(c - ld(hi))
    if exp_lean() == 1:
        # entry j is stored with its high word lowered by j << (20 - ETAB_BITS): the kernel then scales by 2^k with ONE integer
        # multiply-add  hi' + (n << (20 - ETAB_BITS))  (n = 2048 k + j), instead of shift + mask + add
        raw = np.array(tab, dtype=np.float64).view(np.uint64)
        raw = raw - (np.arange(n, dtype=np.uint64) << np.uint64(32 + 20 - ETAB_BITS))
        tab = [float(v) for v in raw.view(np.float64)]
    # speculative range |z| <= zmax as thresholds on the argument's high word (zmax = 700, or 220 in kernels with batched
    # reciprocals: a product of three 1 + e^z must stay finite)
    zhi = int(np.float64(zmax).view(np.uint64) >> np.uint64(32))
    assert np.float64(zmax).view(np.uint64) & np.uint64(0xFFFFFFFF) == 0
    return (f"#define PRB_EXP_ZHI 0x{zhi:08x}\n"
            f"#define PRB_ETAB_BITS {ETAB_BITS}\n#define PRB_ETAB_N {n}\n#define PRB_EXP_LEAN {exp_lean()}\n"
            f"#define PRB_EXP_PINK0 {1 if os.environ.get('PRB_EXP_PINK0', '1') == '1' else 0}\n"
            f"#define PRB_RANGE_SHIFT {1 if os.environ.get('PRB_RANGE_SHIFT', '0') == '1' else 0}\n"
            f"#define PRB_EXP_ESTRIN {1 if os.environ.get('PRB_EXP_ESTRIN', '1') == '1' else 0}\n"
            f"#define PRB_EXP_HI ({hi!r})\n#define PRB_EXP_C3 ({high_word(1.0 / 6.0)!r})\n"
            f"__device__ __constant__ double PRB_EXP_TAB[{n}] = {{" + ", ".join(repr(v) for v in tab) + "};\n"
            f"__device__ __constant__ double PRB_EXP_K[2] = {{{float(ld(n) / ln2)!r}, {lo!r}}};   "
            f"// {n}/ln2, low part of -ln2/{n}\n")


def read_template(name: str) -> str:
    with open(os.path.join(_TPL, name)) as f:
        return f.read()


@dataclass
class InstKernel:
    source: str
    n_state: int
    n_params: int
    n_slots: int
    n_obs: int
    ring_elems: int          # total ring elements per instance
    ring_init: np.ndarray    # [ring_elems] initial contents (physical layout == logical at head 0)
    ring_rolls: List[int]
    ring_lengths: List[int]
    table_values: np.ndarray  # packed tables
    n_hist: int              # delay-differential models: state elements with a device-side history ...
    hist_depth: int          # ... and (fixed-step solvers) the ring depth in steps
    hist_vars: List[int]
    slot_init: np.ndarray
    param_defaults: np.ndarray
    block: int
    solver: str
    precision: str
    evals_per_step: int
    kernel_name: str = "prb_integrate"


def _lit(v: float, f32: bool) -> str:
    if math.isnan(v):
        return "(real)NAN" if False else ("__int_as_float(0x7fc00000)" if f32 else "__longlong_as_double(0x7ff8000000000000LL)")
    if math.isinf(v):
        s = "-" if v < 0 else ""
        return f"({s}__int_as_float(0x7f800000))" if f32 else f"({s}__longlong_as_double(0x7ff0000000000000LL))"
    if f32:
        r = repr(float(np.float32(v)))
        if "e" not in r and "." not in r:
            r += ".0"
        r += "f"
    else:
        r = repr(float(v))
        if "e" not in r and "." not in r:
            r += ".0"
    return f"({r})" if r.startswith("-") else r


def emit_rhs_body(rhs: ScalarRhs, f32: bool, const_div_to_mul: bool, float_time: bool = False,
                  fast_math: bool = False, spec: bool = False, round_dy_f32: bool = False,
                  slot_consts: Optional[dict] = None, model_f32: bool = False,
                  hoist: Optional[List[str]] = None) -> List[str]:
    """Straight-line statements computing dy[] (+ ring / slot stores) from y[], T.  With ``hoist`` (a list that receives
    statements for prb_thread_init), thread-invariant sub-expressions — literals and swept parameters only, grouped by
    ``reassoc.fold_invariants`` — are computed once per thread into ``T.q[k]`` with exact arithmetic."""
    g = rhs.graph
    ring_off = np.cumsum([0] + [r.rows * r.length for r in rhs.rings])
    tab_off = np.cumsum([0] + [int(np.prod(t.shape)) for t in rhs.tables])
    lines: List[str] = []
    lit_table: List[float] = []          # fast_math: doubles that do not fit an instruction immediate -> __constant__ table

    def lit(v: float) -> str:
        """fp64 instructions take a 32-bit immediate (the high word); any other double costs two UMOV / IMAD.MOV per use
        when written as a literal (ncu: 35 % of the JRC loop's issue slots).  Read from a __constant__ table instead: the
        compiler loads it once into uniform registers (LDCU) or uses the c[bank][offset] operand form."""
        if not fast_math or f32 or not math.isfinite(v) or (np.float64(v).view(np.uint64) & np.uint64(0xFFFFFFFF)) == 0:
            return _lit(v, f32)
        if v not in lit_table:
            lit_table.append(v)
        return f"PRB_LIT[{lit_table.index(v)}]"

    qslot: dict = {}
    invariant: dict = {}
    if hoist is not None:
        from .reassoc import invariant_nodes
        invariant = invariant_nodes(rhs)

    def ref(i: int, want_real: bool = True) -> str:
        op = g.ops[i]
        if i in qslot:
            return f"T.q[{qslot[i]}]"
        if op == "const":
            v = g.const_value(i)
            if g.isint[i]:
                return lit(float(v)) if want_real else (f"({int(v)}LL)" if v < 0 else f"{int(v)}LL")
            return lit(float(v))
        if op == "state":
            return f"y[{g.payload[i]}]"
        if op == "param":
            return f"T.p[{g.payload[i]}]"
        if op == "t":
            if not want_real and float_time:
                raise UnsupportedModelError("the model indexes arrays with the time step; adaptive solvers pass a float time")
            return "((areal)t)" if want_real else "t"
        name = f"v{i}"
        if g.isint[i] and want_real:
            return f"((areal){name})"
        return name

    def const_chain(i: int):
        """fast_math: (base node, factor) such that node i == base * factor through a chain of multiplications /
        divisions by compile-time constants (x * c1 * c2, x / c1 * c2 ...): emitted as ONE multiply by the folded constant
        (reassociation: <= 1 ulp per folded factor)."""
        f = 1.0
        while not g.isint[i] and g.ops[i] in ("mul", "div"):
            x, y_ = g.args[i]
            if g.ops[i] == "mul" and g.is_const(y_) and not g.is_const(x):
                f, i = f * float(g.const_value(y_)), x
            elif g.ops[i] == "mul" and g.is_const(x) and not g.is_const(y_):
                f, i = f * float(g.const_value(x)), y_
            elif g.ops[i] == "div" and g.is_const(y_) and not g.is_const(x) and float(g.const_value(y_)) != 0.0:
                f, i = f / float(g.const_value(y_)), x
            else:
                break
        return i, f

    def safe_divisor(i: int) -> bool:
        if g.ops[i] != "add":
            return False
        x, y_ = g.args[i]
        for c, e in ((x, y_), (y_, x)):
            if g.is_const(c) and g.ops[e] == "exp" and 1e-200 <= float(g.const_value(c)) <= 1e200:
                return True
        return False

    fused_exp: dict = {}            # add node -> (exp node, constant) emitted as one prb_spec_exp_plus
    live = list(rhs.live_nodes())
    if fast_math:
        # nodes that only fed a folded chain are dead now: recompute liveness over the rewritten expressions
        needed, stack = set(), [n for n in rhs.dy] + [n for *_, n in rhs.ring_writes] + [n for _, n in rhs.slot_writes]
        while stack:
            i = stack.pop()
            if i in needed:
                continue
            needed.add(i)
            if hoist is not None and invariant.get(i):
                stack.extend(g.args[i])          # hoisted nodes are emitted as plain expressions of their operands
                continue
            if not g.isint[i] and g.ops[i] in ("mul", "div"):
                base, f = const_chain(i)
                if base != i and math.isfinite(f) and f != 0.0:
                    stack.append(base)
                    continue
            stack.extend(g.args[i])
            if g.ops[i] == "table":
                stack.append(g.payload[i][1])
        live = [i for i in live if i in needed]

    # fusing `c + exp(x)` into one call and an Estrin-form polynomial shorten the dependency chain of a sigmoid by two
    # instructions — measured A/B on the JRC sweep: 199.3 ms (neither, default) / 200.1 (Estrin) / 201.8 (fused) / 207.2 (both)
    if spec and os.environ.get("PRB_EXP_PLUS", "0") == "1":
        uses: dict = {}
        for i in live:
            for x in g.args[i]:
                uses[x] = uses.get(x, 0) + 1
        for n in list(rhs.dy) + [n for *_, n in rhs.ring_writes] + [n for _, n in rhs.slot_writes]:
            uses[n] = uses.get(n, 0) + 1
        for i in live:
            if g.ops[i] == "add" and not g.isint[i]:
                x, y_ = g.args[i]
                for c, e in ((x, y_), (y_, x)):
                    if g.is_const(c) and g.ops[e] == "exp" and uses.get(e, 0) == 1 and abs(float(g.const_value(c))) <= 1e200:
                        fused_exp[i] = (e, float(g.const_value(c)))
        skip_exp = {e for e, _ in fused_exp.values()}
    else:
        skip_exp = set()
    litscaled: dict = {}
    regfirst = bool(spec and hoist is not None and os.environ.get("PRB_MAT_REGFIRST", "1") == "1")
    # Batched reciprocals (Montgomery's trick): k bare reciprocals 1 / d_j whose divisors are `c + exp(.)` share ONE
    # reciprocal of the product — 1 / d_1 = R d_2 d_3 with R = 1 / (d_1 d_2 d_3): 3 (k - 1) multiplies replace k - 1 seeds
    # (MUFU + zeroed low word) and cubic steps, the same number of fp64 instructions and two instructions of other pipes less
    # per saved reciprocal.  The product must not overflow: the speculative range of every exp of such a kernel shrinks to
    # |z| <= 220 (PRB_EXP_ZHI; arguments beyond it take the exact path like any other range violation) and the constants
    # are confined to [1e-30, 1e30], so the product stays inside [1e-90, 5e286], the fast range of the reciprocal.
    # Opt-in (PRB_RCP_BATCH=1).  Measured on the JRC sweep (profiles/r02_tune_lean9.jsonl): 82 instead of 96 non-fp64
    # instructions per step, 18.07 -> 17.93 ms per 1000 steps (17.82 with the Horner polynomial), trajectories within 9e-16,
    # tests/test_gpu_fullsize.py green with it — 0.8-1.4 % where the issue model predicts 2.2 % (the three sigmoid chains
    # of an evaluation now join in one product, which exposes latency).  Not the default: the narrower fast range and the
    # untested breadth (GPU budget of the round) outweigh 1 %.
    rcp_group: dict = {}            # last member (node id) -> [member ids]; every member -> its group's leader
    rcp_member: dict = {}
    if spec and os.environ.get("PRB_RCP_BATCH", "0") == "1":
        def _batchable(i):
            if g.ops[i] != "div" or g.isint[i] or i in (invariant or {}) and invariant.get(i):
                return False
            a0, a1 = g.args[i]
            if not (g.is_const(a0) and float(g.const_value(a0)) == 1.0 and safe_divisor(a1)):
                return False
            c = next(float(g.const_value(c_)) for c_ in g.args[a1] if g.is_const(c_))
            return 1e-30 <= c <= 1e30
        cand = [i for i in live if _batchable(i)]
        users: dict = {}
        for i in live:
            for x in g.args[i]:
                users.setdefault(x, []).append(i)
            if fast_math and not g.isint[i] and g.ops[i] in ("mul", "div"):
                base = const_chain(i)[0]
                if base != i:
                    users.setdefault(base, []).append(i)                  # folded constant chains reference their base
        for k0 in range(0, len(cand), 3):
            grp = cand[k0:k0 + 3]
            if len(grp) < 2:
                continue
            last = max(grp)
            # every consumer of a member must come after the group's last member in emission order, and no member's divisor
            # may depend on another member (divisors are c + exp(state terms): check direct use only)
            if all(u > last for m_ in grp for u in users.get(m_, [])) and not any(m2 in g.args[g.args[m_][1]] for m_ in grp for m2 in grp):
                rcp_group[last] = grp
                for m_ in grp:
                    rcp_member[m_] = last
    lone_mul: set = set()            # register products that START a sum of literal-coefficient terms: kept a plain multiply
    if regfirst:
        def _litscaled(n):
            if g.isint[n] or g.ops[n] not in ("mul", "div"):
                return False
            base, f = const_chain(n)
            return base != n and math.isfinite(f) and f != 0.0
        for i in live:
            if g.ops[i] == "add" and not g.isint[i]:
                for x, o in ((g.args[i][0], g.args[i][1]), (g.args[i][1], g.args[i][0])):
                    if _litscaled(x) and g.ops[o] == "mul" and not _litscaled(o) and \
                            (invariant.get(g.args[o][0]) or invariant.get(g.args[o][1])):
                        lone_mul.add(o)
    for i in live:
        op, a = g.ops[i], g.args[i]
        if op in ("const", "state", "param", "t") or i in skip_exp:
            continue
        if i in rcp_member:
            if i not in rcp_group:
                continue                                                  # emitted with the group's last member
            grp = rcp_group[i]
            d = [ref(g.args[m_][1]) for m_ in grp]
            if len(grp) == 2:
                lines += [f"const areal w{i}p = __dmul_rn({d[0]}, {d[1]});", f"const areal w{i}r = prb_spec_rcp_nc(w{i}p);",
                          f"const areal v{grp[0]} = __dmul_rn(w{i}r, {d[1]});", f"const areal v{grp[1]} = __dmul_rn(w{i}r, {d[0]});"]
            else:
                lines += [f"const areal w{i}a = __dmul_rn({d[0]}, {d[1]});", f"const areal w{i}p = __dmul_rn(w{i}a, {d[2]});",
                          f"const areal w{i}r = prb_spec_rcp_nc(w{i}p);", f"const areal v{grp[2]} = __dmul_rn(w{i}r, w{i}a);",
                          f"const areal w{i}t = __dmul_rn(w{i}r, {d[2]});", f"const areal v{grp[0]} = __dmul_rn(w{i}t, {d[1]});",
                          f"const areal v{grp[1]} = __dmul_rn(w{i}t, {d[0]});"]
            continue
        if i in fused_exp:
            e, c = fused_exp[i]
            lines.append(f"const areal v{i} = prb_spec_exp_plus({ref(g.args[e][0])}, {lit(c)}, T.ek0, T.etab_s, bad);")
            continue
        if hoist is not None and invariant.get(i) and not g.isint[i]:
            # thread-invariant: once per thread, exact arithmetic (prb_thread_init)
            plain = {"add": "{0} + {1}", "sub": "{0} - {1}", "mul": "{0} * {1}", "div": "{0} / {1}", "neg": "-{0}",
                     "pow": "pow({0}, {1})", "maximum": "prb_max({0}, {1})", "minimum": "prb_min({0}, {1})",
                     "sign": "prb_sign({0})"}
            fmt = plain.get(op) or (_FUNC[op] + "({0})" if op in _FUNC else None)
            if fmt is not None:
                qslot[i] = len(qslot)
                hoist.append(f"T.q[{qslot[i]}] = " + fmt.format(*[ref(x) for x in a]) + ";")
                continue
        if op == "hist":
            # delayed state: the query time follows NumPy's promotion in the reference-generated text (base_backend.py:442-444)
            hv, scale, dkey, strong = g.payload[i]
            if isinstance(dkey, tuple):                                 # swept delay: per-instance parameter
                d, df = f"(double)T.p[{dkey[1]}]", f"(float)T.p[{dkey[1]}]"
            else:
                d = repr(float.fromhex(dkey))
                df = repr(float(np.float32(float.fromhex(dkey)))) + "f"
            if scale is not None:
                tq = f"((double)t * {float(scale)!r} - {d})"            # `t*dt - d`: int step * Python float -> float64
            elif strong and model_f32:
                tq = f"(double)((float)t - {df})"                       # Python-float t - float32 array: a float32 operation
            else:
                tq = f"((double)t - {d})"
            lines.append(f"const areal v{i} = (areal)prb_hist(T, {hv}, {tq});")
            continue
        if fast_math and not g.isint[i] and op in ("mul", "div"):
            base, f = const_chain(i)
            if f32:
                f = float(np.float32(f))
            if base != i and math.isfinite(f) and f != 0.0:
                lines.append(f"const areal v{i} = {ref(base)} * {lit(f)};")
                litscaled[i] = (base, f)
                continue
        if g.isint[i]:
            x, y_ = (ref(a[0], False), ref(a[1], False) if len(a) > 1 else None)
            expr = {"add": f"{x} + {y_}", "sub": f"{x} - {y_}", "mul": f"{x} * {y_}", "neg": f"-{x}"}[op]
            lines.append(f"const long long v{i} = {expr};")
            continue
        if op == "slot" and slot_consts and g.payload[i] in slot_consts:
            expr = lit(slot_consts[g.payload[i]])
        elif op == "slot":
            expr = f"T.slots[{g.payload[i]}LL * T.B + T.b]"
        elif op == "table":
            tab, tnode, col, ncols = g.payload[i]
            expr = f"T.tables[{int(tab_off[tab])}LL + {ref(tnode, False)} * {ncols}LL + {col}LL]"
        elif op == "interp":
            tx, tf, n_pts = g.payload[i]
            expr = f"prb_interp(T.tables + {int(tab_off[tx])}LL, T.tables + {int(tab_off[tf])}LL, {n_pts}, {ref(a[0])})"
        elif op == "ring":
            ring, row, col, shift = g.payload[i]
            L = rhs.rings[ring].length
            expr = (f"T.rings[({int(ring_off[ring]) + row * L}LL + prb_ring_pos(T.head[{ring}], {col - shift}, {L})) "
                    f"* T.B + T.b]")
        elif op == "add" and regfirst and (a[1] in litscaled or a[0] in litscaled):
            # explicit contraction: the literal-coefficient product joins by FMA (immediate operand, two register reads); left
            # to the compiler, `q*s + y*c` may fuse the register product instead and multiply y*c separately (three reads)
            x, o = (a[1], a[0]) if a[1] in litscaled else (a[0], a[1])
            expr = f"fma({ref(litscaled[x][0])}, {lit(litscaled[x][1])}, {ref(o)})"
        elif op == "mul" and i in lone_mul:
            expr = f"__dmul_rn({ref(a[0])}, {ref(a[1])})"
        elif op == "add":
            expr = f"{ref(a[0])} + {ref(a[1])}"
        elif op == "sub":
            expr = f"{ref(a[0])} - {ref(a[1])}"
        elif op == "mul":
            expr = f"{ref(a[0])} * {ref(a[1])}"
        elif op == "div":
            if const_div_to_mul and g.is_const(a[1]) and float(g.const_value(a[1])) != 0.0:
                c = float(g.const_value(a[1]))
                inv = float(np.float32(1.0) / np.float32(c)) if f32 else 1.0 / c
                expr = f"{ref(a[0])} * {lit(inv)}"
            elif fast_math and f32:
                expr = f"__fdividef({ref(a[0])}, {ref(a[1])})"          # rcp.approx + mul: 2 ulp (|divisor| < 2^126)
            elif fast_math and not f32 and spec and g.is_const(a[0]) and float(g.const_value(a[0])) == 1.0:
                # bare reciprocal: no final multiply; no range check when the divisor is c + exp(.) with a moderate c > 0
                # (the exp's own check already confines it to [c, c + e^700], inside the reciprocal's fast range)
                expr = f"prb_spec_rcp_nc({ref(a[1])})" if safe_divisor(a[1]) else f"prb_spec_rcp({ref(a[1])}, bad)"
            elif fast_math and not f32:
                expr = f"prb_spec_div({ref(a[0])}, {ref(a[1])}, bad)" if spec else f"prb_fast_div({ref(a[0])}, {ref(a[1])})"
            else:
                expr = f"{ref(a[0])} / {ref(a[1])}"
        elif op == "neg":
            expr = f"-{ref(a[0])}"
        elif op == "pow":
            expr = f"pow({ref(a[0])}, {ref(a[1])})"
        elif op == "maximum":
            expr = f"prb_max({ref(a[0])}, {ref(a[1])})"
        elif op == "minimum":
            expr = f"prb_min({ref(a[0])}, {ref(a[1])})"
        elif op == "sign":
            expr = f"prb_sign({ref(a[0])})"
        elif op == "exp" and fast_math and f32:
            # ex2.approx(x * log2 e): ~2 ulp + |x| 2^-24.  The flush-to-zero form (default) drops __expf's handling of results
            # below 2^-126 — a compare and two predicated multiplies per call, 36 of the 208 instructions of a JRC RK4 step
            expr = (f"prb_expf_ftz({ref(a[0])})" if os.environ.get("PRB_EXPF_FTZ", "1") == "1" else f"__expf({ref(a[0])})")
        elif op == "exp2":                                               # reassoc.py: exp(x) of float32 fast_math kernels
            expr = f"prb_ex2_ftz({ref(a[0])})" if f32 else f"exp2({ref(a[0])})"
        elif op == "exp" and fast_math and not f32:
            expr = f"prb_spec_exp({ref(a[0])}, T.ek0, T.etab_s, bad)" if spec else f"prb_fast_exp({ref(a[0])}, T.etab)"
        elif op in _FUNC:
            expr = f"{_FUNC[op]}({ref(a[0])})"
        else:
            raise NotImplementedError(f"no CUDA lowering for op `{op}`")
        lines.append(f"const areal v{i} = {expr};")
    for j, n in enumerate(rhs.dy):
        lines.append(f"dy[{j}] = (areal)(float)({ref(n)});" if round_dy_f32 else f"dy[{j}] = {ref(n)};")
    for ring, row, col, node in rhs.ring_writes:
        spec = rhs.rings[ring]
        L = spec.length
        lines.append(f"T.rings[({int(ring_off[ring]) + row * L}LL + prb_ring_pos(T.head[{ring}], "
                     f"{col - spec.rolls_per_eval}, {L})) * T.B + T.b] = {ref(node)};")
    for slot, node in rhs.slot_writes:
        lines.append(f"T.slots[{slot}LL * T.B + T.b] = {ref(node)};")
    for k, r in enumerate(rhs.rings):
        if r.rolls_per_eval:
            lines.append(f"T.head[{k}] = (T.head[{k}] + {r.rolls_per_eval % r.length}) % {r.length};")
    if lit_table:
        lines.insert(0, "// PRB_LIT = {" + ", ".join(repr(v) for v in lit_table) + "}")
    return lines


def emit_inst_kernel(rhs: ScalarRhs, *, solver: str, precision: str, observers: Sequence[int],
                     block: int = 128, const_div_to_mul: bool = False, min_blocks: int = 1,
                     fast_math: bool = False, rk_method: str = "RK45", reduce: bool = False,
                     hist_max_delay: Optional[float] = None, y_smem: bool = False, burst: Optional[bool] = None) -> InstKernel:
    """Generate the full CUDA C source of the fused rhs+solver instance kernel.  ``reduce``: the kernel keeps running
    statistics of the observers (mean, var, min, max, last) instead of storing every sample (fixed-step solvers).
    ``burst``: time loop in bursts between stored samples (32-bit counter, nothing else in the loop) — the default for
    kernels with per-instance parameters (sweeps: many warps, issue-bound); single runs keep the plain per-step loop."""
    if solver not in SOLVER_IDS:
        raise ValueError(f"unknown solver `{solver}`; supported: {sorted(SOLVER_IDS)}")
    f32 = precision == "float32"
    ns = rhs.n_state
    nrings = len(rhs.rings)
    adaptive = solver in ADAPTIVE_SOLVERS
    slot_consts: dict = {}
    if adaptive and rhs.slots and not rhs.rings:
        # mutable buffers that are fully rewritten before they are read in the same evaluation (the vectorised frontend
        # routes inputs through `u[target_idx] = interp(...)`) carry no state between evaluations: plain temporaries
        g_, seen, stack = rhs.graph, set(), list(rhs.dy)
        while stack:
            i = stack.pop()
            if i not in seen:
                seen.add(i)
                stack.extend(g_.args[i])
        written = {slot for slot, _ in rhs.slot_writes}
        read_first = {g_.payload[i] for i in seen if g_.ops[i] == "slot"}        # elements read before any write
        if not (read_first & written):
            # elements that are read but never written are constants (their initial value); the rest are temporaries
            import copy
            slot_consts = {k: float(rhs.slots[k][2]) for k in read_first}
            rhs = copy.copy(rhs)
            rhs.slots, rhs.slot_writes = [], []
    n_hist = len(rhs.hist_vars)
    hist_depth = 0
    if n_hist:
        if adaptive != (solver == "dopri5_dde") or adaptive != (rhs.hist_dt is None):
            raise UnsupportedModelError("delay-differential models: fixed-step solvers need the `hist(t*dt - d)` form, "
                                        "solver='scipy' the `hist(t - d)` form of the vector field (rebuild for the solver)")
        if not adaptive:
            dmax = max(rhs.hist_delays)
            if rhs.hist_delay_params:
                # swept delays: the ring must hold the longest delay of the sweep (the caller's bound on the parameter table)
                if hist_max_delay is None:
                    raise UnsupportedModelError("a swept history delay needs `hist_max_delay` (upper bound of the swept values)")
                dmax = max(dmax, float(hist_max_delay))
            need = int(math.ceil(dmax / rhs.hist_dt)) + 3
            hist_depth = 1 << max(2, (need - 1).bit_length())            # power of two: slot = step & (depth - 1)
    elif solver == "dopri5_dde":
        raise UnsupportedModelError("solver 'dopri5_dde' is the adaptive path of delay-differential models")
    step_tables = any(rhs.graph.ops[i] == "table" for i in rhs.live_nodes())
    if adaptive and (rhs.rings or step_tables or rhs.slots):
        raise UnsupportedModelError("adaptive integration (solver='scipy') needs a pure vector field: delay ring buffers, "
                                    "step-indexed input tables and mutable buffers are fixed-step constructs")
    pure = not (rhs.rings or rhs.slots or rhs.ring_writes or rhs.slot_writes)
    spec = bool(fast_math and not f32 and pure)
    if adaptive and f32:
        fast_math = False
    # adaptive + float32 model: SciPy integrates the state in float64 and hands float64 `y` to the generated function, so the
    # reference's arithmetic is float64 on float32-valued constants with the result rounded into the float32 `dy` buffer
    # (base.py check_arguments, computegraph.py:1220-1228).  A pure-float32 rhs is too noisy for the error estimator (the
    # reference's own test_3_2_qif_theta then fails with "step size too small"); reproduce the reference instead.
    arith64 = adaptive and f32
    body_f32 = f32 and not arith64
    hoist: Optional[List[str]] = None
    fast_rhs = rhs
    if (spec or (fast_math and f32 and pure)) and not adaptive and rhs.n_out is None:
        # FP64-pipe-bound sweeps: group thread-invariant factors (swept parameters x literals) and compute them once per thread
        from .reassoc import fold_invariants
        fast_rhs, hoist = fold_invariants(rhs), []
    body = emit_rhs_body(fast_rhs, body_f32, const_div_to_mul, float_time=adaptive, fast_math=fast_math, spec=spec,
                         round_dy_f32=arith64, slot_consts=slot_consts, model_f32=f32, hoist=hoist)
    exact_body = emit_rhs_body(rhs, body_f32, const_div_to_mul, float_time=adaptive, round_dy_f32=arith64,
                               slot_consts=slot_consts, model_f32=f32) if spec else None
    spec = spec and any("prb_spec_" in l for l in body)
    use_etab = any("prb_fast_exp(" in l or "prb_spec_exp(" in l or "prb_spec_exp_plus(" in l for l in body)
    np_ = len(rhs.params)
    src: List[str] = []
    src.append("// Generated by pyrates_b200.emit_inst — do not edit.  Instance execution model (1 thread = 1 instance).")
    src.append(read_template("prb_kernel_abi.cuh"))
    src.append(f"typedef {'float' if f32 else 'double'} real;")
    src.append(f"typedef {'double' if (adaptive and f32) else 'real'} areal;        // arithmetic type of the generated rhs")
    src.append(f"typedef {'double' if adaptive else 'long long'} prb_time;     // adaptive solvers hand the rhs a float time")
    if (bool(rhs.params) if burst is None else burst) and not adaptive:
        src.append("#define PRB_BURST_LOOP 1")
    src.append(f"#define PRB_NS {ns}")
    src.append(f"#define PRB_NDY {rhs.n_out if rhs.n_out is not None else ns}        // outputs of prb_rhs (Jacobian programs: n*n)")
    if rhs.n_out is not None:
        if rhs.rings or rhs.slots:
            raise UnsupportedModelError("Jacobian functions of models with delay ring buffers / mutable buffers")
        src.append("#define PRB_EVAL_ONLY 1               // no integrator: only the single-evaluation kernel")
    src.append(f"#define PRB_NP {np_}")
    src.append(f"#define PRB_NOBS {len(observers)}")
    src.append(f"#define PRB_NRINGS {nrings}")
    src.append(f"#define PRB_SOLVER {SOLVER_IDS[solver]}")
    src.append(f"#define PRB_BLOCK {block}")
    src.append(f"#define PRB_MIN_BLOCKS {int(min_blocks)}")
    if y_smem:
        if solver != "rk4" or precision != "float64" or reduce or n_hist:
            raise ValueError("y_smem is implemented for the float64 fixed-step rk4 kernels without reduce / history")
        src.append("#define PRB_Y_SMEM 1")
    if reduce:
        if adaptive:
            raise UnsupportedModelError("on-device observer reduction is implemented for the fixed-step solvers")
        src.append("#define PRB_REDUCE 1")
    src.append(f"#define PRB_NQ {len(hoist) if hoist else 0}")
    src.append(f"#define PRB_NH {n_hist}")
    src.append(f"#define PRB_HDEPTH {hist_depth}")
    if adaptive and solver != "dopri5_dde":
        if rk_method == "DOP853":
            from .dop853 import cuda_tables as dop853_tables
            src.append(dop853_tables())
        else:
            from .rk45 import cuda_method_tables
            src.append(cuda_method_tables(rk_method))
    if ns > 64:
        src.append("#define PRB_UNROLL_NS")        # big states: keep loops rolled (state spills to local memory)
    src.append("""
struct PrbThread {
    long long b, B;
    real p[PRB_NP > 0 ? PRB_NP : 1];
    const real* __restrict__ pbase;       // params[0][0]: the exact fallback of fast_math kernels re-reads its parameters
    areal q[PRB_NQ > 0 ? PRB_NQ : 1];     // thread-invariant sub-expressions (fast_math: reassoc.fold_invariants)
    int head[PRB_NRINGS > 0 ? PRB_NRINGS : 1];
    real* __restrict__ slots;
    real* __restrict__ rings;
    const real* __restrict__ tables;
    const double* etab;            // shared-memory copy of PRB_EXP_TAB (fast_math kernels)
    unsigned int etab_s;           // the same as a 32-bit shared-window address (ld.shared without S2UR per use)
    double ek0;                    // 2048/ln2, pinned in a register pair (prb_spec_exp)
    unsigned int ysm_s;            // PRB_Y_SMEM: shared-window address of this thread's column of the base state
#if PRB_NH > 0
#if PRB_HDEPTH > 0                 // fixed-step history: ring [PRB_HDEPTH][PRB_NH][B] of the states after steps 0..hn
    real* __restrict__ hist;
    long long hn;                  // latest stored step
    double hdt;                    // step size of the history times (i + 1) * dt
#else                              // adaptive history: ring of (time, state) entries [cap][B] / [cap][PRB_NH][B]
    double* __restrict__ ht;
    double* __restrict__ hy;
    long long hn, hcap;            // entries so far (entry 0 = the initial condition), ring capacity
    double ht0, hy0[PRB_NH];       // entry 0 is never evicted: the pre-history
    bool hovf;                     // a lookup needed an entry that was already overwritten
#endif
#endif
};
#if PRB_NH > 0
// Delayed state for delay-differential models — the device-side counterpart of DDEHistory.__call__
// (pyrates/backend/base/base_backend.py:165-177): the initial condition for t <= t0, the latest stored state for
// t >= the last stored time, else linear interpolation between the two stored states that bracket t
// (idx = bisect_right(times, t) - 1).
#if PRB_HDEPTH > 0
// fixed-step solvers: the history holds the state after every step, at times T_0 = 0, T_j = j * dt
// (BaseBackend._solve_euler / _solve_heun: args[0].update((i + 1) * dt, y), :737-738 / :766-767)
__device__ __forceinline__ real prb_hist(const PrbThread& T, const int k, const double tq) {
    const long long n = T.hn;
    const double dt = T.hdt;
    const real* __restrict__ H = T.hist + (long long)k * T.B;
    const long long stride = (long long)PRB_NH * T.B;
    if (tq <= 0.0) return H[0];                                             // slot of step 0 (depth > max delay / dt + 2)
    if (tq >= (double)n * dt) return H[(n & (PRB_HDEPTH - 1)) * stride];
    long long j = (long long)(tq / dt);
    j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
    while (j + 1 < n && (double)(j + 1) * dt <= tq) ++j;                    // bisect_right(times, tq) - 1 on T_j = fl(j * dt)
    while (j > 0 && (double)j * dt > tq) --j;
    const double t0 = (double)j * dt, t1 = (double)(j + 1) * dt;
    const double alpha = (tq - t0) / (t1 - t0);
    const real a = H[(j & (PRB_HDEPTH - 1)) * stride], c = H[((j + 1) & (PRB_HDEPTH - 1)) * stride];
    return a + (real)alpha * (c - a);
}
#else
__device__ __forceinline__ double prb_hist(PrbThread& T, const int k, const double tq) {
    if (tq <= T.ht0) return T.hy0[k];
    const long long n = T.hn, last = n - 1;
    const long long sy = (long long)PRB_NH * T.B;
    const double* __restrict__ HY = T.hy + (long long)k * T.B;
    if (tq >= T.ht[(last % T.hcap) * T.B]) return HY[(last % T.hcap) * sy];
    long long a = n > T.hcap ? n - T.hcap : 0, c = last;
    if (T.ht[(a % T.hcap) * T.B] > tq) { T.hovf = true; return HY[(a % T.hcap) * sy]; }
    while (c - a > 1) {                                                     // invariant: time[a] <= tq < time[c]
        const long long m = (a + c) >> 1;
        if (T.ht[(m % T.hcap) * T.B] <= tq) a = m; else c = m;
    }
    const double t0 = T.ht[(a % T.hcap) * T.B], t1 = T.ht[((a + 1) % T.hcap) * T.B];
    const double alpha = (tq - t0) / (t1 - t0);
    const double y0 = HY[(a % T.hcap) * sy], y1 = HY[((a + 1) % T.hcap) * sy];
    return y0 + alpha * (y1 - y0);
}
#endif
#endif
// ---- fast_math (fp64): cheaper exp / division for FP64-pipe-bound kernels (the JRC sweep spends ~70 % of its fp64
// instructions in 3 sigmoids per evaluation).  exp: x = (2048 k + j) ln2/2048 + r, |r| <= ln2/4096,
// e^x = 2^k * 2^(j/2048) * (1 + r + c2 r^2 + r^3/6): 8 fp64 instructions + one shared-memory table read instead of ~20;
// <= 2.1 ulp for |x| <= 700 (validated against long-double exp on the host, tests/test_compile_inst.py), library exp
// outside.  Division: rcp.approx.ftz.f64 + two Newton steps (5 fp64 instructions), IEEE division for zero / denormal /
// huge / non-finite divisors.  Both stay far inside the 1e-9 trajectory tolerance; off by default in CudaBackend.run.
@@EXP_TABLES@@
__device__ __forceinline__ float prb_ex2_ftz(const float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// float32 fast_math exp: one multiply + MUFU.EX2; results below 2^-126 flush to zero (absolute error < 1.2e-38)
__device__ __forceinline__ float prb_expf_ftz(const float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
__device__ __noinline__ double prb_exp_slow(const double z) { return exp(z); }
__device__ __noinline__ double prb_div_slow(const double a, const double b) { return a / b; }
// Range violations of one speculative evaluation (or step).  exp arguments are checked through the running signed AND unsigned
// maxima of their high words — |z| <= 700 and not NaN  <=>  hi <= 0x4085e000 as a signed int (covers z >= 0; every negative z
// passes) and hi <= 0xc085e000 as an unsigned int (covers z < 0; every positive z passes) — i.e. two 3-input min/max
// instructions per three calls instead of one mask per call plus the maximum.
struct PrbBad {
    int smax; unsigned int umax; bool flag;
    __device__ __forceinline__ PrbBad() : smax((int)0x80000000), umax(0u), flag(false) {}
    __device__ __forceinline__ PrbBad& operator|=(const bool b) { flag |= b; return *this; }
#if PRB_RANGE_SHIFT
    // variant: drop the sign with a shift (an IMAD.SHL, nearly free next to the FP64 pipe) and keep ONE unsigned maximum
    __device__ __forceinline__ void exp_arg(const int hi) { umax = max(umax, (unsigned int)hi << 1); }
    __device__ __forceinline__ operator bool() const { return flag | (umax > ((unsigned int)PRB_EXP_ZHI << 1)); }
#else
    __device__ __forceinline__ void exp_arg(const int hi) { smax = max(smax, hi); umax = max(umax, (unsigned int)hi); }
    __device__ __forceinline__ operator bool() const { return flag | (smax > (int)PRB_EXP_ZHI) | (umax > (0x80000000u | PRB_EXP_ZHI)); }
#endif
};
#if PRB_EXP_LEAN
// shared-memory copy of the table at file scope: the accessors see a __shared__ array
__shared__ double prb_etab[PRB_ETAB_N + 2];
#define PRB_ETAB_FILE_SCOPE 1
#endif
#if PRB_EXP_LEAN == 1
// the entry — stored with its high word lowered by j << 9 — is scaled by 2^k with one integer multiply-add:
// hi' + (n << 9) = hi(2^(j/2048)) + (k << 20) for n = 2048 k + j
__device__ __forceinline__ double prb_etab_scaled(const int n) {
    const double tj = prb_etab[n & (PRB_ETAB_N - 1)];
    return __hiloint2double(__double2hiint(tj) + (n << (20 - PRB_ETAB_BITS)), __double2loint(tj));
}
#define PRB_ETAB_SCALED(tab, n) prb_etab_scaled(n)
// 2^k T (1 + q) + c
__device__ __forceinline__ double prb_exp_finish(const int n, const double q, const double c, const bool plus) {
    const double ts = prb_etab_scaled(n);
    return plus ? fma(ts, q, ts + c) : fma(ts, q, ts);
}
#elif PRB_EXP_LEAN == 2
// plain table; the exponent bits are prepared from n alone (no dependence on the load) and join AFTER the FMA, so the load's
// first consumer sits behind the polynomial
__device__ __forceinline__ double prb_exp_finish(const int n, const double q, const double c, const bool plus) {
    const double tj = prb_etab[n & (PRB_ETAB_N - 1)];
    const int kb = n & ~(PRB_ETAB_N - 1);
    const double e = fma(tj, q, tj);
    const double es = __hiloint2double(__double2hiint(e) + (kb << (20 - PRB_ETAB_BITS)), __double2loint(e));
    return plus ? es + c : es;
}
#else
__device__ __forceinline__ double prb_etab_scaled_s(const unsigned int tab, const int n) {
    double tj;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(tab + ((unsigned int)(n & (PRB_ETAB_N - 1)) << 3)));
    return __hiloint2double(__double2hiint(tj) + ((n >> PRB_ETAB_BITS) << 20), __double2loint(tj));
}
#define PRB_ETAB_SCALED(tab, n) prb_etab_scaled_s(tab, n)
#endif
// q(r) = r + r^2/2 + c3 r^3 on |r| <= ln2/4096
__device__ __forceinline__ double prb_exp_poly(const double r) {
#if PRB_EXP_ESTRIN
    const double r2 = r * r;                                                 // Estrin: q = r + r^2 (1/2 + c3 r), two levels deep
    return fma(r2, fma(r, PRB_EXP_C3, 0.5), r);
#else
    double q = fma(r, PRB_EXP_C3, 0.5);                                      // every coefficient an instruction immediate
    q = fma(q, r, 1.0);
    return q * r;
#endif
}
__device__ __forceinline__ double prb_fast_exp(const double z, const double* __restrict__ tab) {
    if (!(fabs(z) <= 700.0)) return prb_exp_slow(z);
    const double t = fma(z, PRB_EXP_K[0], 6755399441055744.0);               // 2048/ln2, 1.5 * 2^52 (an immediate)
    const int n = __double2loint(t);
    const double nd = t - 6755399441055744.0;
    double r = fma(nd, PRB_EXP_HI, z);                                       // -ln2/2048, high word (an immediate)
    r = fma(nd, PRB_EXP_K[1], r);                                            //            low part
    const double q = prb_exp_poly(r);
#if PRB_EXP_LEAN
    return prb_exp_finish(n, q, 0.0, false);                                 // 2^k * 2^(j/2048) (1 + q): |x| <= 700 keeps it normal
#else
    const double tj = tab[n & (PRB_ETAB_N - 1)];
    const double e = fma(tj, q, tj);
    return __hiloint2double(__double2hiint(e) + ((n >> PRB_ETAB_BITS) << 20), __double2loint(e));
#endif
}
// speculative variants: no branch per call — the caller accumulates the range violations of one rhs evaluation in `bad` and
// re-evaluates the whole (pure) vector field with the library functions in the rare case an argument left the fast range.
// The range checks are integer work on the high word (an fp64 compare would occupy the FP64 pipe the kernel is bound by).
// k0 = 2048/ln2 from a register the caller pins for the whole launch (T.ek0): as a __constant__ read the compiler keeps it in a
// uniform register and copies it into a register pair before EVERY use (the multiply-add already has an immediate operand)
__device__ __forceinline__ double prb_spec_exp(const double z, const double k0, const unsigned int tab, PrbBad& bad) {
#if PRB_EXP_LEAN
    bad.exp_arg(__double2hiint(z));                                          // |z| > 700 (or NaN)
#else
    bad |= (unsigned int)(__double2hiint(z) & 0x7fffffff) > (unsigned int)PRB_EXP_ZHI;
#endif
    const double t = fma(z, k0, 6755399441055744.0);
    const int n = __double2loint(t);
    const double nd = t - 6755399441055744.0;                                // (an I2F on the conversion pipe instead was measured: no gain)
    double r = fma(nd, PRB_EXP_HI, z);
    r = fma(nd, PRB_EXP_K[1], r);
    const double q = prb_exp_poly(r);
#if PRB_EXP_LEAN
    return prb_exp_finish(n, q, 0.0, false);
#else
    double tj;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(tab + ((unsigned int)(n & (PRB_ETAB_N - 1)) << 3)));
    const double e = fma(tj, q, tj);
    return __hiloint2double(__double2hiint(e) + ((n >> PRB_ETAB_BITS) << 20), __double2loint(e));
#endif
}
// c + e^z in one go (sigmoid denominators): the table value is scaled by 2^k first (integer work, off the critical path), so
// the sum joins the final FMA: c + T (1 + q) = fma(T, q, T + c) — one dependent instruction less than exp followed by an add
__device__ __forceinline__ double prb_spec_exp_plus(const double z, const double c, const double k0, const unsigned int tab, PrbBad& bad) {
#if PRB_EXP_LEAN
    bad.exp_arg(__double2hiint(z));
#else
    bad |= (unsigned int)(__double2hiint(z) & 0x7fffffff) > (unsigned int)PRB_EXP_ZHI;
#endif
    const double t = fma(z, k0, 6755399441055744.0);
    const int n = __double2loint(t);
    const double nd = t - 6755399441055744.0;
#if PRB_EXP_LEAN
    double r = fma(nd, PRB_EXP_HI, z);
    r = fma(nd, PRB_EXP_K[1], r);
    return prb_exp_finish(n, prb_exp_poly(r), c, true);
#else
    const double ts = PRB_ETAB_SCALED(tab, n);
    const double s = ts + c;
    double r = fma(nd, PRB_EXP_HI, z);
    r = fma(nd, PRB_EXP_K[1], r);
    return fma(ts, prb_exp_poly(r), s);
#endif
}
__device__ __forceinline__ double prb_spec_div(const double a, const double b, PrbBad& bad) {
    bad |= (((unsigned int)__double2hiint(b) & 0x7ff00000u) - 0x04000000u) >= 0x77c00000u;   // |b| outside ~[1e-289, 1e288]
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const double e = fma(-b, r, 1.0);                  // r0 = (1 - e) / b with |e| <= ~2^-20
    r = fma(r, fma(e, e, e), r);                       // cubic step: r0 (1 + e + e^2) = (1 - e^3) / b
    return a * r;
}
__device__ __forceinline__ double prb_spec_rcp_nc(const double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const double e = fma(-b, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double prb_spec_rcp(const double b, PrbBad& bad) {
    bad |= (((unsigned int)__double2hiint(b) & 0x7ff00000u) - 0x04000000u) >= 0x77c00000u;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const double e = fma(-b, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double prb_fast_div(const double a, const double b) {
    const double ab = fabs(b);
    if (!(ab > 1e-290 && ab < 1e290)) return prb_div_slow(a, b);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    return a * r;
}
__device__ __forceinline__ areal prb_max(areal a, areal b) { return (a != a || b != b) ? a + b : (a > b ? a : b); }
__device__ __forceinline__ areal prb_min(areal a, areal b) { return (a != a || b != b) ? a + b : (a < b ? a : b); }
__device__ __forceinline__ areal prb_sign(areal a) { return (areal)((a > (areal)0) - (a < (areal)0)); }
// numpy.interp(x, xp, fp) for one x (timed inputs of the adaptive solvers): clamp outside [xp[0], xp[n-1]], else
// slope * (x - xp[j]) + fp[j] with xp[j] <= x < xp[j+1] (the expression NumPy's C loop evaluates)
__device__ __forceinline__ areal prb_interp(const real* __restrict__ xp, const real* __restrict__ fp, const int n, const areal x) {
    if (!(x > (areal)__ldg(xp))) return (areal)__ldg(fp);
    if (x >= (areal)__ldg(xp + n - 1)) return (areal)__ldg(fp + n - 1);
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((areal)__ldg(xp + mid) <= x) lo = mid; else hi = mid;
    }
    const areal x0 = (areal)__ldg(xp + lo), f0 = (areal)__ldg(fp + lo);
    const areal slope = ((areal)__ldg(fp + lo + 1) - f0) / ((areal)__ldg(xp + lo + 1) - x0);
    return slope * (x - x0) + f0;
}
// physical ring position of logical column `col` (may be negative by the pending roll) at head `head`
__device__ __forceinline__ long long prb_ring_pos(int head, int col, int L) {
    int p = col - head;
    p += (p < 0) ? L : 0;
    p += (p < 0) ? L : 0;
    return (long long)p;
}
""".replace("@@EXP_TABLES@@\n", exp_tables(220.0 if any("prb_spec_rcp_nc(w" in l for l in body) else 700.0)))
    if use_etab:
        src.append("#define PRB_USE_ETAB 1")
    if body and body[0].startswith("// PRB_LIT = {"):
        vals = body.pop(0)[len("// PRB_LIT = {"):-1]
        src.append(f"__device__ __constant__ double PRB_LIT[{vals.count(',') + 1}] = {{{vals}}};")
    init = ["__device__ __forceinline__ void prb_thread_init(PrbThread& T, const prb_kernel_args& A, long long b) {",
            "    T.b = b; T.B = A.n_inst; T.etab = nullptr; T.etab_s = 0u; T.ek0 = 0.0; T.pbase = (const real*)A.params;",
            "    T.slots = (real*)A.slots; T.rings = (real*)A.rings; T.tables = (const real*)A.tables;"]
    if np_:
        init.append("    const real* __restrict__ P = (const real*)A.params;")
        init.append("    _Pragma(\"unroll\") for (int k = 0; k < PRB_NP; ++k) T.p[k] = P[(long long)k * T.B + b];")
    for line in hoist or []:
        init.append("    " + line)
    if hoist:
        # keep the invariants in registers: without the barrier the compiler rematerialises cheap ones (q = 100 * q') inside
        # the time loop, i.e. puts the multiplies we just hoisted back on the FP64 pipe (seen in SASS)
        init.append("    _Pragma(\"unroll\") for (int k = 0; k < PRB_NQ; ++k) asm volatile(\"\" : \"+%s\"(T.q[k]));"
                    % ("f" if body_f32 else "d"))
    for k, r in enumerate(rhs.rings):
        init.append(f"    T.head[{k}] = (int)((A.eval0 * {r.rolls_per_eval}LL) % {r.length}LL);")
    if n_hist and not adaptive:
        init.append("    T.hist = (real*)A.work + b; T.hn = A.step0; T.hdt = A.dt;")
    elif n_hist:
        init.append("    const double* __restrict__ O = (const double*)A.consts;")
        init.append("    T.hcap = (long long)O[8]; T.ht = (double*)A.work + b; T.hy = (double*)A.work + T.hcap * T.B + b;")
        init.append("    T.hn = 1; T.ht0 = O[2]; T.hovf = false;")
        init.append("    _Pragma(\"unroll\") for (int k = 0; k < PRB_NH; ++k) { T.hy0[k] = O[9 + k]; T.hy[(long long)k * T.B] = O[9 + k]; }")
        init.append("    T.ht[0] = O[2];")
    init.append("}")
    if n_hist and not adaptive:
        push = ["// DDEHistory.update((i + 1) * dt, y): the state after step n - 1 becomes history entry n",
                "__device__ __forceinline__ void prb_hist_push(PrbThread& T, const real (&y)[PRB_NS], const long long n) {",
                "    real* __restrict__ H = T.hist + (n & (PRB_HDEPTH - 1)) * (long long)PRB_NH * T.B;"]
        push += [f"    H[{k}LL * T.B] = y[{e}];" for k, e in enumerate(rhs.hist_vars)]
        push += ["    T.hn = n;", "}"]
        init.append("\n".join(push))
    elif n_hist:
        push = ["// solout(t, y) -> DDEHistory.update(t, y): one entry per accepted step and per call start",
                "__device__ __forceinline__ void prb_hist_push(PrbThread& T, const double t, const double (&y)[PRB_NS]) {",
                "    const long long s = T.hn % T.hcap;",
                "    T.ht[s * T.B] = t;",
                "    double* __restrict__ H = T.hy + s * (long long)PRB_NH * T.B;"]
        push += [f"    H[{k}LL * T.B] = y[{e}];" for k, e in enumerate(rhs.hist_vars)]
        push += ["    ++T.hn;", "}"]
        init.append("\n".join(push))
    init.append("__device__ __forceinline__ void prb_thread_exit(PrbThread&, const prb_kernel_args&, long long) {}")
    src.append("\n".join(init))
    if spec:
        src.append("__device__ __noinline__ void prb_rhs_exact(const prb_time t, const areal (&y)[PRB_NS], areal (&dy)[PRB_NDY], PrbThread& T) {")
        src.extend("    " + l for l in exact_body)
        src.append("}")
    if spec:
        # speculative evaluation: `bad` is only accumulated; the caller decides when to look at it.  The fixed-step solver
        # template checks it ONCE PER STEP (PRB_SPEC_STEP): the stages of a step then form one basic block (one reconvergence
        # point per step instead of one per evaluation — a branch construct costs ~5 issue cycles next to a saturated FP64
        # pipe, profiles/r02_issue_probe2.txt), and a step that saw an out-of-range argument is recomputed by prb_step_exact.
        # Default since round 2 (PRB_SPEC_PER_STEP=0 restores the per-evaluation check).  History: the round-1 form of this
        # option handed y / yn to the out-of-line exact step by reference, which put both arrays in local memory (8 x
        # STL.128 per step on the FAST path) and copied yn -> y at the join; it was measured slower then.  Now the step
        # returns its increment and both outcomes update y in place (inst_kernel.cuh).
        if os.environ.get("PRB_SPEC_PER_STEP", "1") == "1":
            src.append("#define PRB_SPEC_STEP 1")
        # integrator chains (dy_j == y_b, e.g. dV/dt = I of every second-order synapse): RK4 needs no stage accumulator for
        # such a state — V_new = V + dt I + dt^2/6 (k1 + k2 + k3)_I exactly — which saves one FP64 instruction per chain and
        # step (fast_math sweeps only: algebraically exact, different rounding at the 1e-16 level).  Opt-in
        # (PRB_INTEG_PAIRS=1): measured on the JRC sweep 260 instead of 264 fp64 instructions but 3 more integer ones and 4
        # more three-operand FMAs — 18.16 vs 18.07 ms per 1000 steps, no gain (profiles/r02_tune_lean8.jsonl)
        integ = [(int(fast_rhs.graph.payload[n]) if fast_rhs.graph.ops[n] == "state" else -1) for n in fast_rhs.dy]
        if solver == "rk4" and not f32 and os.environ.get("PRB_INTEG_PAIRS", "0") == "1" and any(b >= 0 for b in integ) \
                and len(integ) == ns and os.environ.get("PRB_SPEC_PER_STEP", "1") == "1":
            src.append("#define PRB_INTEG_PAIRS 1")
            src.append(f"__device__ constexpr int PRB_INTEG[PRB_NS] = {{{', '.join(str(b) for b in integ)}}};")
        if os.environ.get("PRB_BURST_UNROLL", "1") != "1":
            src.append(f"#define PRB_BURST_UNROLL _Pragma(\"unroll {int(os.environ['PRB_BURST_UNROLL'])}\")")
        src.append("__device__ __forceinline__ void prb_rhs_spec(const prb_time t, const areal (&y)[PRB_NS], areal (&dy)[PRB_NDY], PrbThread& T, PrbBad& bad) {")
        src.extend("    " + l for l in body)
        src.append("}")
    src.append("__device__ __forceinline__ void prb_rhs(const prb_time t, const areal (&y)[PRB_NS], areal (&dy)[PRB_NDY], PrbThread& T) {")
    if spec:
        src.append("    PrbBad bad;")
        src.append("    prb_rhs_spec(t, y, dy, T, bad);")
        src.append("    if (bad) {                  // an exp / division argument left the fast range: redo this evaluation exactly")
        src.append("        areal yl[PRB_NS], dl[PRB_NDY];")
        src.append("        PrbThread Tl = T;")
        # the swept parameters are re-read from global memory on this (rare) path: keeping T.p alive across the time loop
        # for it costs 2 registers per parameter in a kernel whose fast path only needs the folded invariants T.q
        src.append("        _Pragma(\"unroll\") for (int k = 0; k < PRB_NP; ++k) Tl.p[k] = T.pbase[(long long)k * T.B + T.b];")
        src.append("        _Pragma(\"unroll\") for (int j = 0; j < PRB_NS; ++j) yl[j] = y[j];")
        src.append("        prb_rhs_exact(t, yl, dl, Tl);")
        src.append("        _Pragma(\"unroll\") for (int j = 0; j < PRB_NDY; ++j) dy[j] = dl[j];")
        src.append("    }")
    else:
        src.extend("    " + l for l in body)
    src.append("}")
    rec = ["__device__ __forceinline__ void prb_record(const real (&y)[PRB_NS], real* __restrict__ row, long long b, long long B) {"]
    for o, idx in enumerate(observers):
        if not 0 <= idx < ns:
            raise ValueError(f"observer index {idx} out of range")
        rec.append(f"    row[{o}LL * B + b] = y[{idx}];")
    rec.append("}")
    src.append("\n".join(rec))
    src.append(read_template("inst_kernel.cuh"))

    ring_init = np.concatenate([r.init.reshape(-1) for r in rhs.rings]) if rhs.rings else np.zeros(0)
    tabs = np.concatenate([np.asarray(t.value, dtype=np.float64).reshape(-1) for t in rhs.tables]) if rhs.tables else np.zeros(0)
    evals = {"euler": 1, "heun": 2, "heun_ref": 2, "heun_true": 2, "rk4": 4, "scipy": 6, "rk45": 6, "dopri5_dde": 6}[solver]
    if adaptive and rk_method == "RK23":
        evals = 3
    if adaptive and rk_method == "DOP853":
        evals = 12
    return InstKernel("\n".join(src) + "\n", ns, np_, len(rhs.slots), len(observers), int(ring_init.size), ring_init,
                      [r.rolls_per_eval for r in rhs.rings], [r.length for r in rhs.rings], tabs,
                      n_hist, hist_depth, list(rhs.hist_vars),
                      np.array([s[2] for s in rhs.slots], dtype=np.float64),
                      np.array([p.default for p in rhs.params], dtype=np.float64), block, solver, precision, evals)
