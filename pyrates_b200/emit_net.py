"""CUDA C emitter for the *network* execution model: phased kernel IR (netcompile.NetIR) -> source.

Generates ``prb_eval<STAGE>()`` — one rhs evaluation as a sequence of phases (full reductions, warp-per-row
reductions with the element-wise tail fused behind them, thread-per-element loops, grid barriers in
between) — and splices it into the hand-written solver template ``csrc/templates/net_kernel.cuh``.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

from .emit_inst import ADAPTIVE_SOLVERS, SOLVER_IDS, _FUNC, _lit, read_template
from .devgen import is_device_generated
from .netcompile import MemArray, NetIR, Reduction, Store
from .scalarize import UnsupportedModelError

__all__ = ["NetKernel", "emit_net_kernel"]

LL_REDUNDANT_MAX = 4096           # multi-GPU: up to this many peer-owned elements per barrier every CTA polls them all itself
WCACHE_MAX_BYTES = 192 * 1024     # dynamic shared memory for W-stationary row groups (227 KB per CTA minus static use)
FULL_TWO_LEVEL = 8192        # full reductions at least this long run in two levels (partials per CTA + grid barrier)
FULL_PART_STRIDE = 256       # partial slots per rank: the persistent grid has one CTA per SM (148 on B200), never more than this


@dataclass
class NetKernel:
    source: str
    n_state: int
    n_obs: int
    work_elems: int            # total `work` buffer elements (user arrays + 6 solver vectors)
    work_user: int
    work_init: np.ndarray
    const_init: np.ndarray
    iconst_init: np.ndarray
    ring_init: np.ndarray
    table_init: np.ndarray
    block: int
    solver: str
    precision: str
    evals_per_step: int
    n_phases: int
    kernel_name: str = "prb_integrate_net"
    smem_bytes: int = 0        # dynamic shared memory (tcgen05 operand ring)
    tc: Optional[dict] = None  # tcgen05 coupling-GEMM configuration, None = SIMT row reductions
    ll_slots: int = 0          # multi-GPU: 16-byte packet slots per parity of the receive mirror (work | rings | heartbeats)
    const_elems: int = 0       # size of the constant space on the device (>= const_init.size when parts are device-generated)
    const_fills: tuple = ()    # (element offset, devgen generator, row_lo, row_hi): slices filled on the device


class _Emitter:
    def __init__(self, ir: NetIR, f32: bool, rank: int = 0, world: int = 1, nb: int = 1):
        self.ir, self.g, self.f32 = ir, ir.graph, f32
        self.rank, self.world, self.nb = rank, world, nb
        self.roles: Dict[int, frozenset] = {}
        self.red_work: Dict[int, MemArray] = {}
        self.ring_index = {name: k for k, name in enumerate(a.name for a in ir.arrays.values() if a.space == "ring")}
        self.pending: List[str] = []      # multi-GPU: unpack calls for the regions stored since the last grid barrier
        self.max_fetch = 0                # largest number of peer-owned elements fetched in front of one barrier

    # ---- multi-GPU exchange bookkeeping (net_kernel.cuh: prb_ll_*): every store another rank must see is sent as an LL
    # packet by its owner; the receiving grid fetches the regions its peers own right before the next grid barrier
    def note_region(self, base: str, n: int, ring: bool = False):
        """`base`: C expression of the region's first element (already in batched element units), n: index-space length."""
        if self.world > 1:
            lo, hi = self.bounds(n)
            if hi - lo < n:
                item = ("prb_ll_fetch_ring" if ring else "prb_ll_fetch", f"{base}, ", (n - (hi - lo)) * self.nb,
                        f", {lo * self.nb}LL, {(hi - lo) * self.nb}LL")
                if item not in self.pending:
                    self.pending.append(item)

    def note_work(self, offset: int, n: int):
        self.note_region(f"{offset * self.nb}LL", n)

    def note_scatter(self, arr_offset: int, iarr_offset: int, n: int):
        if self.world > 1:
            lo, hi = self.bounds(n)
            if hi - lo < n:
                item = ("prb_ll_fetch_scatter", f"{arr_offset}LL, C.iconsts + {iarr_offset}LL, {self.nb}LL, ",
                        (n - (hi - lo)) * self.nb, f", {lo}LL, {hi - lo}LL")
                if item not in self.pending:
                    self.pending.append(item)

    def flush(self, indent: str = "    ") -> List[str]:
        """One loop over ALL elements the peers own of the regions stored since the last barrier.  The loop bounds and the
        barrier calls are placeholders: emit_net_kernel decides afterwards whether the grid shares the polling (cooperative
        form) or every CTA polls everything itself between an early arrival and the wait (redundant form)."""
        if not self.pending:
            return []
        total = sum(cnt for _, _, cnt, _ in self.pending)
        self.max_fetch = max(self.max_fetch, total)
        out = [f"{indent}@@ARRIVE@@",
               f"{indent}for (long long q = @@Q0@@; q < {total}LL; q += @@QS@@) {{   // fetch what the peers own "
               f"({len(self.pending)} region(s))"]
        start = 0
        for k, (fn, head, cnt, tail) in enumerate(self.pending):
            cond = "" if k == len(self.pending) - 1 else f"if (q < {start + cnt}LL) "
            out.append(f"{indent}    {'else ' if k else ''}{cond}{fn}(C, {head}q - {start}LL{tail});")
            start += cnt
        out.append(f"{indent}}}")
        self.pending = []
        return out

    def source_ptr(self, r: Reduction) -> str:
        """Pointer to element j = 0 (instance 0) of the source vector of a simple dot reduction."""
        xarr = self.ir.arrays[r.simple[1]]
        scale = f" * {self.nb}LL" if self.nb > 1 else ""
        if xarr.space == "ring":
            rows, L = xarr.shape
            k = self.ring_index[xarr.name]
            col = r.simple[2][1]
            return f"C.rings + ({xarr.offset}LL + prb_ring_pos(C.head[{k}], {col - xarr.rolls}, {L}) * {rows}LL){scale}"
        off = r.simple[2] + (xarr.offset if xarr.space != "y" else 0)
        return f"{self._ptr(xarr)} + {off}LL{scale}"

    def bidx(self, index: str) -> str:
        """Element index -> address inside a batched array ([element][instance], instance fastest)."""
        return f"({index}) * {self.nb}LL + b" if self.nb > 1 else index

    def bounds(self, n: int) -> Tuple[int, int]:
        """Rows [lo, hi) of an index space of length n owned by this rank (== sweep.shard_bounds)."""
        base, rem = divmod(int(n), self.world)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)

    # ---- role analysis: which loop indices a node depends on
    def role(self, i: int) -> frozenset:
        r = self.roles.get(i)
        if r is not None:
            return r
        g = self.g
        op, pl = g.ops[i], g.payload[i]
        if op == "ld":
            r = frozenset(pl[2])
        elif op in ("gather", "ldringg"):
            r = frozenset(pl[3])
        elif op == "ldring":
            r = frozenset(pl[2])
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
                r = r | self.role(a)
        self.roles[i] = r
        return r

    # ---- addressing
    def _ptr(self, arr: MemArray) -> str:
        return {"y": "C.yin", "const": "C.consts", "work": "C.work", "iconst": "C.iconsts", "table": "C.tables",
                "ring": "C.rings"}[arr.space]

    def _load(self, arr: MemArray, index: str) -> str:
        base = self._ptr(arr)
        off = f"{arr.offset}LL + " if arr.space != "y" else ""
        if arr.space in ("const", "iconst", "table"):
            return f"__ldg({base} + {off}{index})"
        return f"{base}[{self.bidx(off + index)}]"

    def ref(self, i: int, cur: Tuple[int, int, bool], want_real=True) -> str:
        g = self.g
        op = g.ops[i]
        if op == "const":
            v = g.const_value(i)
            if g.isint[i] and not want_real:
                return f"({int(v)}LL)"
            return _lit(float(v), self.f32)
        if op == "t":
            return "((real)C.t)" if want_real else "C.t"
        if op == "red":
            rid, kind = g.payload[i]
            red = self.ir.reductions[rid]
            if kind == "s":
                return f"sfull{rid}[b]" if self.nb > 1 else f"red{rid}"
            if cur == (red.phase, red.n, True):
                return f"red{rid}"
            arr = self.red_work[rid]
            return f"C.work[{self.bidx(f'{arr.offset}LL + i')}]"
        name = f"v{i}"
        if g.isint[i] and want_real:
            return f"((real){name})"
        return name

    def expr(self, i: int, cur) -> Optional[str]:
        """C expression computing node i from its operands (None for leaf refs that need no temp)."""
        g, ir = self.g, self.ir
        op, a, pl = g.ops[i], g.args[i], g.payload[i]
        r = lambda k, real=True: self.ref(k, cur, real)
        if op in ("const", "t", "red"):
            return None
        if op == "param":
            return f"__ldg(C.params + {pl}LL * {self.nb}LL + b)"
        if op == "ld":
            arr = ir.arrays[pl[0]]
            return self._load(arr, f"{pl[1]}LL + {pl[2]}")
        if op == "lds":
            arr = ir.arrays[pl[0]]
            return self._load(arr, f"{pl[1]}LL")
        if op == "ldm":
            arr = ir.arrays[pl[0]]
            return self._load(arr, f"(i - {self.bounds(arr.shape[0])[0]}LL) * {pl[2]}LL + j")
        if op == "gather":
            arr, iarr = ir.arrays[pl[0]], ir.arrays[pl[2]]
            return self._load(arr, f"{pl[1]}LL + (long long){self._load(iarr, pl[3])}")
        if op == "ldring":
            arr = ir.arrays[pl[0]]
            rows, L = arr.shape
            k = self.ring_index[pl[0]]
            return ("C.rings[" + self.bidx(f"{arr.offset}LL + prb_ring_pos(C.head[{k}], {pl[1] - arr.rolls}, {L}) * {rows}LL + {pl[2]}") + "]")
        if op in ("interp", "interprow"):
            # numpy.interp over two replicated constant tables; `C.t` is the float time of adaptive runs
            xa, fa = ir.arrays[pl[0]], ir.arrays[pl[1]]
            if op == "interp":
                return f"prb_interp(C.consts + {xa.offset}LL, C.consts + {fa.offset}LL, 1, {pl[2]}, {r(a[0])})"
            return f"prb_interp(C.consts + {xa.offset}LL, C.consts + {fa.offset}LL + {pl[4]}, {pl[3]}, {pl[2]}, {r(a[0])})"
        if op == "ldringg":
            # per-edge delayed read buf[source_idx[e], delays[e]]: the column is addressed relative to the ring head
            arr, rarr, carr = ir.arrays[pl[0]], ir.arrays[pl[1]], ir.arrays[pl[2]]
            rows, L = arr.shape
            k = self.ring_index[pl[0]]
            col = f"(int){self._load(carr, pl[3])} - {arr.rolls}"
            row = f"(long long){self._load(rarr, pl[3])}"
            return ("C.rings[" + self.bidx(f"{arr.offset}LL + prb_ring_pos(C.head[{k}], {col}, {L}) * {rows}LL + {row}") + "]")
        if op == "table":
            arr = ir.arrays[pl[0]]
            col = "i" if pl[2] == "i" else f"{pl[2]}LL"
            return self._load(arr, f"{r(pl[1], False)} * {pl[3]}LL + {col}")
        if g.isint[i]:
            x = r(a[0], False)
            y = r(a[1], False) if len(a) > 1 else None
            return {"add": f"{x} + {y}", "sub": f"{x} - {y}", "mul": f"{x} * {y}", "neg": f"-{x}"}[op]
        if op == "add":
            return f"{r(a[0])} + {r(a[1])}"
        if op == "sub":
            return f"{r(a[0])} - {r(a[1])}"
        if op == "mul":
            return f"{r(a[0])} * {r(a[1])}"
        if op == "div":
            return f"{r(a[0])} / {r(a[1])}"
        if op == "neg":
            return f"-{r(a[0])}"
        if op == "pow":
            return f"pow({r(a[0])}, {r(a[1])})"
        if op == "maximum":
            return f"prb_max({r(a[0])}, {r(a[1])})"
        if op == "minimum":
            return f"prb_min({r(a[0])}, {r(a[1])})"
        if op == "sign":
            return f"prb_sign({r(a[0])})"
        if op in _FUNC:
            return f"{_FUNC[op]}({r(a[0])})"
        raise NotImplementedError(f"no CUDA lowering for op `{op}`")

    def order(self, roots: Sequence[int]) -> List[int]:
        g = self.g
        seen: Set[int] = set()
        stack = list(roots)
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            stack.extend(g.args[i])
            if g.ops[i] == "table":
                stack.append(g.payload[i][1])
        return sorted(seen)

    def emit_nodes(self, nodes: Sequence[int], cur, emitted: Set[int], indent: str, only_roles=None) -> List[str]:
        out = []
        for i in nodes:
            if i in emitted:
                continue
            if only_roles is not None and not (self.role(i) <= only_roles):
                continue
            e = self.expr(i, cur)
            emitted.add(i)
            if e is None:
                continue
            ty = "long long" if self.g.isint[i] else "real"
            out.append(f"{indent}const {ty} v{i} = {e};")
        return out

    # ---- stores
    def store_stmt(self, st: Store, cur, stage_tpl: str) -> str:
        val = self.ref(st.node, cur)
        t = st.target
        B = self.bidx
        if t[0] == "dy":
            self.note_region(f"prb_stage_target<STAGE>(C) + {t[1] * self.nb}LL", st.n)
            return f"prb_emit_k<STAGE>(C, {B(f'{t[1]}LL + i')}, {val});"
        if t[0] == "work":
            arr = self.ir.arrays[t[1]]
            self.note_work(arr.offset, st.n)
            return f"prb_st_work(C, {B(f'{arr.offset}LL + i')}, {val});"
        if t[0] == "work_at":
            arr = self.ir.arrays[t[1]]
            self.note_work(arr.offset + t[2], st.n)
            return f"prb_st_work(C, {B(f'{arr.offset + t[2]}LL')}, {val});"
        if t[0] == "scatter":
            arr, iarr = self.ir.arrays[t[1]], self.ir.arrays[t[2]]
            self.note_scatter(arr.offset, iarr.offset, st.n)
            return f"prb_st_work(C, {B(f'{arr.offset}LL + (long long){self._load(iarr, chr(105))}')}, {val});"
        if t[0] == "ring":
            arr = self.ir.arrays[t[1]]
            rows, L = arr.shape
            k = self.ring_index[t[1]]
            col = f"{arr.offset}LL + prb_ring_pos(C.head[{k}], {t[2] - arr.rolls}, {L}) * {rows}LL"
            self.note_region(f"({col}) * {self.nb}LL", st.n, ring=True)
            pos = f"{col} + i"
            return f"prb_st_ring(C, {B(pos)}, {val});"
        raise NotImplementedError(t[0])


def tc_config(ir: NetIR, precision: str, n_batch: int, world: int, tc_max_acc: int = 128) -> Optional[dict]:
    """tcgen05 (3xTF32) configuration for the sweep-batched dense couplings of ``ir``, or None when the tensor-core
    path does not apply: it needs float32, a batch that is a multiple of 32 instances, one GPU and row reductions that
    are all plain ``dot(W, x)`` contractions."""
    rows = [r for r in ir.reductions if r.kind == "row"]
    if precision != "float32" or world != 1 or not rows or any(r.simple is None for r in rows):
        return None
    per_group = max(sum(1 for r in rows if (r.phase, r.n) == key) for key in {(r.phase, r.n) for r in rows})
    e = 2                                               # 8 epilogue warps; <= 64 fp32 accumulator registers per thread
    limit = tc_max_acc if per_group == 1 else min(tc_max_acc, 64)     # several live accumulator sets: keep the tail spill-free
    bt = next((c for c in (256, 128, 64, 32) if n_batch % c == 0 and per_group * c // e <= limit), None)
    if bt is None:
        return None
    # K values per shared-memory stage: half slabs (16) when a full-slab stage (2 x 16 KB of W + 2 x bt x 128 B of x) would
    # leave fewer than 3 stages inside ~200 KB, i.e. for the 256-instance tiles
    ck = 32
    if (200 * 1024) // (2 * 128 * ck * 4 + 2 * bt * ck * 4) < 3 and os.environ.get("PRB_TC_CK", "16") == "16":
        ck = 16
    stage = 2 * 128 * ck * 4 + 2 * bt * ck * 4
    return {"bt": bt, "e": e, "nstage": max(2, min(4, (200 * 1024) // stage)), "kc": 8, "block": 128 * (1 + e), "ck": ck}


def emit_net_kernel(ir: NetIR, *, solver: str, precision: str, observers: Sequence[int], block: int = 512,
                    rank: int = 0, world: int = 1, n_batch: int = 1, tensor_core: bool = False,
                    tc_max_acc: int = 128, split_rows: bool = True, dot_deep: bool = True,
                    w_stationary: bool = False, ll_redundant: bool = False) -> NetKernel:
    if solver not in SOLVER_IDS:
        raise ValueError(f"unknown solver `{solver}`")
    f32 = precision == "float32"
    g = ir.graph
    ns = ir.n_state
    adaptive = solver in ADAPTIVE_SOLVERS
    if adaptive:
        if f32 or int(n_batch) != 1 or world != 1:
            raise UnsupportedModelError("adaptive integration of networks (solver='scipy') is implemented for float64, one "
                                        "instance, one GPU")
        if any(a.space in ("ring", "table") for a in ir.arrays.values()):
            raise UnsupportedModelError("adaptive integration (solver='scipy') needs a pure vector field: delay ring buffers "
                                        "and step-indexed input tables are fixed-step constructs")
    n_solver_vecs = 11 if adaptive else 6          # RK45: Y, YN, two stage inputs, K0..K6
    # reduction results get a work array so that consumers outside the fused row task can read them
    work_size = ir.work_size
    nb = int(n_batch)
    R_T, B_T = 2, 8                       # batched row task: 2 rows x 8 instances per warp
    if nb > 1 and nb % B_T:
        raise ValueError(f"batched network kernels need a multiple of {B_T} instances (pad the sweep), got {nb}")
    em = _Emitter(ir, f32, rank, world, nb)
    tc = tc_config(ir, precision, nb, world, tc_max_acc) if (tensor_core and nb > 1) else None
    if tensor_core and tc is None:
        raise UnsupportedModelError("the tcgen05 coupling path needs a float32 model, a batch that is a multiple of 32 "
                                    "instances, one GPU and dense dot(W, x) couplings only")
    if tc is not None:
        block = tc["block"]
    tc_consts: Dict[str, Tuple[int, int, int]] = {}     # W array -> (hi offset, lo offset, K slabs) inside consts
    tc_xt: Dict[str, Tuple[int, int]] = {}              # source pointer expr -> (hi, lo) element offsets inside work
    tc_work = 0
    gemm_smem = 0
    tc_extra: List[Tuple[int, np.ndarray]] = []         # (const offset, tiled hi / lo image)
    # multi-GPU: 2-d constants (connectivity) are row-sharded -> re-pack the constant space with shard sizes.
    # Device-generated constants (devgen.DeviceUniform) go to the END of the space: the host image (const_init) covers only
    # what is uploaded, the session allocates the full space and lets each generator fill its (row-sharded) slice.
    const_size = ir.const_size
    const_host = const_size
    const_fills: List[Tuple[int, object, int, int]] = []      # (element offset, generator, row_lo, row_hi)
    lazy_arrays = [a for a in ir.arrays.values() if a.space == "const" and is_device_generated(a.init)]
    if world > 1 or lazy_arrays:
        const_size = 0
        consts = [a for a in ir.arrays.values() if a.space == "const"]
        for a in [a for a in consts if a not in lazy_arrays] + lazy_arrays:
            if a is (lazy_arrays[0] if lazy_arrays else None):
                const_host = const_size
            rows = a.shape[0] if a.shape else 1
            lo, hi = em.bounds(rows) if (len(a.shape) == 2 and not a.replicate) else (0, rows)
            size = (hi - lo) * (a.shape[1] if len(a.shape) == 2 else 1) if a.shape else 1
            a.offset = const_size
            const_size += ((max(size, 1) + 31) // 32) * 32
            if a in lazy_arrays:
                const_fills.append((a.offset, a.init.astype(np.float32 if f32 else np.float64), lo, hi))
        if not lazy_arrays:
            const_host = const_size
    full_part: Dict[int, MemArray] = {}
    for red in ir.reductions:
        if red.kind == "row":
            arr = MemArray(f"__red{red.id}", "work", (red.n,), work_size, np.zeros(red.n))
            work_size += ((red.n + 31) // 32) * 32
            ir.arrays[arr.name] = arr
            em.red_work[red.id] = arr
        elif nb == 1 and red.m >= FULL_TWO_LEVEL:
            # per-CTA partial sums of a two-level full reduction: FULL_PART_STRIDE slots per rank (grid <= 256 CTAs)
            arr = MemArray(f"__part{red.id}", "work", (world * FULL_PART_STRIDE,), work_size, np.zeros(world * FULL_PART_STRIDE))
            work_size += world * FULL_PART_STRIDE
            ir.arrays[arr.name] = arr
            full_part[red.id] = arr
    # observer index table goes to the end of iconsts
    obs = np.asarray(list(observers), dtype=np.int32)
    all_state = obs.size == ns and np.array_equal(obs, np.arange(ns))
    obs_off = ir.iconst_size
    iconst_size = ir.iconst_size + (0 if all_state else ((obs.size + 31) // 32) * 32)

    def row_split(reds, n) -> int:
        """Warps that share one row of a plain dot(W, x) reduction group (1 = the classic warp per row): the largest power
        of two that still gives every team a row, assuming one CTA per SM on the 148 SMs of a B200."""
        if nb != 1 or not split_rows or any(r.simple is None for r in reds) or block % 32:
            return 1
        lo_, hi_ = em.bounds(n)
        rows, wpb = hi_ - lo_, block // 32
        S = 1
        m_min = min(r.m for r in reds)
        while S * 2 <= wpb and rows * S * 2 <= 148 * wpb and m_min // (S * 2) >= 128:      # >= 2 double2 per lane and warp
            S *= 2
        return S if rows > 0 else 1

    wcache_plan: List[dict] = []         # row groups whose W rows live in shared memory for the whole launch
    wcache_used = 0                      # bytes of dynamic shared memory taken by them

    body: List[str] = []
    full_reds = [r for r in ir.reductions if r.kind == "full"]
    for r in full_reds:
        body.append(f"    __shared__ real sfull{r.id}[{nb}];" if nb > 1 else f"    real red{r.id} = (real)0;")
    if nb == 1:
        body.append("    const int b = 0; (void)b;")
    n_phases = ir.n_phases
    for p in range(n_phases):
        body.append(f"    // ---------------- phase {p}")
        # (1) full reductions.  Long ones (one instance, m >= FULL_TWO_LEVEL) run in two levels: every CTA of every rank sums
        # its share of the rows this rank owns and publishes ONE partial (multi-GPU: as an LL packet to the peers), a grid
        # barrier, then every CTA adds the world x 256 partials — instead of every CTA walking the whole vector (N = 65536:
        # 148 x 512 KB of L2 reads per evaluation) and, across GPUs, instead of every rank reducing its complete replica.
        two_level = [r for r in full_reds if r.phase == p and nb == 1 and r.m >= FULL_TWO_LEVEL]
        for r in two_level:
            cur = (p, -1, False)
            emitted = set()
            nodes = em.order([r.node])
            part = full_part[r.id]
            lo, hi = em.bounds(r.m)
            body.append(f"    {{   // full reduction {r.id}, level 1: rows {lo}..{hi} of {r.m}, one partial per CTA")
            body.append("        real acc = (real)0;")
            body.append(f"        for (long long j = {lo}LL + C.gtid; j < {hi}LL; j += C.gsize) {{")
            body.extend(em.emit_nodes(nodes, cur, emitted, "            "))
            body.append(f"            acc += {em.ref(r.node, cur)};")
            body.append("        }")
            body.append("        acc = prb_block_sum(acc);")
            body.append(f"        if (threadIdx.x == 0) prb_st_work(C, {part.offset + rank * FULL_PART_STRIDE}LL + blockIdx.x, acc);")
            body.append(f"        if (blockIdx.x == 0 && threadIdx.x >= gridDim.x && threadIdx.x < {FULL_PART_STRIDE})   // unused slots: zeros")
            body.append(f"            prb_st_work(C, {part.offset + rank * FULL_PART_STRIDE}LL + threadIdx.x, (real)0);")
            body.append("    }")
            em.note_work(part.offset, world * FULL_PART_STRIDE)
        if two_level:
            body.extend(em.flush())
            body.append("    @@BARRIER@@")
            for r in two_level:
                part = full_part[r.id]
                body.append(f"    {{   // level 2: sum of the {world} x {FULL_PART_STRIDE} partials")
                body.append("        real acc = (real)0;")
                body.append(f"        for (int q = threadIdx.x; q < {world * FULL_PART_STRIDE}; q += blockDim.x) acc += C.work[{part.offset}LL + q];")
                body.append(f"        red{r.id} = prb_block_sum(acc);")
                body.append("    }")
        for r in [r for r in full_reds if r.phase == p and r not in two_level]:
            cur = (p, -1, False)
            emitted: Set[int] = set()
            nodes = em.order([r.node])
            body.append(f"    for (int b = 0; b < {nb}; ++b) {{" if nb > 1 else "    {")
            body.append("        real acc = (real)0;")
            body.append(f"        for (long long j = threadIdx.x; j < {r.m}LL; j += blockDim.x) {{")
            body.extend(em.emit_nodes(nodes, cur, emitted, "            "))
            body.append(f"            acc += {em.ref(r.node, cur)};")
            body.append("        }")
            if nb > 1:
                body.append("        acc = prb_block_sum(acc);")
                body.append(f"        if (threadIdx.x == 0) sfull{r.id}[b] = acc;")
                body.append("    }")
                body.append("    __syncthreads();")
            else:
                body.append(f"        red{r.id} = prb_block_sum(acc);")
                body.append("    }")
        # (2) groups by length
        stores_p = [s for s in ir.stores if s.phase == p]
        rows_p = [r for r in ir.reductions if r.kind == "row" and r.phase == p]
        lengths = sorted({s.n for s in stores_p} | {r.n for r in rows_p})
        for n in lengths:
            reds = [r for r in rows_p if r.n == n]
            sts = [s for s in stores_p if s.n == n]
            if reds and nb > 1 and tc is not None:
                # sweep-batched dense coupling on the tensor cores: R[i, b] = sum_j W[i, j] x[j, b] as 128 x BT tiles of
                # tcgen05.mma kind::tf32 (3xTF32), element-wise tail + solver update fused into the TMEM epilogue
                cur = (p, n, True)
                BT, CPT = tc["bt"], tc["bt"] // tc["e"]
                n_mb, n_nbk = (n + 127) // 128, nb // BT
                tc_base = (work_size + n_solver_vecs * ns) * nb
                plan = []
                for r in reds:
                    wname = r.simple[0]
                    KS = (r.m + 31) // 32
                    if wname not in tc_consts:
                        from .tcgemm import split_tf32, tile_operand
                        w_hi, w_lo = split_tf32(np.asarray(ir.arrays[wname].init, dtype=np.float32).reshape(r.n, r.m))
                        t_hi, t_lo = tile_operand(w_hi, 128), tile_operand(w_lo, 128)
                        tc_consts[wname] = (const_size, const_size + t_hi.size, KS)
                        tc_extra.append((const_size, t_hi))
                        tc_extra.append((const_size + t_hi.size, t_lo))
                        const_size += t_hi.size + t_lo.size
                    xp = em.source_ptr(r)
                    if xp not in tc_xt:
                        sz = nb * KS * 32
                        tc_xt[xp] = (tc_base + tc_work, tc_base + tc_work + sz, r.m, KS)
                        tc_work += 2 * sz
                    plan.append((r, tc_consts[wname], tc_xt[xp], KS))
                body.append(f"    // tcgen05 coupling GEMM: {len(reds)} reduction(s), {n_mb} row block(s) x {n_nbk} instance block(s) of {BT}, n = {n}")
                for xp, (o_hi, o_lo, m_src, KS) in tc_xt.items():
                    if any(pl[2][0] == o_hi for pl in plan):
                        body.append(f"    prb_tc_prep(C.gtid, C.gsize, {xp}, {m_src}, C.work + {o_hi}LL, C.work + {o_lo}LL, {KS});")
                # stores that do not consume a coupling sum of this group run thread-per-(element, instance) on the
                # whole grid (coalesced along the instances); only the dependent tail is fused into the TMEM epilogue
                rids = {r.id for r in reds}
                free_sts = [s_ for s_ in sts if not any(_uses_red(g, em, s_.node, rid) for rid in rids)]
                sts = [s_ for s_ in sts if s_ not in free_sts]
                if free_sts:
                    cur_e = (p, n, False)
                    emitted_e: Set[int] = set()
                    body.append(f"    for (long long q = C.gtid; q < {n * nb}LL; q += C.gsize) {{   // stores independent of the coupling, n = {n}")
                    body.append(f"        const long long i = q / {nb}; const int b = (int)(q - i * {nb});")
                    body.extend(em.emit_nodes(em.order([s_.node for s_ in free_sts]), cur_e, emitted_e, "        "))
                    for s_ in free_sts:
                        body.append("        " + em.store_stmt(s_, cur_e, "STAGE"))
                    body.append("    }")
                body.append("    prb_grid_barrier(C);")
                body.append("    {")
                body.append("        const int warp = (int)(threadIdx.x >> 5);")
                body.append(f"        for (int tile = (int)blockIdx.x; tile < {n_mb * n_nbk}; tile += (int)gridDim.x) {{")
                body.append(f"            const int mbk = tile / {n_nbk}, nbk = tile - mbk * {n_nbk};")
                body.append("            if (warp == 0) {")
                body.append("                if (C.lane == 0) {")
                body.append("                    prb_fence_proxy_async();")
                for r, (a_hi, a_lo, _), (x_hi, x_lo, _, _), KS in plan:
                    body.append(f"                    prb_tc_produce(C.tc, C.consts + {a_hi}LL + (long long)mbk * {KS * 4096}LL, "
                                f"C.consts + {a_lo}LL + (long long)mbk * {KS * 4096}LL, "
                                f"C.work + {x_hi}LL + (long long)nbk * {KS * BT * 32}LL, "
                                f"C.work + {x_lo}LL + (long long)nbk * {KS * BT * 32}LL, {KS});")
                body.append("                }")
                body.append("            } else if (warp == 1) {")
                body.append("                if (C.lane == 0) {")
                for r, _, _, KS in plan:
                    body.append(f"                    prb_tc_mma(C.tc, {KS});")
                body.append("                }")
                body.append("            } else if (warp >= 4) {")
                body.append("                const int q = warp & 3, cg = (warp - 4) >> 2;")
                for r, _, _, KS in plan:
                    body.append(f"                float acc{r.id}[{CPT}];")
                    body.append("#pragma unroll")
                    body.append(f"                for (int e = 0; e < {CPT}; ++e) acc{r.id}[e] = 0.0f;")
                    body.append(f"                prb_tc_accumulate(C.tc, {KS}, acc{r.id}, q, cg);")
                body.append(f"                const long long i = (long long)mbk * 128 + q * 32 + C.lane;")
                body.append(f"                if (i < {n}LL) {{")
                body.append("#pragma unroll")
                body.append(f"                    for (int e = 0; e < {CPT}; ++e) {{")
                body.append(f"                        const int b = nbk * {BT} + cg * {CPT} + e;")
                for r, _, _, _ in plan:
                    body.append(f"                        const real red{r.id} = acc{r.id}[e];")
                emitted = set()
                for r in reds:
                    need = any(_uses_red(g, em, s_.node, r.id) and (s_.phase, s_.n) != (p, n) for s_ in ir.stores) or \
                        any(_uses_red(g, em, x.node, r.id) for x in ir.reductions)
                    if need:
                        body.append(f"                        prb_st_work(C, {em.bidx(f'{em.red_work[r.id].offset}LL + i')}, red{r.id});")
                nodes = em.order([s_.node for s_ in sts])
                body.extend(em.emit_nodes(nodes, cur, emitted, "                        "))
                for s_ in sts:
                    body.append("                        " + em.store_stmt(s_, cur, "STAGE"))
                body.append("                    }")
                body.append("                }")
                body.append("            }")
                body.append("        }")
                body.append("    }")
            elif reds and nb > 1 and nb % 32 == 0 and block == 512 and all(r.simple is not None for r in reds):
                # sweep-batched dense coupling as a shared-memory tiled GEMM on the FMA pipes (prb_gemm_batch), results to
                # the reductions' work arrays; the element-wise tail follows behind a grid barrier, thread per (element, instance)
                lo, hi = em.bounds(n)
                # instances per tile = 32 * cb: the widest tile that still gives every SM a tile, else the narrowest
                cands = [c for c in (4, 2, 1) if nb % (32 * c) == 0]
                n_tiles = lambda c: ((hi - lo + 127) // 128) * (nb // (32 * c))
                cb = next((c for c in cands if n_tiles(c) >= 120), cands[-1])
                item = 4 if f32 else 8
                gemm_smem = max(gemm_smem, 2 * 16 * (128 + 2 + 32 * cb) * 4 if f32 else 2 * (128 * 20 + 16 * (32 * cb + 4)) * 8)
                for r in reds:
                    warr = ir.arrays[r.simple[0]]
                    body.append(f"    prb_gemm_batch<{cb}>(C, C.consts + {warr.offset}LL, {lo}LL, {hi - lo}, {r.m}, {em.source_ptr(r)}, "
                                f"{em.red_work[r.id].offset}LL);      // n = {n}")
                    em.note_work(em.red_work[r.id].offset, n)
                body.extend(em.flush())
                body.append("    @@BARRIER@@")
                if sts:
                    cur = (p, n, False)
                    emitted = set()
                    body.append(f"    for (long long q = {lo * nb}LL + C.gtid; q < {hi * nb}LL; q += C.gsize) {{   // tail, thread per (element, instance), n = {n}")
                    body.append(f"        const long long i = q / {nb}; const int b = (int)(q - i * {nb});")
                    body.extend(em.emit_nodes(em.order([s_.node for s_ in sts]), cur, emitted, "        "))
                    for s_ in sts:
                        body.append("        " + em.store_stmt(s_, cur, "STAGE"))
                    body.append("    }")
            elif reds and nb > 1:
                # batched sweep: one warp task = R_T rows x B_T instances; W rows are read once per B_T instances
                cur = (p, n, True)
                lo, hi = em.bounds(n)
                if any(r.simple is None for r in reds):
                    raise UnsupportedModelError("batched network sweeps support dense dot couplings (and pair-space "
                                                "couplings that separate into them) only")
                n_bt = nb // B_T
                body.append(f"    for (long long task = C.gwarp; task < {((hi - lo + R_T - 1) // R_T) * n_bt}LL; task += C.nwarps) {{"
                            f"   // {R_T} rows x {B_T} instances per warp, n = {n}")
                body.append(f"        const long long i0 = {lo}LL + (task / {n_bt}) * {R_T};")
                body.append(f"        const int b0 = (int)(task % {n_bt}) * {B_T};")
                body.append(f"        const int rows = (int)(({hi}LL - i0) < {R_T} ? ({hi}LL - i0) : {R_T});")
                for r in reds:
                    warr = ir.arrays[r.simple[0]]
                    body.append(f"        real acc{r.id}[{R_T}][{B_T}];")
                    body.append(f"        prb_rows_dot_batch<{R_T}, {B_T}>(C.consts + {warr.offset}LL + (i0 - {lo}LL) * {r.m}LL, {r.m}LL, rows, "
                                f"{em.source_ptr(r)} + b0, {r.m}, C.lane, acc{r.id});")
                body.append(f"        const long long i = i0 + (C.lane / {B_T});")
                body.append(f"        const int b = b0 + (C.lane % {B_T});")
                for r in reds:
                    body.append(f"        const real red{r.id} = prb_tile_pick<{R_T}, {B_T}>(acc{r.id}, C.lane);")
                body.append(f"        if (C.lane < {R_T * B_T} && i < {hi}LL) {{")
                emitted = set()
                for r in reds:
                    need = any(_uses_red(g, em, s_.node, r.id) and (s_.phase, s_.n) != (p, n) for s_ in ir.stores) or \
                        any(_uses_red(g, em, x.node, r.id) for x in ir.reductions)
                    if need:
                        em.note_work(em.red_work[r.id].offset, n)
                        body.append(f"            prb_st_work(C, {em.bidx(f'{em.red_work[r.id].offset}LL + i')}, red{r.id});")
                nodes = em.order([s_.node for s_ in sts])
                body.extend(em.emit_nodes(nodes, cur, emitted, "            "))
                for s_ in sts:
                    body.append("            " + em.store_stmt(s_, cur, "STAGE"))
                body.append("        }")
                body.append("    }")
            elif reds and row_split(reds, n) > 1:
                # few rows per GPU (row-sharded shards, small networks): S consecutive warps of a CTA share one row — each
                # takes a contiguous segment of the W row, the partial sums meet in shared memory and the team's first warp runs
                # the element-wise tail.  With one warp per row a 1024-row shard keeps only 1024 x 4 loads in flight (~2 MB):
                # ~30 % of what the HBM / L2 stream needs.  The summation order differs from the one-warp form (1e-16).
                cur = (p, n, True)
                lo, hi = em.bounds(n)
                S = row_split(reds, n)
                wpb = block // 32
                tpb = wpb // S
                tagp = f"{p}_{n}"
                fused_done: Set[int] = set()
                pairs = []
                for r in reds:
                    if r.id in fused_done:
                        continue
                    mate = next((q for q in reds if q.id != r.id and q.id not in fused_done and q.simple[0] == r.simple[0]
                                 and q.m == r.m), None)
                    fused_done.add(r.id)
                    if mate is not None:
                        fused_done.add(mate.id)
                    pairs.append((r, mate))
                # W-stationary rows (opt-in, w_stationary=True): when every team of the (148-CTA) grid owns at most ONE row of
                # this group and the rows a CTA owns fit its shared memory, each CTA copies them there once per launch
                # (prb_net_stage_consts) and every evaluation reads W from shared memory instead of streaming it from L2 again.
                # Measured on B200 (profiles/r02_wcache.jsonl): correct but SLOWER than the 8-deep L2 stream (C4 2 x 1024: 16.0
                # vs 13.7 us per step, C3 N=1024: 4.02 vs 3.47 us) — these networks are bound by the barrier / reduction latency
                # chain, not by the W bytes, and the 128 KB carve-out costs L1.  Off by default.
                row_elems = sum(r.m for r, _ in pairs)
                cache_bit = None
                if (w_stationary and not f32 and hi - lo <= 148 * tpb and all(r.m % 2 == 0 for r, _ in pairs)
                        and wcache_used + tpb * row_elems * 8 <= WCACHE_MAX_BYTES and len(wcache_plan) < 30):
                    cache_bit = len(wcache_plan)
                    offs, o = [], 0
                    for r, _ in pairs:
                        offs.append(o)
                        o += r.m
                    wcache_plan.append({"bit": cache_bit, "S": S, "tpb": tpb, "lo": lo, "hi": hi, "base": wcache_used // 8,
                                        "row_elems": row_elems, "rows": [(ir.arrays[r.simple[0]].offset, r.m, off) for (r, _), off in zip(pairs, offs)]})
                    wcache_used += tpb * row_elems * 8
                body.append(f"    {{   // {S} warps per row, n = {n} ({hi - lo} rows on this GPU)"
                            + (f", W rows cached in shared memory (bit {cache_bit})" if cache_bit is not None else ""))
                body.append(f"        __shared__ real prb_part{tagp}[{len(reds)}][{wpb}];")
                body.append(f"        const int wib = (int)(threadIdx.x >> 5), tib = wib / {S}, wit = wib - tib * {S};")
                body.append(f"        const long long nteams = (long long)gridDim.x * {tpb}, team = (long long)blockIdx.x * {tpb} + tib;")
                if cache_bit is not None:
                    body.append("        extern __shared__ __align__(16) unsigned char prb_dyn_smem[];")
                    body.append(f"        const real* wsm = (const real*)prb_dyn_smem + {wcache_plan[-1]['base']}LL + (long long)tib * {row_elems}LL;")
                    body.append(f"        const bool wcached = (C.wcache >> {cache_bit}) & 1u;")
                body.append(f"        for (long long i0 = {lo}LL; i0 < {hi}LL; i0 += nteams) {{")
                body.append(f"            const long long i = i0 + team;")
                body.append(f"            const bool act = i < {hi}LL;")
                body.append(f"            if (act) {{")
                for pi_, (r, mate) in enumerate(pairs):
                    warr = ir.arrays[r.simple[0]]
                    seg = -(-(-(-r.m // S)) // 4) * 4                     # elements per warp, a multiple of 4: keeps the 16-byte alignment
                    body.append(f"                {{")
                    body.append(f"                    const int s0 = wit * {seg}, len = ({r.m} - s0) < {seg} ? (({r.m} - s0) > 0 ? ({r.m} - s0) : 0) : {seg};")
                    wptr = f"C.consts + {warr.offset}LL + (i - {lo}LL) * {r.m}LL + s0"
                    wsp = f"wsm + {wcache_plan[-1]['rows'][pi_][2]}LL + s0" if cache_bit is not None else None
                    if mate is not None:
                        body.append(f"                    real pa, pb;")
                        if wsp is not None:
                            body.append(f"                    if (wcached) prb_row_dot2_sm({wsp}, {em.source_ptr(r)} + s0, {em.source_ptr(mate)} + s0, len, C.lane, pa, pb);")
                            body.append(f"                    else")
                        body.append(f"                    prb_row_dot2({wptr}, {em.source_ptr(r)} + s0, {em.source_ptr(mate)} + s0, len, C.lane, pa, pb);")
                        body.append(f"                    if (C.lane == 0) {{ prb_part{tagp}[{reds.index(r)}][wib] = pa; prb_part{tagp}[{reds.index(mate)}][wib] = pb; }}")
                    else:
                        if wsp is not None:
                            body.append(f"                    const real pa = wcached ? prb_row_dot_sm({wsp}, {em.source_ptr(r)} + s0, len, C.lane)")
                            body.append(f"                                            : prb_row_dot({wptr}, {em.source_ptr(r)} + s0, len, C.lane);")
                        else:
                            body.append(f"                    const real pa = prb_row_dot({wptr}, {em.source_ptr(r)} + s0, len, C.lane);")
                        body.append(f"                    if (C.lane == 0) prb_part{tagp}[{reds.index(r)}][wib] = pa;")
                    body.append(f"                }}")
                body.append(f"            }}")
                body.append(f"            __syncthreads();")
                body.append(f"            if (act && wit == 0) {{")
                emitted = set()
                for k, r in enumerate(reds):
                    terms = " + ".join(f"prb_part{tagp}[{k}][tib * {S} + {w}]" for w in range(S))
                    body.append(f"                const real red{r.id} = {terms};")
                    need = any(_uses_red(g, em, s_.node, r.id) and (s_.phase, s_.n) != (p, n) for s_ in ir.stores) or \
                        any(_uses_red(g, em, x.node, r.id) for x in ir.reductions)
                    if need:
                        em.note_work(em.red_work[r.id].offset, n)
                        body.append(f"                if (C.lane == 0) prb_st_work(C, {em.red_work[r.id].offset}LL + i, red{r.id});")
                body.extend(em.emit_nodes(em.order([s_.node for s_ in sts]), cur, emitted, "                "))
                if sts:
                    body.append("                if (C.lane == 0) {")
                    for s_ in sts:
                        body.append("                    " + em.store_stmt(s_, cur, "STAGE"))
                    body.append("                }")
                body.append(f"            }}")
                body.append(f"            __syncthreads();")
                body.append(f"        }}")
                body.append(f"    }}")
            elif reds:
                cur = (p, n, True)
                lo, hi = em.bounds(n)
                body.append(f"    for (long long i = {lo}LL + C.gwarp; i < {hi}LL; i += C.nwarps) {{      // warp per row, n = {n}")
                emitted = set()
                def xptr_of(r):
                    return em.source_ptr(r)
                fused_done: Set[int] = set()
                for r in reds:
                    if r.id in fused_done:
                        continue
                    if r.simple is not None:
                        warr = ir.arrays[r.simple[0]]
                        wptr = f"C.consts + {warr.offset}LL + (i - {lo}LL) * {r.m}LL"
                        mate = next((q for q in reds if q.id != r.id and q.id not in fused_done and q.simple is not None
                                     and q.simple[0] == r.simple[0] and q.m == r.m), None)
                        if mate is not None:        # same W: both dot products share one pass over the row
                            fused_done.add(mate.id)
                            body.append(f"        real red{r.id}, red{mate.id};")
                            body.append(f"        prb_row_dot2({wptr}, {xptr_of(r)}, {xptr_of(mate)}, {r.m}, C.lane, red{r.id}, red{mate.id});")
                        else:
                            body.append(f"        const real red{r.id} = prb_row_dot({wptr}, {xptr_of(r)}, {r.m}, C.lane);")
                        for q in ([r] if mate is None else [r, mate]):
                            need = any(_uses_red(g, em, s.node, q.id) and (s.phase, s.n) != (p, n) for s in ir.stores) or \
                                any(_uses_red(g, em, x.node, q.id) for x in ir.reductions)
                            if need:
                                em.note_work(em.red_work[q.id].offset, n)
                                body.append(f"        if (C.lane == 0) prb_st_work(C, {em.red_work[q.id].offset}LL + i, red{q.id});")
                        continue
                    else:
                        nodes = em.order([r.node])
                        body.extend(em.emit_nodes(nodes, cur, emitted, "        ", only_roles=frozenset("i")))
                        body.append(f"        real acc{r.id}a = (real)0, acc{r.id}b = (real)0;")
                        inner: Set[int] = set(emitted)
                        body.append(f"        for (long long j = C.lane; j < {r.m}LL; j += 64) {{")
                        body.extend(em.emit_nodes(nodes, cur, inner, "            "))
                        body.append(f"            acc{r.id}a += {em.ref(r.node, cur)};")
                        body.append(f"            const long long jb = j + 32;")
                        body.append(f"            if (jb < {r.m}LL) {{")
                        body.append(f"                const long long j = jb;")
                        inner2: Set[int] = set(emitted)
                        body.extend(em.emit_nodes(nodes, cur, inner2, "                "))
                        body.append(f"                acc{r.id}b += {em.ref(r.node, cur)};")
                        body.append("            }")
                        body.append("        }")
                        body.append(f"        const real red{r.id} = prb_warp_sum(acc{r.id}a + acc{r.id}b);")
                    need_store = any(_uses_red(g, em, s.node, r.id) and (s.phase, s.n) != (p, n) for s in ir.stores) or \
                        any(_uses_red(g, em, q.node, r.id) for q in ir.reductions)
                    if need_store:
                        em.note_work(em.red_work[r.id].offset, n)
                        body.append(f"        if (C.lane == 0) prb_st_work(C, {em.red_work[r.id].offset}LL + i, red{r.id});")
                roots = [s.node for s in sts]
                nodes = em.order(roots)
                body.extend(em.emit_nodes(nodes, cur, emitted, "        "))
                if sts:
                    body.append("        if (C.lane == 0) {")
                    for s in sts:
                        body.append("            " + em.store_stmt(s, cur, "STAGE"))
                    body.append("        }")
                body.append("    }")
            elif sts:
                cur = (p, n, False)
                roots = [s.node for s in sts]
                nodes = em.order(roots)
                emitted = set()
                lo, hi = em.bounds(n)
                if nb > 1:
                    body.append(f"    for (long long q = {lo * nb}LL + C.gtid; q < {hi * nb}LL; q += C.gsize) {{   // thread per (element, instance), n = {n}")
                    body.append(f"        const long long i = q / {nb}; const int b = (int)(q - i * {nb});")
                else:
                    body.append(f"    for (long long i = {lo}LL + C.gtid; i < {hi}LL; i += C.gsize) {{        // thread per element, n = {n}")
                body.extend(em.emit_nodes(nodes, cur, emitted, "        "))
                for s in sts:
                    body.append("        " + em.store_stmt(s, cur, "STAGE"))
                body.append("    }")
        # multi-GPU: fetch what the peers stored in this phase into the local replicas (the last phase's regions are
        # fetched here as well: the barrier that ends the evaluation follows in prb_stage)
        last_fetch = em.flush()
        body.extend(last_fetch)
        if p < n_phases - 1:
            body.append("    @@BARRIER@@")
        elif not last_fetch:
            body.append("    @@ARRIVE@@")       # redundant form: the evaluation always ends with the early arrival

    rings = [a for a in ir.arrays.values() if a.space == "ring"]
    src: List[str] = []
    src.append("// Generated by pyrates_b200.emit_net — do not edit.  Network execution model (persistent cooperative grid).")
    src.append(read_template("prb_kernel_abi.cuh"))
    src.append(f"typedef {'float' if f32 else 'double'} real;")
    src.append(f"typedef {'double' if adaptive else 'long long'} prb_time;")
    if adaptive:
        from .rk45 import cuda_tables
        src.append(cuda_tables())
    src.append(f"#define PRB_NS {ns}LL")
    src.append(f"#define PRB_NOBS {obs.size}LL")
    src.append(f"#define PRB_NRINGS {len(rings)}")
    src.append(f"#define PRB_SOLVER {SOLVER_IDS[solver]}")
    src.append(f"#define PRB_BLOCK {block}")
    src.append(f"#define PRB_WORK_USER {work_size}LL")
    src.append(f"#define PRB_NB {nb}")
    src.append(f"#define PRB_DOT_DEEP {1 if dot_deep else 0}")
    src.append(f"#define PRB_RANK {rank}")
    src.append(f"#define PRB_WORLD {world}")
    src.append("#define PRB_LL_REDUNDANT @@LLR@@")
    # multi-GPU receive mirror (net_kernel.cuh: prb_ll_*): one 16-byte packet slot per work / ring element, two parities
    src.append(f"#define PRB_LL_WORK {(work_size + n_solver_vecs * ns) * nb + tc_work}LL")
    src.append(f"#define PRB_LL_RING {ir.ring_size * nb}LL")
    src.append("#define PRB_OBS_INDEX(o) (o)" if all_state else f"#define PRB_OBS_INDEX(o) ((long long)A.iconsts[{obs_off}LL + (o)])")
    src.append("struct PrbNet;")
    src.append("__device__ __forceinline__ void prb_net_init_heads(PrbNet& C, long long eval0);")
    src.append("__device__ __forceinline__ void prb_net_advance_heads(PrbNet& C);")
    src.append("__device__ __forceinline__ void prb_net_stage_consts(PrbNet& C);")
    if tc is not None:
        src.append("#define PRB_TC 1")
        src.append(f"#define PRB_TC_BT {tc['bt']}")
        src.append(f"#define PRB_TC_E {tc['e']}")
        src.append(f"#define PRB_TC_NSTAGE {tc['nstage']}")
        src.append(f"#define PRB_TC_KC {tc['kc']}")
        src.append(f"#define PRB_TC_CK {tc['ck']}")
        src.append(read_template("tc_gemm.cuh"))
    src.append(read_template("net_kernel.cuh"))
    heads_i, heads_a = [], []
    for k, a in enumerate(rings):
        L = a.shape[1]
        heads_i.append(f"    C.head[{k}] = (int)((eval0 * {a.rolls}LL) % {L}LL);")
        if a.rolls:
            heads_a.append(f"    C.head[{k}] = (C.head[{k}] + {a.rolls % L}) % {L};")
    src.append("__device__ __forceinline__ void prb_net_init_heads(PrbNet& C, long long eval0) {\n" + "\n".join(heads_i) + "\n}")
    src.append("__device__ __forceinline__ void prb_net_advance_heads(PrbNet& C) {\n" + "\n".join(heads_a) + "\n}")
    # multi-GPU exchange form.  Cooperative: the grid shares the polling of the peers' packets, then meets at the barrier
    # (remote arrival and local barrier in series).  Redundant (small exchanges only): every CTA arrives EARLY, polls all
    # peer-owned elements itself (identical values, benign duplicate stores into the replica) and only then waits for the
    # local arrivals — the local barrier overlaps the NVLink flight instead of following it.  Opt-in (ll_redundant=True):
    # measured on 2 B200s (profiles/r02_redundant_n2.jsonl) it is worth 2-4 % for a single small region (C3 N=1024: 5.03 vs
    # 5.26 us per step) and LOSES where many regions meet one barrier (QIF-SFA 2 x 512, 13 regions: 21.4 vs 16.0 us — every
    # CTA walks the whole region list); the cooperative form stays the default.
    redundant = world > 1 and 0 < em.max_fetch <= LL_REDUNDANT_MAX and ll_redundant
    fixed: List[str] = []
    armed = False
    for line in body:
        if "@@ARRIVE@@" in line:
            if redundant:
                fixed.append(line.replace("@@ARRIVE@@", "prb_barrier_arrive(C);"))
                armed = True
            continue
        if "@@BARRIER@@" in line:
            fixed.append(line.replace("@@BARRIER@@", "prb_barrier_wait(C);" if (redundant and armed) else "prb_grid_barrier(C);"))
            armed = False
            continue
        fixed.append(line.replace("@@Q0@@", "(long long)threadIdx.x" if redundant else "C.gtid")
                     .replace("@@QS@@", "(long long)blockDim.x" if redundant else "C.gsize"))
    body = fixed
    stage = ["// once per launch: every CTA copies the W rows its warp teams own into shared memory (W-stationary small networks)",
             "__device__ __forceinline__ void prb_net_stage_consts(PrbNet& C) {", "    C.wcache = 0u;"]
    if wcache_plan:
        stage.append("    extern __shared__ __align__(16) unsigned char prb_dyn_smem[];")
        stage.append("    const int wib = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);")
    for g_ in wcache_plan:
        S_, tpb_ = g_["S"], g_["tpb"]
        stage.append(f"    if ((long long)gridDim.x * {tpb_} >= {g_['hi'] - g_['lo']}LL) {{      // one row per team: the cached form is valid")
        stage.append(f"        C.wcache |= {1 << g_['bit']}u;")
        stage.append(f"        const int tib = wib / {S_}, wit = wib - tib * {S_};")
        stage.append(f"        const long long i = {g_['lo']}LL + (long long)blockIdx.x * {tpb_} + tib;")
        stage.append(f"        if (i < {g_['hi']}LL) {{")
        for woff, m_, off in g_["rows"]:
            stage.append(f"            {{ const double2* src_ = (const double2*)(C.consts + {woff}LL + (i - {g_['lo']}LL) * {m_}LL);")
            stage.append(f"              double2* dst_ = (double2*)((real*)prb_dyn_smem + {g_['base']}LL + (long long)tib * {g_['row_elems']}LL + {off}LL);")
            stage.append(f"              for (int j = wit * 32 + lane; j < {m_ // 2}; j += {S_ * 32}) dst_[j] = prb_ld_stream(src_ + j); }}")
        stage.append("        }")
        stage.append("    }")
    if wcache_plan:
        stage.append("    __syncthreads();")
    stage.append("}")
    src.append("\n".join(stage))
    src.append("template <int STAGE> __device__ void prb_eval(PrbNet& C) {")
    src.extend(body)
    src.append("}")

    real = np.float32 if f32 else np.float64

    def pack(space, size, dtype):
        buf = np.zeros(max(size, 1), dtype=dtype)
        for a in ir.arrays.values():
            if a.space == space and a.init is not None and not is_device_generated(a.init):
                v = np.asarray(a.init)
                if space == "ring":
                    v = v.T                        # device layout [L][rows]: a delay column is contiguous
                if space == "const" and world > 1 and v.ndim == 2 and not a.replicate:
                    lo, hi = em.bounds(v.shape[0])
                    v = v[lo:hi]                   # this rank's row shard of the connectivity
                buf[a.offset:a.offset + v.size] = v.reshape(-1)          # casts on assignment, no temporary
        return buf

    if lazy_arrays and tc_extra:
        raise UnsupportedModelError("device-generated connectivities are not supported on the tcgen05 sweep path")
    const_buf = pack("const", const_host if lazy_arrays else const_size, real)
    for off, img in tc_extra:
        const_buf[off:off + img.size] = img
    iconst = pack("iconst", iconst_size, np.int32)
    if not all_state:
        iconst[obs_off:obs_off + obs.size] = obs
    evals = {"euler": 1, "heun": 2, "heun_ref": 2, "heun_true": 2, "rk4": 4, "scipy": 6, "rk45": 6}[solver]
    smem = max(gemm_smem, wcache_used)
    if tc is not None:
        smem = tc["nstage"] * (2 * 128 * tc["ck"] * 4 + 2 * tc["bt"] * tc["ck"] * 4)
        tc = dict(tc, smem=smem, xt_elems=tc_work)
    src = [x.replace("@@LLR@@", "1" if redundant else "0") for x in src]
    return NetKernel("\n".join(src) + "\n", ns, int(obs.size), (work_size + n_solver_vecs * ns) * nb + tc_work, work_size,
                     pack("work", work_size, real), const_buf, iconst,
                     pack("ring", ir.ring_size, real), pack("table", ir.table_size, real), block, solver, precision,
                     evals, n_phases, smem_bytes=smem, tc=tc,
                     ll_slots=((work_size + n_solver_vecs * ns) * nb + tc_work + ir.ring_size * nb + 32) if world > 1 else 0,
                     const_elems=max(const_size, 1), const_fills=tuple(const_fills))


def _uses_red(g, em: _Emitter, node: int, rid: int) -> bool:
    for i in em.order([node]):
        if g.ops[i] == "red" and g.payload[i][0] == rid:
            return True
    return False
