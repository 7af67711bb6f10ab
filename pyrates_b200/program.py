"""Vector-field program: the backend-neutral description of one PyRates model's rhs.

A ``Program`` is exactly what the reference hands to a backend:
the ordered assignment stream that ``ComputeGraph._to_str`` issues through
``backend.add_var_update`` (pyrates/backend/computegraph.py:1109-1139, 1141-1218) plus the
argument tuple ``(t, y, dy, *constants)`` that ``ComputeGraph.to_func`` returns
(pyrates/backend/computegraph.py:471-483).  The statements are kept in the NumPy-syntax text
form the reference produces from sympy (``lhs[idx] = rhs``), so a Program can be built

* live, by ``CudaBackend`` while the reference's compute graph drives its hooks, or
* from a reference-generated ``vector_field`` source file + its argument arrays (this is how
  the golden fixtures under ``tests/golden`` travel to the GPU box, where the reference
  itself is not available).

The Program is pure data (no CUDA, no reference imports); the compilers in
``pyrates_b200.scalarize`` / ``pyrates_b200.netcompile`` lower it to CUDA C.
"""
from __future__ import annotations

import ast
import io
import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

__all__ = ["Arg", "Stmt", "Program"]


@dataclass
class Arg:
    """One entry of the generated function's signature ``(t, y, dy, *constants)``.

    ``kind``: 't' (step counter / time), 'y' (state vector), 'dy' (vector-field buffer),
    'const' (read-only parameter / index array / input table), 'var' (mutable non-state
    buffer that persists across rhs evaluations because the reference mutates it in place,
    e.g. delay ring buffers, pyrates/ir/circuit.py:400-425) or 'hist' (the state-history callable of
    delay-differential models, base_backend.py:88-177; its value is the initial state it starts from).
    """
    name: str
    kind: str
    value: np.ndarray
    label: str = ""          # frontend name (``node/op/var``) when known

    @property
    def is_int(self) -> bool:
        return np.issubdtype(self.value.dtype, np.integer)


@dataclass
class Stmt:
    """``lhs[lhs_idx] = rhs`` with NumPy semantics; ``lhs_idx`` is None for a plain binding."""
    lhs: str
    rhs: str
    lhs_idx: Optional[str] = None

    def text(self) -> str:
        return f"{self.lhs}[{self.lhs_idx}] = {self.rhs}" if self.lhs_idx is not None else f"{self.lhs} = {self.rhs}"


@dataclass
class Program:
    args: List[Arg]
    stmts: List[Stmt]
    state_vars: Dict[str, Union[int, Tuple[int, int]]] = field(default_factory=dict)
    float_precision: str = "float64"
    func_name: str = "vector_field"

    # ------------------------------------------------------------------ basic views
    @property
    def is_complex(self) -> bool:
        return "complex" in str(self.float_precision)

    @property
    def n_state(self) -> int:
        """Number of REAL state variables (a complex state occupies two: re, im)."""
        return int(self.arg_by_kind("y").value.size) * (2 if self.is_complex else 1)

    def arg(self, name: str) -> Arg:
        for a in self.args:
            if a.name == name:
                return a
        raise KeyError(name)

    def arg_by_kind(self, kind: str) -> Arg:
        for a in self.args:
            if a.kind == kind:
                return a
        raise KeyError(kind)

    @property
    def y0(self) -> np.ndarray:
        y = np.asarray(self.arg_by_kind("y").value).reshape(-1)
        if self.is_complex:
            return np.stack([y.real, y.imag], axis=1).reshape(-1).astype(self.real_dtype)
        return y

    @property
    def is_dde(self) -> bool:
        return any(a.kind == "hist" for a in self.args)

    @property
    def real_dtype(self):
        return np.float32 if self.float_precision in ("float32", "complex64") else np.float64

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_source(cls, source: str, args: List[Tuple[str, np.ndarray]], *,
                    state_vars: Optional[dict] = None, float_precision: Optional[str] = None,
                    func_name: str = "vector_field", labels: Optional[List[str]] = None) -> "Program":
        """Build a Program from reference-generated NumPy source (``def vector_field(t,y,dy,...)``).

        ``args`` is the ordered list of ``(signature_name, array)`` matching the function
        signature.  Only the body assignments are used; import lines and helper function
        definitions (``sigmoid``, ``wsum``...) are implied by the op vocabulary
        (pyrates/backend/base/base_funcs.py:135-177).
        """
        tree = ast.parse(source.replace("\t", "    "))
        fdef = None
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name == func_name:
                fdef = node
        if fdef is None:
            raise ValueError(f"no function `{func_name}` in source")
        sig = [a.arg for a in fdef.args.args]
        if len(sig) != len(args):
            raise ValueError(f"signature has {len(sig)} names but {len(args)} argument arrays were given")
        stmts: List[Stmt] = []
        assigned = set()
        for node in fdef.body:
            if isinstance(node, ast.Return):
                continue
            if not isinstance(node, ast.Assign) or len(node.targets) != 1:
                raise NotImplementedError(f"unsupported statement in generated source: {ast.unparse(node)}")
            tgt = node.targets[0]
            if isinstance(tgt, ast.Name):
                stmts.append(Stmt(tgt.id, ast.unparse(node.value)))
                assigned.add(tgt.id)
            elif isinstance(tgt, ast.Subscript) and isinstance(tgt.value, ast.Name):
                full = ast.unparse(tgt)               # `name[idx]`; unparsing the slice alone breaks tuples
                stmts.append(Stmt(tgt.value.id, ast.unparse(node.value), full[len(tgt.value.id) + 1:-1]))
                assigned.add(tgt.value.id)
            else:
                raise NotImplementedError(f"unsupported assignment target: {ast.unparse(tgt)}")
        out_args = []
        has_hist = len(sig) > 2 and sig[2] == "hist"       # DDE signature: (t, y, hist, dy, ...) base_backend.py:505-509
        for i, (name, (aname, val)) in enumerate(zip(sig, args)):
            if has_hist and i == 2:
                # the history callable does not travel; its value is the float64 initial state it was built from
                # (ComputeGraph.to_func: get_hist_func(np.asarray(state_vec[:])), computegraph.py:471-478)
                out_args.append(Arg(name, "hist", np.array(getattr(val, "_y", [args[1][1]])[0], dtype=np.float64, copy=True)))
                continue
            val = np.asarray(val)
            if i == 0:
                kind = "t"
            elif i == 1:
                kind = "y"
            elif i == (3 if has_hist else 2):
                kind = "dy"
            else:
                kind = "var" if name in assigned else "const"
            out_args.append(Arg(name, kind, val, labels[i] if labels else (aname if aname != name else "")))
        if float_precision is None:
            float_precision = str(np.asarray(args[1][1]).dtype)
        return cls(out_args, stmts, dict(state_vars or {}), float_precision, func_name)

    # ------------------------------------------------------------------ text form
    def source(self) -> str:
        """Regenerate the NumPy-syntax function text (same shape as the reference's output,
        pyrates/backend/base/base_backend.py:483-535)."""
        lines = [f"def {self.func_name}({','.join(a.name for a in self.args)}):"]
        for s in self.stmts:
            lines.append("\t" + s.text())
        lines.append(f"\treturn {self.arg_by_kind('dy').name}")
        return "\n".join(lines) + "\n"

    # ------------------------------------------------------------------ (de)serialisation
    def to_npz_dict(self) -> dict:
        meta = {
            "func_name": self.func_name,
            "float_precision": self.float_precision,
            "state_vars": {k: (list(v) if isinstance(v, (tuple, list)) else int(v)) for k, v in self.state_vars.items()},
            "args": [{"name": a.name, "kind": a.kind, "label": a.label} for a in self.args],
            "stmts": [[s.lhs, s.rhs, s.lhs_idx] for s in self.stmts],
        }
        out = {"__program__": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
        for i, a in enumerate(self.args):
            out[f"arg{i}"] = np.asarray(a.value)
        return out

    @classmethod
    def from_npz_dict(cls, d) -> "Program":
        meta = json.loads(bytes(np.asarray(d["__program__"]).tobytes()).decode())
        args = [Arg(m["name"], m["kind"], np.asarray(d[f"arg{i}"]), m.get("label", ""))
                for i, m in enumerate(meta["args"])]
        stmts = [Stmt(l, r, i) for l, r, i in meta["stmts"]]
        sv = {k: (tuple(v) if isinstance(v, list) else int(v)) for k, v in meta["state_vars"].items()}
        return cls(args, stmts, sv, meta["float_precision"], meta["func_name"])

    def save(self, path: str, **extra) -> None:
        d = self.to_npz_dict()
        d.update(extra)
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path: str) -> "Program":
        with np.load(path, allow_pickle=False) as d:
            return cls.from_npz_dict(d)

    def clone(self) -> "Program":
        buf = io.BytesIO()
        np.savez(buf, **self.to_npz_dict())
        buf.seek(0)
        with np.load(buf, allow_pickle=False) as d:
            return Program.from_npz_dict(d)
