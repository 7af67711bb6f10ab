"""Test helper: evaluate a scalarised rhs DAG with NumPy scalars (CPU, slow, tests only).

Lets the `-m "not gpu"` suite validate the scalarising front end (``pyrates_b200.scalarize``)
against the oracle without a GPU.  Not part of the product path.
"""
import numpy as np


class DagMachine:
    def __init__(self, rhs, real=np.float64):
        self.rhs = rhs
        self.real = real
        self.rings = [np.array(r.init, dtype=real) for r in rhs.rings]      # physical (rows, L)
        self.heads = [0 for _ in rhs.rings]
        self.slots = np.array([s[2] for s in rhs.slots], dtype=real)
        self.params = np.array([p.default for p in rhs.params], dtype=real)
        self.order = rhs.live_nodes()
        self.hist = None              # delay-differential programs: oracle History over the FULL state (set by the drivers)

    def eval(self, t, y):
        g, real = self.rhs.graph, self.real
        val = {}
        with np.errstate(all="ignore"):
            for i in self.order:
                op, a = g.ops[i], g.args[i]
                if op == "const":
                    v = g.const_value(i)
                    val[i] = int(v) if g.isint[i] else real(v)
                elif op == "state":
                    val[i] = real(y[g.payload[i]])
                elif op == "param":
                    val[i] = self.params[g.payload[i]]
                elif op == "slot":
                    val[i] = self.slots[g.payload[i]]
                elif op == "t":
                    val[i] = int(t) if g.isint[i] else float(t)
                elif op == "hist":
                    # the reference-generated lookup hist(t*dt - d)[idx] / hist(t - d)[idx] (base_backend.py:437-444)
                    hv, scale, dkey, strong = g.payload[i]
                    d = float(self.params[dkey[1]]) if isinstance(dkey, tuple) else float.fromhex(dkey)
                    if scale is not None:
                        tq = int(t) * scale - d
                    elif strong and real is np.float32:
                        tq = float(np.float32(t) - np.float32(d))
                    else:
                        tq = float(t) - d
                    val[i] = real(self.hist(tq)[self.rhs.hist_vars[hv]])
                elif op == "table":
                    tab, tnode, col, ncols = g.payload[i]
                    arr = np.asarray(self.rhs.tables[tab].value).reshape(-1)
                    val[i] = real(arr[int(val[tnode]) * ncols + col])
                elif op == "ring":
                    ring, row, col, shift = g.payload[i]
                    L = self.rhs.rings[ring].length
                    val[i] = self.rings[ring][row, (col - (self.heads[ring] + shift)) % L]
                elif op == "add":
                    val[i] = val[a[0]] + val[a[1]]
                elif op == "sub":
                    val[i] = val[a[0]] - val[a[1]]
                elif op == "mul":
                    val[i] = val[a[0]] * val[a[1]]
                elif op == "div":
                    val[i] = real(val[a[0]]) / real(val[a[1]])
                elif op == "neg":
                    val[i] = -val[a[0]]
                elif op == "pow":
                    val[i] = np.power(real(val[a[0]]), real(val[a[1]]))
                elif op in ("maximum", "minimum"):
                    val[i] = getattr(np, op)(val[a[0]], val[a[1]])
                else:
                    val[i] = real(getattr(np, {"abs": "abs"}.get(op, op))(real(val[a[0]])))
        dy = np.array([val[n] for n in self.rhs.dy], dtype=real)
        for ring, row, col, node in self.rhs.ring_writes:
            spec = self.rhs.rings[ring]
            self.rings[ring][row, (col - (self.heads[ring] + spec.rolls_per_eval)) % spec.length] = val[node]
        for k, r in enumerate(self.rhs.rings):
            self.heads[k] = (self.heads[k] + r.rolls_per_eval) % r.length
        for slot, node in self.rhs.slot_writes:
            self.slots[slot] = val[node]
        return dy

    def euler(self, y0, t0, n_steps, dt, store_step=1, hist=None):
        y = np.array(y0, dtype=self.real)
        self.hist = hist
        rec = []
        for i in range(n_steps):
            if i % store_step == 0:
                rec.append(y.copy())
            y = y + self.real(dt) * self.eval(i + t0, y)
            if hist is not None:
                hist.update((i + 1) * dt, y)
        return np.array(rec)
