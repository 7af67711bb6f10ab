"""Adaptive sweep throughput: JRC, solver='scipy' (RK45 per thread), B instances on a (C, u) grid, T = 1.
Prints one JSON line: step attempts/s (6 rhs evaluations each), node-steps/s, per-instance attempt statistics, and the
oracle (SciPy restatement, one instance on the host) for the same problem."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program

side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rtol, atol, T, dt, dts = 1e-6, 1e-8, 1.0, 1e-4, 1e-3
prog = Program.load(os.path.join(ROOT, "tests", "golden", "jrc_rk45.npz"))
B = side * side
m = InstanceModel(prog, solver="scipy", sweep=bench.SWEEP, observers=[0])
table = bench.param_table(side, side, 0, B)
order = [(p.arg, p.elem) for p in m.rhs.params]
params = np.stack([table[bench.SWEEP.index(o)] for o in order])
s = m.session(B, int(round(T / dts)))
s.reset(params=params)
s.launch_adaptive(T, dt, dts, rtol=rtol, atol=atol)            # warm-up (JIT, clocks)
s.reset(params=params)
e0, e1 = rt.Event(), rt.Event()
e0.record(); s.launch_adaptive(T, dt, dts, rtol=rtol, atol=atol); e1.record(); e1.synchronize()
ms = e0.elapsed_ms(e1)
att = s.attempts
line = {"workload": f"JRC RK45 (solver='scipy', rtol={rtol}, atol={atol}), T={T}, {side}x{side} (C,u) grid", "instances": B,
        "ms": ms, "attempts_total": int(att.sum()), "attempts_min_mean_max": [int(att.min()), float(att.mean()), int(att.max())],
        "attempts_per_s": float(att.sum() / (ms * 1e-3)), "node_steps_per_s": float(3 * att.sum() / (ms * 1e-3)),
        "warp_divergence_overhead": float(att.reshape(-1, 32).max(axis=1).sum() * 32 / att.sum())}
from oracle import numpy_oracle as orc
fx = orc.load_fixture(os.path.join(ROOT, "tests", "golden", "jrc_rk45.npz"))
f = orc.build_vector_field(fx["source"]); n = [0]
def g(*a):
    n[0] += 1
    return f(*a)
args = [np.array(a, copy=True) for a in fx["args"]]; args[0] = args[0][()]
t0 = time.perf_counter(); orc.run_adaptive(g, args, T, dt, dts, rtol=rtol, atol=atol); cpu = time.perf_counter() - t0
line["cpu_oracle"] = {"attempts_per_s": (n[0] - 1) / 6 / cpu, "instances": 1, "kind": "port (SciPy RK45 restatement around the reference rhs)"}
print(json.dumps(line))
s.free()
