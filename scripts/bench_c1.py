"""C1 (BASELINE.json configs[0]): Montbrio QIF mean field, Euler dt=1e-3, T=100 (1e5 steps), ONE instance — the
reference's own CPU-runnable case.  A single thread integrates it (latency-bound, no roofline point): kernel time,
the end-to-end InstanceModel.run call (H2D + launch + D2H of the [1e5, 2] record) and the NumPy oracle on the host."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program
from oracle import numpy_oracle as orc

path = os.path.join(ROOT, "tests", "golden", "qif.npz")
prog = Program.load(path)
T, dt = 100.0, 1e-3
m = InstanceModel(prog, solver="euler")
m.module
s = m.session(1, 100000)
s.reset(); s.launch(1000, dt, 1); rt.synchronize(); s.reset()
e0, e1 = rt.Event(), rt.Event()
e0.record(); s.launch(100000, dt, 1); e1.record(); e1.synchronize()
kernel_ms = e0.elapsed_ms(e1)
s.free()
t0 = time.perf_counter(); out, _ = m.run(T, dt, None); e2e = time.perf_counter() - t0
fx = orc.load_fixture(path)
t0 = time.perf_counter(); rec, _ = orc.run_fixture(fx, solver="euler", T=T, dt=dt, dts=None); cpu = time.perf_counter() - t0
err = float(np.max(np.abs(out[:, :, 0] - rec) / (np.max(np.abs(rec), axis=0) + 1e-300)))
print(json.dumps({"config": "C1 QIF Euler dt=1e-3 T=100, 1 instance", "steps": 100000, "kernel_ms": kernel_ms,
                  "ns_per_step": kernel_ms * 1e6 / 1e5, "e2e_ms": e2e * 1e3, "node_steps_per_s_kernel": 1e5 / (kernel_ms * 1e-3),
                  "cpu_oracle_ms": cpu * 1e3, "speedup_kernel": cpu * 1e3 / kernel_ms, "speedup_e2e": cpu / e2e,
                  "max_scaled_err_vs_oracle": err}))
