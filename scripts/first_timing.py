"""Scratch timing of the C2 instance kernel (JRC sweep, RK4) — superseded by bench.py."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program

print(rt.device_info())
prog = Program.load("tests/golden/jrc.npz")
sweep = [("u", 0), ("weight_v2", 0), ("weight_v2", 1), ("weight", 0), ("weight_v1", 0)]
for B in (2**16, 2**20):
  for solver in ("euler", "heun", "rk4"):
    for block in (128, 256):
      for cdm in (False, True):
        m = InstanceModel(prog, solver=solver, sweep=sweep, observers=[0], block=block, const_div_to_mul=cdm)
        t0 = time.time(); mod = m.module; tj = time.time() - t0
        steps, ss = 2000, 10
        s = m.session(B, steps // ss)
        Cv = np.linspace(68, 1350, B); u = np.linspace(120, 320, B)
        vals = {("u", 0): u, ("weight_v2", 0): Cv, ("weight_v2", 1): Cv/4, ("weight", 0): 0.8*Cv, ("weight_v1", 0): Cv/4}
        params = np.stack([vals[(p.arg, p.elem)] for p in m.rhs.params])
        s.reset(params=params)
        s.launch(100, 1e-4, ss); rt.synchronize()
        s.reset(params=params)
        e0, e1 = rt.Event(), rt.Event()
        e0.record(); s.launch(steps, 1e-4, ss); e1.record(); e1.synchronize()
        ms = e0.elapsed_ms(e1)
        info = mod.kernel_info("prb_integrate", block)
        print(json.dumps({"B": B, "solver": solver, "block": block, "const_div_to_mul": cdm, "ms": ms, "jit_s": round(tj, 2),
                          "inst_steps_per_s": B * steps / ms * 1e3, "regs": info["regs"], "occ_blocks": info["max_blocks_per_sm"]}))
        s.free()
