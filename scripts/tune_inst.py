"""Occupancy / launch-bounds sweep of the C2 instance kernel (JRC, RK4, 2^20 instances)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program

prog = Program.load("tests/golden/jrc.npz")
sweep = [("u", 0), ("weight_v2", 0), ("weight_v2", 1), ("weight", 0), ("weight_v1", 0)]
B, steps, ss = 2**20, 1000, 10
Cv = np.linspace(68, 1350, B); u = np.linspace(120, 320, B)
out_ref = None
SHAPES = [(64, 7, True), (64, 8, True), (64, 10, True), (128, 4, True), (128, 5, True), (96, 6, True), (32, 16, True), (32, 20, True), (256, 2, True)]
YS = {}
ENVS = []
if len(sys.argv) > 1:
    SHAPES = []
    for a in sys.argv[1:]:                       # e.g. 64x7  64x10y  (y = base state in shared memory)  64x7:PRB_EXP_LEAN=0,PRB_SPEC_PER_STEP=1
        a, _, env = a.partition(":")
        shape = tuple(int(x) for x in a.rstrip("y").split("x")) + (True,)
        SHAPES.append(shape + ((a.endswith("y"),)))
        ENVS.append(dict(kv.split("=") for kv in env.split(",")) if env else {})
SHAPES = [s_ if len(s_) == 4 else s_ + (False,) for s_ in SHAPES]
ENVS = ENVS or [{}] * len(SHAPES)
_env0 = dict(os.environ)
for (block, mb, fm, ysm), env in zip(SHAPES, ENVS):
    os.environ.clear(); os.environ.update(_env0); os.environ.update(env)
    m = InstanceModel(prog, solver="rk4", sweep=sweep, observers=[0], block=block, const_div_to_mul=True, min_blocks=mb, fast_math=fm,
                      y_smem=ysm)
    vals = {("u", 0): u, ("weight_v2", 0): Cv, ("weight_v2", 1): Cv/4, ("weight", 0): 0.8*Cv, ("weight_v1", 0): Cv/4}
    params = np.stack([vals[(p.arg, p.elem)] for p in m.rhs.params])
    s = m.session(B, steps // ss)
    s.reset(params=params); s.launch(100, 1e-4, ss); rt.synchronize(); s.reset(params=params)
    all_ms = []
    for _ in range(3):
        s.reset(params=params)
        e0, e1 = rt.Event(), rt.Event()
        e0.record(); s.launch(steps, 1e-4, ss); e1.record(); e1.synchronize()
        all_ms.append(round(e0.elapsed_ms(e1), 3))
    ms = min(all_ms)
    info = m.module.kernel_info("prb_integrate", block)
    out = s.download()[:, 0, ::4099]
    if out_ref is None:
        out_ref = out
    dev = float(np.max(np.abs(out - out_ref)) / np.max(np.abs(out_ref)))
    print(json.dumps({"env": env, "block": block, "min_blocks": mb, "fast_math": fm, "y_smem": ysm, "max_dev_vs_first": dev, "ms": round(ms, 3), "all_ms": all_ms, "inst_steps_per_s": B*steps/ms*1e3, "regs": info["regs"],
                      "local": info["local_bytes"], "occ_blocks": info["max_blocks_per_sm"], "warps_per_sm": info["max_blocks_per_sm"]*block//32}))
    s.free()
