"""JRC RK4 sweep in the reference's DEFAULT precision (float_precision='float32'): 2^20 instances x 1000 steps, kernel only."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, bench_configs as bc
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program
from oracle import numpy_oracle as orc

prog64 = Program.load("tests/golden/jrc.npz")
prog = bc.to_float32(prog64)
B, steps, ss = 2**20, 1000, 10
table = bench.param_table(1024, 1024, 0, B)
SHAPES = [(128, 8, True, False), (128, 6, True, True), (128, 8, True, True), (128, 10, True, True), (128, 12, True, True), (256, 4, True, True)]
if len(sys.argv) > 1:                      # e.g. 128x8 256x4  (fast_math kernels only)
    SHAPES = [tuple(int(x) for x in a.split("x")) + (True, True) for a in sys.argv[1:]]
for block, mb, cdm, fm in SHAPES:
    m = InstanceModel(prog, solver="rk4", sweep=bench.SWEEP, observers=[0], block=block, const_div_to_mul=cdm, min_blocks=mb, fast_math=fm)
    order = [(p.arg, p.elem) for p in m.rhs.params]
    params = np.stack([table[bench.SWEEP.index(k)] for k in order])
    s = m.session(B, steps // ss)
    s.reset(params=params); s.launch(100, 1e-4, ss); rt.synchronize(); s.reset(params=params)
    e0, e1 = rt.Event(), rt.Event()
    e0.record(); s.launch(steps, 1e-4, ss); e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1)
    out = s.download()
    info = m.module.kernel_info("prb_integrate", block)
    # parity of two instances against the float32 oracle (tolerance 1e-4)
    errs = []
    f = orc.build_vector_field(prog.source()); names = [a.name for a in prog.args]
    for b in (0, B - 1):
        u, C, C4, C08, _ = table[:, b]
        args = [np.array(a.value, copy=True) for a in prog.args]
        args[names.index("u")] = np.float32(u); args[names.index("weight_v2")] = np.array([C, C4], dtype=np.float32)
        args[names.index("weight")] = np.float32(C08); args[names.index("weight_v1")] = np.float32(C4)
        rec, _ = orc.run(f, args, steps * 1e-4, 1e-4, 1e-3, "rk4")
        errs.append(float(np.max(np.abs(out[:, 0, b] - rec[:, 0])) / np.max(np.abs(rec[:, 0]))))
    print(json.dumps({"expf_ftz": os.environ.get("PRB_EXPF_FTZ", "1"), "precision": "float32", "block": block, "min_blocks": mb, "const_div_to_mul": cdm, "fast_math": fm, "ms": round(ms, 3),
                      "inst_steps_per_s": B * steps / ms * 1e3, "node_steps_per_s": 3 * B * steps / ms * 1e3, "regs": info["regs"],
                      "occ_blocks": info["max_blocks_per_sm"], "max_scaled_err_vs_f32_oracle": errs}))
    s.free()
