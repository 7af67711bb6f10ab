"""Static A/B of emitter variants of the C2 (JRC, RK4, fast_math) sweep kernel, no GPU needed: for every `BLOCKxMINBLOCKS[:ENV=V,...]`
argument print the instruction mix of the executed time loop (fp64 / other / three-read fp64 instructions — the terms of the
issue model  2 F + O + R3  of DESIGN.md 2.1).  The same argument syntax drives `scripts/tune_inst.py` on the GPU.
    python scripts/sass_variants.py 64x7 64x7:PRB_SPEC_PER_STEP=0 64x7:PRB_EXP_LEAN=0,PRB_MAT_REGFIRST=0"""
import os, sys, json, tempfile
os.environ.setdefault('PRB_CACHE_DIR', 'off')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program
import sass_rf_model as M, collections
prog = Program.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "jrc.npz"))
sweep = [("u", 0), ("weight_v2", 0), ("weight_v2", 1), ("weight", 0), ("weight_v1", 0)]
env0 = dict(os.environ)
for spec in sys.argv[1:]:
    a, _, env = spec.partition(":")
    block, mb = (int(x) for x in a.split("x"))
    os.environ.clear(); os.environ.update(env0)
    if env: os.environ.update(dict(kv.split("=") for kv in env.split(",")))
    m = InstanceModel(prog, solver="rk4", sweep=sweep, observers=[0], block=block, const_div_to_mul=True, min_blocks=mb, fast_math=True)
    with tempfile.NamedTemporaryFile(suffix='.cubin', delete=False) as fh:
        fh.write(m.module.cubin)
    body = M.loop_body(M.parse(fh.name)); os.remove(fh.name)
    rows = M.rf_reads(body)
    c = collections.Counter(M.operands(s)[0].split('.')[0] for a_, s in body)
    F, R3 = len(rows), sum(1 for *_, n in rows if n == 3)
    print(spec, "model cycles", 2 * F + (len(body) - F) + R3, "instr", len(body), "fp64", len(rows), "other", len(body)-len(rows), "3-read", sum(1 for *_, n in rows if n == 3),
          "LDL/STL", c["LDL"] + c["STL"], dict(c.most_common(8)))
