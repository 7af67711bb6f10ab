"""SASS of the time loop of an instance kernel, no GPU needed: instruction mix (fp64 pipe / MUFU / LDS / local memory /
total) for launch-shape and emitter variants of the C2 (JRC, RK4, fast_math) sweep kernel.
    python scripts/sass_loop.py [block min_blocks ...]"""
import collections, os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PRB_CACHE_DIR", "off")
import bench
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program


def loop_ops(cubin: bytes, fun: str = "prb_integrate"):
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(cubin)
    try:
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, f.name], capture_output=True, text=True).stdout
    finally:
        os.remove(f.name)
    lines = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    addr = lambda l: int(re.search(r"/\*([0-9a-f]{4,})\*/", l).group(1), 16)
    loop = None
    for l in lines:
        if "BRA" in l:
            t = re.search(r"0x([0-9a-f]+)", l.split("BRA")[1])
            if t and int(t.group(1), 16) < addr(l) and (loop is None or addr(l) - int(t.group(1), 16) > loop[1] - loop[0]):
                loop = (int(t.group(1), 16), addr(l))
    body = [l for l in lines if loop[0] <= addr(l) <= loop[1]]
    ops = [re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l) for l in body]
    return [o.group(1) for o in ops if o], body


def summarize(ops):
    base = collections.Counter(o.split(".")[0] for o in ops)
    fp64 = sum(base[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    return {"total": len(ops), "fp64": fp64, "MUFU": base["MUFU"], "LDS": base["LDS"], "local": base["STL"] + base["LDL"],
            "top": base.most_common(12)}


if __name__ == "__main__":
    prog = Program.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "jrc.npz"))
    shapes = [(64, 8), (128, 5)] if len(sys.argv) < 3 else [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
    for block, mb in shapes:
        m = InstanceModel(prog, solver="rk4", sweep=bench.SWEEP, observers=[0], block=block, const_div_to_mul=True, min_blocks=mb, fast_math=True)
        mod = rt.compile_module(m.kernel.source, "c2.cu", verbose_ptxas=True)
        regs = [l.strip() for l in mod.log.splitlines() if "Used" in l][:1]
        ops, body = loop_ops(mod.cubin)
        print(block, mb, regs, summarize(ops))
        if os.environ.get("DUMP"):
            print("\n".join(body))
