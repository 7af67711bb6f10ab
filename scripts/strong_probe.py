"""One GPU's share of the strong-scaling C2 record (2^20 grid points over 8 GPUs = 131072 instances): time chunk sizes.
    python scripts/strong_probe.py [instances] [chunk_samples ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyrates_b200 import runtime as rt
from pyrates_b200.program import Program
from pyrates_b200.sweep import BatchedSweep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 17
chunks = [int(x) for x in sys.argv[2:]] or [25, 50, 100, 200, 500]
prog = Program.load(os.path.join(bench.ROOT, "tests", "golden", "jrc.npz"))
table = bench.param_table(1024, 1024, 0, B)
for cs in chunks:
    sw = BatchedSweep(prog, bench.SWEEP, observers=[0], solver="rk4", T=1.0, dt=1e-4, dts=1e-3, n_inst=B, chunk_samples=cs)
    sw.model.compile()
    sw.stage_params(table)
    for _ in range(3):
        sw.upload(); sw.integrate(read_back=False)
    rt.synchronize()
    ev = [(rt.Event(), rt.Event()) for _ in range(5)]
    n0 = rt.launch_count()
    for k in range(5):
        sw.upload()
        ev[k][0].record(sw.compute); sw.integrate(read_back=False); ev[k][1].record(sw.compute)
    rt.synchronize()
    ms = sum(a.elapsed_ms(b) for a, b in ev) / 5
    launches = (rt.launch_count() - n0) // 5
    for _ in range(2):
        sw.run(table)
    t0 = time.perf_counter()
    for _ in range(5):
        sw.run(table)
    e2e = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"instances": B, "chunk_samples": cs, "launches_per_sim": launches, "ms_per_sim": round(ms, 3), "e2e_ms_per_sim": round(e2e, 3),
                      "ideal_ms": round(182.8 * B / 2 ** 20, 3)}))
    sw.free()
