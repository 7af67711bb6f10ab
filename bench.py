#!/usr/bin/env python
"""bench.py — headline benchmark of the PyRates numerical-integration hot path on B200.

Workload (BASELINE.json configs[1]): Jansen-Rit 3-population circuit (8 state
variables), classic RK4, dt=1e-4, T=1 (1e4 solver steps), grid_search over a 1024x1024 (C, u) grid =
2^20 parameter sets per GPU, observer pc/rpo_e_in/v sampled every 1e-3 -> [1000, 2^20] fp64.
One bench "step" = one complete simulation of all instances of this rank (1e4 solver steps each).
Instances shard across GPUs with no data-path collective (weak scaling: 2^20 instances per GPU).

Metric: state-updates/s = nodes x sweep-instances x solver-steps / s  (nodes = 3 for the JRC).

  value : device-resident inputs, kernels only (CUDA events on the launching stream)
  e2e   : the public BatchedSweep.run() call — host parameter table in, host trajectories out
          (pinned H2D of parameters + initial state, chunked D2H of all observers overlapped with compute)
  --impl reference : the reference algorithm on the host CPU (oracle port of the NumPy backend running
          the reference-generated vectorised grid program, tests/golden/jrc_grid_b1024.npz)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NODES = 3                      # JRC: pc, ein, iin
B_ALG = (2 * 8 + 5) * 8 + 8 / 10   # (2*n_sv + n_pp)*s + n_obs*s/store_step = 168.8 B / instance-step
SWEEP = [("u", 0), ("weight_v2", 0), ("weight_v2", 1), ("weight", 0), ("weight_v1", 0)]
# dram__bytes_read.sum + dram__bytes_write.sum of one prb_integrate launch (2^20 instances x 1000 steps, 100 samples)
# from the `ncu --set full` capture summarised in profiles/r01_c2_fastmath_prb_integrate_ncu_full.txt (110.79 + 847.40 MB;
# the exact kernel, profiles/r01_c2_prb_integrate_ncu_full.txt: 114.07 + 848.98 MB — the observer store either way).
NCU_TRAFFIC_BYTES_PER_LAUNCH = 958.2e6
# per-launch DRAM bytes by instance-steps per launch: 1000-step launches (--chunk-samples 100) and the default 250-step
# launches (--chunk-samples 25; profiles/r01_c2_fastmath_chunk25_ncu_full.txt)
NCU_TRAFFIC = {2 ** 20 * 1000: NCU_TRAFFIC_BYTES_PER_LAUNCH, 2 ** 20 * 250: 331.16e6}    # 109.19 + 221.97 MB (r01_c2_reassoc capture)
FP64_INSTR = {True: 265, False: 505}     # fp64-pipe instructions per instance-step in the loop (cuobjdump, fast / exact kernel)
METRIC = "state-updates/sec (nodes x sweeps x steps/s), JRC RK4 grid_search"
UNIT = "node-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def param_table(n_c: int, n_u: int, lo: int, hi: int) -> np.ndarray:
    """Rows follow SWEEP; instance order = pyrates.utility.linearize_grid(permute=True) (meshgrid, C fastest)."""
    C = np.linspace(68.0, 1350.0, n_c)
    u = np.linspace(120.0, 320.0, n_u)
    idx = np.arange(lo, hi)
    Cv, uv = C[idx % n_c], u[(idx // n_c) % n_u]
    return np.stack([uv, Cv, Cv / 4, 0.8 * Cv, Cv / 4])


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.3 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference(steps: int, warmup: int, sim_T: float):
    """Reference algorithm on the host: oracle port (NumPy) of BaseBackend's loop around the
    reference-generated *vectorised* grid program (B=1024, what pyrates.grid_search builds)."""
    from oracle import numpy_oracle as orc
    fx = orc.load_fixture(os.path.join(ROOT, "tests", "golden", "jrc_grid_b1024.npz"))
    B = 1024
    n_steps = int(round(sim_T / 1e-4))
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        orc.run_fixture(fx, solver="rk4", T=sim_T, dt=1e-4, dts=1e-3)
        if i >= warmup:
            times.append(time.perf_counter() - t)
    sec = float(np.mean(times))
    val = NODES * B * n_steps / sec
    sample = (f"B=1024 (32x32 (C,u) sub-grid), T={sim_T} ({n_steps} RK4 steps), reference-generated vectorised rhs "
              f"executed by the NumPy oracle port; NumPy elementwise kernels are single-threaded")
    return val, sec, {"value": val, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": "port", "sample": sample}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=1024, help="grid side: instances per GPU = grid*grid")
    ap.add_argument("--sim-time", type=float, default=1.0)
    ap.add_argument("--cpu-sim-time", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk-samples", type=int, default=25,
                    help="samples per launch: the D2H of chunk c overlaps the kernel of chunk c+1, the last chunk's copy is the e2e tail")
    ap.add_argument("--strict-div", action="store_true",
                    help="exact kernels: IEEE divisions, library exp (default: reciprocal multiplies, table exp, rcp + cubic "
                         "step, folded constant chains — each <= 2 ulp, trajectories within 1e-15 of the exact kernel)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    config = {"workload": "C2: Jansen-Rit 3-population circuit, RK4 dt=1e-4 T=%g, grid_search over %dx%d (C,u) grid per GPU, "
                          "observer pc/rpo_e_in/v every 1e-3" % (a.sim_time, a.grid, a.grid),
              "instances_per_gpu": a.grid * a.grid, "solver_steps": int(round(a.sim_time / 1e-4)), "n_state": 8,
              "nodes_per_instance": NODES, "sharding": f"instances x{world} (no collective)",
              "l2": "per-step working set (observer tensor 8 B x samples x instances) >> 126 MB L2; no flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        val, sec, cb = cpu_reference(a.steps, a.warmup, a.cpu_sim_time)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    from pyrates_b200 import runtime as rt
    from pyrates_b200.program import Program
    from pyrates_b200.sweep import BatchedSweep

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rt.set_device(local_rank)
    numa_node = None
    if world > 1:
        from pyrates_b200.distributed import bind_to_gpu_numa_node
        numa_node = bind_to_gpu_numa_node(local_rank)      # pinned result buffers next to this GPU's PCIe root

    def barrier():
        if dist is not None:
            dist.barrier()
        rt.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    prog = Program.load(os.path.join(ROOT, "tests", "golden", "jrc.npz"))
    B = a.grid * a.grid
    t_jit = time.perf_counter()
    sw = BatchedSweep(prog, SWEEP, observers=[0], solver="rk4", T=a.sim_time, dt=1e-4, dts=1e-3, n_inst=B,
                      chunk_samples=a.chunk_samples, const_div_to_mul=not a.strict_div)
    sw.model.module                      # JIT (NVRTC) happens here, outside every timed region
    jit_s = time.perf_counter() - t_jit
    table = param_table(a.grid, a.grid, rank * B, (rank + 1) * B)     # weak scaling: each rank its own grid block
    sw.stage_params(table)
    sw.upload()
    rt.synchronize()

    # ---- device-resident timing (value): inputs in HBM, kernels only
    for _ in range(W):
        sw.upload()
        sw.integrate(read_back=False)
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    wall0 = time.time()
    launches0 = rt.launch_count()
    ev = [(rt.Event(), rt.Event()) for _ in range(a.steps)]
    barrier()
    for k in range(a.steps):
        sw.upload()                      # re-initialise state on the device stream (not timed: before the start event)
        ev[k][0].record(sw.compute)
        sw.integrate(read_back=False)
        ev[k][1].record(sw.compute)
    barrier()
    kernel_ms = sum(e0.elapsed_ms(e1) for e0, e1 in ev)
    launches = rt.launch_count() - launches0
    kernel_ms = max_over_ranks(kernel_ms)
    n_steps = sw.steps
    value = NODES * B * world * n_steps * a.steps / (kernel_ms * 1e-3)

    # ---- end-to-end timing (e2e): host table in, host trajectories out, copies inside the timed region
    for _ in range(2):
        sw.run(table)
    barrier()
    t0 = time.perf_counter()
    for k in range(a.steps):
        out = sw.run(table)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    wall1 = time.time()
    e2e_val = NODES * B * world * n_steps * a.steps / e2e_s
    checksum = float(out[-1, 0, :8].sum())

    # ---- the same sweep with the observer reduced on the device (grid_search(reduce=True)): host table in, per-instance
    # mean / var / min / max / last out.  Informational — the headline e2e above returns every trajectory like the reference.
    red = BatchedSweep(prog, SWEEP, observers=[0], solver="rk4", T=a.sim_time, dt=1e-4, dts=1e-3, n_inst=B,
                       const_div_to_mul=not a.strict_div, reduce=True)
    red.model.module
    for _ in range(2):
        red.run(table)
    barrier()
    t0 = time.perf_counter()
    for k in range(a.steps):
        stats = red.run(table)
    barrier()
    red_s = max_over_ranks(time.perf_counter() - t0)
    red_info = {"value": NODES * B * world * n_steps * a.steps / red_s, "unit": UNIT, "ms_per_step": red_s * 1e3 / a.steps,
                "h2d_bytes_per_step": red.h2d_bytes, "d2h_bytes_per_step": red.d2h_bytes,
                "api": "pyrates_b200.sweep.BatchedSweep(reduce=True).run", "stats": ["mean", "var", "min", "max", "last"],
                "mean_matches_full_run": bool(np.allclose(stats[0, 0, :64], out[:, 0, :64].mean(axis=0), rtol=1e-10, atol=0))}
    red.free()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    hbm, which = peaks()
    inst_steps_per_launch = B * n_steps * a.steps / launches        # launches of the device-resident (value) timing
    launch_s = kernel_ms * 1e-3 / launches
    achieved = B_ALG * inst_steps_per_launch / launch_s / 1e9
    info = sw.model.module.kernel_info("prb_integrate", sw.model.kernel.block)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
        "ms_per_step": kernel_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "instance_steps_per_s": value / NODES,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": sw.h2d_bytes, "d2h_bytes_per_step": sw.d2h_bytes,
                "ms_per_step": e2e_s * 1e3 / a.steps, "api": "pyrates_b200.sweep.BatchedSweep.run", "checksum": checksum},
        "e2e_reduced": red_info,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": (NCU_TRAFFIC.get(int(inst_steps_per_launch)) if a.grid == 1024 else None),
                     "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_c2_reassoc_prb_integrate_ncu_full.txt)",
                     "algorithmic_bytes_per_launch": B_ALG * inst_steps_per_launch, "peak_source": which, "kernel": "prb_integrate (fused JRC rhs + RK4 + observer store)",
                     "algorithmic_bytes_per_instance_step": B_ALG, "launch_ms": launch_s * 1e3,
                     "note": "state/stages are register-resident across the launch: real DRAM traffic is only the "
                             "observer store, so the SURVEY 8d round-trip model can exceed 1.0; the true bound is the FP64 pipe",
                     "fp64_pipe": {"instr_per_instance_step": FP64_INSTR[not a.strict_div],
                                   "frac": FP64_INSTR[not a.strict_div] * 2.0 * (value / NODES / world / 32) / (148 * 4 * 1.965e9),
                                   "model": "fp64-pipe warp instructions (DFMA/DMUL/DADD/DSETP) in the SASS time loop x 2 issue "
                                            "cycles (16 lanes/SMSP) / (148 SMs x 4 SMSPs x 1.965 GHz)"}},
        "kernel": {"regs": info["regs"], "local_bytes": info["local_bytes"], "blocks_per_sm": info["max_blocks_per_sm"],
                   "block": sw.model.kernel.block, "const_div_to_mul": not a.strict_div, "fast_math": not a.strict_div},
        "jit_s": jit_s, "numa_node_rank0": numa_node,
        "clocks": clocks.stop(wall0, wall1) if clocks else None,
    }
    if world == 1 and not a.no_cpu_baseline:
        _, _, cb = cpu_reference(1, 0, a.cpu_sim_time)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    sw.free()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
