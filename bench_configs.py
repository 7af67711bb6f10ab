#!/usr/bin/env python
"""Secondary benchmark: the coupled-network configs of BASELINE.json (C3 WC dense, C4 QIF-SFA delays, C5 Kuramoto)
on the network execution model, each next to the NumPy oracle port on the host CPU.

Programs at the named sizes are obtained by *resizing* the golden fixture programs (tests/golden/*_n8.npz, captured
from the reference at N=8): the generated statements do not depend on N except through slice bounds and array
shapes, so every dimension equal to the fixture's N is mapped to the requested N and the arrays are re-drawn from
the seeded distributions SURVEY.md §8d prescribes.  (The reference itself is not available on the GPU box.)

    python bench_configs.py --config c3 --size 8192 --solver euler --steps 200
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pyrates_b200.program import Arg, Program, Stmt   # noqa: E402


def resize_program(prog: Program, n_old: int, n_new: int, fill) -> Program:
    """Map every array dimension == n_old to n_new and rescale the y/dy slice bounds accordingly.
    ``fill(name, shape, old_value)`` returns the new array for argument `name`."""
    f = n_new // n_old
    assert n_new % n_old == 0

    def scale_slices(text):
        return re.sub(r"\b(y|dy)\[(\d+):(\d+)\]", lambda m: f"{m.group(1)}[{int(m.group(2)) * f}:{int(m.group(3)) * f}]", text)

    stmts = [Stmt(s.lhs, scale_slices(s.rhs), None if s.lhs_idx is None else
                  re.sub(r"^(\d+):(\d+)$", lambda m: f"{int(m.group(1)) * f}:{int(m.group(2)) * f}", s.lhs_idx))
             for s in prog.stmts]
    args = []
    for a in prog.args:
        v = np.asarray(a.value)
        if a.kind in ("y", "dy"):
            shape = (v.size * f,)
        else:
            shape = tuple(n_new if d == n_old else d for d in v.shape)
        new = fill(a.name, shape, v)
        if hasattr(new, "fill") and hasattr(new, "host"):           # device-generated constant: no host array
            args.append(Arg(a.name, a.kind, new, a.label))
        else:
            args.append(Arg(a.name, a.kind, np.asarray(new, dtype=v.dtype).reshape(shape), a.label))
    return Program(args, stmts, {}, prog.float_precision, prog.func_name)


def _tile(old, shape):
    old = np.asarray(old)
    if old.shape == tuple(shape):
        return old
    if old.ndim == 0:
        return np.full(shape, old)
    reps = [int(np.ceil(s / o)) for s, o in zip(shape, old.shape)]
    return np.tile(old, reps)[tuple(slice(0, s) for s in shape)]


def make_c3(n: int) -> Program:
    """WC N-node dense network: W diag 15, off-diag N(0, 1/sqrt(N)), seed 42 (examples/benchmark_population_connectivity.py:31-39)."""
    base = Program.load(os.path.join(ROOT, "tests", "golden", "wc_n8.npz"))
    rng = np.random.default_rng(42)

    def fill(name, shape, old):
        if name == "weight":
            W = rng.normal(0.0, 1.0 / np.sqrt(n), shape)
            np.fill_diagonal(W, 15.0)
            return W
        if name == "y":
            return rng.uniform(0.05, 0.6, shape)
        if name == "theta":
            return rng.normal(2.0, 0.5, shape)
        return _tile(old, shape)
    return resize_program(base, 8, n, fill)


def make_c4(n: int) -> Program:
    """QIF-SFA, two populations of n units: p1->p2 gamma-kernel delay (cascade n=4), p2->p1 ring-buffer delay of 3 steps
    (SURVEY §8d C4 (iii)); eta ~ N(-5,1), W ~ N(0, 1/sqrt(n)), seed 42."""
    base = Program.load(os.path.join(ROOT, "tests", "golden", "qsfa_both_n8.npz"))
    rng = np.random.default_rng(42)

    def fill(name, shape, old):
        if name.startswith("weight"):
            return rng.normal(0.0, 1.0 / np.sqrt(n), shape)
        if name.startswith("eta"):
            return rng.normal(-5.0, 1.0, shape)
        if name == "r_buffer":
            return np.zeros(shape)
        return _tile(old, shape)
    return resize_program(base, 8, n, fill)


def make_c5(n: int, weight=None) -> Program:
    """Kuramoto, N phase oscillators, dense W ~ U(0, 2/N) with sin(theta_j - theta_i) pair coupling.
    ``weight``: a replacement for the matrix (e.g. a device-generated constant, pyrates_b200.devgen.DeviceUniform)."""
    base = Program.load(os.path.join(ROOT, "tests", "golden", "kuramoto_n8.npz"))

    def fill(name, shape, old):
        if name == "weight":
            if weight is not None:
                return weight
            return np.random.default_rng(2).uniform(0.0, 2.0 / n, shape)
        if name == "y":
            return np.random.default_rng(1).uniform(0, 2 * np.pi, shape)
        if name == "omega":
            return np.random.default_rng(1).normal(10, 1, shape)
        return _tile(old, shape)
    return resize_program(base, 8, n, fill)


# ----------------------------------------------------------------------------- the configs through the REAL front end
def import_pyrates():
    """An importable PyRates for the front end (templates -> IR -> compute graph), with backend='cuda' registered.
    On the GPU box that is the unmodified reference installed under baseline/_ref (plus the `ruamel.yaml` stand-in of
    baseline/shims: the image lacks that dependency of the reference).  Returns the module or None."""
    import importlib
    try:
        importlib.import_module("pyrates")
    except Exception:
        for sub in ("shims", "_ref"):
            d = os.path.join(ROOT, "baseline", sub)
            if os.path.isdir(d) and d not in sys.path:
                sys.path.insert(0, d)
        try:
            importlib.import_module("pyrates")
        except Exception:
            return None
    import pyrates_b200.register as reg
    if not reg.register():
        return None
    return sys.modules["pyrates"]


def _capture(circuit, dt: float, solver: str = "euler", precision: str = "float64") -> Program:
    """CircuitTemplate -> Program through the reference's own pipeline driving CudaBackend (the drop-in boundary)."""
    from pyrates.ir.node import clear_ir_caches
    func, args, _, _ = circuit.get_run_func("vector_field", step_size=dt, backend="cuda", solver=solver,
                                            float_precision=precision, verbose=False, clear=False)
    prog = func.program(args)
    try:
        circuit.clear()
    except Exception:
        pass
    clear_ir_caches()
    return prog


def circuit_c3(n: int):
    """WC population, dense connectivity: examples/benchmark_population_connectivity.py:31-39,86-106 (W diag 15, off-diagonal
    N(0, 1/sqrt(N)), seed 42); random initial rates / thresholds so that the trajectory leaves the template's fixed point."""
    from pyrates import CircuitTemplate, NodeTemplate
    from pyrates.frontend.template.population import Connectivity, PopulationTemplate
    rng = np.random.default_rng(42)
    W = rng.normal(0.0, 1.0 / np.sqrt(n), (n, n))
    np.fill_diagonal(W, 15.0)
    exc = NodeTemplate.from_yaml("model_templates.neural_mass_models.wilsoncowan.exc_pop")
    pop = PopulationTemplate(name="e", node=exc, n=n, params={"rate_op/r": list(rng.uniform(0.05, 0.6, n)),
                                                              "se_op/theta": list(rng.normal(2.0, 0.5, n))})
    return CircuitTemplate(name=f"wc_n{n}", populations={"e": pop},
                           connections=[Connectivity(source="e/rate_op/r", target="e/se_op/r_in", weights=W)])


def circuit_c4(n: int):
    """QIF-SFA, two populations of n units: p1 -> p2 gamma-kernel delay (mean 4 ms, spread 2 ms: cascade
    of 4 stages), p2 -> p1 ring-buffer delay of 3 steps; eta ~ N(-5, 1), W ~ N(0, 1/sqrt(n)), seed 42."""
    from pyrates import CircuitTemplate, NodeTemplate
    from pyrates.frontend.template.population import Connectivity, PopulationTemplate
    rng = np.random.default_rng(42)
    eta = rng.normal(-5.0, 1.0, n)
    W = rng.normal(0.0, 1.0 / np.sqrt(n), (n, n))
    W2 = rng.normal(0.0, 1.0 / np.sqrt(n), (n, n))
    qs = NodeTemplate.from_yaml("model_templates.neural_mass_models.qif.qif_sfa_pop")
    p1 = PopulationTemplate(name="p1", node=qs, n=n, params={"qif_sfa_op/eta": list(eta)})
    p2 = PopulationTemplate(name="p2", node=qs, n=n, params={"qif_sfa_op/eta": list(eta[::-1])})
    c12 = Connectivity(source="p1/qif_sfa_op/r", target="p2/qif_sfa_op/r_in", weights=W, delays=4e-3, spread=2e-3)
    c21 = Connectivity(source="p2/qif_sfa_op/r", target="p1/qif_sfa_op/r_in", weights=W2, delays=3e-3)
    return CircuitTemplate(name="qsfa_both", populations={"p1": p1, "p2": p2}, connections=[c12, c21])


def circuit_c5(n: int, dense: bool = True, W=None):
    """Kuramoto phase oscillators with sin(theta_j - theta_i) coupling (model_templates/oscillators/kuramoto.yaml:20-27,75-78):
    dense W ~ U(0, 2/N) (C5b) or the scalar all-to-all weight 1/N (C5a, lowered to a global `vsum`)."""
    from pyrates import CircuitTemplate, EdgeTemplate, NodeTemplate
    from pyrates.frontend.template.population import Connectivity, PopulationTemplate
    phase_pop = NodeTemplate.from_yaml("model_templates.oscillators.kuramoto.phase_pop")
    sin_edge = EdgeTemplate.from_yaml("model_templates.oscillators.kuramoto.sin_edge")
    theta = np.random.default_rng(1).uniform(0, 2 * np.pi, n)
    omega = np.random.default_rng(1).normal(10, 1, n)
    pop = PopulationTemplate(name="e", node=phase_pop, n=n, params={"phase_op/theta": list(theta), "phase_op/omega": list(omega)})
    if W is None:
        W = np.random.default_rng(2).uniform(0, 2.0 / n, (n, n)) if dense else 1.0 / n
    conn = Connectivity(source="e/phase_op/theta", target="e/phase_op/s_in", weights=W, edge=sin_edge,
                        edge_var_map={"theta_s": "source", "theta_t": "e/phase_op/theta"})
    return CircuitTemplate(name="kuramoto", populations={"e": pop}, connections=[conn])


def build_program(config: str, n: int, *, dt: float = 1e-3, front_end: str = "auto", device_w: bool = False):
    """Program of a BASELINE.json network config at size n -> (Program, how).  ``front_end``: 'reference' builds the circuit
    with PopulationTemplate + Connectivity and captures the program the reference's compute graph hands to CudaBackend;
    'resized' re-sizes the committed N = 8 fixture of the same circuit (no PyRates needed); 'auto' prefers the former.
    ``device_w`` (c5): the dense weight matrix becomes a device-generated U(0, 2/N) constant (devgen.DeviceUniform) — the
    32 GiB matrix of N = 65536 never exists in host memory; the circuit is built at N = 8 and re-sized."""
    if config == "c5" and device_w:
        from pyrates_b200.devgen import DeviceUniform
        prog = make_c5(n, weight=DeviceUniform((n, n), 2.0 / n, 2))
        return prog, "N=8 program of the reference re-sized, W generated on the device (splitmix64 U(0, 2/N))"
    if front_end in ("auto", "reference") and import_pyrates() is not None:
        circuit = {"c3": circuit_c3, "c4": circuit_c4, "c5": circuit_c5, "c5a": lambda k: circuit_c5(k, dense=False)}[config](n)
        return _capture(circuit, dt), "reference front end (PopulationTemplate + Connectivity -> ComputeGraph -> CudaBackend)"
    if front_end == "reference":
        raise RuntimeError("PyRates is not importable (expected the reference under baseline/_ref)")
    prog = {"c3": make_c3, "c4": make_c4, "c5": make_c5, "c5a": make_c5a}[config](n)
    return prog, "N=8 program of the reference re-sized (tests/golden fixtures)"


def make_c5a(n: int) -> Program:
    """Kuramoto with the scalar all-to-all weight 1/N: the coupling is a global `vsum`."""
    base = Program.load(os.path.join(ROOT, "tests", "golden", "kuramoto_scalar_n128.npz"))

    def fill(name, shape, old):
        if name == "y":
            return np.random.default_rng(1).uniform(0, 2 * np.pi, shape)
        if name == "omega":
            return np.random.default_rng(1).normal(10, 1, shape)
        old = np.asarray(old)
        if old.ndim == 0 and np.isclose(float(old), 1.0 / 128):
            return np.asarray(1.0 / n)
        return _tile(old, shape)
    return resize_program(base, 128, n, fill)


def to_float32(prog: Program) -> Program:
    """The same program at the reference's default float_precision='float32' (real arguments cast, text unchanged)."""
    args = [Arg(a.name, a.kind, np.asarray(a.value).astype(np.float32) if np.asarray(a.value).dtype == np.float64
                else np.asarray(a.value), a.label) for a in prog.args]
    return Program(args, list(prog.stmts), {}, "float32", prog.func_name)


CONFIGS = {"c3": (make_c3, 1, 1), "c4": (make_c4, 12, 2), "c5": (make_c5, 1, 1)}   # (builder, n_sv per unit, populations)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(CONFIGS))
    ap.add_argument("--size", dest="n", type=int, required=True)
    ap.add_argument("--solver", default="euler")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--check", action="store_true", help="compare the first cpu-steps against the oracle")
    ap.add_argument("--batch", type=int, default=1, help="sweep instances sharing the connectivity (grid_search over a network)")
    ap.add_argument("--sweep", default="", help="comma-separated swept scalar arguments (batch > 1)")
    ap.add_argument("--precision", default="float64", choices=["float64", "float32"])
    ap.add_argument("--tensor-core", default="auto", choices=["auto", "off"])
    ap.add_argument("--tc-max-acc", type=int, default=128, help="fp32 accumulator registers per epilogue thread (128 -> 256-wide tiles)")
    ap.add_argument("--no-split", action="store_true", help="one warp per row even when the GPU has fewer rows than warps")
    ap.add_argument("--no-deep", action="store_true", help="4 instead of 8 independent loads per lane in the fp64 row dot")
    ap.add_argument("--redundant", action="store_true", help="multi-GPU: small exchanges in the redundant-fetch form (every CTA polls)")
    ap.add_argument("--wcache", action="store_true", help="cache each CTA's W rows in shared memory when they fit (W-stationary; measured slower)")
    a = ap.parse_args()
    from pyrates_b200 import runtime as rt
    from pyrates_b200.netengine import NetworkModel
    rank, world, local_rank = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dist = None
    if world > 1:
        # one process per GPU; rows of W are sharded, stage vectors are all-gathered INSIDE the kernel over NVLink
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rt.set_device(local_rank)
    build, _, pops = CONFIGS[a.config]
    prog = build(a.n)
    if a.precision == "float32":
        prog = to_float32(prog)
    dt = 1e-3
    sweep = [x for x in a.sweep.split(",") if x]
    # multi-GPU: a single network shards the ROWS of W (in-kernel all-gather); a sweep over a network shards its
    # INSTANCES instead — every rank integrates its own `--batch` instances, no communication (weak scaling)
    shard_instances = world > 1 and a.batch > 1
    m = NetworkModel(prog, solver=a.solver, observers=[0], block=a.block, rank=0 if shard_instances else rank,
                     world=1 if shard_instances else world, n_batch=a.batch,
                     sweep=sweep, tensor_core="auto" if a.tensor_core == "auto" else False, tc_max_acc=a.tc_max_acc,
                     split_rows=not a.no_split, dot_deep=not a.no_deep, w_stationary=a.wcache, ll_redundant=a.redundant)
    params = None
    if sweep:
        params = m.param_defaults[:, None] * (1.0 + 0.1 * np.random.default_rng(3 + rank).uniform(-1, 1, (len(sweep), a.batch))) \
            + 0.05 * np.random.default_rng(4 + rank).uniform(0, 1, (len(sweep), a.batch))
    t0 = time.perf_counter()
    m.module
    jit_s = time.perf_counter() - t0
    s = m.session(a.steps)
    s.reset(params=params)
    s.launch(min(a.steps, 20), dt, 1)
    rt.synchronize()
    s.check()
    s.reset(params=params)
    if dist is not None:
        rt.synchronize()
        dist.barrier()
    e0, e1 = rt.Event(), rt.Event()
    e0.record()
    s.launch(a.steps, dt, 1)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_ms(e1)
    s.check()
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    evals = m.kernel.evals_per_step
    nodes = a.n * pops
    wbytes = sum(arr.size for arr in m.ir.arrays.values() if arr.space == "const" and len(arr.shape) == 2) * np.dtype(m.real).itemsize
    hbm = 6548.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = float(json.load(open(p))["hbm_gbs"])
    wbytes_rank = wbytes / (1 if shard_instances else world)
    inst_total = a.batch * (world if shard_instances else 1)
    achieved = wbytes_rank * evals * a.steps / (ms * 1e-3) / 1e9
    if rank != 0:
        s.free()
        dist.destroy_process_group()
        return
    line = {"config": a.config, "n": a.n, "solver": a.solver, "n_gpus": world, "steps": a.steps, "ms_per_step": ms / a.steps,
            "node_steps_per_s": nodes * inst_total * a.steps / (ms * 1e-3), "evals_per_step": evals, "W_bytes": wbytes,
            "sharding": ("instances (no collective)" if shard_instances else "rows of W (in-kernel all-gather)") if world > 1 else "none",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "model": "this rank's W shard bytes x evaluations per step (state vectors negligible); per GPU"},
            "jit_s": jit_s, "phases": m.kernel.n_phases, "block": m.kernel.block, "batch": a.batch, "precision": a.precision,
            "split_rows": not a.no_split, "dot_deep": not a.no_deep, "w_stationary": a.wcache and "W rows cached" in m.kernel.source,
            "exchange": ("redundant" if "#define PRB_LL_REDUNDANT 1" in m.kernel.source else "cooperative") if world > 1 and not shard_instances else None}
    if a.batch > 1 and params is not None:
        # end to end through the host API of a network sweep: swept parameters + initial state from host memory in,
        # recorded trajectories [n_samples, n_obs, B] back in host memory (constants stay resident, like a model would
        # across grid_search chunks)
        e2e_steps = a.steps
        t_e = time.perf_counter()
        s.reset(params=params)
        s.samples_done = 0
        s.launch(e2e_steps, dt, 1)
        host = s.download()
        e2e_s = time.perf_counter() - t_e
        line["e2e"] = {"value": nodes * a.batch * e2e_steps / e2e_s, "unit": "node-steps/s", "steps": e2e_steps,
                       "h2d_bytes_per_step": int((params.nbytes + m.n_state * a.batch * np.dtype(m.real).itemsize) / e2e_steps),
                       "d2h_bytes_per_step": int(host[:e2e_steps].nbytes / e2e_steps), "ms_per_step": e2e_s * 1e3 / e2e_steps,
                       "api": "NetworkSession.reset / launch / download (what NetworkSweep.run does per call)"}
    if a.batch > 1:
        # sweep-batched coupling is GEMM-shaped: 2 flops per (target, source, instance) and evaluation
        flops = sum(2.0 * r.n * r.m for r in m.ir.reductions if r.kind == "row") * a.batch * evals * a.steps     # per GPU
        tf = flops / (ms * 1e-3) / 1e12
        if m.kernel.tc is not None:
            peak = 1388.2
            if os.path.exists(p):
                peak = float(json.load(open(p)).get("bf16_tflops_sustained", peak))
            line["roofline"] = {"bound": "tensor", "achieved": 3 * tf, "peak": peak / 2, "unit": "TFLOP/s", "frac": 3 * tf / (peak / 2),
                                "model": "tf32 tensor-pipe flops = 3 x (2 N_t N_s B) per evaluation (3xTF32 error-compensated "
                                         "products); peak = measured sustained bf16 dense / 2 (tf32 runs at half the bf16 rate)",
                                "fp32_equivalent_tflops": tf}
            line["tc"] = m.kernel.tc
        else:
            fma_peak = 148 * 128 * 2 * 1.965e9 / 1e12 / (1 if a.precision == "float32" else 2)
            line["roofline"] = {"bound": "fma", "achieved": tf, "peak": fma_peak, "unit": "TFLOP/s", "frac": tf / fma_peak,
                                "model": "2 N_t N_s B flops per evaluation on the FMA pipe (nominal 148 SMs x 128 (64 fp64) FMA/clk x 1.965 GHz)"}
    if a.cpu_steps and a.batch == 1:
        from oracle import numpy_oracle as orc
        f = orc.build_vector_field(prog.source())
        args = [np.array(x.value, copy=True) for x in prog.args]
        args[0] = args[0][()]
        t0 = time.perf_counter()
        rec, _ = orc.run(f, args, a.cpu_steps * dt, dt, None, a.solver if a.solver != "heun" else "heun")
        cpu_s = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": nodes * a.cpu_steps / cpu_s, "unit": "node-steps/s", "kind": "port",
                                "cores": os.cpu_count(), "sample": f"{a.cpu_steps} steps, NumPy/OpenBLAS default threads"}
        if a.check:
            out = s.download()[: a.cpu_steps, 0, 0]
            line["max_abs_diff_vs_oracle_obs0"] = float(np.max(np.abs(out - rec[:, 0])))
    print(json.dumps(line))
    s.free()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
