"""torchrun worker of tests/test_gpu_multigpu.py: one process per GPU, rows of every index space sharded across the
ranks, stage vectors exchanged inside the persistent kernel (LL packets over NVLink, net_kernel.cuh: prb_ll_*).
Every rank holds a complete replica of the state, so EVERY rank compares its full recorded state with the oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import bench_configs as bc
    from conftest import golden_path, traj_rel_err
    from oracle import numpy_oracle as orc
    from pyrates_b200 import runtime as rt
    from pyrates_b200.netengine import NetworkModel
    from pyrates_b200.program import Program

    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt.set_device(local)

    def oracle_run(prog, T, dt, dts, solver):
        f = orc.build_vector_field(prog.source())
        args = [np.array(a.value, copy=True) for a in prog.args]
        args[0] = args[0][()]
        return orc.run(f, args, T, dt, dts, solver)[0]

    cases = []
    # (name, program, solver, T, dt, dts, tolerance)
    cases.append(("c3_wc_n1024_rk4", bc.make_c3(1024), "rk4", 12e-3, 1e-3, None, 1e-9))            # dense dot(W, r)
    cases.append(("c3_wc_n1024_euler", bc.make_c3(1024), "euler", 30e-3, 1e-3, None, 1e-9))
    for name, solver in (("kuramoto_scalar_n128", "rk4"),      # global vsum coupling: block-redundant full reduction over the replica
                         ("kuramoto_n128", "euler"),           # separable pair coupling: 2 phases, materialised cos / sin vectors
                         ("qsfa_both_n96", "heun"),            # gamma cascade + delay ring buffer (ring stores exchanged)
                         ("wc_n128", "heun_true"),
                         ("jrc_grid_b16", "heun"),             # gather / scatter edges of a vectorised circuit
                         ("jrc_2delay", "rk4"),                # scatter stores u[target_idx] = ...
                         ("net10", "heun"),                    # legacy per-edge delays: gathered ring reads buf[source_idx, delays]
                         ("net13", "euler")):
        fx = orc.load_fixture(golden_path(name))
        r = fx["run"]
        cases.append((name, Program.load(golden_path(name)), solver, r["T"], r["dt"], r["dts"], 1e-9))
    cases.append(("c4_qsfa_2x512_heun", bc.make_c4(512), "heun", 10e-3, 1e-3, None, 1e-9))
    # global vsum coupling over a long vector: two-level reduction, per-CTA partials of every rank exchanged as LL packets
    cases.append(("c5a_vsum_n8192_rk4", bc.make_c5a(8192), "rk4", 20e-3, 1e-3, None, 1e-9))
    worst = {}
    ok = True
    for name, prog, solver, T, dt, dts, tol in cases:
        ref = oracle_run(prog, T, dt, dts, solver)
        before = rt.launch_count()
        out, y_final = NetworkModel(prog, solver=solver, rank=rank, world=world).run(T, dt, dts)
        assert rt.launch_count() == before + 1
        err = traj_rel_err(out[:, :, 0], ref)
        worst[name] = err
        if not (err <= tol):
            ok = False
            print(f"[rank {rank}] {name}: rel err {err:.3e} > {tol}", flush=True)
    errs = [None] * world
    dist.all_gather_object(errs, worst)
    if rank == 0:
        print("MULTIGPU_RESULT " + json.dumps({"world": world, "ok": ok, "errors_by_rank": errs}), flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
