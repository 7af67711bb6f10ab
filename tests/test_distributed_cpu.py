"""N>1 host logic on CPU: world_size-2 gloo process group — instance sharding covers the grid exactly once and the
result gather reassembles the instance axis in grid order."""
import os
import socket

import numpy as np
import pytest

from pyrates_b200.sweep import shard_bounds


def test_shard_bounds_partition():
    for n in (1, 2, 7, 16, 1000, 2 ** 20, 2 ** 20 + 3):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_inst, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyrates_b200.distributed import gather_instances, my_shard, rank_world
    assert rank_world() == (rank, world)
    lo, hi = my_shard(n_inst)
    # stand-in for the per-rank GPU integration: value encodes (sample, observer, global instance)
    s, o = np.meshgrid(np.arange(5), np.arange(2), indexing="ij")
    local = (s[..., None] * 1000.0 + o[..., None] * 100.0 + np.arange(lo, hi)[None, None, :]).astype(np.float64)
    full = gather_instances(local, n_inst, dst=0)
    if rank == 0:
        q.put(full)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_inst", [16, 17])
def test_gloo_world2_shard_and_gather(n_inst):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_inst, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    s, o = np.meshgrid(np.arange(5), np.arange(2), indexing="ij")
    want = s[..., None] * 1000.0 + o[..., None] * 100.0 + np.arange(n_inst)[None, None, :]
    np.testing.assert_array_equal(full, want)


def test_sweep_refuses_a_horizon_that_overflows_the_result_rows():
    """T not a multiple of the sampling step: the loop would record ceil(steps / store_step) samples into round(T / dts)
    rows.  The reference dies with an IndexError there (base_backend.py:723,730-733); BatchedSweep must refuse before it
    allocates anything (no write past the pinned result buffer)."""
    from conftest import golden_path
    from pyrates_b200.program import Program
    from pyrates_b200.sweep import BatchedSweep
    prog = Program.load(golden_path("jrc"))
    with pytest.raises(IndexError, match="records 4 samples"):
        BatchedSweep(prog, [("u", 0)], observers=[0], solver="euler", T=10.0, dt=1.0, dts=3.0, n_inst=4)


def _compile_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), PRB_CACHE_DIR="off")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyrates_b200 import runtime as rt
    calls = []
    real = rt.lib().prb_compile
    src = 'extern "C" __global__ void k(float* x) { x[threadIdx.x] *= 3.f; }'
    m = rt.compile_module(src, "shared.cu", shared=True)          # rank 0 compiles, rank 1 receives the cubin
    # a source only THIS rank has: keys differ -> every rank compiles locally instead of dead-locking
    own = rt.compile_module(f'extern "C" __global__ void k{rank}(float* x) {{ x[0] = {rank}.f; }}', "own.cu", shared=True)
    # differing cache states: only rank 1 has this module already — the collective must still line up on both ranks
    src3 = 'extern "C" __global__ void k3(float* x) { x[threadIdx.x] += 1.f; }'
    if rank == 1:
        rt.compile_module(src3, "third.cu")
    third = rt.compile_module(src3, "third.cu", shared=True)
    assert third.cubin[:4] == b"\x7fELF"
    q.put((rank, bytes(m.cubin), m.log, len(own.cubin)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_rank0_compiles_and_broadcasts_the_cubin():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_compile_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r, (c, log, n)) for r, c, log, n in (q.get(timeout=180) for _ in range(2)))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0][0] == got[1][0] and got[0][0][:4] == b"\x7fELF" and got[0][2] > 500 and got[1][2] > 500
