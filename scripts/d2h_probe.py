"""Concurrent D2H ceiling of the box: every rank (torchrun, one per GPU) copies NBYTES from device memory into a pinned
host buffer in 210 MB pieces, no kernel running; wall clock, max over ranks.  Pinned buffer from cuMemHostAlloc vs the
huge-page mapping + cuMemHostRegister (runtime.PinnedBuffer), with and without binding to the GPU's NUMA node."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pyrates_b200 import runtime as rt

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
N = int(float(sys.argv[1]) if len(sys.argv) > 1 else 8.388608e9)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt.set_device(local)
chunk = 210 * 2 ** 20
d = rt.DeviceBuffer(chunk)
s = rt.Stream()


def sync_all():
    rt.synchronize()
    if world > 1:
        dist.barrier()


def mx(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"world": world, "bytes_per_rank": N}
for numa in (False, True):
    if numa:
        from pyrates_b200.distributed import bind_to_gpu_numa_node
        out["numa_node"] = bind_to_gpu_numa_node(local)
    for mode in ("cuda", "hugepage"):
        os.environ["PRB_PINNED"] = mode
        t = time.perf_counter()
        h = rt.PinnedBuffer((N,), np.uint8)
        alloc = mx(time.perf_counter() - t)
        best = 1e9
        for rep in range(3):
            sync_all()
            t = time.perf_counter()
            off = 0
            while off < N:
                n = min(chunk, N - off)
                rt._check(rt.lib().prb_memcpy_d2h(h.ptr + off, d.ptr, n, s.handle))
                off += n
            s.synchronize()
            sync_all()
            best = min(best, mx(time.perf_counter() - t))
        out[f"{mode}{'_numa' if numa else ''}"] = {"alloc_s": alloc, "aggregate_gbs": N * world / best / 1e9, "per_gpu_gbs": N / best / 1e9}
        h.free()
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
