"""Multi-GPU plumbing for sweeps: one process per GPU (torchrun), instances sharded contiguously, results gathered.

Parameter sweeps are embarrassingly parallel: each rank integrates its slice of the grid on its own
GPU with no data-path collective; the only communication is the final gather of the recorded observers (and the
max-over-ranks reduction of timings in bench.py).  ``torch.distributed`` is plumbing here: NCCL on GPU boxes, gloo
in the CPU test-suite.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from .sweep import shard_bounds

__all__ = ["rank_world", "my_shard", "gather_instances"]


def rank_world() -> Tuple[int, int]:
    """(rank, world) from an initialised process group, else from torchrun's environment, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def my_shard(n_instances: int) -> Tuple[int, int]:
    r, w = rank_world()
    return shard_bounds(n_instances, w, r)


def gather_instances(local: np.ndarray, n_instances: int, dst: int = 0, device: Optional[str] = None) -> Optional[np.ndarray]:
    """Gather per-rank results ``[..., B_local]`` (instance axis last) to ``dst`` as ``[..., n_instances]``.
    Ranks hold the contiguous slices given by ``shard_bounds``; slices may differ by one instance, so the
    exchange is padded to the largest slice.  Returns None on the other ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    rank, world = dist.get_rank(), dist.get_world_size()
    bounds = [shard_bounds(n_instances, world, r) for r in range(world)]
    bmax = max(hi - lo for lo, hi in bounds)
    lead = local.shape[:-1]
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    pad = np.zeros(lead + (bmax,), dtype=local.dtype)
    pad[..., : local.shape[-1]] = local
    t = torch.from_numpy(np.ascontiguousarray(pad)).to(device)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    if dist.get_backend() == "nccl":
        # NCCL has no gather-to-one in older torch builds: all_gather is equivalent for this small control path
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t)
    else:
        dist.gather(t, bufs, dst=dst)
    if rank != dst:
        return None
    out = np.empty(lead + (n_instances,), dtype=local.dtype)
    for r, (lo, hi) in enumerate(bounds):
        out[..., lo:hi] = bufs[r].cpu().numpy()[..., : hi - lo]
    return out


def bind_to_gpu_numa_node(local_rank: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated
    afterwards (first touch) sit next to the GPU's PCIe root.  With 8 ranks streaming 8.4 GB each per sweep the
    D2H path is otherwise limited by cross-socket traffic.  Best effort: returns the node or None."""
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
