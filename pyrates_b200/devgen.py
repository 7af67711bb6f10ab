"""Constants that are generated on the device instead of being uploaded from host memory.

A ``DeviceUniform`` stands in for a large read-only argument of a Program (typically a dense connectivity): it has the
array attributes the network compiler looks at (``shape``, ``dtype``, ``ndim``, ``size``, ``nbytes``) but owns no memory.
``NetworkSession`` allocates the constant space on the device and lets the object fill its (row-sharded) slice with
``prb_fill_uniform`` (csrc/templates/fill_kernel.cuh) — BASELINE.json's Kuramoto configuration at N = 65536 has a 32 GiB
fp64 weight matrix that would otherwise be drawn on the host and pushed over PCIe by every rank.

The generator is counter-based (element e of the flattened array depends only on ``seed`` and ``e``), so ``host()`` — a
NumPy restatement of the same integer arithmetic — returns bit-identical values; the tests use it to pin the device
generator and to hand the same matrix to the oracle at sizes the oracle can run.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import runtime as rt

__all__ = ["DeviceUniform", "is_device_generated"]

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def is_device_generated(x) -> bool:
    return isinstance(x, DeviceUniform)


class DeviceUniform:
    """``scale * U[0, 1)`` of the given shape: value[e] = scale * (splitmix64(seed + (e + 1) * golden) >> 11) * 2**-53."""

    def __init__(self, shape: Tuple[int, ...], scale: float, seed: int, dtype=np.float64):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.scale = float(scale)
        self.seed = int(seed) & 0x7FFFFFFFFFFFFFFF
        self.ndim = len(self.shape)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize

    def astype(self, dtype) -> "DeviceUniform":
        return DeviceUniform(self.shape, self.scale, self.seed, dtype)

    # ------------------------------------------------------------------ NumPy twin (tests / oracle side)
    def host(self, row_lo: int = 0, row_hi: Optional[int] = None) -> np.ndarray:
        rows = self.shape[0] if self.shape else 1
        row_hi = rows if row_hi is None else row_hi
        cols = self.size // max(rows, 1)
        e = np.arange(row_lo * cols, row_hi * cols, dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = np.uint64(self.seed) + (e + np.uint64(1)) * _GOLDEN
            z = (z ^ (z >> np.uint64(30))) * _M1
            z = (z ^ (z >> np.uint64(27))) * _M2
            z = z ^ (z >> np.uint64(31))
        u = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        return (self.scale * u).astype(self.dtype).reshape((row_hi - row_lo,) + self.shape[1:])

    # ------------------------------------------------------------------ device side
    def fill(self, dptr: int, row_lo: int = 0, row_hi: Optional[int] = None, stream=None):
        """Write rows [row_lo, row_hi) of the array to device memory at ``dptr`` (asynchronous on ``stream``)."""
        rows = self.shape[0] if self.shape else 1
        row_hi = rows if row_hi is None else row_hi
        cols = self.size // max(rows, 1)
        n = (row_hi - row_lo) * cols
        if n <= 0:
            return
        a = rt.KernelArgs()
        a.n_steps, a.store_step, a.step0, a.t0 = 0, 1, int(row_lo) * cols, self.seed
        a.n_inst, a.dt, a.y = int(n), self.scale, int(dptr)
        sms = rt.device_info()["sm_count"]
        _fill_module(self.dtype).run("prb_fill_uniform", a, block=256, grid=int(min(sms * 8, -(-n // 256))), stream=stream)


_modules = {}


def _fill_module(dtype) -> rt.Module:
    key = np.dtype(dtype).name
    if key not in _modules:
        from .emit_inst import read_template
        real = "float" if np.dtype(dtype) == np.float32 else "double"
        src = read_template("prb_kernel_abi.cuh") + f"\ntypedef {real} real;\n" + read_template("fill_kernel.cuh")
        _modules[key] = rt.compile_module(src, f"prb_fill_{key}.cu")
    return _modules[key]
