"""Host-side driver of the network execution model (persistent cooperative kernel, one network instance).

Counterpart of ``engine.InstanceModel`` for models whose state does not fit one thread: builds the phased
kernel (``netcompile`` + ``emit_net``), owns the device buffers (packed constants incl. the connectivity
matrices, work/stage vectors, ring buffers, grid-barrier word) and launches ``prb_integrate_net`` through
the C ABI with ``cooperative=1``.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np

from . import runtime as rt
from .emit_net import NetKernel, emit_net_kernel, tc_config
from .engine import ADAPTIVE_SOLVERS, FIXED_STEP_SOLVERS, n_steps_for
from .netcompile import NetIR, netcompile
from .program import Program
from .scalarize import UnsupportedModelError

__all__ = ["NetworkModel", "NetworkSession", "build_check", "build_check_sweeps"]


class NetworkModel:
    def __init__(self, prog: Program, *, solver: str = "euler", observers: Optional[Sequence[int]] = None,
                 block: int = 512, fmad: bool = True, blocks_per_sm: int = 1, rank: int = 0, world: int = 1,
                 separable: bool = True, n_batch: int = 1, sweep: Sequence[str] = (), tensor_core="auto",
                 tc_max_acc: int = 128, split_rows: bool = True, dot_deep: bool = True, w_stationary: bool = False,
                 ll_redundant: bool = False):
        """``n_batch`` > 1 integrates that many sweep instances of the network together (shared connectivity);
        ``sweep`` names the scalar / uniform-vector arguments that vary per instance.  ``tensor_core``: run the batched
        dense couplings as tcgen05 3xTF32 GEMMs (float32 models, batch a multiple of 32; "auto" = whenever that
        applies), otherwise as fp32/fp64 FMA register tiles (one pass over W per 8 instances)."""
        if solver not in FIXED_STEP_SOLVERS + ADAPTIVE_SOLVERS:
            raise rt.CudaBackendError(f"Backend `CudaBackend` does not support solver `{solver}`. "
                                      f"Supported solvers: {list(FIXED_STEP_SOLVERS + ADAPTIVE_SOLVERS)}.")
        self.prog = prog
        self.solver = solver
        self.adaptive = solver in ADAPTIVE_SOLVERS
        if prog.is_complex:
            raise UnsupportedModelError("complex-valued models are supported by the instance execution model only")
        self.precision = prog.float_precision
        self.real = prog.real_dtype
        self.n_batch = int(n_batch)
        self.ir: NetIR = netcompile(prog, separable=separable, sweep=sweep)
        self.param_names = [n for n, _ in self.ir.params]
        self.param_defaults = np.array([d for _, d in self.ir.params], dtype=np.float64)
        self.observers = list(range(prog.n_state)) if observers is None else [int(o) for o in observers]
        self.rank, self.world = int(rank), int(world)
        if not 0 <= self.rank < self.world or self.world > 32:
            raise ValueError(f"rank {rank} / world {world}: row-sharded network runs support 1..32 GPUs of one NVLink domain "
                             f"(32 heartbeat slots in the receive mirror)")
        if tensor_core == "auto":
            tensor_core = self.n_batch > 1 and tc_config(self.ir, self.precision, self.n_batch, self.world) is not None
        self.kernel: NetKernel = emit_net_kernel(self.ir, solver=solver, precision=self.precision,
                                                 observers=self.observers, block=block, rank=self.rank, world=self.world,
                                                 n_batch=self.n_batch, tensor_core=bool(tensor_core), tc_max_acc=tc_max_acc,
                                                 split_rows=split_rows, dot_deep=dot_deep, w_stationary=w_stationary,
                                                 ll_redundant=ll_redundant)
        self.fmad = fmad
        self.blocks_per_sm = blocks_per_sm
        self._module: Optional[rt.Module] = None
        t0 = np.asarray(prog.arg_by_kind("t").value)
        self.t0 = float(t0) if self.adaptive else int(t0)

    @property
    def module(self) -> rt.Module:
        if self._module is None:
            self._module = rt.compile_module(self.kernel.source, f"prb_net_{self.solver}.cu", fmad=self.fmad)
        return self._module

    @property
    def n_state(self) -> int:
        return self.kernel.n_state

    def session(self, n_samples: int, stream=None) -> "NetworkSession":
        return NetworkSession(self, n_samples, stream)

    def run(self, T: float, dt: float, dts: Optional[float] = None, *, y0: Optional[np.ndarray] = None,
            params: Optional[np.ndarray] = None, **solver_kwargs):
        """Returns ``(out[n_samples, n_obs, n_batch], y_final[n_state, n_batch])``.  ``solver='scipy'`` accepts SciPy's
        ``rtol`` / ``atol`` / ``max_step`` / ``method='RK45'`` (base_backend.py:802-812)."""
        steps, store_step, n_samples = n_steps_for(T, dt, dts)
        if self.adaptive:
            n_samples = int(round(T / (dts if dts else dt)))
        s = self.session(n_samples)
        try:
            s.reset(y0, params)
            if self.adaptive:
                s.launch_adaptive(T, dt, dts, **solver_kwargs)
            else:
                s.launch(steps, dt, store_step)
            out = s.download()
            y_final = s.download_state()
        finally:
            s.free()
        return out, y_final


class NetworkSession:
    def __init__(self, model: NetworkModel, n_samples: int, stream=None):
        self.m = model
        k = model.kernel
        real = model.real
        self.stream = stream
        self.n_samples = int(n_samples)
        item = np.dtype(real).itemsize
        self.B = B = model.n_batch
        self.d_y = rt.DeviceBuffer(max(k.n_state * B, 1) * item)
        self.d_out = rt.DeviceBuffer(max(self.n_samples * k.n_obs * B, 1) * item)
        self.d_params = rt.DeviceBuffer(max(len(model.param_names) * B, 1) * item)
        self.d_consts = rt.DeviceBuffer(max(k.const_elems, k.const_init.size, 1) * item)
        self.d_consts.upload(np.asarray(k.const_init, dtype=real))
        for off, gen, lo, hi in k.const_fills:                # device-generated connectivity: this rank's rows only
            gen.fill(self.d_consts.ptr + off * item, lo, hi, stream)
        self.d_iconsts = rt.DeviceBuffer.from_host(k.iconst_init.astype(np.int32))
        self.d_tables = rt.DeviceBuffer.from_host(k.table_init.astype(real))
        self.d_work = rt.DeviceBuffer(max(k.work_elems, 1) * item)
        self.d_rings = rt.DeviceBuffer(max(k.ring_init.size * B, 1) * item)
        self.d_sync = rt.DeviceBuffer(8 * 16)
        self.sm_count = rt.device_info()["sm_count"]
        if self.sm_count * model.blocks_per_sm > 256 and "__part" in "".join(model.ir.arrays):
            raise rt.CudaBackendError("two-level full reductions hold 256 per-CTA partials per rank; the persistent grid is larger")
        self.steps_done = 0
        self.samples_done = 0
        self.d_peers = None
        self.d_mirror = None
        self._peer_ptrs = []
        if model.world > 1:
            self._connect_peers()

    def _connect_peers(self):
        """Multi-GPU (one process per GPU, rows sharded): allocate this rank's receive mirror (two parities of one 16-byte
        packet slot per work / ring element, net_kernel.cuh: prb_ll_*), exchange CUDA IPC handles of (mirror, sync) with the
        other ranks and build the device table the kernels use for NVLink peer stores.  torch.distributed is only the
        control plane here."""
        import torch.distributed as dist
        m = self.m
        if not dist.is_initialized() or dist.get_world_size() != m.world or dist.get_rank() != m.rank:
            raise rt.CudaBackendError("multi-GPU network runs need an initialised torch.distributed group matching rank/world")
        self.d_mirror = rt.DeviceBuffer(2 * 16 * m.kernel.ll_slots)
        mine = [rt.ipc_handle(b) for b in (self.d_mirror, self.d_sync)]
        everyone = [None] * m.world
        dist.all_gather_object(everyone, mine)
        table = np.zeros(2 * m.world, dtype=np.uint64)
        own = (self.d_mirror.ptr, self.d_sync.ptr)
        for p in range(m.world):
            for k in range(2):
                if p == m.rank:
                    ptr = own[k]
                else:
                    ptr = rt.ipc_open(everyone[p][k])
                    self._peer_ptrs.append(ptr)
                table[k * m.world + p] = ptr
        self.d_peers = rt.DeviceBuffer.from_host(table)
        # JIT + module load happen here, on every rank, BEFORE any launch-time barrier: the in-kernel waits are bounded, and
        # a rank that still compiles while its peers already spin would otherwise trip them
        m.module.kernel_info(m.kernel.kernel_name, m.kernel.block, m.kernel.smem_bytes)
        dist.barrier()

    def reset(self, y0: Optional[np.ndarray] = None, params: Optional[np.ndarray] = None):
        """Upload initial state ([n_state] or [n_state, B]), swept parameters ([n_params, B]) and buffer images;
        every batched array is laid out [element][instance]."""
        k, real, B = self.m.kernel, self.m.real, self.B
        y0 = self.m.prog.y0 if y0 is None else np.asarray(y0)
        if y0.ndim == 1:
            y0 = np.repeat(y0[:, None], B, axis=1)
        if y0.shape != (k.n_state, B):
            raise ValueError(f"y0 must have shape ({k.n_state},) or ({k.n_state}, {B})")
        self.d_y.upload(np.ascontiguousarray(y0, dtype=real), self.stream)
        n_par = len(self.m.param_names)
        if n_par:
            if params is None:
                params = np.repeat(self.m.param_defaults[:, None], B, axis=1)
            params = np.ascontiguousarray(params, dtype=real)
            if params.shape != (n_par, B):
                raise ValueError(f"params must have shape ({n_par}, {B}) in the order {self.m.param_names}")
            self.d_params.upload(params, self.stream)
        self.d_work.zero(self.stream)                        # solver vectors / operand tiles: cleared on the device
        if k.work_init.size:
            self.d_work.upload(np.ascontiguousarray(np.repeat(k.work_init, B), dtype=real), self.stream)
        self.d_rings.upload(np.ascontiguousarray(np.repeat(k.ring_init, B), dtype=real), self.stream)
        self.steps_done = 0
        self.samples_done = 0

    def launch(self, n_steps: int, dt: float, store_step: int = 1, events=None):
        """``events``: optional ``(start, stop)`` runtime.Event pair recorded on the launch stream right before / after the
        kernel (behind the host-side rendezvous of a multi-GPU launch), for device-side timing."""
        k = self.m.kernel
        new_samples = 0
        if n_steps > 0:
            first = (-self.steps_done) % store_step
            if first < n_steps:
                new_samples = (n_steps - 1 - first) // store_step + 1
        if self.samples_done + new_samples > self.n_samples:
            raise ValueError("launch would overflow the recording buffer")
        self.d_sync.zero(self.stream)
        if self.m.world > 1:
            import torch.distributed as dist
            self.d_mirror.zero(self.stream)     # packet tags restart with every launch
            if self.stream is not None:
                self.stream.synchronize()
            rt.synchronize()
            dist.barrier()                      # every rank's mirror / barrier words are zero before any kernel sends
        a = rt.KernelArgs()
        a.n_steps, a.store_step, a.step0 = int(n_steps), int(store_step), int(self.steps_done)
        a.t0, a.eval0 = self.m.t0, int(self.steps_done) * k.evals_per_step
        a.n_inst, a.sample0, a.dt = self.B, int(self.samples_done), float(dt)
        a.y, a.out = self.d_y.ptr, self.d_out.ptr
        a.params, a.slots = self.d_params.ptr, None
        a.rings, a.tables = self.d_rings.ptr, self.d_tables.ptr
        a.consts, a.iconsts = self.d_consts.ptr, self.d_iconsts.ptr
        a.work, a.sync = self.d_work.ptr, self.d_sync.ptr
        a.peers = self.d_peers.ptr if self.d_peers is not None else None
        a.rank, a.world = self.m.rank, self.m.world
        if events is not None:
            events[0].record(self.stream)
        self.m.module.run(k.kernel_name, a, block=k.block, grid=self.sm_count * self.m.blocks_per_sm,
                          smem=k.smem_bytes, cooperative=True, stream=self.stream)
        if events is not None:
            events[1].record(self.stream)
        self.steps_done += int(n_steps)
        self.samples_done += new_samples
        return new_samples

    def launch_adaptive(self, T: float, dt: float, dts: Optional[float] = None, *, rtol: float = 1e-3, atol: float = 1e-6,
                        max_step: float = np.inf, method: str = "RK45", max_attempts: int = 10_000_000, **unknown):
        """solve_ivp(method='RK45', first_step=dt, t_eval=linspace(0, T, round(T/dts), endpoint=False)) for the whole
        network in ONE persistent launch (grid-wide error norm, one step size for all units)."""
        if method not in (None, "RK45"):
            raise rt.CudaBackendError(f"solver='scipy' on the CUDA backend implements method='RK45'; got method={method!r}")
        if unknown:
            raise TypeError(f"unsupported solve_ivp options on the CUDA backend: {sorted(unknown)}")
        if np.ndim(atol) or np.ndim(rtol):
            raise rt.CudaBackendError("per-component tolerances are not supported; pass scalar rtol / atol")
        k = self.m.kernel
        step = dts if dts else dt
        n_samples = int(round(T / step))
        if n_samples > self.n_samples:
            raise ValueError(f"session holds {self.n_samples} samples, the run records {n_samples}")
        rtol = max(float(rtol), 100 * np.finfo(float).eps)
        opts = np.zeros(16)
        opts[:8] = [rtol, float(atol), self.m.t0, float(T), float(dt), (float(T) / n_samples) if n_samples else 0.0,
                    float(max_step), float(max_attempts)]
        if getattr(self, "d_opts", None) is not None:           # buffer of an earlier launch_adaptive call
            self.d_opts.free()
        self.d_opts = rt.DeviceBuffer.from_host(opts)
        self.d_sync.zero(self.stream)
        a = rt.KernelArgs()
        a.n_steps, a.store_step, a.step0, a.t0, a.eval0 = n_samples, 1, 0, 0, 0
        a.n_inst, a.sample0, a.dt = self.B, 0, float(dt)
        a.y, a.out = self.d_y.ptr, self.d_out.ptr
        a.params, a.slots = self.d_params.ptr, self.d_opts.ptr
        a.rings, a.tables = self.d_rings.ptr, self.d_tables.ptr
        a.consts, a.iconsts = self.d_consts.ptr, self.d_iconsts.ptr
        a.work, a.sync, a.peers = self.d_work.ptr, self.d_sync.ptr, None
        a.rank, a.world = 0, 1
        self.m.module.run(k.kernel_name, a, block=k.block, grid=self.sm_count * self.m.blocks_per_sm,
                          smem=k.smem_bytes, cooperative=True, stream=self.stream)
        self.d_opts.download(opts, self.stream)
        if self.stream is not None:
            self.stream.synchronize()
        self.check()
        self.samples_done = n_samples
        self.attempts = int(opts[12])
        if int(opts[11]) != 0:
            why = {1: "Required step size is less than spacing between numbers.", 2: "attempt budget exhausted"}
            raise rt.CudaBackendError(f"adaptive integration failed: {why.get(int(opts[11]), 'unknown')}")
        return n_samples

    def check(self):
        """Raise if a barrier timed out inside the kernel (lost peer / launch failure on another rank)."""
        words = np.zeros(8, dtype=np.uint64)
        self.d_sync.download(words, self.stream)
        if self.stream is not None:
            self.stream.synchronize()
        if words[2] != 0:
            what = {101: "tcgen05 producer (smem slot never freed)", 102: "tcgen05 MMA issuer (operand tile never arrived)",
                    103: "tcgen05 epilogue (accumulator never completed)", 104: "tcgen05 MMA issuer (TMEM buffer never drained)"
                    }.get(int(words[2]), "grid barrier / peer exchange (a CTA or peer rank did not arrive)")
            raise rt.CudaBackendError(f"network kernel wait timed out: {what}")

    def download(self) -> np.ndarray:
        out = np.empty((self.n_samples, self.m.kernel.n_obs, self.B), dtype=self.m.real)
        if out.size:
            self.d_out.download(out, self.stream)
        if self.stream is not None:
            self.stream.synchronize()
        self.check()
        return out

    def download_state(self) -> np.ndarray:
        y = np.empty((self.m.kernel.n_state, self.B), dtype=self.m.real)
        self.d_y.download(y, self.stream)
        if self.stream is not None:
            self.stream.synchronize()
        return y

    def free(self):
        if self.m.world > 1:
            import torch.distributed as dist
            rt.synchronize()
            for p in self._peer_ptrs:
                rt.ipc_close(p)
            self._peer_ptrs = []
            if dist.is_initialized():
                dist.barrier()                  # nobody frees memory a peer may still have mapped
        for n in ("d_y", "d_out", "d_consts", "d_iconsts", "d_tables", "d_work", "d_rings", "d_sync", "d_peers", "d_params", "d_opts", "d_mirror"):
            b = getattr(self, n, None)
            if b is not None:
                b.free()


# ----------------------------------------------------------------------------- driver hooks
_CHECK = [("wc_n128", "rk4"), ("qsfa_both_n96", "heun"), ("kuramoto_n128", "euler"), ("jrc_grid_b1024", "heun")]


def build_check(golden_dir: str) -> None:
    """NVRTC-compile (sm_100a) one network kernel per coupling family; used by __graft_entry__.build()."""
    for name, solver in _CHECK:
        m = NetworkModel(Program.load(os.path.join(golden_dir, name + ".npz")), solver=solver)
        cubin = m.module.cubin
        assert cubin[:4] == b"\x7fELF", name
        print(f"[build] net {name}/{solver}: {m.kernel.n_phases} phase(s), cubin {len(cubin)} bytes")


_CHECK_SWEEP = [("wc_n300_f32", "heun", 256, ["r_ext"]),        # float32 sweep -> tcgen05 3xTF32 coupling GEMM
                ("wc_n128", "rk4", 128, ["r_ext"])]            # float64 sweep -> DMMA coupling GEMM


def build_check_sweeps(golden_dir: str) -> None:
    for name, solver, nb, sweep in _CHECK_SWEEP:
        m = NetworkModel(Program.load(os.path.join(golden_dir, name + ".npz")), solver=solver, n_batch=nb, sweep=sweep)
        cubin = m.module.cubin
        assert cubin[:4] == b"\x7fELF", name
        kind = f"tcgen05 tiles 128x{m.kernel.tc['bt']}" if m.kernel.tc else "DMMA / FMA tiles"
        print(f"[build] net sweep {name}/{solver} x{nb}: {kind}, cubin {len(cubin)} bytes")
