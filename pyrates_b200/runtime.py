"""ctypes binding of libprb (include/prb.h) — the only way the Python host side touches the GPU.

No CPU execution path exists: if ``libprb.so`` is missing or the CUDA driver cannot be loaded, every
compute call raises ``CudaBackendError``.  ``prb_compile`` (NVRTC) works without a GPU, which is what
the CPU test-suite uses to check that every generated kernel builds for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import threading
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

__all__ = ["CudaBackendError", "lib", "Module", "DeviceBuffer", "compile_module", "KernelArgs", "RunDesc",
           "device_info", "launch_count", "Stream", "Event", "synchronize", "PinnedBuffer", "set_device", "release_pinned_pool"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libprb.so")
ARCH = "sm_100a"


class CudaBackendError(RuntimeError):
    """Raised for every libprb failure (mirrors PyRatesException of the reference,
    pyrates/backend/__init__.py:42)."""


class KernelArgs(C.Structure):
    _fields_ = [("n_steps", C.c_int64), ("store_step", C.c_int64), ("step0", C.c_int64), ("t0", C.c_int64),
                ("eval0", C.c_int64), ("n_inst", C.c_int64), ("sample0", C.c_int64), ("dt", C.c_double),
                ("y", C.c_void_p), ("out", C.c_void_p), ("params", C.c_void_p), ("slots", C.c_void_p),
                ("rings", C.c_void_p), ("tables", C.c_void_p), ("consts", C.c_void_p), ("iconsts", C.c_void_p),
                ("work", C.c_void_p), ("sync", C.c_void_p), ("peers", C.c_void_p), ("rank", C.c_int32),
                ("world", C.c_int32)]


class RunDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("cooperative", C.c_int32), ("grid_dim", C.c_uint32),
                ("block_dim", C.c_uint32), ("smem_bytes", C.c_uint32), ("reserved", C.c_uint32),
                ("stream", C.c_void_p), ("args", KernelArgs)]


class DeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("sm_count", C.c_int32), ("max_smem_per_block_optin", C.c_int32), ("l2_bytes", C.c_int32),
                ("clock_khz", C.c_int32), ("total_mem", C.c_int64)]


class KernelInfo(C.Structure):
    _fields_ = [("regs", C.c_int32), ("static_smem", C.c_int32), ("local_bytes", C.c_int32),
                ("max_threads", C.c_int32), ("max_blocks_per_sm", C.c_int32)]


# exported symbols of include/prb.h with (restype, argtypes); tests check that all of them resolve
SYMBOLS = {
    "prb_abi_version": (C.c_int, []),
    "prb_last_error": (C.c_char_p, []),
    "prb_launch_count": (C.c_int64, []),
    "prb_nvrtc_version": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "prb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "prb_set_device": (C.c_int, [C.c_int]),
    "prb_get_device_info": (C.c_int, [C.POINTER(DeviceInfo)]),
    "prb_device_synchronize": (C.c_int, []),
    "prb_compile": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_void_p)]),
    "prb_module_from_cubin": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "prb_module_cubin": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "prb_module_log": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p)]),
    "prb_module_free": (C.c_int, [C.c_void_p]),
    "prb_kernel_get_info": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(KernelInfo)]),
    "prb_run": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(RunDesc)]),
    "prb_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "prb_free": (C.c_int, [C.c_void_p]),
    "prb_memset": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "prb_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "prb_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "prb_memcpy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "prb_ipc_get_handle": (C.c_int, [C.c_void_p, C.c_char_p]),
    "prb_ipc_open_handle": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "prb_ipc_close_handle": (C.c_int, [C.c_void_p]),
    "prb_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "prb_host_free": (C.c_int, [C.c_void_p]),
    "prb_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "prb_host_unregister": (C.c_int, [C.c_void_p]),
    "prb_stream_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "prb_stream_destroy": (C.c_int, [C.c_void_p]),
    "prb_stream_synchronize": (C.c_int, [C.c_void_p]),
    "prb_stream_wait_event": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prb_event_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "prb_event_destroy": (C.c_int, [C.c_void_p]),
    "prb_event_record": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prb_event_synchronize": (C.c_int, [C.c_void_p]),
    "prb_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
}

_lib = None
_lib_lock = threading.Lock()


def lib():
    """Load libprb.so (built in-tree by ``__graft_entry__.build()`` / ``pyrates_b200.build``)."""
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                if not os.path.exists(_LIB_PATH):
                    raise CudaBackendError(
                        f"{_LIB_PATH} is missing — build it with `python -m pyrates_b200.build`; the CUDA backend "
                        f"has no CPU fallback")
                l = C.CDLL(_LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(l, name)
                    fn.restype, fn.argtypes = res, args
                if l.prb_abi_version() != 2:
                    raise CudaBackendError("libprb ABI version mismatch; rebuild the library")
                _lib = l
    return _lib


def _check(code: int):
    if code != 0:
        raise CudaBackendError(lib().prb_last_error().decode(errors="replace"))


def set_device(ordinal: int):
    _check(lib().prb_set_device(int(ordinal)))


def device_info() -> dict:
    info = DeviceInfo()
    _check(lib().prb_get_device_info(C.byref(info)))
    return {"name": info.name.decode(), "cc": (info.cc_major, info.cc_minor), "sm_count": info.sm_count,
            "max_smem_optin": info.max_smem_per_block_optin, "l2_bytes": info.l2_bytes,
            "clock_khz": info.clock_khz, "total_mem": info.total_mem}


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().prb_device_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    return int(lib().prb_launch_count())


def synchronize():
    _check(lib().prb_device_synchronize())


# ----------------------------------------------------------------------------- memory
class DeviceBuffer:
    """Caller-owned device allocation (the library never frees caller memory)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _check(lib().prb_malloc(C.byref(p), self.nbytes))
        self.ptr = p.value

    @classmethod
    def from_host(cls, arr: np.ndarray, stream=None) -> "DeviceBuffer":
        arr = np.ascontiguousarray(arr)
        buf = cls(max(arr.nbytes, 1))
        buf.upload(arr, stream)
        return buf

    def upload(self, arr: np.ndarray, stream=None, offset: int = 0):
        arr = np.ascontiguousarray(arr)
        if arr.nbytes + offset > self.nbytes:
            raise ValueError("upload exceeds buffer size")
        _check(lib().prb_memcpy_h2d(self.ptr + offset, arr.ctypes.data, arr.nbytes, _s(stream)))
        if stream is None:
            synchronize()

    def download(self, out: np.ndarray, stream=None, offset: int = 0):
        if not out.flags.c_contiguous or out.nbytes + offset > self.nbytes:
            raise ValueError("download target must be contiguous and fit the buffer")
        _check(lib().prb_memcpy_d2h(out.ctypes.data, self.ptr + offset, out.nbytes, _s(stream)))
        if stream is None:
            synchronize()
        return out

    def zero(self, stream=None):
        _check(lib().prb_memset(self.ptr, 0, self.nbytes, _s(stream)))

    def free(self):
        if getattr(self, "ptr", None):
            try:
                lib().prb_free(self.ptr)
            finally:
                self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ipc_handle(buf: "DeviceBuffer") -> bytes:
    h = C.create_string_buffer(64)
    _check(lib().prb_ipc_get_handle(buf.ptr, h))
    return h.raw


def ipc_open(handle: bytes) -> int:
    p = C.c_void_p()
    _check(lib().prb_ipc_open_handle(handle, C.byref(p)))
    return p.value


def ipc_close(ptr: int):
    _check(lib().prb_ipc_close_handle(ptr))


class PinnedBuffer:
    """Pinned host memory exposed as a NumPy array (for async H2D/D2H in the end-to-end path).

    Small buffers come from ``cuMemHostAlloc``.  Large ones (>= 64 MiB: the multi-GB result array of a sweep) are an
    anonymous mapping advised to transparent huge pages, first-touched by a few threads and then page-locked with
    ``cuMemHostRegister``: measured on the B200 box for the 8.39 GB result of the 2^20-instance sweep, 3.35 s for
    cuMemHostAlloc vs 0.3 s registration + the (parallel) first touch, same 56 GB/s D2H rate
    (profiles/r02_pinned_probe.txt).  ``PRB_PINNED=cuda`` forces the driver allocator."""

    HUGE_MIN = 64 << 20
    _map_bytes: Dict[int, int] = {}          # id(mapping) -> registered bytes

    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._map = None
        n = max(self.nbytes, 1)
        if n >= self.HUGE_MIN and os.environ.get("PRB_PINNED", "hugepage") != "cuda":
            spare = _take_spare(n)
            if spare is not None:
                self._map, self.ptr = spare          # a released result buffer of a previous sweep: already registered
            else:
                self._map, self.ptr = _map_huge(n)
                try:
                    _check(lib().prb_host_register(self.ptr, n))
                except CudaBackendError:
                    self._map.close()
                    self._map = None
                    raise
                self._map_bytes[id(self._map)] = n
            self._raw = (C.c_char * n).from_address(self.ptr)
        else:
            p = C.c_void_p()
            _check(lib().prb_host_alloc(C.byref(p), n))
            self.ptr = p.value
            self._raw = (C.c_char * n).from_address(self.ptr)
        self.array = np.frombuffer(self._raw, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def detach(self) -> np.ndarray:
        """Hand the memory over to the array: the pinned allocation is released when the last NumPy view of it dies
        (results of a sweep are returned without a second multi-GB host copy)."""
        import weakref
        arr, ptr = self.array, self.ptr
        weakref.finalize(self._raw, _free_pinned, ptr, self._map)
        self.ptr, self.array, self._raw, self._map = None, None, None, None
        return arr

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            self._raw = None
            ptr, mp = self.ptr, self._map
            self.ptr, self._map = None, None
            _free_pinned(ptr, mp)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _map_huge(n: int):
    """Anonymous private mapping of n bytes, 2 MiB aligned, MADV_HUGEPAGE, first-touched in parallel."""
    import mmap
    from concurrent.futures import ThreadPoolExecutor
    two = 2 << 20
    mp = mmap.mmap(-1, n + two, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = C.addressof(C.c_char.from_buffer(mp))
    ptr = (base + two - 1) & ~(two - 1)
    try:
        libc = C.CDLL(None, use_errno=True)
        libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        libc.madvise(ptr, n, 14)                                     # MADV_HUGEPAGE (advisory: 4 KB pages work as well)
    except (OSError, AttributeError):
        pass
    view = np.frombuffer(mp, dtype=np.uint8, count=n, offset=ptr - base)
    workers = max(1, min(8, (os.cpu_count() or 1), n // (256 << 20)))
    step = -(-n // workers)
    step += (-step) % two

    def touch(k):
        view[k * step:min(n, (k + 1) * step):4096] = 0                # the page faults are what costs; NumPy releases the GIL
    if workers > 1:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(touch, range(workers)))
    else:
        touch(0)
    del view
    return mp, ptr


# One released huge-page buffer is kept registered for the next sweep of a similar size (repeated grid_search calls: the
# 8.4 GB result buffer costs ~0.5 s to map, touch and register, nothing to reuse).  PRB_PINNED_POOL=0 disables it.
_spare = []          # [(mapping, ptr, bytes)] — at most one entry
_spare_lock = threading.Lock()


def _take_spare(n: int):
    with _spare_lock:
        if _spare and n <= _spare[0][2] <= 2 * n:
            mp, ptr, _ = _spare.pop()
            return mp, ptr
    return None


def release_pinned_pool():
    """Unregister and unmap the spare pinned buffer (if any)."""
    with _spare_lock:
        items, _spare[:] = list(_spare), []
    for mp, ptr, _ in items:
        try:
            lib().prb_host_unregister(ptr)
            PinnedBuffer._map_bytes.pop(id(mp), None)
            mp.close()
        except Exception:
            pass


def _free_pinned(ptr, mapping=None):
    try:
        if mapping is None:
            lib().prb_host_free(ptr)
            return
        size = PinnedBuffer._map_bytes.get(id(mapping), 0)
        if size and os.environ.get("PRB_PINNED_POOL", "1") != "0":
            with _spare_lock:
                if not _spare or _spare[0][2] < size:
                    old, _spare[:] = list(_spare), [(mapping, ptr, size)]
                    mapping = None
                else:
                    old = []
            for mp, p_, _ in old:
                lib().prb_host_unregister(p_)
                PinnedBuffer._map_bytes.pop(id(mp), None)
                mp.close()
            if mapping is None:
                return
        lib().prb_host_unregister(ptr)
        PinnedBuffer._map_bytes.pop(id(mapping), None)
        mapping.close()
    except Exception:
        pass


def _s(stream):
    return stream.handle if isinstance(stream, Stream) else stream


class Stream:
    def __init__(self):
        p = C.c_void_p()
        _check(lib().prb_stream_create(C.byref(p)))
        self.handle = p.value

    def synchronize(self):
        _check(lib().prb_stream_synchronize(self.handle))

    def wait_event(self, event: "Event"):
        _check(lib().prb_stream_wait_event(self.handle, event.handle))

    def __del__(self):
        try:
            if self.handle:
                lib().prb_stream_destroy(self.handle)
        except Exception:
            pass


class Event:
    def __init__(self):
        p = C.c_void_p()
        _check(lib().prb_event_create(C.byref(p)))
        self.handle = p.value

    def record(self, stream=None):
        _check(lib().prb_event_record(self.handle, _s(stream)))

    def synchronize(self):
        _check(lib().prb_event_synchronize(self.handle))

    def elapsed_ms(self, end: "Event") -> float:
        ms = C.c_float()
        _check(lib().prb_event_elapsed_ms(self.handle, end.handle, C.byref(ms)))
        return float(ms.value)

    def __del__(self):
        try:
            if self.handle:
                lib().prb_event_destroy(self.handle)
        except Exception:
            pass


# ----------------------------------------------------------------------------- modules
class Module:
    """A JIT-compiled cubin (``prb_module``); kernels are loaded lazily on first launch."""

    def __init__(self, handle: int, log: str = "", source: str = ""):
        self.handle = handle
        self.log = log
        self.source = source

    @property
    def cubin(self) -> bytes:
        data, size = C.c_void_p(), C.c_size_t()
        _check(lib().prb_module_cubin(self.handle, C.byref(data), C.byref(size)))
        return C.string_at(data.value, size.value)

    def kernel_info(self, kernel: str, block: int = 0, smem: int = 0) -> dict:
        info = KernelInfo()
        _check(lib().prb_kernel_get_info(self.handle, kernel.encode(), block, smem, C.byref(info)))
        return {"regs": info.regs, "static_smem": info.static_smem, "local_bytes": info.local_bytes,
                "max_threads": info.max_threads, "max_blocks_per_sm": info.max_blocks_per_sm}

    def run(self, kernel: str, args: KernelArgs, *, block: int, grid: int = 0, smem: int = 0,
            cooperative: bool = False, stream=None):
        d = RunDesc()
        d.struct_size = C.sizeof(RunDesc)
        d.cooperative = 1 if cooperative else 0
        d.grid_dim, d.block_dim, d.smem_bytes = int(grid), int(block), int(smem)
        d.stream = _s(stream)
        d.args = args
        _check(lib().prb_run(self.handle, kernel.encode(), C.byref(d)))

    def __del__(self):
        try:
            if self.handle:
                lib().prb_module_free(self.handle)
                self.handle = None
        except Exception:
            pass


_module_cache: Dict[str, Module] = {}
_toolchain = None


def _toolchain_tag() -> str:
    """NVRTC version + libprb ABI version: part of the cubin cache key, so a toolkit upgrade cannot reuse stale cubins."""
    global _toolchain
    if _toolchain is None:
        major, minor = C.c_int(0), C.c_int(0)
        lib().prb_nvrtc_version(C.byref(major), C.byref(minor))
        _toolchain = f"nvrtc{major.value}.{minor.value}/abi{lib().prb_abi_version()}"
    return _toolchain
DEFAULT_OPTS = ("--std=c++17", "-lineinfo", "--extra-device-vectorization")


def _cache_dir() -> Optional[str]:
    d = os.environ.get("PRB_CACHE_DIR")
    if d is None:
        d = os.path.join(_HERE, "_jit_cache")
    if d in ("", "0", "off"):
        return None
    try:
        os.makedirs(d, exist_ok=True)
        return d
    except OSError:
        return None


def _compile_shared(key: str, build) -> Optional[Module]:
    """One NVRTC run for all ranks of a torch.distributed group: rank 0 builds (or takes the module from its cache), the
    cubin is broadcast to the ranks that lack it.  Collective — every rank must call it at the same point; ranks whose
    source differs (mismatching keys) fall back to local compilation, and the cache state may differ between ranks."""
    try:
        import torch.distributed as dist
    except ImportError:
        return None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    have = key in _module_cache
    state = [None] * dist.get_world_size()
    dist.all_gather_object(state, (key, have))
    if any(k != key for k, _ in state):
        return None
    if all(h for _, h in state):
        return _module_cache[key]
    box = [None]
    if dist.get_rank() == 0:
        box[0] = (_module_cache[key] if have else build()).cubin
    dist.broadcast_object_list(box, src=0)
    if have or dist.get_rank() == 0:
        return _module_cache.get(key)               # rank 0: build() put it into the cache
    handle = C.c_void_p()
    _check(lib().prb_module_from_cubin(box[0], len(box[0]), C.byref(handle)))
    return Module(handle.value, "", "")


def compile_module(source: str, name: str = "prb_kernel.cu", opts: Sequence[str] = (), *, fmad: bool = True,
                   verbose_ptxas: bool = False, shared: bool = False) -> Module:
    """NVRTC-compile ``source`` for sm_100a.  Cached in memory and on disk by source hash
    (cf. the reference's SHA-256 module cache, base_backend.py:557)."""
    all_opts = list(DEFAULT_OPTS) + list(opts) + [f"--fmad={'true' if fmad else 'false'}"]
    if verbose_ptxas:
        all_opts += ["--ptxas-options=-v"]
    if shared:
        key = hashlib.sha256(("\0".join(all_opts) + "\0" + ARCH + "\0" + _toolchain_tag() + "\0" + source).encode()).hexdigest()
        m = _compile_shared(key, lambda: compile_module(source, name, opts, fmad=fmad))
        if m is not None:
            m.source = m.source or source
            _module_cache[key] = m
            return m
    key = hashlib.sha256(("\0".join(all_opts) + "\0" + ARCH + "\0" + _toolchain_tag() + "\0" + source).encode()).hexdigest()
    m = _module_cache.get(key)
    if m is not None:
        return m
    l = lib()
    cdir = _cache_dir()
    path = os.path.join(cdir, key + ".cubin") if cdir else None
    handle = C.c_void_p()
    if path and os.path.exists(path) and not verbose_ptxas:
        with open(path, "rb") as f:
            blob = f.read()
        _check(l.prb_module_from_cubin(blob, len(blob), C.byref(handle)))
        m = Module(handle.value, "", source)
    else:
        arr = (C.c_char_p * len(all_opts))(*[o.encode() for o in all_opts])
        code = l.prb_compile(source.encode(), name.encode(), ARCH.encode(), arr, len(all_opts), C.byref(handle))
        log = ""
        if handle.value:
            p = C.c_char_p()
            l.prb_module_log(handle.value, C.byref(p))
            log = (p.value or b"").decode(errors="replace")
        if code != 0:
            msg = l.prb_last_error().decode(errors="replace")
            if handle.value:
                l.prb_module_free(handle.value)
            raise CudaBackendError(msg)
        m = Module(handle.value, log, source)
        if path:
            try:
                tmp = path + f".{os.getpid()}.tmp"
                with open(tmp, "wb") as f:
                    f.write(m.cubin)
                os.replace(tmp, path)
            except OSError:
                pass
    _module_cache[key] = m
    return m
