"""How fast can the box provide a pinned 8.4 GB result buffer, and how fast does D2H run into it?
(a) cuMemHostAlloc, (b) anonymous mmap + MADV_HUGEPAGE + first touch + cuMemHostRegister, (c) the same without huge pages."""
import ctypes as C, mmap, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrates_b200 import runtime as rt

N = int(float(sys.argv[1]) if len(sys.argv) > 1 else 8.388608e9)
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| defrag:", open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip())
rt.set_device(0)
chunk = 210 * 2 ** 20
d = rt.DeviceBuffer(chunk)
s = rt.Stream()

def d2h(ptr):
    best = 1e9
    for rep in range(2):
        t = time.perf_counter(); off = 0
        while off < N:
            n = min(chunk, N - off)
            rt._check(rt.lib().prb_memcpy_d2h(ptr + off, d.ptr, n, s.handle)); off += n
        s.synchronize(); best = min(best, time.perf_counter() - t)
    return N / best / 1e9

t = time.perf_counter(); p = C.c_void_p(); rt._check(rt.lib().prb_host_alloc(C.byref(p), N)); ta = time.perf_counter() - t
print(f"cuMemHostAlloc {N/1e9:.2f} GB: {ta:.3f} s; D2H {d2h(p.value):.1f} GB/s")
t = time.perf_counter(); rt._check(rt.lib().prb_host_free(p)); print(f"  free {time.perf_counter()-t:.3f} s")
libc = C.CDLL(None, use_errno=True)
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
for huge in (True, False):
    t = time.perf_counter()
    mm = mmap.mmap(-1, N + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = C.addressof(C.c_char.from_buffer(mm))
    ptr = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
    if huge:
        rc = libc.madvise(ptr, N, 14)
    t1 = time.perf_counter()
    arr = np.frombuffer(mm, dtype=np.uint8, count=N, offset=ptr - base)
    arr[::4096] = 1                                       # first touch
    t2 = time.perf_counter()
    rt._check(rt.lib().prb_host_register(ptr, N)); t3 = time.perf_counter()
    print(f"mmap{' + MADV_HUGEPAGE' if huge else ''}: map {t1-t:.3f} s, touch {t2-t1:.3f} s, register {t3-t2:.3f} s, total {t3-t:.3f} s; D2H {d2h(ptr):.1f} GB/s")
    ah = int(open('/proc/meminfo').read().split('AnonHugePages:')[1].split()[0])
    print(f"  AnonHugePages {ah/2**20:.2f} GB")
    t = time.perf_counter(); rt._check(rt.lib().prb_host_unregister(ptr)); del arr; print(f"  unregister {time.perf_counter()-t:.3f} s")
    # register WITHOUT touching first (the driver faults the pages in)
    mm2 = mmap.mmap(-1, N + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base2 = C.addressof(C.c_char.from_buffer(mm2)); ptr2 = (base2 + (2 << 20) - 1) & ~((2 << 20) - 1)
    if huge:
        libc.madvise(ptr2, N, 14)
    t = time.perf_counter(); rt._check(rt.lib().prb_host_register(ptr2, N)); print(f"  register untouched: {time.perf_counter()-t:.3f} s; D2H {d2h(ptr2):.1f} GB/s")
    rt._check(rt.lib().prb_host_unregister(ptr2))
