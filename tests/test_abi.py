"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/prb.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from pyrates_b200 import runtime as rt


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "prb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    names = _declared_symbols()
    assert len(names) >= 25
    l = ctypes.CDLL(os.path.join(ROOT, "pyrates_b200", "libprb.so"))
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, missing
    assert set(names) == set(rt.SYMBOLS), set(names) ^ set(rt.SYMBOLS)


def test_struct_layout_matches_header():
    assert ctypes.sizeof(rt.KernelArgs) == 8 * 8 + 10 * 8 + 16
    assert ctypes.sizeof(rt.RunDesc) == 6 * 4 + 8 + ctypes.sizeof(rt.KernelArgs)
    assert rt.lib().prb_abi_version() == 2


def test_documented_ctypes_stub_matches_the_header(tmp_path):
    """INTEGRATION.md section 2 documents the binding a maintainer would write: its structures must have the size gcc
    computes from include/prb.h (field offsets of the members the stub touches included) and agree with runtime.py."""
    import subprocess
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "class KernelArgs(C.Structure)" in b)
    ns = {}
    exec(stub, ns)
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "prb.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(prb_kernel_args), sizeof(prb_run_desc), offsetof(prb_kernel_args, peers), offsetof(prb_kernel_args, world), '
                   'offsetof(prb_run_desc, stream), offsetof(prb_run_desc, args)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    ka, rd, o_peers, o_world, o_stream, o_args = map(int, subprocess.check_output([str(exe)]).split())
    for K, R in ((ns["KernelArgs"], ns["RunDesc"]), (rt.KernelArgs, rt.RunDesc)):
        assert (ctypes.sizeof(K), ctypes.sizeof(R)) == (ka, rd)
        assert (K.peers.offset, K.world.offset, R.stream.offset, R.args.offset) == (o_peers, o_world, o_stream, o_args)
    assert [f[0] for f in ns["KernelArgs"]._fields_] == [f[0] for f in rt.KernelArgs._fields_]


def test_nvrtc_compiles_without_gpu_and_reports_errors(monkeypatch):
    monkeypatch.setenv("PRB_CACHE_DIR", "off")
    m = rt.compile_module('extern "C" __global__ void k(float* x) { x[threadIdx.x] *= 2.f; }', "t.cu")
    assert len(m.cubin) > 500 and m.cubin[:4] == b"\x7fELF"
    with pytest.raises(rt.CudaBackendError, match="error"):
        rt.compile_module('extern "C" __global__ void k(float* x) { x[0] = undefined_symbol; }', "bad.cu")


def test_compute_calls_fail_loudly_without_device():
    """No CPU fallback: on a box without the CUDA driver every compute entry point raises."""
    try:
        n = rt.device_count()
    except rt.CudaBackendError as e:
        assert "no CPU execution path" in str(e)
        with pytest.raises(rt.CudaBackendError):
            rt.DeviceBuffer(16)
        return
    assert n >= 1
