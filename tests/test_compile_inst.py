"""Every unrollable fixture's generated instance kernel builds for sm_100a (NVRTC, no GPU needed)."""
import os

import pytest

from conftest import golden_path
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program

CASES = [("qif", "euler"), ("qif_f32", "heun"), ("jrc", "rk4"), ("net12", "heun_true"), ("qif_input", "rk4"),
         ("jrc_2delay", "euler"), ("qsfa_both_n8", "heun")]


@pytest.mark.parametrize("name,solver", CASES)
def test_instance_kernel_compiles(name, solver):
    os.environ["PRB_CACHE_DIR"] = "off"
    m = InstanceModel(Program.load(golden_path(name)), solver=solver)
    src = m.kernel.source
    assert "prb_integrate" in src and "__grid_constant__" in src
    mod = rt.compile_module(src, f"{name}.cu", verbose_ptxas=True)
    assert mod.cubin[:4] == b"\x7fELF"
    assert "registers" in mod.log


def test_jrc_rk4_stays_in_registers():
    os.environ["PRB_CACHE_DIR"] = "off"
    m = InstanceModel(Program.load(golden_path("jrc")), solver="rk4",
                      sweep=[("u", 0), ("weight", 0), ("weight_v1", 0), ("weight_v2", 0), ("weight_v2", 1)],
                      observers=[0])
    mod = rt.compile_module(m.kernel.source, "jrc_rk4.cu", verbose_ptxas=True)
    line = [l for l in mod.log.splitlines() if "spill" in l][-1]
    assert "0 bytes spill stores" in line and "0 bytes stack frame" in line, mod.log


def test_unknown_solver_is_rejected():
    with pytest.raises(rt.CudaBackendError, match="does not support solver"):
        InstanceModel(Program.load(golden_path("qif")), solver="scipy")
