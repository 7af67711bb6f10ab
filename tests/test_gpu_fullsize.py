"""Parity at the BASELINE.json sizes through size-independent checks (the oracle cannot run 2^20 instances):

* C2 — 2^20-instance JRC sweep through the end-to-end engine (BatchedSweep): random instances of the grid against the
  oracle integrating that single parameter set; duplicated parameter columns give bit-identical trajectories; the
  streamed (chunked, overlapped D2H) result is bit-identical to a single-launch run.
* C3 / C4 / C5 — networks at N in the thousands (programs resized from the N=8 goldens, bench_configs.py) against the
  oracle on the full state vector for a few steps."""
import numpy as np
import pytest

import bench
import bench_configs as bc
from conftest import golden_path, traj_rel_err
from oracle import numpy_oracle as orc
from pyrates_b200.engine import InstanceModel
from pyrates_b200.netengine import NetworkModel
from pyrates_b200.program import Program
from pyrates_b200.sweep import BatchedSweep

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_c2_full_grid_sweep_spot_checks(gpu):
    prog = Program.load(golden_path("jrc"))
    grid = 1024
    B = grid * grid
    table = bench.param_table(grid, grid, 0, B)
    rng = np.random.default_rng(11)
    picks = sorted(rng.choice(B, 6, replace=False).tolist()) + [0, B - 1]
    dup_src, dup_dst = picks[0], picks[1] + 1
    table[:, dup_dst] = table[:, dup_src]
    T, dt, dts = 0.05, 1e-4, 1e-3
    sw = BatchedSweep(prog, bench.SWEEP, observers=[0, 2], solver="rk4", T=T, dt=dt, dts=dts, n_inst=B,
                      chunk_samples=16, const_div_to_mul=False)
    try:
        out = np.array(sw.run(table), copy=True)                    # [50, 2, 2^20], 4 chunks streamed
        assert out.shape == (50, 2, B) and np.all(np.isfinite(out))
        assert np.array_equal(out[:, :, dup_src], out[:, :, dup_dst])
        f = orc.build_vector_field(prog.source())
        names = [a.name for a in prog.args]
        for b in picks:
            u, C, C4, C08, _ = table[:, b]
            args = [np.array(a.value, copy=True) for a in prog.args]
            args[names.index("u")] = np.float64(u)
            args[names.index("weight_v2")] = np.array([C, C4])
            args[names.index("weight")] = np.float64(C08)
            args[names.index("weight_v1")] = np.float64(C4)
            rec, _ = orc.run(f, args, T, dt, dts, "rk4")
            assert traj_rel_err(out[:, :, b], rec[:, [0, 2]]) <= 1e-9, b
        # one un-chunked launch of a 2^16 slice must reproduce the streamed result bit for bit
        m = InstanceModel(prog, solver="rk4", sweep=bench.SWEEP, observers=[0, 2], min_blocks=5)
        order = [(p.arg, p.elem) for p in m.rhs.params]
        sl = slice(3 * 2 ** 16, 4 * 2 ** 16)
        params = np.stack([table[bench.SWEEP.index(k), sl] for k in order])
        single, _ = m.run(T, dt, dts, params=params)
        assert np.array_equal(single, out[:, :, sl])
    finally:
        sw.free()


@pytest.mark.parametrize("solver", ["euler", "rk4"])
def test_c3_wc_dense_n2048_full_state(gpu, solver):
    prog = bc.make_c3(2048)
    f = orc.build_vector_field(prog.source())
    args = [np.array(a.value, copy=True) for a in prog.args]
    args[0] = args[0][()]
    ref, _ = orc.run(f, args, 12e-3, 1e-3, None, solver)
    out, y_final = NetworkModel(prog, solver=solver).run(12e-3, 1e-3, None)
    assert traj_rel_err(out[:, :, 0], ref) <= 1e-9
    np.testing.assert_allclose(y_final[:, 0], args[1], rtol=1e-9, atol=1e-12)     # oracle mutated y0 in place to y(T)


def test_c4_qif_sfa_two_populations_n1024_full_state(gpu):
    prog = bc.make_c4(1024)
    f = orc.build_vector_field(prog.source())
    args = [np.array(a.value, copy=True) for a in prog.args]
    args[0] = args[0][()]
    ref, _ = orc.run(f, args, 10e-3, 1e-3, None, "heun")
    out, _ = NetworkModel(prog, solver="heun").run(10e-3, 1e-3, None)
    assert out.shape[1] == 12 * 1024
    assert traj_rel_err(out[:, :, 0], ref) <= 1e-9


def test_c5_kuramoto_n1024_separable_full_state(gpu):
    prog = bc.make_c5(1024)
    f = orc.build_vector_field(prog.source())
    args = [np.array(a.value, copy=True) for a in prog.args]
    args[0] = args[0][()]
    ref, _ = orc.run(f, args, 8e-3, 1e-3, None, "rk4")
    out, _ = NetworkModel(prog, solver="rk4").run(8e-3, 1e-3, None)
    assert traj_rel_err(out[:, :, 0], ref) <= 1e-9


def test_c2_bench_kernel_fast_math_full_grid_spot_checks(gpu):
    """The exact configuration bench.py times (const_div_to_mul + fast_math, 2^20 instances): random grid points against
    the oracle at the fp64 tolerance, over a horizon long enough for the JRC oscillation to develop (T = 0.3)."""
    prog = Program.load(golden_path("jrc"))
    grid = 1024
    B = grid * grid
    table = bench.param_table(grid, grid, 0, B)
    T, dt, dts = 0.3, 1e-4, 1e-3
    sw = BatchedSweep(prog, bench.SWEEP, observers=[0], solver="rk4", T=T, dt=dt, dts=dts, n_inst=B, chunk_samples=100,
                      const_div_to_mul=True)
    assert "prb_spec_exp(" in sw.model.kernel.source and "prb_rhs_exact" in sw.model.kernel.source
    try:
        out = np.array(sw.run(table), copy=True)
        assert out.shape == (300, 1, B) and np.all(np.isfinite(out))
        f = orc.build_vector_field(prog.source())
        names = [a.name for a in prog.args]
        for b in [0, B - 1] + sorted(np.random.default_rng(3).choice(B, 5, replace=False).tolist()):
            u, C, C4, C08, _ = table[:, b]
            args = [np.array(a.value, copy=True) for a in prog.args]
            args[names.index("u")] = np.float64(u)
            args[names.index("weight_v2")] = np.array([C, C4])
            args[names.index("weight")] = np.float64(C08)
            args[names.index("weight_v1")] = np.float64(C4)
            rec, _ = orc.run(f, args, T, dt, dts, "rk4")
            assert traj_rel_err(out[:, :, b], rec[:, [0]]) <= 1e-9, b
    finally:
        sw.free()


def test_fast_math_falls_back_to_exact_evaluation_outside_the_fast_range(gpu):
    """exp arguments beyond +-700 (overflow / underflow) and huge divisors: the speculative kernel must re-evaluate with
    the library functions and agree with the exact kernel bit for bit there."""
    prog = Program.load(golden_path("jrc"))
    y0 = np.array(prog.y0, dtype=np.float64)
    outs = []
    for fm in (False, True):
        m = InstanceModel(prog, solver="euler", fast_math=fm, const_div_to_mul=fm)
        # v = y0[1]-ish potentials of +-3 drive 560*(0.006 - v) to about -+1700: exp over/underflows
        y = np.stack([y0 + np.array([0, 0, 0, 0, 3.0, -3.0, 0, 0]), y0 + np.array([3.0, 0, -9.0, 0, 0, 0, 0, 0])], axis=1)
        out, _ = m.run(2e-3, 1e-4, None, y0=y, n_inst=2)
        outs.append(out)
    assert np.all(np.isfinite(outs[0])) and traj_rel_err(outs[1][:, :, 0], outs[0][:, :, 0]) <= 1e-12
    assert traj_rel_err(outs[1][:, :, 1], outs[0][:, :, 1]) <= 1e-12


def test_sweep_results_can_stay_on_the_device_as_torch_tensors(gpu):
    """BatchedSweep.run_torch: parameters from a CUDA tensor, trajectories left in a CUDA tensor the kernels wrote —
    bit-identical to the host path, and usable by torch on the same stream without synchronisation calls."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("torch sees no CUDA device")
    prog = Program.load(golden_path("jrc"))
    B = 4096
    table = bench.param_table(64, 64, 0, B)
    sw = BatchedSweep(prog, bench.SWEEP, observers=[0, 2], solver="rk4", T=0.05, dt=1e-4, dts=1e-3, n_inst=B, chunk_samples=10)
    try:
        host = np.array(sw.run(table), copy=True)
        dev = sw.run_torch(torch.as_tensor(table, device="cuda"))
        assert dev.is_cuda and tuple(dev.shape) == host.shape
        peak = dev.abs().amax(dim=0)                        # consumer on torch's stream, right behind the kernel
        assert np.array_equal(dev.cpu().numpy(), host)
        np.testing.assert_array_equal(peak.cpu().numpy(), np.abs(host).max(axis=0))
    finally:
        sw.free()
