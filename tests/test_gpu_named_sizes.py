"""Parity AT the sizes BASELINE.json names (VERDICT r01 "next round" item 2), under `-m gpu`:

* C3  Wilson-Cowan N = 8192 dense: full state, 8 Euler + 2 RK4 steps vs the oracle — the program comes from the real
      front end (PopulationTemplate + Connectivity -> ComputeGraph -> CudaBackend) when PyRates is importable (baseline/_ref
      on the GPU box), and the public CircuitTemplate.run(backend='cuda') is compared with the reference's own NumPy backend
      on the same circuit (reference loops mirrored: /root/reference/tests/test_population_connectivity.py:38,94);
* C4  QIF-SFA 2 x 2048 (gamma cascade + delay ring, :361-415, :422-476), reference-"heun", full state vs oracle;
* C5  Kuramoto N = 16384 dense sin coupling (:238-283), full state vs an independent fp64 NumPy evaluation; N = 4096 vs the
      oracle's direct pair-space evaluation; the device-generated W of the N = 65536 bench vs its NumPy twin;
* the tcgen05 (3xTF32) sweep GEMM at N = 8192 (K = 8192: 32 promotion chunks), B = 64, and the DMMA sweep GEMM at N = 2048,
      B = 64: sampled instances vs the fp64 oracle;
* C2  the fast-math kernel bench.py times, full horizon T = 1.0, sampled grid points vs the oracle.
"""
import numpy as np
import pytest

import bench
import bench_configs as bc
from conftest import golden_path, traj_rel_err
from oracle import numpy_oracle as orc
from pyrates_b200.netengine import NetworkModel
from pyrates_b200.program import Arg, Program
from pyrates_b200.sweep import BatchedSweep

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _oracle(prog, T, dt, solver, dts=None):
    f = orc.build_vector_field(prog.source())
    args = [np.array(a.value, copy=True) for a in prog.args]
    args[0] = args[0][()]
    return orc.run(f, args, T, dt, dts, solver)[0]


@pytest.fixture(scope="module")
def c3_8192():
    return bc.build_program("c3", 8192)


@pytest.mark.parametrize("solver,steps", [("euler", 8), ("rk4", 2)])
def test_c3_wc_n8192_full_state(gpu, c3_8192, solver, steps):
    prog, how = c3_8192
    assert prog.n_state == 8192 and prog.arg("weight").value.shape == (8192, 8192)
    ref = _oracle(prog, steps * 1e-3, 1e-3, solver)
    out, _ = NetworkModel(prog, solver=solver).run(steps * 1e-3, 1e-3, None)
    assert out.shape == (steps, 8192, 1)
    assert np.ptp(ref[-1]) > 1e-3                                   # a non-trivial trajectory, not the fixed point
    assert traj_rel_err(out[:, :, 0], ref) <= 1e-9, how


def test_c3_n8192_public_run_matches_reference_numpy_backend(gpu, tmp_path, monkeypatch):
    """CircuitTemplate.run(backend='cuda') vs run(backend='default') of the UNMODIFIED reference, same circuit, N = 8192."""
    if bc.import_pyrates() is None:
        pytest.skip("PyRates front end not importable (baseline/_ref missing)")
    monkeypatch.chdir(tmp_path)
    from pyrates.ir.node import clear_ir_caches
    kw = dict(simulation_time=8e-3, step_size=1e-3, solver="euler", outputs={"r": "e/rate_op/r"}, float_precision="float64",
              verbose=False, clear=True)
    res = {}
    for backend in ("cuda", "default"):
        clear_ir_caches()
        res[backend] = bc.circuit_c3(8192).run(backend=backend, **kw)
    assert res["cuda"].shape == res["default"].shape == (8, 8192)
    assert list(res["cuda"].columns) == list(res["default"].columns)
    assert traj_rel_err(res["cuda"].values, res["default"].values) <= 1e-9


def test_c4_qif_sfa_2x2048_full_state(gpu):
    prog, how = bc.build_program("c4", 2048)
    assert prog.n_state == 12 * 2048
    ref = _oracle(prog, 10e-3, 1e-3, "heun")
    out, _ = NetworkModel(prog, solver="heun").run(10e-3, 1e-3, None)
    assert traj_rel_err(out[:, :, 0], ref) <= 1e-9, how


def test_c5_kuramoto_n16384_full_state_vs_independent_fp64(gpu):
    """dtheta_i = omega_i + K_i * sum_j W_ij sin(theta_j - theta_i) (kuramoto.yaml:20-27,75-78) integrated with RK4 by a
    plain NumPy loop (angle-addition form: two BLAS matvecs) — independent of both the oracle and the kernels."""
    n = 16384
    prog, how = bc.build_program("c5", n)
    W = np.asarray(prog.arg("weight").value)
    omega, K, s_ext = (np.asarray(prog.arg(k).value, dtype=np.float64) for k in ("omega", "K", "s_ext"))
    assert W.shape == (n, n)

    def f(th):
        s, c = np.sin(th), np.cos(th)
        return omega + s_ext + K * (c * (W @ s) - s * (W @ c))
    y = np.array(prog.y0, dtype=np.float64)
    ref = []
    dt = 1e-3
    for _ in range(3):
        ref.append(y.copy())
        k1 = f(y); k2 = f(y + 0.5 * dt * k1); k3 = f(y + 0.5 * dt * k2); k4 = f(y + dt * k3)
        y = y + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    out, y_final = NetworkModel(prog, solver="rk4").run(3e-3, dt, None)
    assert traj_rel_err(out[:, :, 0], np.array(ref)) <= 1e-9, how
    np.testing.assert_allclose(y_final[:, 0], y, rtol=1e-9)


def test_c5_kuramoto_n4096_vs_oracle_pair_space(gpu):
    prog, how = bc.build_program("c5", 4096)
    ref = _oracle(prog, 3e-3, 1e-3, "euler")
    out, _ = NetworkModel(prog, solver="euler").run(3e-3, 1e-3, None)
    assert traj_rel_err(out[:, :, 0], ref) <= 1e-9, how


def test_c5a_global_vsum_n65536_two_level_reduction(gpu):
    """Scalar all-to-all weight -> `weight * sum(theta)` (pyrates/ir/circuit.py:752-761) at the named size: the full
    reduction runs in two levels (one partial per CTA, grid barrier, sum of the partials)."""
    prog, how = bc.build_program("c5a", 65536, front_end="resized")
    m = NetworkModel(prog, solver="rk4")
    assert "level 2: sum of the 1 x 256 partials" in m.kernel.source
    ref = _oracle(prog, 5e-3, 1e-3, "rk4")
    out, _ = m.run(5e-3, 1e-3, None)
    assert traj_rel_err(out[:, :, 0], ref) <= 1e-9, how


def test_c5_device_generated_connectivity_equals_its_numpy_twin(gpu):
    """The N = 65536 bench generates W on the device (devgen.DeviceUniform): same program with the generator's NumPy twin
    uploaded from the host must give bit-identical trajectories, and the oracle (fed the twin) agrees at 1e-9."""
    n = 2048
    lazy, _ = bc.build_program("c5", n, device_w=True)
    gen = lazy.arg("weight").value
    Wh = gen.host()
    assert Wh.shape == (n, n) and 0.0 <= Wh.min() and Wh.max() < 2.0 / n and abs(Wh.mean() * n - 1.0) < 1e-2
    host = Program([Arg(a.name, a.kind, Wh if a.name == "weight" else a.value, a.label) for a in lazy.args], lazy.stmts, {},
                   lazy.float_precision, lazy.func_name)
    a, _ = NetworkModel(lazy, solver="rk4").run(4e-3, 1e-3, None)
    b, _ = NetworkModel(host, solver="rk4").run(4e-3, 1e-3, None)
    assert np.array_equal(a, b)
    assert traj_rel_err(a[:, :, 0], _oracle(host, 4e-3, 1e-3, "rk4")) <= 1e-9


def _sweep_instances(prog, m, params, sweep, picks, T, dt, solver):
    f = orc.build_vector_field(prog.source())
    names = [a.name for a in prog.args]
    for b in picks:
        args = [np.array(a.value, copy=True) for a in prog.args]
        args[0] = args[0][()]
        for k, nme in enumerate(sweep):
            v = args[names.index(nme)]
            args[names.index(nme)] = np.full(v.shape, params[k, b], dtype=v.dtype) if v.ndim else np.asarray(params[k, b], dtype=v.dtype)[()]
        yield b, orc.run(f, args, T, dt, None, solver)[0]


def test_tcgen05_sweep_n8192_k32_chunks_vs_fp64_oracle(gpu, c3_8192):
    """K = 8192 = 32 promotion chunks of 256: the fp32 partial sums of the 3xTF32 products are added outside the tensor
    core (PRB_TC_KC), which is what keeps the K = 8192 accumulation inside the 1e-4 float32 tolerance."""
    prog64, _ = c3_8192
    prog = bc.to_float32(prog64)
    B, steps, sweep = 64, 5, ["r_ext", "k"]
    m = NetworkModel(prog, solver="euler", n_batch=B, sweep=sweep)
    assert m.kernel.tc is not None and m.kernel.tc["kc"] * 32 * 32 == 8192
    params = m.param_defaults[:, None] + np.random.default_rng(9).uniform(-0.2, 0.4, (2, B))
    out, _ = m.run(steps * 1e-3, 1e-3, None, params=params)
    assert out.shape == (steps, 8192, B) and np.all(np.isfinite(out))
    for b, rec in _sweep_instances(prog64, m, params, sweep, (0, 21, 42, B - 1), steps * 1e-3, 1e-3, "euler"):
        assert traj_rel_err(out[:, :, b], rec) <= 1e-4, b


def test_dmma_sweep_n2048_vs_fp64_oracle(gpu):
    prog, _ = bc.build_program("c3", 2048)
    B, steps, sweep = 64, 6, ["r_ext", "k"]
    m = NetworkModel(prog, solver="rk4", n_batch=B, sweep=sweep)
    assert m.kernel.tc is None and "prb_gemm_batch<" in m.kernel.source.split("prb_eval(PrbNet& C) {")[1]
    params = m.param_defaults[:, None] + np.random.default_rng(10).uniform(-0.2, 0.4, (2, B))
    out, _ = m.run(steps * 1e-3, 1e-3, None, params=params)
    for b, rec in _sweep_instances(prog, m, params, sweep, (0, 17, 40, B - 1), steps * 1e-3, 1e-3, "rk4"):
        assert traj_rel_err(out[:, :, b], rec) <= 1e-9, b


def test_c2_bench_kernel_full_horizon_T1(gpu):
    """The kernel and horizon of the headline number: fast_math + const_div_to_mul, RK4, dt = 1e-4, T = 1.0 (1e4 steps),
    1000 samples; four grid points of a 2^16-instance block against the oracle at 1e-9."""
    prog = Program.load(golden_path("jrc"))
    grid = 256
    B = grid * grid
    table = bench.param_table(grid, grid, 0, B)
    T, dt, dts = 1.0, 1e-4, 1e-3
    sw = BatchedSweep(prog, bench.SWEEP, observers=[0], solver="rk4", T=T, dt=dt, dts=dts, n_inst=B, chunk_samples=25,
                      const_div_to_mul=True)
    assert "prb_spec_exp(" in sw.model.kernel.source
    try:
        out = np.array(sw.run(table), copy=True)
        assert out.shape == (1000, 1, B)
        f = orc.build_vector_field(prog.source())
        names = [a.name for a in prog.args]
        for b in (0, 12345, 40000, B - 1):
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


def test_public_grid_search_2e16_returns_the_pinned_result_without_a_copy(gpu, tmp_path, monkeypatch):
    """pyrates.grid_search(backend='cuda', as_arrays=True) on a 256 x 256 (C, u) grid: the timings dict is filled, the result
    matches BatchedSweep, and the '<name>_<i>' index of the returned parameter frame follows utility.py:248-252."""
    pyrates = bc.import_pyrates()
    if pyrates is None:
        pytest.skip("PyRates front end not importable (baseline/_ref missing)")
    monkeypatch.chdir(tmp_path)
    side = 256
    grid, pmap = bench.jrc_grid(side)
    tm = {}
    res, pg = pyrates.grid_search(bench.JRC_YAML, grid, pmap, step_size=1e-4, simulation_time=0.02, sampling_step_size=1e-3,
                                  outputs={"v": "pc/rpo_e_in/v"}, solver="rk4", backend="cuda", float_precision="float64",
                                  as_arrays=True, timings=tm, verbose=False)
    assert res.shape == (20, 1, side * side) and set(tm) >= {"probe_build_s", "jit_s", "run_s", "frame_s"}
    assert list(pg.index[:2]) == ["JRC_0", "JRC_1"] and pg.index[-1] == f"JRC_{side * side - 1}"
    prog = Program.load(golden_path("jrc"))
    sw = BatchedSweep(prog, bench.SWEEP, observers=[0], solver="rk4", T=0.02, dt=1e-4, dts=1e-3, n_inst=side * side)
    try:
        want = sw.run(bench.param_table(side, side, 0, side * side))
        assert traj_rel_err(res[:, 0, :], want[:, 0, :]) <= 1e-12
    finally:
        sw.free()
    # the frame path: same values, the reference's column labels
    frame, _ = pyrates.grid_search(bench.JRC_YAML, grid.iloc[:64].copy(), pmap, step_size=1e-4, simulation_time=0.02,
                                   sampling_step_size=1e-3, outputs={"v": "pc/rpo_e_in/v"}, solver="rk4", backend="cuda",
                                   float_precision="float64", verbose=False)
    assert frame.shape == (20, 64) and tuple(frame.columns[3]) == ("v", "JRC_3", "pc", "rpo_e_in/v")
    np.testing.assert_array_equal(frame.values, res[:, 0, :64])
