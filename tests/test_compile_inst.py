"""Every unrollable fixture's generated instance kernel builds for sm_100a (NVRTC, no GPU needed)."""
import os

import pytest

from conftest import golden_path
from pyrates_b200 import runtime as rt
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program

CASES = [("qif", "euler"), ("qif_f32", "heun"), ("jrc", "rk4"), ("net12", "heun_true"), ("qif_input", "rk4"),
         ("jrc_2delay", "euler"), ("qsfa_both_n8", "heun"),
         # delay-differential models: history ring (fixed step) / dopri5 + (time, state) history (solver='scipy')
         ("dde_mg", "rk4"), ("dde_net16_f32", "heun"), ("dde_mg_dopri5", "scipy"), ("dde_net16_dopri5_f32", "scipy"),
         ("dde_jrc_pair_dopri5", "scipy"), ("jrc_dop853", "scipy")]


@pytest.mark.parametrize("name,solver", CASES)
def test_instance_kernel_compiles(name, solver):
    os.environ["PRB_CACHE_DIR"] = "off"
    m = InstanceModel(Program.load(golden_path(name)), solver=solver, rk_method="DOP853" if "dop853" in name else None)
    src = m.kernel.source
    assert "prb_integrate" in src and "__grid_constant__" in src and ("PRB_RK_DOP853 1" in src) == ("dop853" in name)
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


def test_delay_differential_kernels_carry_a_history():
    fixed = InstanceModel(Program.load(golden_path("dde_mg")), solver="euler")
    k = fixed.kernel
    # delays 0.0173 / 0.0091 at dt = 1e-4: 173 steps back, interpolation needs two neighbours -> ring of 256 steps
    assert k.n_hist == 2 and k.hist_vars == [0, 1] and k.hist_depth == 256 and fixed.rhs.hist_dt == 1e-4
    assert "prb_hist_push(T, y, A.step0 + i + s + 1)" in k.source and "#define PRB_SOLVER 0" in k.source
    adaptive = InstanceModel(Program.load(golden_path("dde_mg_dopri5")), solver="scipy")
    assert adaptive.kernel_solver == "dopri5_dde" and adaptive.kernel.hist_depth == 0 and "#define PRB_SOLVER 5" in adaptive.kernel.source
    assert list(adaptive.hist_y0) == [1.2, 0.4]
    # a vector field generated for the other solver family is refused (its history lookups use the other time argument)
    from pyrates_b200.scalarize import UnsupportedModelError
    with pytest.raises(UnsupportedModelError):
        InstanceModel(Program.load(golden_path("dde_mg")), solver="scipy")
    with pytest.raises(UnsupportedModelError):
        InstanceModel(Program.load(golden_path("dde_mg_dopri5")), solver="heun")


def test_c2_bench_kernel_instruction_mix():
    """The kernel bench.py times (JRC RK4 sweep, fast_math + invariant folding, 64 x 7 launch bounds): its time loop must
    keep the FP64-pipe instruction count DESIGN.md / bench.py quote (264 per instance-step; 317 before the folding pass) and
    the lean integer side of round 2 (<= 105 instructions of other pipes per step, was 181; one speculation check per step;
    no local-memory access and no call on the executed path) — counted in the SASS like scripts/sass_rf_model.py does, no
    GPU needed."""
    import collections
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    os.environ["PRB_CACHE_DIR"] = "off"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "scripts"))
    import bench
    import sass_rf_model as sm
    m = InstanceModel(Program.load(golden_path("jrc")), solver="rk4", sweep=bench.SWEEP, observers=[0], const_div_to_mul=True,
                      fast_math=True, block=64, min_blocks=7)
    assert "#define PRB_SPEC_STEP 1" in m.kernel.source and "#define PRB_EXP_LEAN 1" in m.kernel.source
    mod = rt.compile_module(m.kernel.source, "c2_mix.cu")
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"c2_mix_{os.getpid()}.cubin")
    with open(path, "wb") as f:
        f.write(mod.cubin)
    try:
        ins = sm.parse(path)
    finally:
        os.remove(path)
    body = sm.loop_body(ins)                       # executed path of the innermost loop around the fp64 work
    ops = collections.Counter(sm.operands(s)[0].split(".")[0] for _, s in body)
    fp64 = sum(ops[k] for k in sm.FP64)
    assert fp64 == bench.FP64_INSTR[True], (fp64, bench.FP64_INSTR)
    other = len(body) - fp64
    assert ops["MUFU"] == 12 and ops["LDS"] == 12 and other <= 105, (other, ops)
    assert ops["STL"] + ops["LDL"] + ops["CALL"] == 0, "local-memory access or call on the hot path of the time loop"
    assert ops["BSSY"] == 1, "one reconvergence point (the speculation check) per step"
    # the exact re-evaluation of a step sits out of line behind that check: exactly one call in the loop's address range
    lo, hi = body[0][0], body[-1][0]
    assert sum("CALL" in s for a, s in ins if lo <= a <= hi) == 1
    # register-file reads (scripts/sass_rf_model.py): fp64 instructions with three distinct register operands issue in three
    # cycles instead of two on B200 (profiles/r02_issue_probe.txt) — the emitter keeps them rare
    assert sum(n == 3 for *_, n in sm.rf_reads(body)) <= 24


def test_unknown_solver_is_rejected():
    with pytest.raises(rt.CudaBackendError, match="does not support solver"):
        InstanceModel(Program.load(golden_path("qif")), solver="lsoda")


@pytest.mark.parametrize("name", ["qif_rk45", "jrc_rk45", "wc_n8_rk45"])
def test_adaptive_rk45_kernel_compiles_without_spills(name):
    """solver='scipy' (solve_ivp RK45 semantics): seven stage vectors + dense output stay in registers (a few spilled
    words at the 255-register limit are tolerated)."""
    os.environ["PRB_CACHE_DIR"] = "off"
    m = InstanceModel(Program.load(golden_path(name)), solver="scipy")
    assert "typedef double prb_time" in m.kernel.source and "#define PRB_SOLVER 4" in m.kernel.source
    mod = rt.compile_module(m.kernel.source, f"{name}.cu", verbose_ptxas=True)
    import re
    assert all(int(re.search(r"(\d+) bytes spill stores", l).group(1)) <= 64 for l in mod.log.splitlines() if "spill" in l), mod.log


def test_adaptive_solver_rejects_fixed_step_constructs():
    from pyrates_b200.scalarize import UnsupportedModelError
    with pytest.raises(UnsupportedModelError, match="adaptive"):
        InstanceModel(Program.load(golden_path("qsfa_ring_n8")), solver="scipy")        # delay ring buffer
    with pytest.raises(UnsupportedModelError, match="adaptive"):
        InstanceModel(Program.load(golden_path("qif_input")), solver="scipy")           # step-indexed input table


def test_rk45_tableau_is_scipys_and_reaches_the_kernel_exactly():
    import numpy as np
    from oracle import numpy_oracle as orc
    from pyrates_b200 import rk45
    for mine, theirs in ((rk45.A, orc.RK45_A), (rk45.B, orc.RK45_B), (rk45.C, orc.RK45_C), (rk45.E, orc.RK45_E), (rk45.P, orc.RK45_P),
                         (rk45.RK23["B"], orc.RK23_B), (rk45.RK23["C"], orc.RK23_C), (rk45.RK23["E"], orc.RK23_E),
                         (rk45.RK23["P"], orc.RK23_P)):
        assert np.array_equal(np.asarray(mine, dtype=float), theirs)
    assert np.array_equal(np.asarray([r + [0] * (2 - len(r)) for r in rk45.RK23["A"]], dtype=float), orc.RK23_A[:, :2])
    src = InstanceModel(Program.load(golden_path("qif_rk45")), solver="scipy").kernel.source
    line = [l for l in src.splitlines() if l.startswith("#define PRB_RK_P_INIT")][0]
    vals = [float(x) for x in line.split("INIT", 1)[1].replace("{", " ").replace("}", " ").replace(",", " ").split()]
    assert np.array_equal(np.array(vals).reshape(7, 4), orc.RK45_P)          # repr() literals round-trip bit for bit


def test_fast_math_exp_algorithm_is_within_2_ulp_on_the_host():
    """NumPy twin of prb_fast_exp / prb_spec_exp (emit_inst.py): 2^(j/2048) table + degree-3 polynomial, with the very
    constants the kernel source carries."""
    import re
    import numpy as np
    src = InstanceModel(Program.load(golden_path("jrc")), solver="rk4", fast_math=True).kernel.source
    assert "prb_spec_exp(" in src and "prb_spec_div(" in src and "#define PRB_USE_ETAB 1" in src and "void prb_rhs_exact(" in src
    assert "void prb_rhs_spec(" in src and "#define PRB_SPEC_STEP 1" in src            # one speculation check per step
    tab = np.array([float(x) for x in re.search(r"PRB_EXP_TAB\[2048\] = \{([^}]*)\}", src).group(1).split(",")])
    # lean integer side (round 2): entry j carries its high word lowered by j << 9, so that  hi' + (n << 9)  is the high
    # word of 2^k 2^(j/2048) for n = 2048 k + j — undo it for the NumPy twin and check the identity on the raw words
    assert "#define PRB_EXP_LEAN 1" in src and "hi(tj) + (n << (20 - PRB_ETAB_BITS))" in src.replace("__double2hiint", "hi")
    raw = tab.view(np.uint64) + (np.arange(2048, dtype=np.uint64) << np.uint64(41))
    tab = raw.view(np.float64).copy()
    assert np.array_equal(tab, np.exp2(np.arange(2048, dtype=np.longdouble) / 2048).astype(np.float64))
    inv, mllo = [float(x) for x in re.search(r"PRB_EXP_K\[2\] = \{([^}]*)\}", src).group(1).split(",")]
    mlhi = float(re.search(r"#define PRB_EXP_HI \(([^)]*)\)", src).group(1))
    c3 = float(re.search(r"#define PRB_EXP_C3 \(([^)]*)\)", src).group(1))
    assert tab.size == 2048 and abs(tab[1024] - 2 ** 0.5) < 1e-15 and abs(inv * np.log(2) / 2048 - 1) < 1e-15
    # instruction immediates: the low words of the high part of -ln2/2048 and of the cubic coefficient are zero
    assert all((np.float64(v).view(np.uint64) & np.uint64(0xFFFFFFFF)) == 0 for v in (mlhi, c3)) and abs(c3 * 6 - 1) < 1e-6
    assert abs((mlhi + mllo) * 2048 / np.log(2) + 1) < 1e-15
    z = np.random.default_rng(0).uniform(-700, 700, 400_000)
    t = z * inv + 6755399441055744.0
    n = (t.view(np.int64) & 0xFFFFFFFF).astype(np.int64)
    n = np.where(n >= 2 ** 31, n - 2 ** 32, n)
    # the kernel's one-instruction scaling on the stored words (32-bit wrap-around arithmetic) == ldexp(2^(j/2048), k)
    hi_c = ((raw - (np.arange(2048, dtype=np.uint64) << np.uint64(41))) >> np.uint64(32)).astype(np.int64)
    hi_k = (hi_c[n & 2047] + (n << 9)) & 0xFFFFFFFF
    scaled = ((hi_k.astype(np.uint64) << np.uint64(32)) | (raw[n & 2047] & np.uint64(0xFFFFFFFF))).view(np.float64)
    assert np.array_equal(scaled, np.ldexp(tab[n & 2047], n >> 11))
    # PrbBad: |z| <= 700 and not NaN  <=>  hi <= 0x4085e000 as a signed AND hi <= 0xc085e000 as an unsigned 32-bit integer
    # (high-word granularity: the accepted range is |z| < 700 + 2^-11)
    zz = np.concatenate([z[:1000], [700.0, -700.0, 700.0 + 2.0 ** -11, -700.0 - 2.0 ** -11, 0.0, -0.0, 1e300, -1e300,
                                    np.inf, -np.inf, np.nan, -np.nan, 5e-324, -5e-324, 709.0, -745.0]])
    hi = (zz.view(np.uint64) >> np.uint64(32)).astype(np.uint32)
    ok = (hi.view(np.int32) <= 0x4085e000) & (hi <= np.uint32(0xc085e000))
    assert np.array_equal(ok, np.abs(zz) < 700.0 + 2.0 ** -11) and ok[1000] and ok[1001] and not ok[1002] and not ok[1003]
    nd = t - 6755399441055744.0
    assert np.array_equal(nd * mlhi, (nd.astype(np.longdouble) * np.longdouble(mlhi)).astype(np.float64))     # exact product
    r = (z + nd * mlhi) + nd * mllo
    q = r * (1.0 + r * (0.5 + r * c3))
    got = np.ldexp(tab[n & 2047] + tab[n & 2047] * q, n >> 11)
    ref = np.exp(z.astype(np.longdouble))
    assert float(np.max(np.abs(got.astype(np.longdouble) - ref) / ref)) < 2.5 * 2.0 ** -53
    # Estrin form of the speculative kernels and the fused 1 + e^z of sigmoid denominators (prb_spec_exp_plus)
    q2 = r + (r * r) * (0.5 + r * c3)
    got2 = np.ldexp(tab[n & 2047] + tab[n & 2047] * q2, n >> 11)
    assert float(np.max(np.abs(got2.astype(np.longdouble) - ref) / ref)) < 2.5 * 2.0 ** -53
    ts = np.ldexp(tab[n & 2047], n >> 11)
    plus = ts * q2 + (ts + 1.0)
    assert float(np.max(np.abs(plus.astype(np.longdouble) - (ref + 1)) / (ref + 1))) < 2.5 * 2.0 ** -53
    assert "prb_spec_exp_plus(" in src          # defined; used only with PRB_EXP_PLUS=1 (measured: no gain)
    # models with mutable buffers keep the branching variant (a speculative re-evaluation would replay side effects)
    src2 = InstanceModel(Program.load(golden_path("jrc_2delay")), solver="heun", fast_math=True).kernel.source
    assert "prb_fast_div(" in src2 and "#define PRB_SPEC_STEP" not in src2 and "void prb_rhs_exact(" not in src2


@pytest.mark.parametrize("env", [{"PRB_EXP_LEAN": "0"}, {"PRB_EXP_LEAN": "2"}, {"PRB_SPEC_PER_STEP": "0"}, {"PRB_RANGE_SHIFT": "1"},
                                 {"PRB_INTEG_PAIRS": "1"}, {"PRB_MAT_REGFIRST": "0", "PRB_EXP_ESTRIN": "0", "PRB_EXP_PINK0": "0"},
                                 {"PRB_EXP_PLUS": "1"}, {"PRB_BURST_UNROLL": "2"}, {"PRB_RCP_BATCH": "1"}])
def test_measured_but_not_adopted_emitter_variants_still_build(env, monkeypatch):
    """DESIGN.md 2.1 lists emitter variants that were measured on B200 and kept as opt-in switches; each must keep
    compiling for sm_100a (they are the A/B partners of scripts/tune_inst.py)."""
    monkeypatch.setenv("PRB_CACHE_DIR", "off")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sweep = [("u", 0), ("weight_v2", 0), ("weight_v2", 1), ("weight", 0), ("weight_v1", 0)]
    m = InstanceModel(Program.load(golden_path("jrc")), solver="rk4", sweep=sweep, observers=[0], const_div_to_mul=True,
                      fast_math=True, block=64, min_blocks=7)
    assert m.module.cubin[:4] == b"\x7fELF"
    src = m.kernel.source
    if env.get("PRB_SPEC_PER_STEP") == "0":
        assert "#define PRB_SPEC_STEP 1" not in src
    if env.get("PRB_RCP_BATCH") == "1":
        body = src[src.index("void prb_rhs_spec("):src.index("void prb_rhs(")]
        # one reciprocal of the product of the three sigmoid denominators; the fast exp range shrinks so that it stays finite
        assert body.count("prb_spec_rcp_nc(") == 1 and body.count("__dmul_rn(") >= 6 and "#define PRB_EXP_ZHI 0x406b8000" in src
    else:
        assert "#define PRB_EXP_ZHI 0x4085e000" in src
    if env.get("PRB_INTEG_PAIRS") == "1":
        assert "PRB_INTEG[PRB_NS] = {1, -1, 3, -1, 6, 7, -1, -1}" in src      # dV/dt = I chains of the three JRC populations


def test_float32_fast_math_exp_is_a_flushing_exp2_with_folded_log2e(monkeypatch):
    """float32 sweeps: exp(x) -> ex2.approx.ftz(x log2 e) with log2 e absorbed by the argument's coefficients (-560 log2 e)."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench_configs as bc
    sweep = [("u", 0), ("weight_v2", 0), ("weight_v2", 1), ("weight", 0), ("weight_v1", 0)]
    prog = bc.to_float32(Program.load(golden_path("jrc")))
    src = InstanceModel(prog, solver="rk4", sweep=sweep, observers=[0], const_div_to_mul=True, fast_math=True).kernel.source
    body = src[src.index("void prb_rhs("):src.index("void prb_record(")]
    assert body.count("prb_ex2_ftz(") == 3 and "__expf(" not in body and "exp(" not in body.replace("prb_ex2_ftz(", "")
    c = float(np.float32(-560.0) * np.float32(1.4426950408889634))
    assert f"({c!r}f)" in body or f"({float(np.float32(c))!r}f)" in body
    monkeypatch.setenv("PRB_EXPF_FTZ", "0")
    src0 = InstanceModel(prog, solver="rk4", sweep=sweep, observers=[0], const_div_to_mul=True, fast_math=True).kernel.source
    assert "__expf(" in src0[src0.index("void prb_rhs("):src0.index("void prb_record(")]


def test_burst_time_loop_keeps_the_sampling_schedule_of_the_per_step_loop():
    """Python twins of the two time loops of inst_kernel.cuh (PRB_BURST_LOOP and the plain one): for chained launches with
    arbitrary step offsets and decimations they must record the same samples at the same steps and advance the same
    counters — the rule of BaseBackend._solve (row k = state at step k * store_step, sampled BEFORE the update)."""
    import re
    import numpy as np
    src = open(os.path.join(os.path.dirname(rt.__file__), "csrc", "templates", "inst_kernel.cuh")).read()
    # the twins below restate these lines; fail when the template drifts away from them
    for line in ("long long until_store = (store_step - (A.step0 % store_step)) % store_step;", "if (until_store < left) left = until_store;",
                 "if (left > 0x40000000LL) left = 0x40000000LL;", "until_store -= burst;", "for (int s = 0; s < burst; ++s, ++t) {", "i += burst;"):
        assert line in src, line
    assert re.search(r"#ifdef PRB_BURST_LOOP\s+long long i = 0;\s+while \(i < A.n_steps\) \{\s+#else\s+for \(long long i = 0; i < A.n_steps; \+\+i, \+\+t\)", src)

    def plain(step0, n_steps, store_step, sample0):
        until, sample, t, rec, steps = (store_step - step0 % store_step) % store_step, sample0, step0, [], []
        for i in range(n_steps):
            if until == 0:
                rec.append((sample, t)); sample += 1; until = store_step
            until -= 1
            steps.append(t); t += 1
        return rec, steps, sample, t

    def burst(step0, n_steps, store_step, sample0, cap=0x40000000):
        until, sample, t, rec, steps, i = (store_step - step0 % store_step) % store_step, sample0, step0, [], [], 0
        while i < n_steps:
            if until == 0:
                rec.append((sample, t)); sample += 1; until = store_step
            left = min(n_steps - i, until, cap)
            assert left >= 1
            until -= left
            for s in range(left):
                steps.append(t); t += 1
            i += left
        return rec, steps, sample, t

    rng = np.random.default_rng(3)
    for _ in range(300):
        store_step = int(rng.integers(1, 12))
        step0, sample0 = 0, 0
        for n_steps in rng.integers(0, 40, size=4):                       # a chain of launches (BatchedSweep chunks)
            a, b = plain(step0, int(n_steps), store_step, sample0), burst(step0, int(n_steps), store_step, sample0, cap=int(rng.integers(1, 9)))
            assert a == b
            step0, sample0 = a[3], a[2]


def test_register_read_model_counts_distinct_operands_and_reuse_hits():
    """scripts/sass_rf_model.py (the R3 term of the issue model): immediates, constants, uniform registers and operands
    served by the reuse cache of the previous instruction's same slot are free; `-R` / `|R|` modifiers are the same register."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import sass_rf_model as sm
    body = [(0x00, "DFMA R2, R4, R6, R8"),                       # three distinct registers
            (0x10, "DFMA R2, R4.reuse, -560, R8"),                # immediate: two
            (0x20, "DFMA R10, R4, R12, R14"),                     # R4 comes from the reuse cache (slot 0): two
            (0x30, "DFMA R10, -R10, R10, 1"),                     # one register read twice: one
            (0x40, "IMAD.MOV.U32 R20, RZ, RZ, RZ"),               # not on the fp64 pipe (and clears the reuse cache)
            (0x50, "DFMA R2, R4, UR26, R8"),                      # uniform register: two
            (0x60, "DADD R2, R4, c[0x3][0x4000]"),                # constant bank: one
            (0x70, "@!P0 DMUL R2, R4, R6")]                       # predicated: two
    assert [n for *_, n in sm.rf_reads(body)] == [3, 2, 2, 1, 2, 1, 2]
    assert sm.operands("@!P0 DMUL R2, R4, R6") == ("DMUL", ["R2", "R4", "R6"])
