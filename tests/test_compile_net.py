"""Network-model kernels build for sm_100a without a GPU (NVRTC): one per coupling family, plus the multi-GPU
variants (row shards, peer stores, cross-GPU barrier) for world sizes 2 and 3."""
import os

import numpy as np
import pytest

from conftest import golden_path
from pyrates_b200 import runtime as rt
from pyrates_b200.netcompile import netcompile
from pyrates_b200.netengine import NetworkModel
from pyrates_b200.program import Program
from pyrates_b200.scalarize import UnsupportedModelError

CASES = [("wc_n128", "rk4"), ("wc_n128_f32", "euler"), ("qsfa_both_n96", "heun"), ("kuramoto_n128", "heun_true"),
         ("kuramoto_scalar_n128", "euler"), ("jrc_grid_b1024", "rk4"), ("jrc_2delay", "heun"), ("net13", "euler")]


@pytest.mark.parametrize("name,solver", CASES)
def test_network_kernel_compiles(name, solver):
    os.environ["PRB_CACHE_DIR"] = "off"
    m = NetworkModel(Program.load(golden_path(name)), solver=solver)
    assert "prb_integrate_net" in m.kernel.source and "prb_eval" in m.kernel.source
    mod = rt.compile_module(m.kernel.source, f"{name}.cu", verbose_ptxas=True)
    assert mod.cubin[:4] == b"\x7fELF"


def test_phase_structure():
    """State-sourced dense coupling needs no intra-evaluation barrier; gather edges over a computed vector need one;
    read-before-write persistent buffers (WAR) force an extra phase."""
    assert netcompile(Program.load(golden_path("wc_n128"))).n_phases == 1
    assert netcompile(Program.load(golden_path("kuramoto_n128")), separable=False).n_phases == 1
    # separable rewrite: phase 0 materialises sin/cos(theta), phase 1 = two GEMVs sharing one W pass
    ir = netcompile(Program.load(golden_path("kuramoto_n128")))
    assert ir.n_phases == 2 and [r.simple[0] for r in ir.reductions] == ["weight", "weight"]
    assert netcompile(Program.load(golden_path("jrc_grid_b16"))).n_phases == 2
    assert netcompile(Program.load(golden_path("jrc_2delay"))).n_phases == 3


@pytest.mark.parametrize("world", [2, 3])
def test_multi_gpu_variants_compile_and_shard_constants(world):
    os.environ["PRB_CACHE_DIR"] = "off"
    prog = Program.load(golden_path("qsfa_both_n96"))
    sizes = []
    for rank in range(world):
        m = NetworkModel(prog, solver="rk4", rank=rank, world=world)
        assert f"#define PRB_WORLD {world}" in m.kernel.source and "st.volatile.global.v4.u32" in m.kernel.source
        rt.compile_module(m.kernel.source, f"mg{rank}.cu")
        sizes.append(m.kernel.const_init.size)
    full = NetworkModel(prog, solver="rk4").kernel.const_init.size
    assert max(sizes) < 0.6 * full if world == 2 else max(sizes) < 0.45 * full      # W rows are sharded


def test_legacy_per_edge_ring_gather_lowers_to_a_head_relative_gather():
    """`a_buffered = a_buffer[source_idx, a_delays]` (pyrates/ir/circuit.py:598-641, the legacy per-edge delays): one
    gathered ring read per edge, column = delay relative to the circular head, no barrier (every delay >= 1 reads a column
    written in an earlier evaluation)."""
    ir = netcompile(Program.load(golden_path("net10")))
    assert ir.n_phases == 1 and any(op == "ldringg" for op in ir.graph.ops)
    m = NetworkModel(Program.load(golden_path("net10")), solver="heun")
    assert "prb_ring_pos(C.head[0], (int)__ldg(C.iconsts +" in m.kernel.source and m.module.cubin[:4] == b"\x7fELF"


@pytest.mark.parametrize("name,sweep,nb,bt", [("wc_n300_f32", ["r_ext"], 256, 256), ("wc_n300_f32", ["r_ext"], 384, 128), ("qsfa_both_n96_f32", ["I_ext"], 128, 64),
                                              ("kuramoto_n128_f32", ["s_ext"], 96, 32)])
def test_tensor_core_sweep_kernels_compile(name, sweep, nb, bt):
    """float32 sweep-batched networks take the tcgen05 path: the cubin holds UTC*MMA / LDTM / UBLKCP (checked in SASS
    when cuobjdump is available) and the connectivity is packed as hi / lo K-major tiles."""
    os.environ["PRB_CACHE_DIR"] = "off"
    m = NetworkModel(Program.load(golden_path(name)), solver="heun", n_batch=nb, sweep=sweep)
    k = m.kernel
    assert k.tc is not None and k.tc["bt"] == bt and k.smem_bytes == k.tc["smem"] > 48 * 1024
    assert "tcgen05.mma.cta_group::1.kind::tf32" in k.source and "prb_tc_accumulate" in k.source
    mod = rt.compile_module(k.source, f"{name}_tc.cu")
    assert mod.cubin[:4] == b"\x7fELF"
    import shutil, subprocess, tempfile
    if shutil.which("cuobjdump"):
        with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
            f.write(mod.cubin)
            f.flush()
            sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout
        assert "UTCHMMA" in sass and "LDTM" in sass and "UBLKCP" in sass


def test_fp64_sweep_kernel_uses_dmma_and_streaming_gemv_uses_wide_nonallocating_loads():
    """SASS evidence without a GPU: the float64 sweep-batched coupling runs on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64)
    and the single-instance GEMV streams W with 128-bit loads that do not allocate in L1 (LDG.E.128 ... no_allocate)."""
    import shutil, subprocess, tempfile
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    os.environ["PRB_CACHE_DIR"] = "off"

    def sass_of(model, tag):
        mod = rt.compile_module(model.kernel.source, f"{tag}.cu")
        with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
            f.write(mod.cubin)
            f.flush()
            return subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout
    prog = Program.load(golden_path("wc_n128"))
    batched = NetworkModel(prog, solver="rk4", n_batch=128, sweep=["r_ext"])
    assert batched.kernel.tc is None and "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64" in batched.kernel.source
    assert "DMMA" in sass_of(batched, "wc_dmma")
    single = NetworkModel(prog, solver="euler")
    sass = sass_of(single, "wc_gemv")
    assert any("LDG" in l and ".128" in l for l in sass.splitlines()), "expected 128-bit global loads in the W stream"


def test_tensor_core_path_is_float32_only():
    prog = Program.load(golden_path("wc_n128"))
    assert NetworkModel(prog, solver="euler", n_batch=128, sweep=["r_ext"]).kernel.tc is None      # fp64 -> FMA tiles
    with pytest.raises(UnsupportedModelError):
        NetworkModel(prog, solver="euler", n_batch=128, sweep=["r_ext"], tensor_core=True)


def test_operand_tiling_roundtrip_and_split():
    from pyrates_b200 import tcgemm
    rng = np.random.default_rng(0)
    m = rng.normal(size=(300, 70)).astype(np.float32)
    t = tcgemm.tile_operand(m, 128)
    assert t.size == 3 * 3 * 128 * 32
    assert np.array_equal(tcgemm.untile_operand(t, 300, 70, 128), m)
    r, k = 37, 21                                              # documented offset formula (tc_gemm.cuh)
    assert t[((k // 4) * 16 + r // 8) * 32 + (r % 8) * 4 + k % 4] == m[r, k]
    hi, lo = tcgemm.split_tf32(m)
    assert np.array_equal(hi + lo, m) and not np.any(hi.view(np.uint32) & 0x1FFF)
    assert np.max(np.abs(lo)) <= np.max(np.abs(m)) * 2.0 ** -10


@pytest.mark.parametrize("name", ["wc_n128_rk45", "kuramoto_n128_rk45"])
def test_adaptive_network_kernel_compiles(name):
    os.environ["PRB_CACHE_DIR"] = "off"
    m = NetworkModel(Program.load(golden_path(name)), solver="scipy")
    assert "typedef double prb_time" in m.kernel.source and "#define PRB_SOLVER 4" in m.kernel.source
    assert m.kernel.work_elems >= 11 * m.kernel.n_state            # Y, YN, two stage inputs, K0..K6
    mod = rt.compile_module(m.kernel.source, f"{name}.cu", verbose_ptxas=True)
    assert all("0 bytes spill stores" in l for l in mod.log.splitlines() if "spill" in l), mod.log
    with pytest.raises(UnsupportedModelError, match="adaptive"):
        NetworkModel(Program.load(golden_path("wc_n128_f32")), solver="scipy")
    with pytest.raises(UnsupportedModelError, match="adaptive"):
        NetworkModel(Program.load(golden_path("qsfa_both_n96")), solver="scipy")


def test_multi_gpu_exchange_regions_cover_every_shared_store():
    """Row-sharded kernels: every store another rank must see (stage vectors, materialised intermediates, ring columns,
    scatter targets) is sent as an LL packet by its owner and fetched by the receiving grid right before the next grid
    barrier — the emitter must list an unpack call for each stored region, with this rank's row bounds, and no
    system-scope fence may remain on the data path."""
    import re
    cases = {"kuramoto_n128": ["prb_ll_fetch(C, 0LL, q - 0LL, 32LL, 32LL)", "prb_ll_fetch(C, 128LL, q - 96LL, 32LL, 32LL)",
                               "prb_stage_target<STAGE>(C) + 0LL"],
             "qsfa_both_n96": ["prb_ll_fetch_ring(C,", "prb_stage_target<STAGE>(C) + 1056LL"],
             "jrc_2delay": ["prb_ll_fetch_scatter(C,"], "net13": ["prb_ll_fetch_scatter(C,"]}
    for name, needles in cases.items():
        prog = Program.load(golden_path(name))
        m = NetworkModel(prog, solver="heun", rank=1, world=4)
        src = m.kernel.source
        body = src[src.rindex("template <int STAGE> __device__ void prb_eval"):]
        for n in needles:
            assert n in body, (name, n)
        n_dy = len(re.findall(r"prb_emit_k<STAGE>\(", body))
        assert len(re.findall(r"prb_ll_fetch\(C, prb_stage_target<STAGE>", body)) == n_dy       # one region per dy store
        # every fetch loop precedes the barrier of its phase; the last phase's regions are fetched before prb_eval returns
        assert body.rstrip().endswith("}") and "prb_ll_fetch" in body.split("prb_grid_barrier(C);")[-1]
        assert body.count("// fetch what the peers own") == m.kernel.n_phases
        assert "__threadfence_system" not in src and m.kernel.ll_slots > 0
        assert m.module.cubin[:4] == b"\x7fELF"
    single = NetworkModel(Program.load(golden_path("kuramoto_n128")), solver="heun")
    assert "prb_ll_fetch" not in single.kernel.source[single.kernel.source.rindex("template <int STAGE> __device__ void prb_eval"):]


def test_few_rows_per_gpu_share_a_row_between_warps():
    """Row splitting (emit_net.row_split): a shard with fewer rows than warps gives every row to S consecutive warps of a CTA
    (partial sums through shared memory); enough rows -> the classic warp-per-row loop.  Both forms build for sm_100a."""
    import bench_configs as bc
    prog = bc.make_c3(2048)
    one = NetworkModel(prog, solver="euler")                               # 2048 rows on 2368 warps: warp per row
    assert "warp per row" in one.kernel.source and "warps per row" not in one.kernel.source
    eight = NetworkModel(prog, solver="euler", rank=3, world=8)            # 256 rows per GPU: 8 warps per row
    assert "// 8 warps per row, n = 2048 (256 rows on this GPU)" in eight.kernel.source
    assert "prb_part0_2048[1][16]" in eight.kernel.source and eight.module.cubin[:4] == b"\x7fELF"
    off = NetworkModel(prog, solver="euler", rank=3, world=8, split_rows=False, dot_deep=False)
    assert "warps per row" not in off.kernel.source and "#define PRB_DOT_DEEP 0" in off.kernel.source
    c4 = NetworkModel(bc.make_c4(1024), solver="heun")                      # two dots per row, 1024 rows: 2 warps per row
    assert "// 2 warps per row, n = 1024" in c4.kernel.source and c4.module.cubin[:4] == b"\x7fELF"
    assert "prb_row_dot_sm(" not in c4.kernel.source.split("prb_eval(PrbNet& C) {")[1] and c4.kernel.smem_bytes == 0
    # opt-in W-stationary form (measured slower, profiles/r02_wcache.jsonl): 8 teams per CTA x (1024 + 1024) fp64 of W rows =
    # 128 KB of dynamic shared memory, filled once per launch
    ws = NetworkModel(bc.make_c4(1024), solver="heun", w_stationary=True)
    assert "W rows cached in shared memory (bit 0)" in ws.kernel.source and ws.kernel.smem_bytes == 8 * 2048 * 8
    assert "prb_row_dot_sm(wsm +" in ws.kernel.source and "C.wcache |= 1u;" in ws.kernel.source and ws.module.cubin[:4] == b"\x7fELF"
    big = NetworkModel(bc.make_c3(8192), solver="euler", rank=0, world=8, w_stationary=True)   # 8 x 64 KB per CTA: does not fit
    assert "W rows cached" not in big.kernel.source


def test_device_generated_connectivity_twin_and_constant_layout():
    """devgen.DeviceUniform: the NumPy twin is a pure function of (seed, global element index) — any row shard equals the
    slice of the whole — and the emitter puts the generated matrix behind the uploaded constants, row-sharded per rank."""
    import numpy as np
    from pyrates_b200.devgen import DeviceUniform
    from pyrates_b200.program import Arg
    g = DeviceUniform((12, 5), 2.0 / 12, 2)
    full = g.host()
    assert full.shape == (12, 5) and 0.0 <= full.min() and full.max() < 2.0 / 12 and len(np.unique(full)) == 60
    np.testing.assert_array_equal(g.host(3, 9), full[3:9])
    np.testing.assert_array_equal(DeviceUniform((12, 5), 2.0 / 12, 2).host(), full)
    assert not np.array_equal(DeviceUniform((12, 5), 2.0 / 12, 3).host(), full)
    # known answer: SplitMix64 of seed 0 at index 0 is 0xE220A8397B1DCDAF
    assert DeviceUniform((1, 1), 1.0, 0).host()[0, 0] == (0xE220A8397B1DCDAF >> 11) * 2.0 ** -53
    p = Program.load(golden_path("kuramoto_n128"))
    lazy = Program([Arg(a.name, a.kind, DeviceUniform((128, 128), 2.0 / 128, 2) if a.name == "weight" else a.value, a.label)
                    for a in p.args], p.stmts, {}, p.float_precision, p.func_name)
    for world, rank in ((1, 0), (4, 3)):
        k = NetworkModel(lazy, solver="euler", rank=rank, world=world).kernel
        (off, gen, lo, hi), = k.const_fills
        assert off == k.const_init.size and k.const_elems == off + (hi - lo) * 128 and (lo, hi) == ((0, 128) if world == 1 else (96, 128))
    from pyrates_b200.devgen import _fill_module
    assert _fill_module("float32").cubin[:4] == b"\x7fELF"


def test_long_full_reductions_run_in_two_levels():
    """`sum(x)` over >= 8192 elements: per-CTA partial sums + one grid barrier + sum of the partials, instead of every CTA
    walking the whole vector; across GPUs each rank reduces only the rows it owns and the partials travel as LL packets."""
    import bench_configs as bc
    short = NetworkModel(Program.load(golden_path("kuramoto_scalar_n128")), solver="euler")
    assert "level 1" not in short.kernel.source
    prog = bc.make_c5a(8192)
    one = NetworkModel(prog, solver="euler")
    body = one.kernel.source[one.kernel.source.rindex("template <int STAGE> __device__ void prb_eval"):]
    assert "level 1: rows 0..8192 of 8192" in body and "level 2: sum of the 1 x 256 partials" in body and body.count("prb_grid_barrier(C);") == 1
    four = NetworkModel(prog, solver="euler", rank=2, world=4)
    body = four.kernel.source[four.kernel.source.rindex("template <int STAGE> __device__ void prb_eval"):]
    assert "level 1: rows 4096..6144 of 8192" in body and "level 2: sum of the 4 x 256 partials" in body
    assert "prb_ll_fetch(C, 0LL, q - 0LL, 512LL, 256LL)" in body        # partials of ranks 0, 1 | 3 around this rank's 256 slots
    assert four.module.cubin[:4] == b"\x7fELF"


def test_interpolated_inputs_lower_to_interp_nodes_over_replicated_tables():
    """Adaptive runs feed timed inputs as `interp(t, time, u)` / `interp_rows(t, time, u[steps, cols])`
    (frontend/template/circuit.py:1692-1709): one node over two constant tables, evaluated with numpy.interp's expression;
    a vector of interpolated inputs that feeds a dot() is materialised first (one more phase)."""
    a = NetworkModel(Program.load(golden_path("qif_input_rk45")), solver="scipy")
    assert "prb_interp(C.consts +" in a.kernel.source and a.kernel.n_phases == 1
    b = NetworkModel(Program.load(golden_path("qif2_input_rk45")), solver="scipy")
    assert ", 2, 1000, ((real)C.t))" in b.kernel.source and b.kernel.n_phases == 2     # column stride 2, 1000 knots
    assert b.ir.arrays["I_ext_input"].replicate and b.module.cubin[:4] == b"\x7fELF"


def test_small_exchanges_use_the_redundant_fetch_form():
    """Multi-GPU exchange forms.  Default: the grid shares the polling of the peers' packets and meets at the barrier afterwards.
    Opt-in (ll_redundant=True, measured: profiles/r02_redundant_n2.jsonl): up to LL_REDUNDANT_MAX peer-owned elements per barrier
    every CTA arrives EARLY, polls all of them itself and then waits.  Either way every barrier is one arrive + one wait."""
    import re
    import bench_configs as bc
    small = NetworkModel(bc.make_c3(2048), solver="rk4", rank=1, world=8, ll_redundant=True)   # 1792 peer-owned elements per barrier
    src = small.kernel.source
    body = src[src.rindex("template <int STAGE> __device__ void prb_eval"):]
    assert "#define PRB_LL_REDUNDANT 1" in src and "@@" not in src
    assert "prb_barrier_arrive(C);" in body and "q = (long long)threadIdx.x; q < 1792LL; q += (long long)blockDim.x" in body
    assert "prb_grid_barrier(C);" not in body                   # the wait of the last phase sits in prb_stage
    big = NetworkModel(bc.make_c3(8192), solver="rk4", rank=1, world=8, ll_redundant=True)     # 7168 elements: cooperative form
    src = big.kernel.source
    body = src[src.rindex("template <int STAGE> __device__ void prb_eval"):]
    assert "#define PRB_LL_REDUNDANT 0" in src and "q = C.gtid; q < 7168LL; q += C.gsize" in body and "prb_barrier_arrive" not in body
    two = NetworkModel(Program.load(golden_path("kuramoto_n128")), solver="euler", rank=0, world=2, ll_redundant=True)   # two phases
    body = two.kernel.source[two.kernel.source.rindex("template <int STAGE> __device__ void prb_eval"):]
    assert len(re.findall(r"prb_barrier_arrive\(C\);", body)) == 2 and len(re.findall(r"prb_barrier_wait\(C\);", body)) == 1
    off = NetworkModel(bc.make_c3(2048), solver="rk4", rank=1, world=8)                        # the default: cooperative
    assert "#define PRB_LL_REDUNDANT 0" in off.kernel.source
    for m in (small, big, two, off):
        assert m.module.cubin[:4] == b"\x7fELF"
