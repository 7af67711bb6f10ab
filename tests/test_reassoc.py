"""fast_math invariant folding (pyrates_b200/reassoc.py): the rewritten DAG is algebraically the same vector field, its
thread-invariant factors are hoisted into prb_thread_init, and the JRC sweep kernel carries the instruction mix DESIGN.md
quotes."""
import numpy as np
import pytest

from conftest import golden_path
from dag_eval import DagMachine
from pyrates_b200.engine import InstanceModel
from pyrates_b200.program import Program
from pyrates_b200.reassoc import fold_invariants, invariant_nodes
from pyrates_b200.scalarize import scalarize

JRC_SWEEP = [("u", 0), ("weight_v2", 0), ("weight_v2", 1), ("weight", 0), ("weight_v1", 0)]
CASES = [("jrc", JRC_SWEEP, 0.01), ("jrc", [], 0.01), ("qif", [("eta", 0)], 0.5), ("qif_sfa", [("eta", 0), ("alpha", 0)], 0.3),
         ("net13", [], 0.5), ("qif_input", [("eta", 0)], 0.3), ("dde_mg", [("bet", 0), ("tau", 0)], 0.1)]


@pytest.mark.parametrize("name,sweep,scale", CASES)
def test_folded_dag_is_the_same_vector_field(name, sweep, scale):
    prog = Program.load(golden_path(name))
    rhs = scalarize(prog, sweep)
    folded = fold_invariants(rhs)
    a, b = DagMachine(rhs), DagMachine(folded)
    rng = np.random.default_rng(1)
    a.params = b.params = a.params * rng.uniform(0.7, 1.3, a.params.shape)
    if prog.is_dde:                                   # a short shared history: the lookups interpolate in it
        from oracle import numpy_oracle as orc
        hist = orc.History(prog.arg("hist").value)
        for j in range(1, 400):
            hist.update(j * 1e-4, prog.y0 * (1 + 0.01 * np.sin(0.05 * j)))
        a.hist = b.hist = hist
    for k in range(8):
        k = 399 + k if prog.is_dde else k
        y = prog.y0 + scale * rng.normal(size=prog.n_state)
        want, got = a.eval(k, y), b.eval(k, y)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12 * np.max(np.abs(want)))


def test_jrc_sweep_kernel_hoists_parameter_products():
    prog = Program.load(golden_path("jrc"))
    m = InstanceModel(prog, solver="rk4", sweep=JRC_SWEEP, observers=[0], const_div_to_mul=True, fast_math=True, block=64,
                      min_blocks=7)
    src = m.kernel.source
    body = src[src.index("void prb_rhs_spec(const prb_time"):src.index("void prb_rhs(const prb_time")]
    init = src[src.index("void prb_thread_init("):src.index("void prb_thread_exit(")]
    # every swept parameter enters the time loop only through hoisted products (T.q), never as T.p
    assert "T.p[" not in body and body.count("T.q[") == 5 and init.count("T.q[") >= 6 and 'asm volatile("" : "+d"(T.q[k]))' in init
    # sigmoids: table exp, bare reciprocal without a range check (the divisor is 1 + exp), one FMA with the hoisted weight
    assert body.count("prb_spec_exp(") == 3 and body.count("prb_spec_rcp_nc(") == 3 and "prb_spec_div(" not in body
    # the exact re-evaluation path still uses the parameters as the reference wrote them
    exact = src[src.index("void prb_rhs_exact("):src.index("void prb_rhs_spec(const prb_time")]
    assert "T.p[" in exact and "T.q[" not in exact
    # models with mutable buffers / exact kernels are not rewritten
    plain = InstanceModel(prog, solver="rk4", sweep=JRC_SWEEP, observers=[0]).kernel.source
    assert "#define PRB_NQ 0" in plain
    ring = InstanceModel(Program.load(golden_path("jrc_2delay")), solver="heun", fast_math=True).kernel.source
    assert "#define PRB_NQ 0" in ring


def test_invariant_classification():
    prog = Program.load(golden_path("jrc"))
    rhs = fold_invariants(scalarize(prog, JRC_SWEEP))
    inv, g = invariant_nodes(rhs), rhs.graph
    assert any(v and g.ops[i] == "mul" for i, v in inv.items())          # hoistable products exist
    assert all(not inv[i] for i in rhs.live_nodes() if g.ops[i] in ("state", "exp"))
