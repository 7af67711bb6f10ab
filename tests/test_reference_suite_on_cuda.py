"""The reference's OWN per-backend tests, run with backend='cuda' (GPU box, needs the installed reference under
baseline/_ref).  These are the known-answer loops SURVEY.md §4 / §8c lists as the golden vectors of the hot path
(tests/test_backend_simulations.py, tests/test_population_connectivity.py, tests/test_implemented_models.py): each takes a
`backend` argument that upstream's conftest parametrises from VALID_BACKENDS — here it is simply 'cuda'."""
import importlib.util
import os
import warnings

import pytest

from oracle.ref_env import import_reference, reference_root

pyrates = import_reference()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(pyrates is None, reason="reference PyRates not importable"),
              pytest.mark.timeout(300)]
warnings.filterwarnings("ignore")

SUITES = {
    "test_backend_simulations": ["test_2_1_operator", "test_2_2_node", "test_2_3_edge", "test_2_4_solver",
                                 "test_2_5_inputs_outputs", "test_2_6_vectorization", "test_2_7_backends"],
    "test_population_connectivity": ["test_population_n1_matches_scalar", "test_population_n4_identity_weight",
                                     "test_population_heterogeneous_params", "test_connectivity_edge_identity",
                                     "test_connectivity_edge_kuramoto_sine", "test_connectivity_dynamic_edge",
                                     "test_matrix_discrete_delay", "test_matrix_gamma_kernel_delay"],
    "test_implemented_models": ["test_3_1_jansenrit", "test_3_2_qif_theta", "test_3_3_wilson_cowan", "test_3_4_kuramoto"],
    # get_jacobian_func: symbolic Jacobians evaluated on the device (test_jacobian_sparse skips itself: the backend declares
    # SUPPORTS_SPARSE_JACOBIAN = False, like upstream's torch / jax backends)
    "test_jacobian": ["test_jacobian_ode_dimensions", "test_jacobian_ode_vs_finite_difference",
                      "test_jacobian_ode_analytical_values", "test_jacobian_ode_transcendentals",
                      "test_jacobian_stateful_edge", "test_jacobian_sparse", "test_jacobian_dde_structure",
                      "test_jacobian_dde_j0_vs_finite_difference"],
}


def _load(module: str):
    path = os.path.join(reference_root(), "tests", module + ".py")
    if not os.path.exists(path):
        pytest.skip(f"{path} is not part of this reference install")
    spec = importlib.util.spec_from_file_location("pyrates_ref_" + module, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("module,name", [(m, n) for m, names in SUITES.items() for n in names])
def test_reference_test_passes_with_cuda_backend(gpu, tmp_path, monkeypatch, module, name):
    import pyrates_b200.register as reg
    assert reg.register()
    monkeypatch.chdir(tmp_path)
    mod = _load(module)
    from pyrates.ir.node import clear_ir_caches

    def reset():
        # what upstream's autouse fixtures do between tests (tests/test_jacobian.py:16-33): templates are cached and
        # update_template(in_place=True) mutates them
        from pyrates import OperatorTemplate
        from pyrates.frontend.template import template_cache
        from pyrates.ir.circuit import in_edge_indices, in_edge_vars
        OperatorTemplate.cache.clear()
        clear_ir_caches()
        in_edge_indices.clear()
        in_edge_vars.clear()
        template_cache.clear()
    reset()
    try:
        getattr(mod, name)("cuda")
    finally:
        reset()
