"""The reference's own parser-level tests (tests/test_backend_parser.py) with backend='cuda'.  They drive the backend's
graph-construction interface only (ctor, get_var, get_op callables used by ComputeGraph.eval_node for constant folding,
index strings) — no device work, so they run in the CPU suite wherever the reference is importable."""
import importlib.util
import os
import warnings

import pytest

from oracle.ref_env import import_reference, reference_root

pyrates = import_reference()
pytestmark = pytest.mark.skipif(pyrates is None, reason="reference PyRates not importable on this machine")
warnings.filterwarnings("ignore")

NAMES = ["test_1_1_expression_parser_init", "test_1_2_expression_parser_parsing_exceptions",
         "test_1_3_expression_parser_math_ops", "test_1_4_expression_parser_funcs", "test_1_5_expression_parser_indexing",
         "test_1_7_equation_parsing"]


@pytest.mark.parametrize("name", NAMES)
def test_reference_parser_test_passes_with_cuda_backend(tmp_path, monkeypatch, name):
    import pyrates_b200.register as reg
    assert reg.register()
    monkeypatch.chdir(tmp_path)
    path = os.path.join(reference_root(), "tests", "test_backend_parser.py")
    if not os.path.exists(path):
        pytest.skip(f"{path} is not part of this reference install")
    spec = importlib.util.spec_from_file_location("pyrates_ref_test_backend_parser", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from pyrates.ir.node import clear_ir_caches
    clear_ir_caches()
    try:
        getattr(mod, name)("cuda")
    finally:
        clear_ir_caches()
