#!/bin/bash
# Round 2, after the lean C2 kernel: bench line for the records the instance kernels feed (C2 headline, C1, C2 f32, strong
# scaling fields, public API), smoke(), and the GPU tests that run instance kernels through the public boundary.
mkdir -p gpurun_out
timeout 400 python bench.py --configs c1,c2_f32,public_api,d2h > gpurun_out/r02_final2_bench_n1.json 2> gpurun_out/r02_final2_bench_n1.err
echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final2_smoke.txt 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_final2_smoke.txt
timeout 420 python -m pytest tests/test_gpu_named_sizes.py tests/test_backend_boundary.py tests/test_reference_suite_on_cuda.py -x -q -m gpu -p no:cacheprovider --durations=6 > gpurun_out/r02_final2_pytest.txt 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02_final2_pytest.txt
