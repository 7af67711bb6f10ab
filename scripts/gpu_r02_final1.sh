#!/bin/bash
# Round 2 final single-GPU records: the full bench line (all configs), the reference arm, the complete GPU test-suite.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err
echo "ref rc=$?"
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_final_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.txt 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r02_final_smoke.txt
