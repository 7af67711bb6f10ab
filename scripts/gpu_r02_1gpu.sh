#!/bin/bash
# Round 2, first single-GPU call: new parity tests at the named sizes, then the full bench line, then the reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r02_gpu.txt 2>&1
nproc >> gpurun_out/r02_gpu.txt; free -g | head -2 >> gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_named_sizes.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_named_sizes.txt 2>&1
echo "named sizes rc=$?"; tail -5 gpurun_out/r02_named_sizes.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_n1.err; head -c 3000 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
echo "ref rc=$?"; head -c 1500 gpurun_out/r02_bench_ref.json
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_fullsize.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_net_full.txt 2>&1
echo "net/full rc=$?"; tail -3 gpurun_out/r02_net_full.txt
