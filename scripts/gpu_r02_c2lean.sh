#!/bin/bash
# Round 2, lean C2 kernel: headline bench line (C2 only), then one ncu --set full capture of prb_integrate from the same command.
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/r02_lean_bench_c2.json 2> gpurun_out/r02_lean_bench_c2.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_lean_bench_c2.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^prb_integrate$ -s 40 -c 1 -f -o gpurun_out/r02_lean_c2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/r02_lean_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_lean_ncu.log
