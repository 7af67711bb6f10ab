#!/bin/bash
# Round 2 closing records on one B200: the full bench line with the final kernels, then the ncu launch list of the C2-only command.
mkdir -p gpurun_out
timeout 420 python bench.py > gpurun_out/r02_final3_bench_n1.json 2> gpurun_out/r02_final3_bench_n1.err
echo "bench rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_lean_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/r02_lean_launches.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_lean_launches.csv | cut -c1-200
