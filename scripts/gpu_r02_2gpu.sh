#!/bin/bash
# Round 2, 2-GPU call: multi-GPU parity tests (row-sharded kernels, LL exchange), torchrun bench, small-N scaling probes.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q -m gpu -p no:cacheprovider -rs > gpurun_out/r02_multigpu_tests.txt 2>&1
echo "multigpu tests rc=$?"; tail -8 gpurun_out/r02_multigpu_tests.txt
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --configs c3,c3_rk4,c4,c4_2x2048,c5,c5_rk4,c5a,strong,d2h > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "bench n2 rc=$?"; tail -c 600 gpurun_out/r02_bench_n2.err; head -c 600 gpurun_out/r02_bench_n2.json
for cfg in "c3 2048 euler 2000" "c3 1024 euler 2000" "c4 1024 heun 2000" "c3 8192 rk4 200"; do
  set -- $cfg
  timeout 300 $TR --master-port 29512 bench_configs.py --config $1 --size $2 --solver $3 --steps $4 --check --cpu-steps 10 >> gpurun_out/r02_cfg_n2.jsonl 2>> gpurun_out/r02_cfg_n2.err
  timeout 300 python bench_configs.py --config $1 --size $2 --solver $3 --steps $4 --check --cpu-steps 10 >> gpurun_out/r02_cfg_n1.jsonl 2>> gpurun_out/r02_cfg_n1.err
done
timeout 300 python bench_configs.py --config c3 --size 8192 --solver euler --steps 500 --cpu-steps 0 >> gpurun_out/r02_cfg_n1.jsonl 2>> gpurun_out/r02_cfg_n1.err
cut -c1-330 gpurun_out/r02_cfg_n2.jsonl; cut -c1-330 gpurun_out/r02_cfg_n1.jsonl
